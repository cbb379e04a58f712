"""Kernel micro-benchmark: device-resident time per aerobulk_model call for every kernel variant.
usage: python tools/kbench.py [Ni Nj] ; prints one line per (algo, skin, day/night)."""
import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aerobulk_b200 as ab
from aerobulk_b200 import synth

NI = int(sys.argv[1]) if len(sys.argv) > 2 else 1440
NJ = int(sys.argv[2]) if len(sys.argv) > 2 else 720
n = NI * NJ
IN = ("sst", "t_zt", "hum_zt", "U_zu", "V_zu", "slp")
OUT = ("QL", "QH", "Tau_x", "Tau_y", "Evap", "T_s")
f = synth.fields(NI, NJ)
dev = {k: torch.from_numpy(np.ravel(v, order="F").copy()).cuda() for k, v in f.items()}
rsw_day = torch.from_numpy(np.ravel(synth.rad_sw_hour(NI, NJ, 12), order="F").copy()).cuda()
if os.environ.get("KBENCH_SINGLE"):
    rsw_day = dev["rad_sw"]       # the C5 forcing: 400 (0.3 + 0.7 u) W/m2
rsw_night = torch.zeros(n, dtype=torch.float64, device="cuda")
out = {k: torch.empty(n, dtype=torch.float64, device="cuda") for k in OUT}
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
ab.set_stream(st.cuda_stream)
from aerobulk_b200 import model as abm
abm.set_sort(int(os.environ.get('KBENCH_SORT', '1')))
peak = ab.measure_fp64_peak()
print(f"lib {os.environ.get('AEROBULK_GPU_LIB','default')}  DFMA peak {peak/1e12:.2f} T instr/s")
res = []
cases = [(a, False, None) for a in ("ncar", "andreas", "coare3p0", "coare3p6", "ecmwf")] + \
        [(a, True, r) for a in ("coare3p0", "coare3p6", "ecmwf") for r in ("night", "day")]
if os.environ.get("KBENCH_ONE"):
    a_, s_, r_ = os.environ["KBENCH_ONE"].split(",")
    cases = [(a_, s_ == "1", r_ if r_ != "-" else None)]
elif os.environ.get("KBENCH_QUICK"):
    cases = [("ncar", False, None), ("andreas", False, None), ("coare3p6", False, None), ("coare3p6", True, "night"), ("coare3p6", True, "day"), ("ecmwf", True, "day")]
for algo, skin, rad in cases:
    ab.reset(); ab.set_stream(st.cuda_stream)
    kw = dict(Niter=int(os.environ.get("KBENCH_NITER", "5")))
    o = {k: out[k] for k in OUT[:5]}
    if skin:
        kw.update(l_use_skin=True, rad_sw=rsw_day if rad == "day" else rsw_night, rad_lw=dev["rad_lw"])
        o = out
    NT = 12
    ts = []
    single = bool(os.environ.get("KBENCH_SINGLE"))      # every timed call is its own session (jt = Nt = 1), as in BASELINE C5
    ab.set_verbose(False)
    if single:
        ab.set_async(True)
    for jt in range(1, NT + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if single:
            ab.new_session()
        e0.record(st)
        if single:
            ab.aerobulk_model_device(1, 1, algo, 2., 10., *[dev[k] for k in IN], out=o, shape=(NI, NJ), **kw)
        else:
            ab.aerobulk_model_device(jt, NT, algo, 2., 10., *[dev[k] for k in IN], out=o, shape=(NI, NJ), **kw)
        e1.record(st)
        torch.cuda.synchronize()
        if 3 <= jt < NT:
            ts.append(e0.elapsed_time(e1))
    if single:
        ab.synchronize()
        ab.set_async(False)
    ms = float(np.median(ts))
    W = ab.work_per_point(algo, skin, 5)
    print(f"{algo:9s} skin={int(skin)} {rad or '-':5s}: {ms:7.3f} ms  {n/ms/1e3:8.1f} Mpt/s  W-frac {W*n/(ms*1e-3)/peak:5.2f}")
