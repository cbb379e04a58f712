"""One bench case in a loop, for ncu (tools/ncu_bench_cases.sh) -- the SAME calls bench.py times, at 1440x720 (the generator
is latitude-based and scale-free: the same distribution of points as the 12960x6480 grid):
    c5:<algo>[+skin]     8 independent jt = Nt = 1 sessions (BASELINE C5 / C3 / C1 / C4 at nb_iter = 5)
    c4:<algo>:<nb_iter>  the same with another iteration count (BASELINE C4 sweep)
    c2                   two 24-step COARE 3.6 + skin sessions with the diurnal short-wave of BASELINE C2 (state carried)
usage: python tools/prof_case.py <case> [Ni Nj]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aerobulk_b200 as ab
from aerobulk_b200 import synth

case = sys.argv[1]
NI, NJ = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1440, 720)
n = NI * NJ
IN = ("sst", "t_zt", "hum_zt", "U_zu", "V_zu", "slp")
OUT = ("QL", "QH", "Tau_x", "Tau_y", "Evap", "T_s")
f = synth.fields(NI, NJ)
dev = {k: torch.from_numpy(np.ravel(v, order="F").copy()).cuda() for k, v in f.items()}
out = {k: torch.empty(n, dtype=torch.float64, device="cuda") for k in OUT}
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
ab.reset()
ab.set_verbose(False)
ab.set_stream(st.cuda_stream)
ab.set_async(True)
kind, *rest = case.split(":")
if kind in ("c5", "c4"):
    algo, skin = (rest[0][:-5], True) if rest[0].endswith("+skin") else (rest[0], False)
    nb = int(rest[1]) if len(rest) > 1 else 5
    o = out if skin else {k: out[k] for k in OUT[:5]}
    kw = dict(l_use_skin=True, rad_sw=dev["rad_sw"], rad_lw=dev["rad_lw"]) if skin else {}
    for _ in range(8):
        ab.new_session()
        ab.aerobulk_model_device(1, 1, algo, 2., 10., *[dev[k] for k in IN], out=o, Niter=nb, shape=(NI, NJ), **kw)
    ab.synchronize()
elif kind == "c2":
    rsw = [torch.from_numpy(np.ravel(synth.rad_sw_hour(NI, NJ, jt), order="F").copy()).cuda() for jt in range(1, 25)]
    for _ in range(2):
        ab.new_session()
        for jt in range(1, 25):
            ab.aerobulk_model_device(jt, 24, "coare3p6", 2., 10., *[dev[k] for k in IN], out=out, Niter=5, l_use_skin=True,
                                     rad_sw=rsw[jt - 1], rad_lw=dev["rad_lw"], shape=(NI, NJ))
        ab.synchronize()
else:
    raise SystemExit("unknown case " + case)
print("done", case)
