"""Per-call wall time of aerobulk_model with plain (pageable) numpy arrays, then with the SAME arrays page-locked
through aerobulk_gpu_host_register -- what a caller that only relinks gains by registering its fields once."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aerobulk_b200 as ab
from aerobulk_b200 import synth
NI, NJ, NT = 1440, 720, 24
n = NI * NJ
IN = ("sst", "t_zt", "hum_zt", "U_zu", "V_zu", "slp")
OUT = ("QL", "QH", "Tau_x", "Tau_y", "Evap", "T_s")
f = synth.fields(NI, NJ)
np_in = {k: np.array(f[k], order="F") for k in IN + ("rad_sw", "rad_lw")}
np_out = {k: np.zeros((NI, NJ), order="F") for k in OUT}
ab.set_verbose(False)

def session():
    ab.reset()
    ts = []
    for jt in range(1, NT + 1):
        t0 = time.perf_counter()
        ab.aerobulk_model(jt, NT, "coare3p6", 2., 10., *[np_in[k] for k in IN], Niter=5, l_use_skin=True,
                          rad_sw=np_in["rad_sw"], rad_lw=np_in["rad_lw"], out=np_out)
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts[2:-1]))

a = min(session() for _ in range(3))
ref = {k: v.copy() for k, v in np_out.items()}
t0 = time.perf_counter()
for v in list(np_in.values()) + list(np_out.values()):
    ab.host_register(v)
treg = (time.perf_counter() - t0) * 1e3
b = min(session() for _ in range(3))
same = all(np.array_equal(ref[k], np_out[k]) for k in OUT)
for v in list(np_in.values()) + list(np_out.values()):
    ab.host_unregister(v)
c = session()
print(f"bounce={os.environ.get('AEROBULK_GPU_BOUNCE','1')} threads={os.environ.get('AEROBULK_GPU_HOST_THREADS','auto')} chunk={os.environ.get('AEROBULK_GPU_BOUNCE_CHUNK_POINTS','180224')}: pageable {a:.3f} ms/call ({n / a / 1e3:.0f} Mpt/s) | registered {b:.3f} ms/call ({n / b / 1e3:.0f} Mpt/s), "
      f"one-off registration of 14 arrays {treg:.1f} ms, results identical: {same} | after unregister {c:.3f} ms/call")
