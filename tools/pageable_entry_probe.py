"""Wall time per call of aerobulk_gpu_turb / aerobulk_gpu_oce_ice on 1 M-point PAGEABLE numpy arrays (run once with
AEROBULK_GPU_BOUNCE=0 and once without to see what the copy threads + pinned slab buy)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aerobulk_b200 as ab
from aerobulk_b200 import synth
NI, NJ = 1440, 720
f = synth.fields(NI, NJ)
tc = f["sst"] - 273.15
es = 611.2 * np.exp(17.67 * tc / (tc + 243.5))
ssq = np.asfortranarray(0.98 * 0.622 * es / (f["slp"] - 0.378 * es))
theta = np.asfortranarray(f["t_zt"] + 0.0196)
wnd = np.asfortranarray(np.hypot(f["U_zu"], f["V_zu"]))
ab.set_verbose(False)
ab.reset()

def best(fn, reps=6):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)

t_turb = best(lambda: ab.turb("coare3p6", 1, 2.0, 10.0, f["sst"], theta, ssq, f["hum_zt"], wnd))
g = synth.ice_fields(NI * NJ)
t_ice = best(lambda: ab.oce_ice("nemo", "ecmwf", 2.0, 10.0, **g, want=("Tau", "QH", "QL", "Evap", "QH_i", "QL_i", "Tau_i", "QH_w", "QL_w", "Tau_w")))
print(f"bounce={os.environ.get('AEROBULK_GPU_BOUNCE', '1')}: turb coare3p6 (5 in, 8 out) {t_turb:.2f} ms/call | "
      f"oce_ice nemo+ecmwf (7 in, 10 out) {t_ice:.2f} ms/call   [python wrapper included]")
