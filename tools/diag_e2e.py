"""Diagnostic: per-call timings of the host-array and device-resident paths over several sessions."""
import os, sys, time, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aerobulk_b200 as ab
from aerobulk_b200 import synth
from aerobulk_b200 import model as abm

NI, NJ, NT = 1440, 720, 24
n = NI * NJ
IN = ("sst", "t_zt", "hum_zt", "U_zu", "V_zu", "slp")
OUT = ("QL", "QH", "Tau_x", "Tau_y", "Evap", "T_s")
f = synth.fields(NI, NJ)
rsw = [synth.rad_sw_hour(NI, NJ, jt) for jt in range(1, NT + 1)]

def smi():
    try:
        return subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.active,pcie.link.gen.current,pcie.link.width.current",
                                        "--format=csv,noheader"], text=True).strip()
    except Exception as e:
        return str(e)

def pinned(a):
    t = torch.empty(n, dtype=torch.float64).pin_memory()
    t.numpy()[:] = np.ravel(a, order="F")
    return t
host = {k: pinned(f[k]) for k in IN + ("rad_lw",)}
hrsw = [pinned(a) for a in rsw]
hout = {k: torch.empty(n, dtype=torch.float64).pin_memory() for k in OUT}
np_in = {k: v.numpy().reshape((NI, NJ), order="F") for k, v in host.items()}
np_rsw = [t.numpy().reshape((NI, NJ), order="F") for t in hrsw]
np_out = {k: v.numpy().reshape((NI, NJ), order="F") for k, v in hout.items()}
ab.reset()
print("smi idle:", smi())
mode = sys.argv[1] if len(sys.argv) > 1 else "host"
nsess = int(sys.argv[2]) if len(sys.argv) > 2 else 8
if mode == "host":
    for s in range(nsess):
        ts = []
        for jt in range(1, NT + 1):
            t0 = time.perf_counter()
            ab.aerobulk_model(jt, NT, "coare3p6", 2., 10., *[np_in[k] for k in IN], Niter=5, l_use_skin=True,
                              rad_sw=np_rsw[jt - 1], rad_lw=np_in["rad_lw"], out=np_out)
            ts.append((time.perf_counter() - t0) * 1e3)
        print(f"host session {s}: total {sum(ts):.1f} ms; per call:", " ".join(f"{t:.1f}" for t in ts), "|", smi())
else:
    dev = {k: v.cuda() for k, v in host.items()}
    drsw = [t.cuda() for t in hrsw]
    out = {k: torch.empty(n, dtype=torch.float64, device="cuda") for k in OUT}
    st = torch.cuda.current_stream()
    ab.set_stream(st.cuda_stream)
    for s in range(nsess):
        evs = []
        for jt in range(1, NT + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            ab.aerobulk_model_device(jt, NT, "coare3p6", 2., 10., *[dev[k] for k in IN], out=out, Niter=5, l_use_skin=True,
                                     rad_sw=drsw[jt - 1], rad_lw=dev["rad_lw"], shape=(NI, NJ))
            e1.record(st)
            evs.append((e0, e1))
        torch.cuda.synchronize()
        ts = [a.elapsed_time(b) for a, b in evs]
        print(f"dev session {s}: total {sum(ts):.1f} ms; per call:", " ".join(f"{t:.2f}" for t in ts), "|", smi())
