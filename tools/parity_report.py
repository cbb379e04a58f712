"""Worst scaled error |gpu - oracle| / (|oracle| + S_f) per algorithm on the synthetic 1440x720 grid (one call, and a
24-step skin session), with the count of points above 1e-10.  The numbers DESIGN.md quotes."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aerobulk_b200 as ab
from aerobulk_b200 import synth
from oracle.oracle import OracleSession

NI, NJ = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (720, 360)
f = synth.fields(NI, NJ)
IN = ("sst", "t_zt", "hum_zt", "U_zu", "V_zu", "slp")
ab.set_verbose(False)
print(f"grid {NI}x{NJ}; columns: worst error, 2nd worst, points > 1e-10")
for algo in ("ncar", "andreas", "coare3p0", "coare3p6", "ecmwf"):
    ab.reset(); ab.set_verbose(False)
    got = ab.aerobulk_model(1, 1, algo, 2., 10., *[f[k] for k in IN], Niter=5)
    o = OracleSession(threads=os.cpu_count()); ref = o.model(1, 1, algo, 2., 10., *[f[k] for k in IN], Niter=5)
    e = np.max([v for v in synth.parity_errors(got, ref).values()], axis=0).ravel()
    s = np.sort(e)
    print(f"{algo:9s} no skin        {s[-1]:.2e} {s[-2]:.2e} {(e > 1e-10).sum()}")
for algo in ("coare3p0", "coare3p6", "ecmwf"):
    ab.reset(); ab.set_verbose(False)
    o = OracleSession(threads=os.cpu_count())
    worst = np.zeros(NI * NJ)
    for jt in range(1, 25):
        rsw = synth.rad_sw_hour(NI, NJ, jt)
        kw = dict(Niter=5, l_use_skin=True, rad_sw=rsw, rad_lw=f["rad_lw"])
        got = ab.aerobulk_model(jt, 24, algo, 2., 10., *[f[k] for k in IN], **kw)
        ref = o.model(jt, 24, algo, 2., 10., *[f[k] for k in IN], **kw)
        worst = np.maximum(worst, np.max([v for v in synth.parity_errors(got, ref).values()], axis=0).ravel())
    s = np.sort(worst)
    print(f"{algo:9s} skin, 24 steps {s[-1]:.2e} {s[-2]:.2e} {(worst > 1e-10).sum()}")
