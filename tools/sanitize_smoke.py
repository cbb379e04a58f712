"""Small calls through every kernel family, meant to run under compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aerobulk_b200 as ab
from aerobulk_b200 import synth
ab.set_verbose(False)
Ni, Nj = 97, 53            # not a multiple of the block size
f = synth.fields(Ni, Nj)
IN = ("sst", "t_zt", "hum_zt", "U_zu", "V_zu", "slp")
for algo in ("ncar", "andreas", "coare3p0", "coare3p6", "ecmwf"):
    ab.reset(); ab.set_verbose(False)
    ab.aerobulk_model(1, 1, algo, 2., 10., *[f[k] for k in IN], Niter=4)
for algo in ("coare3p6", "ecmwf"):
    ab.reset(); ab.set_verbose(False)
    for jt in (1, 2, 3):
        ab.aerobulk_model(jt, 3, algo, 2., 10., *[f[k] for k in IN], Niter=4, l_use_skin=True, rad_sw=f["rad_sw"], rad_lw=f["rad_lw"])
# AEROBULK_INIT statistics with masked points (stats_fast_kernel flags their blocks, stats_fix_kernel redoes them)
g2 = {k: v.copy() for k, v in f.items()}
g2["sst"][3, 7] -= 273.15
g2["slp"][90, 50] /= 100.0
ab.reset(); ab.set_verbose(False)
ab.aerobulk_model(1, 1, "ncar", 2., 10., *[g2[k] for k in IN], Niter=3)
# a pageable jt == 1 call with two pipeline chunks: copy threads feed the staged pipeline, speculative AEROBULK_INIT per chunk
big = synth.fields(640, 641)
ab.reset(); ab.set_verbose(False)
ab.aerobulk_model(1, 1, "coare3p6", 2., 10., *[big[k] for k in IN], Niter=3, l_use_skin=True, rad_sw=big["rad_sw"], rad_lw=big["rad_lw"])
ab.reset()
ab.set_sort(2)
ab.aerobulk_model(1, 1, "andreas", 10., 10., *[f[k] for k in IN], Niter=3)
ab.set_sort(1)
d = synth.station_series(5, 37)
ab.series("coare3p6", 2., 10., **d)
ab.series("ncar", 2., 10., **d, want=("QL",))
g = synth.ice_fields(301)
for ice in ("nemo", "an05", "lu12", "lg15"):
    ab.oce_ice(ice, "ecmwf", 2., 10., **g)
print("sanitize_smoke done, launches:", ab.launch_count())
