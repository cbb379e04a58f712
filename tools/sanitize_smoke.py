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
