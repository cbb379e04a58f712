"""Regression fixtures of the ORACLE for the paths the reference holds no fixture for (sea ice, station series, sea-ice series, TURB_*
optional outputs).  NOT reference outputs: they freeze today's restatement so that a later edit of oracle/ that changes
a number is noticed (tests/test_oracle_golden.py::test_oracle_regression_fixtures).  Re-run only when such a change
is intended:  python tests/golden/make_regression.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from aerobulk_b200 import synth  # noqa: E402
from oracle.oracle import OracleSession  # noqa: E402


def cases():
    out = {}
    one = lambda v: np.array([float(v)])
    # test_ice.sh scenario (src/ice/test_aerobulk_oce+ice.f90 inputs), every ice algorithm, ECMWF over the leads
    o = OracleSession()
    o.set_nb_iter(20)
    for ice in ("nemo", "easy", "an05", "lu12", "lg15"):
        r = o.oce_ice(ice, "ecmwf", 2.0, 10.0, one(270.15), one(271.35), one(276.15), one(0.004), one(3.0), one(101000.0),
                      one(0.8), cxn=[1.4e-3, 1.4e-3, 1.4e-3])
        out[f"ice_scenario/{ice}"] = {k: float(v[0]) for k, v in r.items()}
    # an unstable ice point
    for ice in ("an05", "lg15"):
        r = o.oce_ice(ice, "ncar", 2.0, 10.0, one(268.15), one(271.35), one(258.15), one(0.0009), one(9.0), one(100500.0),
                      one(0.35))
        out[f"ice_unstable/{ice}"] = {k: float(v[0]) for k, v in r.items()}
    # sea-ice station series (src/ice/test_aerobulk_buoy_series_ice.f90): the test_ice.sh point with radiation, a record
    # below the SIC gate, an unstable one
    sic, sit = np.array([0.8, 0.005, 0.35]), np.array([270.15, 270.15, 268.15])
    tair, qa = np.array([276.15, 276.15, 258.15]), np.array([0.004, 0.004, 0.0009])
    wnd, slp = np.array([3.0, 3.0, 9.0]), np.array([101000.0, 101000.0, 100500.0])
    rsw, rlw = np.array([120.0, 120.0, 0.0]), np.array([250.0, 250.0, 190.0])
    for ice in ("nemo", "an05", "lu12", "lg15"):
        r = o.series_ice(ice, 2.0, 10.0, sic, sit, tair, qa, wnd, slp, rsw, rlw)
        out[f"series_ice/{ice}"] = {k: [float(x) for x in v] for k, v in r.items()}
    # station series: 30 records of 3 stations, two algorithms; the last record and two mid-series values per output
    d = synth.station_series(30, 3)
    for algo in ("coare3p6", "ecmwf"):
        s = OracleSession()
        s.set_nb_iter(20)
        r = s.series(algo, 2.0, 10.0, **d)
        out[f"series/{algo}"] = {k: [float(v[29, 1]), float(v[14, 0]), float(v[8, 2])] for k, v in r.items()}
    # TURB_* optional outputs on 4 points
    f = synth.fields(2, 2)
    from oracle import oracle as om
    L = om.lib()
    ssq = np.array([0.98 * L.abo_q_sat(t, p) for t, p in zip(np.ravel(f["sst"], order="F"), np.ravel(f["slp"], order="F"))])
    th = np.ravel(f["t_zt"], order="F") + 0.0098 * 2.0
    w = np.hypot(np.ravel(f["U_zu"], order="F"), np.ravel(f["V_zu"], order="F"))
    for algo in ("ncar", "andreas", "coare3p0", "coare3p6", "ecmwf"):
        s = OracleSession()
        s.set_nb_iter(8)
        r = s.turb(algo, 1, 2.0, 10.0, np.ravel(f["sst"], order="F"), th, ssq, np.ravel(f["hum_zt"], order="F"), w,
                   want=("CdN", "ChN", "CeN", "xz0", "xu_star", "xL", "xUN10"))
        out[f"turb/{algo}"] = {k: [float(x) for x in v] for k, v in r.items()}
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_regression.json")
    with open(path, "w") as fh:
        json.dump({"_note": "oracle-generated regression values, NOT reference outputs (see make_regression.py)",
                   "cases": cases()}, fh, indent=1)
    print("wrote", path)
