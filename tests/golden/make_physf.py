"""Generates tests/golden/physf.json by IMPORTING the reference's own Python physics module
(/root/reference/python/modules/aerobulk_physf.py: e_sat :65-91, q_air_dp :93-103, Lvap :32-36) and evaluating it on a
fixed set of inputs.  Run in the build container (the reference does not exist on the GPU box):

    python tests/golden/make_physf.py

The fixture pins SURVEY.md 8a rows a4 (q_air_dp) and a5 (e_sat) and L_vap of a7 to reference-held code.  The Python
module uses the triple point 273.16 K where the Fortran uses rt0 = 273.15 K (src/mod_phymbl.f90:793, :590): the test
(tests/test_oracle_physf.py) maps one onto the other exactly (T -> T * 273.15 / 273.16 for e_sat, T -> T - 0.01 for Lvap).
"""
import importlib.util
import json
import os

import numpy as np

REF = "/root/reference/python/modules/aerobulk_physf.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "physf.json")


def main():
    spec = importlib.util.spec_from_file_location("aerobulk_physf", REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    T = np.concatenate([np.linspace(182.0, 330.0, 75), np.array([273.15, 273.16, 288.15, 295.15, 300.0, 305.37])])
    slp = np.array([80000.0, 95000.0, 101000.0, 101325.0, 110000.0])
    dp = np.linspace(230.0, 305.0, 31)
    DP, P = np.meshgrid(dp, slp, indexing="ij")
    fx = {
        "source": "python/modules/aerobulk_physf.py of brodeau/aerobulk (imported, unmodified)",
        "rtt0_python": float(m.rtt0), "rt0_python_Lvap": float(m.rt0), "reps0": float(m.reps0),
        "e_sat": {"T": T.tolist(), "value": np.asarray(m.e_sat(T)).tolist()},
        "Lvap": {"T": T.tolist(), "value": [float(m.Lvap(t)) for t in T]},
        "q_air_dp": {"dp": DP.ravel().tolist(), "slp": P.ravel().tolist(),
                     "value": np.asarray(m.q_air_dp(DP.ravel(), P.ravel())).tolist()},
    }
    with open(OUT, "w") as f:
        json.dump(fx, f, indent=0)
    print("wrote", OUT, len(T), "temperatures,", DP.size, "dew-point/pressure pairs")


if __name__ == "__main__":
    main()
