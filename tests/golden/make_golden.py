#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/.

Run in the build container only (needs /root/reference and SURVEY.md):
    python tests/golden/make_golden.py

Outputs
  ex_ab.json       -- the reference's own captured output doc/ex_ab.dat
                      (stdout of src/tests/example_call_aerobulk.f90, nb_iter=50),
                      parsed verbatim: 7 significant digits, 5 algorithms x 2 points.
                      THIS is the fixture that pins the oracle to the reference.
  survey_kat.json  -- SURVEY.md Appendix B: 16-digit values from the survey's
                      independent Python transcription of the Fortran source
                      (secondary cross-check, NOT reference output).
  readme_toy.json  -- README.md:188-211 toy table (stale by ~1e-3 relative, see
                      SURVEY.md 8c): coarse known-answer only.
"""
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
SURVEY = os.path.join(HERE, "..", "..", "SURVEY.md")

NUM = r"[-+]?\d+\.\d+(?:[eE][-+]?\d+)?"


def parse_ex_ab():
    txt = open(os.path.join(REF, "doc", "ex_ab.dat")).read()
    blocks = re.split(r"\*{5,} ([A-Z0-9 .]+) \*{5,}", txt)
    # blocks = [preamble, name1, body1, name2, body2, ...]
    names = {"COARE 3.0": "coare3p0", "COARE 3.6": "coare3p6", "ECMWF": "ecmwf", "NCAR": "ncar", "ANDREAS": "andreas"}
    out = {"_source": "doc/ex_ab.dat (reference captured output; example_call_aerobulk.f90 inputs, nb_iter=50)",
           "inputs": {"zt": 2.0, "zu": 10.0, "sst_C": [22.0, 22.0], "t_zt_C": [20.0, 25.0], "q_zt": [0.012, 0.012],
                      "U_zu": [5.0, 5.0], "V_zu": [0.0, 0.0], "slp": [101000.0, 101000.0],
                      "rad_sw": [0.0, 0.0], "rad_lw": [350.0, 350.0], "nb_iter": 50, "jt": 1, "Nt": 1,
                      "rt0": 273.15},
           "algos": {}}
    keys = {"Pot. temperature at zt": "theta_zt_C", "Sensible heat flux: QH": "QH", "Latent  heat flux: QL": "QL",
            "Evaporation:     Evap": "Evap_mm_day", "Skin temperature: SSST": "SSST_C",
            "Tau_x": "Tau_x", "Tau_y": "Tau_y"}
    for i in range(1, len(blocks), 2):
        name = names[blocks[i].strip()]
        body = blocks[i + 1]
        m = re.search(r"nb_iter\s*=\s*(\d+)", body)
        assert m and int(m.group(1)) == 50
        d = {}
        for line in body.splitlines():
            for k, kk in keys.items():
                if line.strip().startswith(k):
                    vals = re.findall(NUM, line.split("=", 1)[1])
                    d[kk] = [float(v) for v in vals[:2]]
                    d[kk + "_str"] = vals[:2]
        out["algos"][name] = d
    return out


def md_tables(txt, start_marker, end_marker=None):
    seg = txt.split(start_marker, 1)[1]
    if end_marker and end_marker in seg:
        seg = seg.split(end_marker, 1)[0]
    rows = []
    for line in seg.splitlines():
        if line.startswith("|") and not re.match(r"^\|[-| :]+$", line.strip()):
            rows.append([c.strip() for c in line.strip().strip("|").split("|")])
    return rows


def fl(x):
    x = x.strip()
    return None if x in ("–", "-", "") else float(x)


def parse_survey():
    txt = open(SURVEY).read()
    out = {"_source": "SURVEY.md Appendix B (survey's literal Python transcription; secondary known answers)"}
    b1 = md_tables(txt, "**B.1", "**B.2")
    out["B1"] = [dict(algo=r[0], T_air_C=float(r[1]), QL=fl(r[2]), QH=fl(r[3]), Tau_x=fl(r[4]), Evap=fl(r[5]), T_s=fl(r[6]))
                 for r in b1[1:]]
    b2 = md_tables(txt, "**B.2", "**B.3")
    out["B2"] = [dict(algo=r[0], T_air_C=float(r[1]), QL=fl(r[2]), QH=fl(r[3]), Tau_x=fl(r[4]), Tau_y=fl(r[5]), T_s=fl(r[6]))
                 for r in b2[1:]]
    b3 = md_tables(txt, "**B.3", "**B.4")
    out["B3"] = [dict(algo=r[0], nb_iter=int(r[1]), jt=int(r[2]), dT_wl=fl(r[3]), T_s=fl(r[4]), QL=fl(r[5]), QH=fl(r[6]),
                      Qnt_ac=fl(r[7]), Tau_ac=fl(r[8]), Hz_wl=fl(r[9])) for r in b3[1:]]
    b4 = md_tables(txt, "**B.4")
    cols = b4[0][1:]
    out["B4_psi"] = [dict(zeta=float(r[0]), **{c: float(v) for c, v in zip(cols, r[1:])}) for r in b4[1:] if len(r) == 9]
    m = re.search(r"`e_sat\(295\.15\)` = (\S+) Pa .*?`q_sat\(295\.15, 101000\)` = (\S+) .*?"
                  r"`Theta_from_z_P0_T_q\(2, 101000, 293\.15, 0\.012\)` = (\S+) K", txt, re.S)
    out["B4_blocks"] = dict(e_sat_295p15=float(m.group(1)), q_sat_295p15_101000=float(m.group(2)),
                            theta_2_101000_293p15_0p012=float(m.group(3)))
    return out


def parse_readme():
    lines = open(os.path.join(REF, "README.md")).read().splitlines()[187:212]
    out = {"_source": "README.md:188-211 (toy: zu=10, zt=2, SST=22C, T=20C, q=12 g/kg, U=5, nb_iter=20, SLP=101000, no skin); STALE ~1e-3",
           "algos": ["coare3p0", "coare3p6", "ncar", "ecmwf", "andreas"], "rows": {}}
    for line in lines:
        if "=" in line and not line.strip().startswith("="):
            k, v = line.split("=", 1)
            vals = re.findall(r"[-+]?\d+\.\d+(?:[eE][-+]?\d+)?", v)
            if len(vals) >= 5:
                out["rows"][k.strip()] = [float(x) for x in vals[:5]]
    return out


if __name__ == "__main__":
    for name, fn in (("ex_ab.json", parse_ex_ab), ("survey_kat.json", parse_survey), ("readme_toy.json", parse_readme)):
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(fn(), f, indent=1)
        print("wrote", name)
