"""Command-line front end of the station time-series driver (the reference's test_aerobulk_buoy_series_oce.x with CSV
files in place of NetCDF and flags in place of its prompts):

    python -m aerobulk_b200.series_cli buoy.csv out.csv --algo coare3p6 --zt 2 --zu 10 [--no-skin] [--nb-iter 20] [--rdt 3600]

Input columns (names of the reference, src/mod_const.f90:208-220): time, lon, sst, t_air, one of q_air | rh_air | dp_air,
wndspd or u10 + v10, msl, ssrd, strd.  See include/aerobulk_gpu.h (aerobulk_gpu_series_csv).

With --ice: the sea-ice program test_aerobulk_buoy_series_ice.x (src/ice/test_aerobulk_buoy_series_ice.f90) instead:

    python -m aerobulk_b200.series_cli ice.csv out.csv --ice --algo lg15 --zt 2 --zu 10

Input columns (ERA5 names of that program, :179-211): time, siconc, istl1, t2m, one of q_air | rh_air | d2m, wndspd or
u10 + v10, msl, ssrd, strd; temperatures in K or deg C (values below 100 are taken as deg C, like TO_KELVIN_3D).
Output columns: time, Wind, A, then rho_a Qlat Qsen Qlw QNS Qsw Tau SBLM Cd_i Ch_i z0 Rib_zt Rib_zu CdN in SI units
(the program writes SBLM in mm/day and the coefficients x 1000).
"""
from __future__ import annotations

import argparse
import sys


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m aerobulk_b200.series_cli", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("csv_in")
    ap.add_argument("csv_out")
    ap.add_argument("--algo", default=None, choices=("ncar", "coare3p0", "coare3p6", "ecmwf", "andreas", "nemo", "an05", "lu12", "lg15"),
                    help="default coare3p6, with --ice lg15")
    ap.add_argument("--ice", action="store_true", help="sea-ice station series (nemo | an05 | lu12 | lg15)")
    ap.add_argument("--zt", type=float, default=2.0, help="height of air temperature and humidity [m]")
    ap.add_argument("--zu", type=float, default=10.0, help="height of the wind speed [m]")
    ap.add_argument("--no-skin", action="store_true", help="no cool-skin / warm-layer (COARE, ECMWF)")
    ap.add_argument("--nb-iter", type=int, default=20, help="bulk iterations (the reference program uses 20)")
    ap.add_argument("--rdt", type=float, default=3600.0, help="time step between records [s] (mod_const rdt)")
    a = ap.parse_args(argv)
    if a.zt > 99.0 or a.zu > 99.0:   # src/tests/test_aerobulk_buoy_series_oce.f90:329-331
        print("Be reasonable in your choice of zt or zu, they should not exceed a few tenths of meters!", file=sys.stderr)
        return 2
    ice_algos = ("nemo", "an05", "lu12", "lg15")
    algo = a.algo or ("lg15" if a.ice else "coare3p6")
    if (algo in ice_algos) != a.ice:
        print(f"--algo {algo} {'needs' if algo in ice_algos else 'does not go with'} --ice", file=sys.stderr)
        return 2
    a.algo = algo
    import aerobulk_b200 as ab
    ab.set_nb_iter(a.nb_iter)
    ab.set_rdt(a.rdt)
    if a.ice:
        return _ice(ab, a)
    try:
        ab.series_csv(a.csv_in, a.csv_out, a.algo, a.zt, a.zu, not a.no_skin)
    except ab.AerobulkError as e:
        print(e.message, file=sys.stderr)
        return 1
    print(f" *** {a.csv_out} written ({a.algo}, zt={a.zt} m, zu={a.zu} m, skin={'off' if a.no_skin else 'on'})")
    return 0


ICE_COLUMNS = (("rho_a", "rho_zu"), ("Qlat", "QL"), ("Qsen", "QH"), ("Qlw", "Qlw"), ("QNS", "QNS"), ("Qsw", "Qsw"), ("Tau", "TAU"),
               ("SBLM", "SBLM"), ("Cd_i", "Cd_i"), ("Ch_i", "Ch_i"), ("z0", "z0"), ("Rib_zt", "RiB_zt"), ("Rib_zu", "RiB_zu"),
               ("CdN", "CdN"))


def _ice(ab, a) -> int:
    import numpy as np
    try:
        lines = [l for l in open(a.csv_in) if l.strip() and not l.lstrip().startswith("#")]
    except OSError as e:
        print(f"cannot open {a.csv_in}: {e.strerror}", file=sys.stderr)
        return 1
    names = [c.strip() for c in lines[0].split(",")]
    rows = [[c.strip() for c in l.split(",")] for l in lines[1:]]
    col = lambda k: np.array([float(r[names.index(k)]) for r in rows])
    need = ["time", "siconc", "istl1", "t2m", "msl", "ssrd", "strd"]
    miss = [k for k in need if k not in names]
    hum = [k for k in ("q_air", "rh_air", "d2m") if k in names]
    if miss or len(hum) != 1 or not ("wndspd" in names or ("u10" in names and "v10" in names)):
        print(f"{a.csv_in}: needs the columns {need}, one of q_air | rh_air | d2m, and wndspd or u10 + v10", file=sys.stderr)
        return 1
    kelvin = lambda t: np.where(t < 100.0, t + 273.15, t)
    wind = col("wndspd") if "wndspd" in names else np.hypot(col("u10"), col("v10"))
    h = col(hum[0])
    kind = {"q_air": "q", "rh_air": "rh", "d2m": "dp"}[hum[0]]
    if kind == "dp":
        h = kelvin(h)
    sic = col("siconc")
    try:
        r = ab.series_ice(a.algo, a.zt, a.zu, sic, kelvin(col("istl1")), kelvin(col("t2m")), h, wind, col("msl"),
                          col("ssrd"), col("strd"), hum_kind=kind, want=tuple(k for _, k in ICE_COLUMNS))
    except ab.AerobulkError as e:
        print(e.message, file=sys.stderr)
        return 1
    t = names.index("time")
    with open(a.csv_out, "w") as f:
        f.write(",".join(["time", "Wind", "A"] + [c for c, _ in ICE_COLUMNS]) + "\n")
        for i, row in enumerate(rows):
            f.write(",".join([row[t], repr(float(wind[i])), repr(float(sic[i]))] + [repr(float(r[k][i])) for _, k in ICE_COLUMNS]) + "\n")
    print(f" *** {a.csv_out} written ({a.algo}, zt={a.zt} m, zu={a.zu} m, {len(rows)} records, {int((sic > 0.01).sum())} with ice)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
