"""Command-line front end of the station time-series driver (the reference's test_aerobulk_buoy_series_oce.x with CSV
files in place of NetCDF and flags in place of its prompts):

    python -m aerobulk_b200.series_cli buoy.csv out.csv --algo coare3p6 --zt 2 --zu 10 [--no-skin] [--nb-iter 20] [--rdt 3600]

Input columns (names of the reference, src/mod_const.f90:208-220): time, lon, sst, t_air, one of q_air | rh_air | dp_air,
wndspd or u10 + v10, msl, ssrd, strd.  See include/aerobulk_gpu.h (aerobulk_gpu_series_csv).
"""
from __future__ import annotations

import argparse
import sys


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m aerobulk_b200.series_cli", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("csv_in")
    ap.add_argument("csv_out")
    ap.add_argument("--algo", default="coare3p6", choices=("ncar", "coare3p0", "coare3p6", "ecmwf", "andreas"))
    ap.add_argument("--zt", type=float, default=2.0, help="height of air temperature and humidity [m]")
    ap.add_argument("--zu", type=float, default=10.0, help="height of the wind speed [m]")
    ap.add_argument("--no-skin", action="store_true", help="no cool-skin / warm-layer (COARE, ECMWF)")
    ap.add_argument("--nb-iter", type=int, default=20, help="bulk iterations (the reference program uses 20)")
    ap.add_argument("--rdt", type=float, default=3600.0, help="time step between records [s] (mod_const rdt)")
    a = ap.parse_args(argv)
    if a.zt > 99.0 or a.zu > 99.0:   # src/tests/test_aerobulk_buoy_series_oce.f90:329-331
        print("Be reasonable in your choice of zt or zu, they should not exceed a few tenths of meters!", file=sys.stderr)
        return 2
    import aerobulk_b200 as ab
    ab.set_nb_iter(a.nb_iter)
    ab.set_rdt(a.rdt)
    try:
        ab.series_csv(a.csv_in, a.csv_out, a.algo, a.zt, a.zu, not a.no_skin)
    except ab.AerobulkError as e:
        print(e.message, file=sys.stderr)
        return 1
    print(f" *** {a.csv_out} written ({a.algo}, zt={a.zt} m, zu={a.zu} m, skin={'off' if a.no_skin else 'on'})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
