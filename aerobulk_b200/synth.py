"""Synthetic forcing fields for tests and benchmarks.

Counter-based: every value is splitmix64(seed ^ field_id<<32 ^ global_linear_index) -> u in [0,1),
so any row block of any partition of the (Ni, Nj) grid can be generated independently and is
bit-identical to the same rows of the full grid (row-block sharding needs no exchange).
All values satisfy the reference's sanity ranges (src/mod_const.f90:138-146) and wind is capped
at 35 m/s to stay under the tau > 10 N/m^2 fail-stop (src/mod_phymbl.f90:1250-1253).
"""
from __future__ import annotations

import numpy as np

SEED = 20251017
_F = {"sst": 1, "dT": 2, "hum": 3, "wspd": 4, "wdir": 5, "slp": 6, "rlw": 7, "rsw": 8, "calm": 9}


def _u01(seed: int, field: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = idx.astype(np.uint64) ^ np.uint64(seed) ^ (np.uint64(field) << np.uint64(32))
        x = x + np.uint64(0x9E3779B97F4A7C15)
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def fields(Ni: int, Nj: int, j0: int = 0, j1: int | None = None, seed: int = SEED, humidity: str = "sh",
           jt: int | None = None, order: str = "F") -> dict:
    """Rows j0..j1-1 of the global (Ni, Nj) synthetic grid, as (Ni, j1-j0) Fortran-ordered FP64 arrays.

    humidity: 'sh' specific [kg/kg] (default), 'rh' relative [%], 'dp' dew point [K].
    jt: hour 1..24 -> diurnal rad_sw = max(0, 1000 sin(pi (jt-6)/12)) (0.3+0.7u); None -> 400 (0.3+0.7u).
    """
    j1 = Nj if j1 is None else j1
    nj = j1 - j0
    jj = np.arange(j0, j1, dtype=np.uint64)
    ii = np.arange(Ni, dtype=np.uint64)
    idx = (ii[:, None] + jj[None, :] * np.uint64(Ni))          # (Ni, nj), global linear index, column-major
    lat = (-90.0 + (np.arange(j0, j1) + 0.5) * 180.0 / Nj) * np.pi / 180.0
    lat2 = np.broadcast_to(lat[None, :], (Ni, nj))

    u = lambda name: _u01(seed, _F[name], idx)
    sst = 273.15 + np.clip(-1.8 + 30.0 * np.cos(lat2) ** 2 + 2.0 * (u("sst") - 0.5), -1.8, 32.0)
    t_zt = sst + (-8.0 + 12.0 * u("dT"))
    slp = 101325.0 + 1500.0 * np.sin(3.0 * lat2) + 3000.0 * (u("slp") - 0.5)
    tc = t_zt - 273.15
    frac = 0.55 + 0.43 * u("hum")
    esat = 611.2 * np.exp(17.67 * tc / (tc + 243.5))
    if humidity == "sh":
        hum = frac * 0.622 * esat / slp
    elif humidity == "rh":
        hum = 100.0 * frac
    elif humidity == "dp":
        # dew point of vapour pressure frac*esat (inverse Magnus), always below t_zt
        ln = np.log(frac * esat / 611.2)
        hum = 273.15 + 243.5 * ln / (17.67 - ln)
    else:
        raise ValueError(humidity)
    wspd = np.minimum(35.0, 8.0 * np.sqrt(-np.log(1.0 - u("wspd"))))
    wdir = 2.0 * np.pi * u("wdir")
    calm = u("calm") < (1.0 / 4096.0)
    U = np.where(calm, 0.0, wspd * np.cos(wdir))
    V = np.where(calm, 0.0, wspd * np.sin(wdir))
    rad_lw = 300.0 + 120.0 * u("rlw")
    amp = 400.0 if jt is None else max(0.0, 1000.0 * np.sin(np.pi * (jt - 6) / 12.0))
    rad_sw = amp * (0.3 + 0.7 * u("rsw"))
    out = dict(sst=sst, t_zt=t_zt, hum_zt=hum, U_zu=U, V_zu=V, slp=slp, rad_sw=rad_sw, rad_lw=rad_lw)
    return {k: np.asarray(v, dtype=np.float64, order=order) for k, v in out.items()}


def rad_sw_hour(Ni: int, Nj: int, jt: int, j0: int = 0, j1: int | None = None, seed: int = SEED) -> np.ndarray:
    """Only the diurnal short-wave field of hour jt (the other fields do not depend on jt)."""
    j1 = Nj if j1 is None else j1
    jj = np.arange(j0, j1, dtype=np.uint64)
    ii = np.arange(Ni, dtype=np.uint64)
    idx = (ii[:, None] + jj[None, :] * np.uint64(Ni))
    amp = max(0.0, 1000.0 * np.sin(np.pi * (jt - 6) / 12.0))
    return np.asfortranarray(amp * (0.3 + 0.7 * _u01(seed, _F["rsw"], idx)))


# scaled-relative parity metric of SURVEY.md 8d: err = |a-b| / (|b| + S_f)
PARITY_SCALE = {"QL": 10.0, "QH": 10.0, "Tau_x": 1e-2, "Tau_y": 1e-2, "Evap": 1e-5, "T_s": 1.0}


def parity_errors(got: dict, ref: dict) -> dict:
    """Per-field array of scaled errors."""
    return {k: np.abs(got[k] - ref[k]) / (np.abs(ref[k]) + PARITY_SCALE[k]) for k in ref if k in got}
