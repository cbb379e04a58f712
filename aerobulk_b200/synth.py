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


def station_series(Nt: int, S: int, seed: int = SEED, humidity: str = "q", dt_s: int = 3600, start_s: int = 0) -> dict:
    """Synthetic forcing of S stations over Nt records spaced dt_s seconds, arrays (Nt, S).

    Longitudes cover [-180, 180); short-wave follows the LOCAL solar hour so that the COARE warm layer builds
    by day and its dawn reset (src/mod_skin_coare.f90:159-163) fires; wind, air-sea temperature difference and
    humidity wander slowly (weather) plus per-record noise.  Wind is either an exact calm (1 record in 512) or
    at least 0.5 m/s: between the two the ECMWF iteration is ill-conditioned (a 1e-15 relative input perturbation moves
    the fluxes of the CPU restatement by up to 1e-9 at 0.1 m/s), which would test the conditioning rather than the implementation.  humidity: 'q' [kg/kg], 'rh' [%], 'dp' [K]."""
    jt = np.arange(Nt, dtype=np.uint64)
    st = np.arange(S, dtype=np.uint64)
    idx = jt[:, None] * np.uint64(S) + st[None, :]
    u = lambda name, k=0: _u01(seed + 7919 * (k + 1), _F[name], idx)
    us = lambda name, k=0: _u01(seed + 104729 * (k + 1), _F[name], st)[None, :]      # per station
    lon = -180.0 + 360.0 * (np.arange(S) + 0.5) / S
    secs = start_s + dt_s * np.arange(Nt, dtype=np.int64)
    isd = ((secs % 86400) // 60 * 60).astype(np.int32)
    hour_loc = ((secs[:, None] / 3600.0 + lon[None, :] / 15.0) % 24.0)
    day = secs[:, None] / 86400.0
    sst = 273.15 + 8.0 + 20.0 * us("sst") + 0.3 * np.sin(2 * np.pi * day / 9.0 + 6.28 * us("sst", 1))
    dT = -3.0 + 4.5 * us("dT") + 2.5 * np.sin(2 * np.pi * day / 3.7 + 6.28 * us("dT", 1)) + 0.6 * (u("dT") - 0.5)
    t_zt = sst + dT
    slp = 101325.0 + 1200.0 * np.sin(2 * np.pi * day / 5.3 + 6.28 * us("slp")) + 200.0 * (u("slp") - 0.5)
    frac = np.clip(0.62 + 0.3 * us("hum") + 0.08 * np.sin(2 * np.pi * day / 2.9 + 6.28 * us("hum", 1)) + 0.04 * (u("hum") - 0.5), 0.3, 0.97)
    tc = t_zt - 273.15
    esat = 611.2 * np.exp(17.67 * tc / (tc + 243.5))
    if humidity in ("q", "sh"):
        hum = frac * 0.622 * esat / slp
    elif humidity == "rh":
        hum = 100.0 * frac
    elif humidity == "dp":
        ln = np.log(frac * esat / 611.2)
        hum = 273.15 + 243.5 * ln / (17.67 - ln)
    else:
        raise ValueError(humidity)
    wind = np.clip(1.0 + 10.0 * us("wspd") + 4.0 * np.sin(2 * np.pi * day / 2.3 + 6.28 * us("wspd", 1)) + 1.5 * (u("wspd") - 0.5), 0.5, 24.0)
    wind = np.where(u("calm") < 1.0 / 512.0, 0.0, wind)
    rad_lw = 330.0 + 80.0 * us("rlw") + 20.0 * (u("rlw") - 0.5)
    rad_sw = np.maximum(0.0, 950.0 * np.sin(np.pi * (hour_loc - 6.0) / 12.0)) * (0.35 + 0.65 * u("rsw"))
    c = lambda a: np.array(np.broadcast_to(a, (Nt, S)), dtype=np.float64, order="C")
    return dict(isecday_utc=isd, lon=lon, sst=c(sst), t_zt=c(t_zt), hum_zt=c(hum), wind=c(wind), slp=c(slp),
                rad_sw=c(rad_sw), rad_lw=c(rad_lw))


def ice_fields(n: int, seed: int = SEED, humidity: str = "q") -> dict:
    """Synthetic polar forcing for the sea-ice algorithms: n points, 1-D arrays.
    sit -35..-0.5 degC, leads at -1.9..0.5 degC, air-ice difference -6..+8 K (about half stable), RH 60..98 %, Rayleigh
    wind capped at 28 m/s with 1 exact calm in 2048, ice fraction 0..1 with exact 0 and 1 occurrences."""
    idx = np.arange(n, dtype=np.uint64)
    u = lambda name, k=0: _u01(seed + 15485863 * (k + 1), _F[name], idx)
    sit = 273.15 - 35.0 + 34.5 * u("sst")
    sst = 273.15 - 1.9 + 2.4 * u("sst", 1)
    t_zt = sit + (-6.0 + 14.0 * u("dT"))
    slp = 100000.0 + 3500.0 * (u("slp") - 0.5)
    rh = 60.0 + 38.0 * u("hum")
    tc = t_zt - 273.15
    esat = 611.2 * np.exp(17.67 * tc / (tc + 243.5))
    if humidity in ("q", "sh"):
        hum = 0.01 * rh * 0.622 * esat / slp
    elif humidity == "rh":
        hum = rh
    elif humidity == "dp":
        ln = np.log(0.01 * rh * esat / 611.2)
        hum = 273.15 + 243.5 * ln / (17.67 - ln)
    else:
        raise ValueError(humidity)
    wind = np.minimum(28.0, 7.0 * np.sqrt(-np.log(1.0 - u("wspd"))))
    wind = np.where(u("calm") < 1.0 / 2048.0, 0.0, wind)
    a = u("rsw")
    frice = np.where(a < 0.02, 0.0, np.where(a > 0.9, 1.0, (a - 0.02) / 0.88))
    return dict(sit=sit, sst=sst, t_zt=t_zt, hum_zt=hum, wind=wind, slp=slp, frice=frice)


def fields_parallel(Ni: int, Nj: int, j0: int = 0, j1: int | None = None, threads: int = 8, rows_per_task: int = 32,
                    out: dict | None = None, **kw) -> dict:
    """:func:`fields` for big row blocks: the rows are generated in independent chunks on a thread pool (numpy releases the
    GIL inside its loops) and written straight into (Ni, j1-j0) Fortran-ordered arrays -- `out` may supply them (e.g.
    views of pinned memory).  Bit-identical to :func:`fields` (the generator is counter-based)."""
    from concurrent.futures import ThreadPoolExecutor
    j1 = Nj if j1 is None else j1
    nj = j1 - j0
    keys = ("sst", "t_zt", "hum_zt", "U_zu", "V_zu", "slp", "rad_sw", "rad_lw")
    res = out if out is not None else {k: np.empty((Ni, nj), dtype=np.float64, order="F") for k in keys}

    def task(a):
        b = min(nj, a + rows_per_task)
        f = fields(Ni, Nj, j0=j0 + a, j1=j0 + b, **kw)
        for k in keys:
            if k in res:
                res[k][:, a:b] = f[k]

    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        list(ex.map(task, range(0, nj, rows_per_task)))
    return res
