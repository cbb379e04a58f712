"""ctypes binding of the CPU oracle (oracle/aerobulk_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / ``--impl reference`` legs.  Nothing under
aerobulk_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libaerobulk_oracle.so")

ALGOS = {"coare3p0": 1, "coare3p6": 2, "ncar": 3, "ecmwf": 4, "andreas": 5}


def build(force: bool = False) -> str:
    """Compile the oracle with the recipe in oracle/Makefile (gcc -O2 -ffp-contract=off)."""
    src = os.path.join(_HERE, "aerobulk_oracle.c")
    hdr = os.path.join(_HERE, "aerobulk_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libaerobulk_oracle.so"])
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.abo_new.restype = C.c_void_p
        L.abo_free.argtypes = [C.c_void_p]
        L.abo_set_rdt.argtypes = [C.c_void_p, C.c_double]
        L.abo_set_gdept.argtypes = [C.c_void_p, C.c_double]
        L.abo_set_nb_iter.argtypes = [C.c_void_p, C.c_int]
        L.abo_get_nb_iter.argtypes = [C.c_void_p]
        L.abo_get_use_skin.argtypes = [C.c_void_p]
        L.abo_get_humidity_type.argtypes = [C.c_void_p]
        L.abo_get_humidity_type.restype = C.c_char_p
        L.abo_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.abo_errmsg.argtypes = [C.c_void_p]
        L.abo_errmsg.restype = C.c_char_p
        L.abo_model.restype = C.c_int
        L.abo_model.argtypes = ([C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_double, C.c_double,
                                 C.c_int, C.c_int] + [_dp] * 11 +
                                [C.POINTER(C.c_int), C.POINTER(C.c_int), _dp, _dp, _dp])
        L.abo_get_state.restype = C.c_long
        L.abo_get_state.argtypes = [C.c_void_p, C.c_int, _dp]
        for name, nargs in (("abo_e_sat", 1), ("abo_q_sat", 2), ("abo_theta_from_z_P0_T_q", 4),
                            ("abo_rho_air", 3), ("abo_visc_air", 1), ("abo_L_vap", 1), ("abo_cp_air", 1),
                            ("abo_alpha_sw", 1), ("abo_qlw_net", 2), ("abo_one_on_L", 5), ("abo_Ri_bulk", 6),
                            ("abo_q_air_rh", 3), ("abo_q_air_dp", 2), ("abo_cd_n10_ncar", 1),
                            ("abo_charn_coare3p0", 1), ("abo_charn_coare3p6", 1), ("abo_u_star_andreas", 1)):
            f = getattr(L, name)
            f.restype = C.c_double
            f.argtypes = [C.c_double] * nargs
        L.abo_z0tq_LKB.restype = C.c_double
        L.abo_z0tq_LKB.argtypes = [C.c_int, C.c_double, C.c_double]
        L.abo_delta_skin_layer.restype = C.c_double
        L.abo_delta_skin_layer.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_double]
        L.abo_psi_m.restype = C.c_double
        L.abo_psi_m.argtypes = [C.c_int, C.c_double]
        L.abo_psi_h.restype = C.c_double
        L.abo_psi_h.argtypes = [C.c_int, C.c_double]
        L.abo_turb_noskin.argtypes = [C.c_int, C.c_int] + [C.c_double] * 7 + [_dp]
        L.abo_set_nitend.argtypes = [C.c_void_p, C.c_int]
        L.abo_turb.restype = C.c_int
        L.abo_turb.argtypes = ([C.c_void_p, C.c_char_p, C.c_int, C.c_double, C.c_double, C.c_long] + [_dp] * 5 +
                               [C.c_int, C.c_int] + [_dp] * 6 + [_dp] * 3 + [C.c_int, _dp, C.POINTER(_dp)])
        L.abo_series.restype = C.c_int
        L.abo_series.argtypes = ([C.c_void_p, C.c_char_p, C.c_int, C.c_long, C.c_double, C.c_double, C.POINTER(C.c_int), _dp] +
                                 [_dp] * 3 + [C.c_int] + [_dp] * 4 + [C.c_int, C.POINTER(_dp)])
        L.abo_turb_ice.restype = C.c_int
        L.abo_turb_ice.argtypes = ([C.c_void_p, C.c_char_p, C.c_double, C.c_double, C.c_long] + [_dp] * 7 + [C.c_int] +
                                   [_dp] * 6 + [C.POINTER(_dp)])
        L.abo_series_ice.restype = C.c_int
        L.abo_series_ice.argtypes = ([C.c_void_p, C.c_char_p, C.c_double, C.c_double, C.c_long] + [_dp] * 4 + [C.c_int] +
                                     [_dp] * 4 + [C.POINTER(_dp)])
        L.abo_oce_ice.restype = C.c_int
        L.abo_oce_ice.argtypes = ([C.c_void_p, C.c_char_p, C.c_char_p, C.c_double, C.c_double, C.c_long] + [_dp] * 4 +
                                  [C.c_int] + [_dp] * 4 + [C.c_int, C.POINTER(_dp)])
        for name, nargs in (("abo_e_sat_ice", 1), ("abo_q_sat_ice", 2), ("abo_f_m_louis", 4), ("abo_f_h_louis", 4),
                            ("abo_psi_m_ice", 1), ("abo_psi_h_ice", 1), ("abo_rough_leng_m", 2), ("abo_CdN10_f_LU13", 1),
                            ("abo_CdN_f_LG15_light", 3)):
            f = getattr(L, name)
            f.restype = C.c_double
            f.argtypes = [C.c_double] * nargs
        L.abo_rough_leng_tq.restype = C.c_int
        L.abo_rough_leng_tq.argtypes = [C.c_double] * 3 + [_dp, _dp]
        L.abo_gamma_moist.restype = C.c_double
        L.abo_gamma_moist.argtypes = [C.c_double, C.c_double]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[oracle rc={code}] {msg}")
        self.code = code


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_dp)


class OracleSession:
    """One `aerobulk_model` session (module-level SAVE state of the reference)."""

    def __init__(self, threads: int = 1):
        self._L = lib()
        self._s = C.c_void_p(self._L.abo_new())
        self._L.abo_set_threads(self._s, threads)

    def close(self):
        if self._s:
            self._L.abo_free(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_rdt(self, v): self._L.abo_set_rdt(self._s, float(v))
    def set_gdept(self, v): self._L.abo_set_gdept(self._s, float(v))
    def set_nb_iter(self, v): self._L.abo_set_nb_iter(self._s, int(v))
    def set_threads(self, v): self._L.abo_set_threads(self._s, int(v))

    @property
    def nb_iter(self): return self._L.abo_get_nb_iter(self._s)
    @property
    def use_skin(self): return bool(self._L.abo_get_use_skin(self._s))
    @property
    def humidity_type(self): return self._L.abo_get_humidity_type(self._s).decode()

    def model(self, jt, Nt, calgo, zt, zu, sst, t_zt, hum_zt, U_zu, V_zu, slp,
              Niter=None, l_use_skin=None, rad_sw=None, rad_lw=None):
        """AEROBULK_MODEL; arrays are (Ni,Nj) Fortran-ordered or 1-D.  Returns dict of outputs."""
        arrs = [np.asfortranarray(a, dtype=np.float64) for a in (sst, t_zt, hum_zt, U_zu, V_zu, slp)]
        shape = arrs[0].shape
        Ni = shape[0]
        Nj = shape[1] if len(shape) > 1 else 1
        rs = None if rad_sw is None else np.asfortranarray(rad_sw, dtype=np.float64)
        rl = None if rad_lw is None else np.asfortranarray(rad_lw, dtype=np.float64)
        outs = {k: np.zeros(shape, dtype=np.float64, order="F") for k in ("QL", "QH", "Tau_x", "Tau_y", "Evap")}
        Ts = np.zeros(shape, dtype=np.float64, order="F") if (rs is not None and rl is not None) else None
        ni = None if Niter is None else C.byref(C.c_int(int(Niter)))
        ls = None if l_use_skin is None else C.byref(C.c_int(int(bool(l_use_skin))))
        rc = self._L.abo_model(self._s, int(jt), int(Nt), calgo.encode(), float(zt), float(zu), Ni, Nj,
                               *[_ptr(a) for a in arrs],
                               *[_ptr(outs[k]) for k in ("QL", "QH", "Tau_x", "Tau_y", "Evap")],
                               ni, ls, _ptr(rs), _ptr(rl), _ptr(Ts))
        if rc != 0:
            raise OracleError(rc, self._L.abo_errmsg(self._s).decode())
        if Ts is not None:
            outs["T_s"] = Ts
        return outs

    def set_nitend(self, v): self._L.abo_set_nitend(self._s, int(v))

    TURB_OPT = ("CdN", "ChN", "CeN", "xz0", "xu_star", "xL", "xUN10", "pdT_cs", "pdT_wl", "pHz_wl")

    def turb(self, calgo, kt, zt, zu, T_s, t_zt, q_s, q_zt, U_zu, l_use_cs=False, l_use_wl=False,
             Qsw=None, rad_lw=None, slp=None, isecday_utc=0, plong=None, want=()):
        """Direct TURB_* call; returns dict with T_s, q_s (updated copies), Cd..Ubzu and the wanted optionals."""
        f = lambda a: None if a is None else np.ascontiguousarray(np.ravel(a, order="F"), dtype=np.float64)
        Ts, qs = f(T_s).copy(), f(q_s).copy()
        n = Ts.size
        ins = [f(t_zt), f(q_zt), f(U_zu)]
        outs = {k: np.zeros(n) for k in ("Cd", "Ch", "Ce", "t_zu", "q_zu", "Ubzu")}
        optv = {k: np.zeros(n) for k in want}
        arr = (_dp * 10)(*[_ptr(optv[k]) if k in optv else None for k in self.TURB_OPT])
        rs, rl, sp, pl = f(Qsw), f(rad_lw), f(slp), f(plong)
        rc = self._L.abo_turb(self._s, calgo.encode(), int(kt), float(zt), float(zu), n, _ptr(Ts), _ptr(ins[0]), _ptr(qs),
                              _ptr(ins[1]), _ptr(ins[2]), int(bool(l_use_cs)), int(bool(l_use_wl)),
                              *[_ptr(outs[k]) for k in ("Cd", "Ch", "Ce", "t_zu", "q_zu", "Ubzu")],
                              _ptr(rs), _ptr(rl), _ptr(sp), int(isecday_utc), _ptr(pl), arr)
        if rc != 0:
            raise OracleError(rc, self._L.abo_errmsg(self._s).decode())
        outs.update(optv)
        outs["T_s"], outs["q_s"] = Ts, qs
        return outs

    SERIES_OUT = ("rho_zu", "QL", "QH", "Qlw", "QNS", "Qsw", "dT_cs", "dT_wl", "TAU", "dT", "Hz_wl", "Qnt_ac", "Tau_ac",
                  "Cd", "Ce", "Ch", "theta_zu", "q_zu", "t_zu", "RiB", "z0", "u_star", "L", "UN10", "Ts", "Evap", "q_zt",
                  "theta_zt")

    def series(self, calgo, zt, zu, isecday_utc, lon, sst, t_zt, hum_zt, wind, slp, rad_sw, rad_lw,
               hum_kind=0, l_use_skin=True):
        """Buoy-series time loop (abo_series); inputs (Nt,S) C-ordered; returns all 28 series."""
        isd = np.ascontiguousarray(isecday_utc, dtype=np.int32)
        Nt = isd.shape[0]
        lon = np.ascontiguousarray(lon, dtype=np.float64).reshape(-1)
        S = lon.shape[0]
        ins = [np.ascontiguousarray(a, dtype=np.float64) for a in (sst, t_zt, hum_zt, wind, slp, rad_sw, rad_lw)]
        outs = {k: np.zeros((Nt, S)) for k in self.SERIES_OUT}
        arr = (_dp * 28)(*[_ptr(outs[k]) for k in self.SERIES_OUT])
        rc = self._L.abo_series(self._s, calgo.encode(), Nt, S, float(zt), float(zu),
                                isd.ctypes.data_as(C.POINTER(C.c_int)), _ptr(lon), _ptr(ins[0]), _ptr(ins[1]), _ptr(ins[2]),
                                int(hum_kind), _ptr(ins[3]), _ptr(ins[4]), _ptr(ins[5]), _ptr(ins[6]),
                                int(bool(l_use_skin)), arr)
        if rc != 0:
            raise OracleError(rc, self._L.abo_errmsg(self._s).decode())
        return outs

    ICE_OPT = ("CdN", "ChN", "CeN", "xz0", "xu_star", "xL", "xUN10", "CdN_frm")
    OCE_ICE_OUT = tuple([k + "_i" for k in ("Cd", "Ch", "Ce", "theta_zu", "q_zu", "t_zu", "Ub", "RiB", "z0", "u_star", "L",
                                             "UN10", "rho_zu", "Tau", "QH", "QL", "Evap")] +
                        [k + "_w" for k in ("Cd", "Ch", "Ce", "theta_zu", "q_zu", "Ub", "z0", "u_star", "L", "UN10", "Tau",
                                             "QH", "QL", "Evap")] + ["Tau", "QH", "QL", "Evap"])

    def turb_ice(self, calgo, zt, zu, Ts_i, t_zt, qs_i, q_zt, U_zu, frice=None, cxn=None, per_point_form_drag=False,
                 want=()):
        """TURB_ICE_<calgo> on flat arrays; returns Cd Ch Ce t_zu q_zu Ubzu + wanted optionals."""
        f = lambda a: None if a is None else np.ascontiguousarray(np.ravel(a, order="F"), dtype=np.float64)
        ins = [f(a) for a in (Ts_i, t_zt, qs_i, q_zt, U_zu, frice)]
        n = ins[0].size
        cx = None if cxn is None else np.ascontiguousarray(cxn, dtype=np.float64)
        outs = {k: np.zeros(n) for k in ("Cd", "Ch", "Ce", "t_zu", "q_zu", "Ubzu")}
        optv = {k: np.zeros(n) for k in want}
        arr = (_dp * 8)(*[_ptr(optv[k]) if k in optv else None for k in self.ICE_OPT])
        rc = self._L.abo_turb_ice(self._s, calgo.encode(), float(zt), float(zu), n, *[_ptr(a) for a in ins], _ptr(cx),
                                  int(bool(per_point_form_drag)), *[_ptr(outs[k]) for k in ("Cd", "Ch", "Ce", "t_zu", "q_zu", "Ubzu")],
                                  arr)
        if rc != 0:
            raise OracleError(rc, self._L.abo_errmsg(self._s).decode())
        outs.update(optv)
        return outs

    def oce_ice(self, calgo_ice, calgo_oce, zt, zu, sit, sst, t_zt, hum_zt, wind, slp, frice, hum_kind=0, cxn=None,
                per_point_form_drag=False):
        """Ice + leads workflow (abo_oce_ice); returns the 35 series of OCE_ICE_OUT."""
        f = lambda a: None if a is None else np.ascontiguousarray(np.ravel(a, order="F"), dtype=np.float64)
        ins = [f(a) for a in (sit, sst, t_zt, hum_zt, wind, slp, frice)]
        n = ins[0].size
        cx = None if cxn is None else np.ascontiguousarray(cxn, dtype=np.float64)
        outs = {k: np.zeros(n) for k in self.OCE_ICE_OUT}
        arr = (_dp * 35)(*[_ptr(outs[k]) for k in self.OCE_ICE_OUT])
        rc = self._L.abo_oce_ice(self._s, calgo_ice.encode(), None if calgo_oce is None else calgo_oce.encode(), float(zt),
                                 float(zu), n, _ptr(ins[0]), _ptr(ins[1]), _ptr(ins[2]), _ptr(ins[3]), int(hum_kind),
                                 _ptr(ins[4]), _ptr(ins[5]), _ptr(ins[6]), _ptr(cx), int(bool(per_point_form_drag)), arr)
        if rc != 0:
            raise OracleError(rc, self._L.abo_errmsg(self._s).decode())
        return outs

    SERIES_ICE_OUT = ("rho_zu", "QL", "QH", "Qlw", "QNS", "Qsw", "TAU", "SBLM", "Cd_i", "Ch_i", "Ce_i", "z0", "RiB_zt", "RiB_zu",
                      "CdN", "u_star", "L", "UN10", "theta_zu", "q_zu", "Ublk")

    def series_ice(self, calgo, zt, zu, sic, sit, t_zt, hum_zt, wind, slp, rad_sw, rad_lw, hum_kind=0):
        """Sea-ice station series (abo_series_ice); returns the 21 series of SERIES_ICE_OUT, flattened."""
        f = lambda a: np.ascontiguousarray(np.ravel(a), dtype=np.float64)
        ins = [f(a) for a in (sic, sit, t_zt, hum_zt, wind, slp, rad_sw, rad_lw)]
        n = ins[0].size
        outs = {k: np.zeros(n) for k in self.SERIES_ICE_OUT}
        arr = (_dp * 21)(*[_ptr(outs[k]) for k in self.SERIES_ICE_OUT])
        rc = self._L.abo_series_ice(self._s, calgo.encode(), float(zt), float(zu), n, *[_ptr(a) for a in ins[:4]],
                                    int(hum_kind), *[_ptr(a) for a in ins[4:]], arr)
        if rc != 0:
            raise OracleError(rc, self._L.abo_errmsg(self._s).decode())
        return outs

    def state(self, which: int, n: int):
        out = np.zeros(n, dtype=np.float64)
        got = self._L.abo_get_state(self._s, which, _ptr(out))
        return out if got == n else None


def turb_noskin(algo, nb_iter, zt, zu, sst, tha_zt, ssq, q_zt, U_zu):
    out = np.zeros(13)
    lib().abo_turb_noskin(ALGOS[algo], nb_iter, zt, zu, sst, tha_zt, ssq, q_zt, U_zu, _ptr(out))
    keys = ("Cd", "Ch", "Ce", "t_zu", "q_zu", "Ubzu", "CdN", "ChN", "CeN", "z0", "us", "L", "UN10")
    return dict(zip(keys, out))
