"""CPU checks of the oracle's sea-ice restatement.  The reference holds no ice fixture (its ice test
programs are interactive or need NetCDF forcing): "parity unpinned"; these tests hold the restatement to hand
evaluations of the source formulas, to the published constants, and to its own invariants."""
import ctypes as C
import math

import numpy as np
import pytest

from aerobulk_b200 import synth
from oracle import oracle
from oracle.oracle import OracleSession


def test_goff_ice_and_louis_building_blocks():
    L = oracle.lib()
    # Goff-Gratch over ice: 6.1071 hPa at the triple point (the formula's own anchor), ~2.60 hPa at -10 degC
    assert abs(L.abo_e_sat_ice(273.16) - 610.71) < 1e-9
    assert abs(L.abo_e_sat_ice(263.15) - 259.9) < 1.0
    assert L.abo_e_sat_ice(263.15) < L.abo_e_sat(263.15)          # ice saturation below water saturation under 0 degC
    e = L.abo_e_sat_ice(258.15)
    eps = 287.05 / 461.495
    assert abs(L.abo_q_sat_ice(258.15, 101000.0) - eps * e / (101000.0 - (1 - eps) * e)) < 1e-18
    # Louis (1979), src/mod_phymbl.f90:1419-1479: 1 at neutrality; stable branch 1/(1 + a Ri/sqrt(1+Ri)), a = 10 / 15
    assert L.abo_f_m_louis(10.0, 0.0, 1.4e-3, 1e-3) == 1.0 and L.abo_f_h_louis(10.0, 0.0, 1.4e-3, 1e-3) == 1.0
    assert abs(L.abo_f_m_louis(10.0, 0.1, 1.4e-3, 1e-3) - 1 / (1 + 10 * 0.1 / math.sqrt(1.1))) < 1e-15
    assert abs(L.abo_f_h_louis(10.0, 0.1, 1.4e-3, 1e-3) - 1 / (1 + 15 * 0.1 / math.sqrt(1.1))) < 1e-15
    ri, cdn, z0 = -0.5, 1.4e-3, 1e-3
    ztu = ri / (1 + 3 * 25 * cdn * math.sqrt(abs(-ri * (10.0 / z0 + 1))))
    assert abs(L.abo_f_m_louis(10.0, ri, cdn, z0) - (1 - 10 * ztu)) < 1e-15
    # form drag: LU13 Ce (1-A)**(1 + 1/14), LG15 Eq.46
    assert abs(L.abo_CdN10_f_LU13(0.5) - 2.23e-3 * 0.5 ** (1 + 1 / 14.0)) < 1e-18
    assert L.abo_CdN10_f_LU13(1.0) == 0.0 and L.abo_CdN10_f_LU13(0.0) == 2.23e-3
    r = math.log(10 / 4.54e-4) / math.log(8 / 4.54e-4)
    assert abs(L.abo_CdN_f_LG15_light(8.0, 0.3, 4.54e-4) - 3.46e-3 * r * r * 0.3 * 0.7 ** 1.4) < 1e-17


def test_psi_ice_and_roughness():
    L = oracle.lib()
    assert abs(L.abo_psi_m_ice(0.5) + (0.7 * 0.5 + 0.75 * (0.5 - 14.3) * math.exp(-0.175) + 10.7)) < 1e-14
    assert L.abo_psi_m_ice(0.5) == L.abo_psi_h_ice(0.5)            # same stable branch
    x = (1 + 16 * 0.5) ** 0.25
    assert abs(L.abo_psi_h_ice(-0.5) - 2 * math.log((1 + x * x) / 2)) < 1e-14
    assert abs(L.abo_psi_m_ice(-0.5) - (math.log((1 + x * x) / 2) + 2 * math.log((1 + x) / 2) - 2 * math.atan(x) + math.pi / 2)) < 1e-14
    # psi(+0) is the stable branch at 0: -(0.75*(-14.3) + 10.7) = 0.025
    assert abs(L.abo_psi_m_ice(0.0) - 0.025) < 1e-14
    # Andreas et al. 2005 Eq.19 and the three regimes of the Andreas 1987 table
    us, nu = 0.2, 1.3e-5
    zz = (us - 0.18) / 0.1
    z0 = 0.135 * nu / us + 0.035 * us * us / 9.8 * (5 * math.exp(-zz * zz) + 1)
    assert abs(L.abo_rough_leng_m(us, nu) - z0) < 1e-18
    z0t, z0q = C.c_double(), C.c_double()
    for re_, b in ((0.1, (1.25, 0.0, 0.0)), (1.0, (0.149, -0.550, 0.0)), (10.0, (0.317, -0.565, -0.183))):
        z = re_ * nu / us
        assert L.abo_rough_leng_tq(z, us, nu, C.byref(z0t), C.byref(z0q)) == 0
        lg = math.log(us * z / nu)
        assert abs(z0t.value - z * math.exp(b[0] + b[1] * lg + b[2] * lg * lg)) < 1e-15 * z
    # the gap of the reference: for 2.49999 < R* < 2.5 no regime is selected and it calls ctl_stop
    assert L.abo_rough_leng_tq(2.499995 * nu / us, us, nu, C.byref(z0t), C.byref(z0q)) == 1


def test_ice_scenario_of_test_ice_sh():
    """Inputs of the reference's test_ice.sh (zu 10, zt 2, SIT -3 degC, A 80 %, SST -1.8 degC, t_zt 3 degC, q 4 g/kg,
    3 m/s, 1010 hPa -- nb_iter 20 as the program sets).  No captured output exists in the reference; the checks are
    physical (strongly stable: downward heat flux, deposition clipped to 0) plus regression values of THIS restatement."""
    o = OracleSession()
    o.set_nb_iter(20)
    one = lambda v: np.array([v])
    res = {}
    for ice in ("nemo", "an05", "lg15_io"):
        r = o.oce_ice(ice, "ecmwf", 2.0, 10.0, one(270.15), one(271.35), one(276.15), one(0.004), one(3.0), one(101000.0), one(0.8))
        res[ice] = r
        assert r["QH_i"][0] > 0 and r["QL_i"][0] > 0 and r["Evap_i"][0] == 0.0 and r["RiB_i"][0] > 0
        assert r["Tau"][0] == pytest.approx(0.8 * r["Tau_i"][0] + 0.2 * r["Tau_w"][0], rel=1e-15)
    assert res["nemo"]["Cd_i"][0] == 1.4e-3 and res["nemo"]["UN10_i"][0] == pytest.approx(3.0, rel=1e-12)
    assert res["an05"]["Cd_i"][0] == pytest.approx(1.54e-4, rel=5e-3)       # stable collapse of the AN05 iteration
    assert res["lg15_io"]["Cd_i"][0] == pytest.approx(1.319e-3, rel=1e-3)
    assert res["lg15_io"]["Ch_i"][0] == pytest.approx(9.73e-4, rel=1e-3)
    # the leads do not depend on the ice algorithm
    assert res["nemo"]["QH_w"][0] == res["an05"]["QH_w"][0] == res["lg15_io"]["QH_w"][0]
    lg = o.oce_ice("lg15", "ecmwf", 2.0, 10.0, one(270.15), one(271.35), one(276.15), one(0.004), one(3.0), one(101000.0), one(0.8))
    assert all(np.array_equal(lg[k], res["lg15_io"][k]) for k in lg)


def test_lg15_form_drag_comes_from_the_last_point():
    """Reference behaviour (src/ice/mod_cdn_form_ice.f90:324): every point gets the form drag of the LAST point."""
    f = synth.ice_fields(64)
    o = OracleSession()
    tha = f["t_zt"] + 0.0098 * 2.0
    L = oracle.lib()
    siq = np.array([L.abo_q_sat_ice(t, p) for t, p in zip(f["sit"], f["slp"])])
    a = o.turb_ice("lg15", 2.0, 10.0, f["sit"], tha, siq, f["hum_zt"], f["wind"], frice=f["frice"], want=("CdN_frm",))
    r = math.log(10 / 4.54e-4) / math.log(10 / 4.54e-4)
    A = f["frice"][-1]
    assert np.all(a["CdN_frm"] == a["CdN_frm"][0])
    assert a["CdN_frm"][0] == pytest.approx(3.46e-3 * r * r * A * (1 - A) ** 1.4, rel=1e-14)
    b = o.turb_ice("lg15", 2.0, 10.0, f["sit"], tha, siq, f["hum_zt"], f["wind"], frice=f["frice"], per_point_form_drag=True,
                   want=("CdN_frm",))
    assert b["CdN_frm"][-1] == a["CdN_frm"][-1] and len(np.unique(b["CdN_frm"])) > 10
    # lu12 uses LU13, which is written array-wise in the reference: per point
    c = o.turb_ice("lu12", 2.0, 10.0, f["sit"], tha, siq, f["hum_zt"], f["wind"], frice=f["frice"], want=("CdN_frm", "CdN"))
    assert len(np.unique(c["CdN_frm"])) > 10 and np.all(c["CdN"] >= c["CdN_frm"])


def test_an05_fail_stop_window():
    """With these inputs the roughness Reynolds number of one AN05 iteration falls in ]2.49999, 2.5[ (found by scanning
    the wind speed): the reference stops with 'something wrong with zsmoot, ztrans, zrough'."""
    o = OracleSession()
    one = lambda v: np.array([v])
    with pytest.raises(oracle.OracleError) as e:
        o.turb_ice("an05", 2.0, 10.0, one(263.15), one(265.15), one(0.0018), one(0.0015), one(3.239102012771403))
    assert e.value.code == 10
    o.turb_ice("an05", 2.0, 10.0, one(263.15), one(265.15), one(0.0018), one(0.0015), one(3.3))


def test_ice_all_algorithms_finite_and_ordered():
    f = synth.ice_fields(5000)
    o = OracleSession(threads=4)
    o.set_nb_iter(10)
    for ice in ("nemo", "easy", "an05", "lu12", "lg15"):
        r = o.oce_ice(ice, None, 2.0, 10.0, f["sit"], None, f["t_zt"], f["hum_zt"], f["wind"], f["slp"], f["frice"],
                      cxn=[1.4e-3, 1.3e-3, 1.3e-3], per_point_form_drag=True)
        for k in ("Cd_i", "Ch_i", "Ce_i", "theta_zu_i", "q_zu_i", "t_zu_i", "Tau_i", "QH_i", "QL_i", "Evap_i", "RiB_i"):
            assert np.all(np.isfinite(r[k])), (ice, k)
        assert np.all(r["Cd_i"] > 0) and np.all(r["Tau_i"] >= 0) and np.all(r["Evap_i"] <= 0)
        assert np.all(r["Ub_i"] >= 0.2)
        assert np.all(r["QH_w"] == 0)          # leads skipped
    with pytest.raises(oracle.OracleError):
        o.oce_ice("best", None, 2.0, 10.0, f["sit"], None, f["t_zt"], f["hum_zt"], f["wind"], f["slp"], f["frice"])


def test_series_ice_restatement_against_its_building_blocks():
    """abo_series_ice (src/ice/test_aerobulk_buoy_series_ice.f90:326-470): the bulk part coincides with the ice half of
    abo_oce_ice evaluated with each record's own concentration; the radiative terms, the gate at SIC = 0.01 and RiB at zt
    are checked against hand evaluations of the source formulas."""
    n = 4000
    f = synth.ice_fields(n, seed=21)
    rng = np.random.default_rng(5)
    rsw, rlw = np.maximum(0.0, 400.0 * rng.random(n) - 80.0), 170.0 + 140.0 * rng.random(n)
    sic = f["frice"].copy()
    sic[:8] = [0.0, 0.005, 0.01, 0.0100001, 0.02, 0.5, 1.0, 0.009999]
    o = OracleSession(threads=4)
    o.set_nb_iter(20)
    L = oracle.lib()
    ice = sic > 0.01
    assert not ice[2] and ice[3]
    for algo in ("nemo", "an05", "lu12", "lg15"):
        r = o.series_ice(algo, 2.0, 10.0, sic, f["sit"], f["t_zt"], f["hum_zt"], f["wind"], f["slp"], rsw, rlw)
        b = o.oce_ice(algo, None, 2.0, 10.0, f["sit"], None, f["t_zt"], f["hum_zt"], f["wind"], f["slp"], sic, per_point_form_drag=True)
        for a, c in (("QH", "QH_i"), ("QL", "QL_i"), ("TAU", "Tau_i"), ("SBLM", "Evap_i"), ("Cd_i", "Cd_i"), ("Ch_i", "Ch_i"),
                     ("Ce_i", "Ce_i"), ("z0", "z0_i"), ("RiB_zu", "RiB_i"), ("u_star", "u_star_i"), ("L", "L_i"),
                     ("UN10", "UN10_i"), ("theta_zu", "theta_zu_i"), ("q_zu", "q_zu_i"), ("Ublk", "Ub_i")):
            assert np.array_equal(r[a][ice], b[c][ice]), (algo, a)
        for k in o.SERIES_ICE_OUT:
            if k not in ("Qsw", "RiB_zt"):
                assert np.all(r[k][~ice] == 0.0), (algo, k)
        assert np.array_equal(r["Qsw"], (1.0 - 0.8) * rsw)
        qlw = 0.996 * (rlw - 5.67e-8 * (f["sit"] * f["sit"]) * (f["sit"] * f["sit"]))
        assert np.allclose(r["Qlw"][ice], qlw[ice], rtol=1e-15, atol=0)
        assert np.array_equal(r["QNS"][ice], (r["QH"] + r["QL"] + r["Qlw"])[ice])
        for i in (0, 3, 100, 2000):
            tha = f["t_zt"][i] + L.abo_gamma_moist(f["t_zt"][i], f["hum_zt"][i]) * 2.0
            siq = L.abo_q_sat_ice(f["sit"][i], f["slp"][i])
            ref = L.abo_Ri_bulk(2.0, f["sit"][i], tha, siq, f["hum_zt"][i], max(f["wind"][i], 0.2))
            assert r["RiB_zt"][i] == ref
        # prhoa of BULK_FORMULA: air density at zu from theta_zu - gamma_dry zu, with the pressure lowered by rho g zu
        i = 5
        ta = r["theta_zu"][i] - 9.8 / 1005.0 * 10.0
        rho = L.abo_rho_air(ta, r["q_zu"][i], f["slp"][i])
        assert r["rho_zu"][i] == pytest.approx(L.abo_rho_air(ta, r["q_zu"][i], f["slp"][i] - rho * 9.8 * 10.0), rel=1e-13)
    with pytest.raises(oracle.OracleError):
        o.series_ice("easy", 2.0, 10.0, sic, f["sit"], f["t_zt"], f["hum_zt"], f["wind"], f["slp"], rsw, rlw)
