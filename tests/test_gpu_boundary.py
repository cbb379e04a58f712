"""Boundary behaviour of the C ABI on the GPU: session semantics of AEROBULK_MODEL
(sticky globals, optional arguments, fail-stop conditions), the two C++-bridge symbols and
size-independent properties at BASELINE's full grid sizes."""
import ctypes as C

import numpy as np
import pytest

from aerobulk_b200 import synth

pytestmark = pytest.mark.gpu
IN_KEYS = ("sst", "t_zt", "hum_zt", "U_zu", "V_zu", "slp")


@pytest.fixture(scope="module")
def ab():
    import aerobulk_b200 as ab
    ab.lib()
    return ab


def _ins(f):
    return [f[k] for k in IN_KEYS]


def test_sticky_globals_and_optionals(ab):
    f = synth.fields(64, 32)
    ab.reset()
    assert ab.nb_iter() == 5 and not ab.use_skin() and ab.humidity_type() == "sh"
    o = ab.aerobulk_model(1, 1, "ncar", 2.0, 10.0, *_ins(f))
    assert "T_s" not in o                                  # T_s only with rad_sw & rad_lw
    ab.aerobulk_model(1, 1, "ncar", 2.0, 10.0, *_ins(f), Niter=8)
    assert ab.nb_iter() == 8                                # Niter is sticky (mod_aerobulk.f90:236)
    o8 = ab.aerobulk_model(1, 1, "ncar", 2.0, 10.0, *_ins(f))
    o8b = ab.aerobulk_model(1, 1, "ncar", 2.0, 10.0, *_ins(f), Niter=8)
    assert np.array_equal(o8["QL"], o8b["QL"])
    # radiation given but skin not requested: bulk SST is returned as T_s
    o = ab.aerobulk_model(1, 1, "coare3p6", 2.0, 10.0, *_ins(f), rad_sw=f["rad_sw"], rad_lw=f["rad_lw"])
    assert np.array_equal(o["T_s"], f["sst"]) and not ab.use_skin()
    o = ab.aerobulk_model(1, 1, "coare3p6", 2.0, 10.0, *_ins(f), l_use_skin=True, rad_sw=f["rad_sw"], rad_lw=f["rad_lw"])
    assert ab.use_skin() and not np.array_equal(o["T_s"], f["sst"])
    # l_use_skin_schemes is sticky (mod_aerobulk.f90:74): a later call without the flag still uses the skin
    o2 = ab.aerobulk_model(1, 1, "coare3p6", 2.0, 10.0, *_ins(f), rad_sw=f["rad_sw"], rad_lw=f["rad_lw"])
    assert np.array_equal(o2["T_s"], o["T_s"])
    # ... while ncar ignores it
    o3 = ab.aerobulk_model(1, 1, "ncar", 2.0, 10.0, *_ins(f), rad_sw=f["rad_sw"], rad_lw=f["rad_lw"])
    assert np.array_equal(o3["T_s"], f["sst"])


def test_fail_stop_conditions(ab):
    f = synth.fields(32, 16)
    ab.reset()
    with pytest.raises(ab.AerobulkError) as e:
        ab.aerobulk_model(0, 1, "ncar", 2.0, 10.0, *_ins(f))
    assert e.value.code == 1
    with pytest.raises(ab.AerobulkError) as e:
        ab.aerobulk_model(1, 1, "bogus", 2.0, 10.0, *_ins(f))
    assert e.value.code == 7
    with pytest.raises(ab.AerobulkError) as e:                # skin only for coare*/ecmwf
        ab.aerobulk_model(1, 1, "ncar", 2.0, 10.0, *_ins(f), l_use_skin=True, rad_sw=f["rad_sw"], rad_lw=f["rad_lw"])
    assert e.value.code == 2
    ab.reset()
    with pytest.raises(ab.AerobulkError) as e:                # skin needs radiation
        ab.aerobulk_model(1, 1, "ecmwf", 2.0, 10.0, *_ins(f), l_use_skin=True)
    assert e.value.code == 3
    ab.reset()
    g = dict(f)
    g["sst"] = f["sst"] - 273.15                              # deg C instead of K: everything masked
    with pytest.raises(ab.AerobulkError) as e:
        ab.aerobulk_model(1, 1, "ncar", 2.0, 10.0, *_ins(g))
    assert e.value.code == 4
    g = dict(f)
    g["slp"] = f["slp"].copy()
    g["slp"][3, 3] = 1013.0                                   # one point in hPa: masked, not fatal
    ab.reset()
    ab.aerobulk_model(1, 1, "ncar", 2.0, 10.0, *_ins(g))
    g = dict(f)
    g["hum_zt"] = f["hum_zt"] * 1000.0 + 110.0                # neither sh, rh nor dp
    with pytest.raises(ab.AerobulkError) as e:
        ab.aerobulk_model(1, 1, "ncar", 2.0, 10.0, *_ins(g))
    assert e.value.code == 5
    g = dict(f)
    g["U_zu"] = f["U_zu"] * 0.0 + 48.0                         # |U| <= 50 passes the mask, tau > 10 N/m2 stops
    g["V_zu"] = f["V_zu"] * 0.0
    ab.reset()
    with pytest.raises(ab.AerobulkError) as e:
        ab.aerobulk_model(1, 1, "coare3p6", 2.0, 10.0, *_ins(g))
    assert e.value.code == 8 and "wind stress too strong" in e.value.message
    # double jt==1 of a skin session before jt==Nt: the reference's ALLOCATE fails
    ab.reset()
    kw = dict(l_use_skin=True, rad_sw=f["rad_sw"], rad_lw=f["rad_lw"])
    ab.aerobulk_model(1, 3, "coare3p0", 2.0, 10.0, *_ins(f), **kw)
    with pytest.raises(ab.AerobulkError) as e:
        ab.aerobulk_model(1, 3, "coare3p6", 2.0, 10.0, *_ins(f), **kw)   # 3.0 and 3.6 share mod_skin_coare's arrays
    assert e.value.code == 9
    ab.reset()


def test_empty_and_tiny_inputs(ab):
    ab.reset()
    z = np.zeros((0,), dtype=np.float64)
    with pytest.raises(ab.AerobulkError):      # the reference divides by SUM(mask)=0 -> "whole domain is masked"
        ab.aerobulk_model(1, 1, "ncar", 2.0, 10.0, z, z, z, z, z, z)
    f = synth.fields(1, 1)
    ab.reset()
    o = ab.aerobulk_model(1, 1, "andreas", 2.0, 10.0, *_ins(f))
    assert o["QL"].shape == (1, 1) and np.isfinite(o["QL"]).all()
    f = synth.fields(129, 7)                   # ragged: not a multiple of the block size
    ab.reset()
    o = ab.aerobulk_model(1, 1, "ecmwf", 2.0, 10.0, *_ins(f))
    assert np.isfinite(o["QH"]).all()


def test_calm_points_have_zero_stress(ab):
    f = synth.fields(512, 256)
    calm = (f["U_zu"] == 0) & (f["V_zu"] == 0)
    assert calm.sum() > 0
    ab.reset()
    o = ab.aerobulk_model(1, 1, "coare3p0", 2.0, 10.0, *_ins(f))
    assert (o["Tau_x"][calm] == 0).all() and (o["Tau_y"][calm] == 0).all()   # zWzu <= 1e-3, mod_aerobulk_compute.f90:191
    assert np.isfinite(o["QL"][calm]).all()


def test_cxx_bridge_symbols(ab):
    """aerobulk_cxx_skin / aerobulk_cxx_no_skin exactly as src/aerobulk.cpp:105-108,135-137 calls them."""
    L = ab.lib()
    ab.reset()
    m = 2
    rt0 = 273.15
    arr = lambda *v: np.array(v, dtype=np.float64)
    sst, t, q, U, V, slp = arr(22 + rt0, 22 + rt0), arr(20 + rt0, 25 + rt0), arr(.012, .012), arr(4., 4.), arr(9., 9.), arr(101000., 101000.)
    rsw, rlw = arr(0., 0.), arr(350., 350.)
    outs = [np.zeros(m) for _ in range(6)]
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ci = lambda v: C.byref(C.c_int(v))
    cd = lambda v: C.byref(C.c_double(v))
    algo = b"coare3p6"
    L.aerobulk_cxx_skin(ci(1), ci(1), algo, cd(2.0), cd(10.0), p(sst), p(t), p(q), p(U), p(V), p(slp),
                        *[p(o) for o in outs[:5]], ci(8), C.byref(C.c_bool(True)), p(rsw), p(rlw), p(outs[5]),
                        ci(len(algo)), ci(m))
    ref = ab.aerobulk_model(1, 1, "coare3p6", 2.0, 10.0, sst, t, q, U, V, slp, Niter=8, l_use_skin=True, rad_sw=rsw, rad_lw=rlw)
    for o, k in zip(outs, ("QL", "QH", "Tau_x", "Tau_y", "Evap", "T_s")):
        assert np.array_equal(o, ref[k]), k
    assert outs[5][0] < 22 + rt0          # cool skin at night
    algo = b"ncar"
    outs2 = [np.zeros(m) for _ in range(5)]
    L.aerobulk_cxx_no_skin(ci(1), ci(1), algo, cd(2.0), cd(10.0), p(sst), p(t), p(q), p(U), p(V), p(slp),
                           *[p(o) for o in outs2], ci(8), ci(len(algo)), ci(m))
    ref = ab.aerobulk_model(1, 1, "ncar", 2.0, 10.0, sst, t, q, U, V, slp, Niter=8)
    for o, k in zip(outs2, ("QL", "QH", "Tau_x", "Tau_y", "Evap")):
        assert np.array_equal(o, ref[k]), k


@pytest.mark.parametrize("algo,skin", [("ecmwf", True), ("andreas", False), ("coare3p0", False)])
def test_full_size_properties_4320x2160(ab, algo, skin):
    """BASELINE configs 3/4 at full size (9.3 M points): size-independent properties instead of the oracle.
    (a) a strided sample of the big grid equals the same points computed alone (independence of points,
    chunking/pipelining does not leak between points); (b) outputs finite; (c) rotating the wind by 90 deg
    rotates the stress and leaves the scalar fluxes bit-identical."""
    Ni, Nj = 4320, 2160
    f = synth.fields(Ni, Nj)
    kw = dict(l_use_skin=True, rad_sw=f["rad_sw"], rad_lw=f["rad_lw"]) if skin else {}
    ab.reset()
    big = ab.aerobulk_model(1, 1, algo, 2.0, 10.0, *_ins(f), **kw)
    for k, v in big.items():
        assert np.isfinite(v).all(), k
    sl = (slice(7, None, 97), slice(3, None, 89))
    sub = {k: np.asfortranarray(v[sl]) for k, v in f.items()}
    kws = dict(l_use_skin=True, rad_sw=sub["rad_sw"], rad_lw=sub["rad_lw"]) if skin else {}
    ab.reset()
    small = ab.aerobulk_model(1, 1, algo, 2.0, 10.0, *_ins(sub), **kws)
    for k in small:
        assert np.array_equal(small[k], big[k][sl]), k
    rot = dict(f)
    rot["U_zu"], rot["V_zu"] = np.asfortranarray(-f["V_zu"]), f["U_zu"]
    ab.reset()
    r = ab.aerobulk_model(1, 1, algo, 2.0, 10.0, *_ins(rot), **kw)
    assert np.array_equal(r["QL"], big["QL"]) and np.array_equal(r["QH"], big["QH"]) and np.array_equal(r["Evap"], big["Evap"])
    assert np.array_equal(r["Tau_x"], -big["Tau_y"]) and np.array_equal(r["Tau_y"], big["Tau_x"])
