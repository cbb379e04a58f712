"""aerobulk_gpu_set_devices(n): ONE aerobulk_model call split over n GPUs inside the library (SURVEY.md 8b "C symbols"
row, 8e).  The caller makes the reference's call (whole (Ni,Nj) host fields, src/mod_aerobulk.f90:176-269); the library
cuts the flat point range into contiguous shards (latitude row blocks), one per GPU.

Needs >= 2 visible GPUs (gpurun --gpus 2); with one GPU only the single-device degenerate cases run."""
import ctypes as C

import numpy as np
import pytest

from aerobulk_b200 import synth

pytestmark = pytest.mark.gpu
IN_KEYS = ("sst", "t_zt", "hum_zt", "U_zu", "V_zu", "slp")


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.fixture()
def ab():
    import aerobulk_b200 as ab
    ab.lib()
    ab.reset()
    yield ab
    ab.reset()
    ab.set_devices(1)


def test_set_devices_validates(ab):
    with pytest.raises(ab.AerobulkError):
        ab.set_devices(0)
    with pytest.raises(ab.AerobulkError):
        ab.set_devices(_ngpu() + 1)
    ab.set_devices(1)
    assert ab.get_devices() == 1


def test_split_call_is_bit_identical_no_skin(ab):
    """All five algorithms: the split call returns the bits of the single-GPU call (partition invariance through the
    ONE-call entry), including a grid whose shards are not whole rows (boundaries on multiples of 2048 points)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    nd = min(_ngpu(), 4)
    for Ni, Nj in ((360, 180), (1001, 37)):
        f = synth.fields(Ni, Nj)
        for algo in ("ncar", "andreas", "coare3p0", "coare3p6", "ecmwf"):
            ab.set_devices(1)
            ab.reset()
            one = ab.aerobulk_model(1, 1, algo, 2.0, 10.0, *[f[k] for k in IN_KEYS], Niter=6)
            ab.set_devices(nd)
            many = ab.aerobulk_model(1, 1, algo, 2.0, 10.0, *[f[k] for k in IN_KEYS], Niter=6)
            for k in one:
                assert np.array_equal(one[k], many[k]), (Ni, Nj, algo, k)
    plan = ab.shard_plan(1001 * 37, nd)
    assert plan[0] == 0 and plan[-1] == 1001 * 37 and all(b % 2048 == 0 for b in plan[1:-1])


@pytest.mark.parametrize("algo", ["coare3p6", "ecmwf"])
def test_split_session_carries_state_per_device(ab, algo):
    """A 6-step skin session: every step and the gathered warm-layer state bit-identical to the single-GPU session;
    pageable AND pinned caller arrays (the bounce slab and the zero-copy path exist once per device)."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    Ni, Nj, Nt = 640, 200, 6
    f = synth.fields(Ni, Nj)
    n = Ni * Nj
    nstate = 4 if algo != "ecmwf" else 1

    def run(nd, register):
        ab.set_devices(1)
        ab.reset()
        ab.set_devices(nd)
        arrays = {k: np.array(v, order="F") for k, v in f.items()}
        rsw = [synth.rad_sw_hour(Ni, Nj, 9 + jt) for jt in range(Nt)]
        outs = {k: np.zeros((Ni, Nj), order="F") for k in ("QL", "QH", "Tau_x", "Tau_y", "Evap", "T_s")}
        reg = []
        if register:
            reg = list(arrays.values()) + rsw + list(outs.values())
            for a in reg:
                ab.host_register(a)
        res, states = [], []
        try:
            for jt in range(1, Nt + 1):
                o = ab.aerobulk_model(jt, Nt, algo, 2.0, 10.0, *[arrays[k] for k in IN_KEYS], Niter=5, l_use_skin=True,
                                      rad_sw=rsw[jt - 1], rad_lw=arrays["rad_lw"], out=outs)
                res.append({k: v.copy() for k, v in o.items()})
                if jt < Nt:
                    states.append([ab.get_state(w, n) for w in range(nstate)])
                    assert all(s is not None for s in states[-1]), (nd, jt)
        finally:
            for a in reg:
                ab.host_unregister(a)
        return res, states

    ref, ref_state = run(1, False)
    for nd, register in ((2, False), (min(_ngpu(), 4), True)):
        got, got_state = run(nd, register)
        for jt in range(Nt):
            for k in ref[jt]:
                assert np.array_equal(ref[jt][k], got[jt][k]), (algo, nd, jt + 1, k)
        for jt in range(Nt - 1):
            for w in range(nstate):
                assert np.array_equal(ref_state[jt][w], got_state[jt][w]), (algo, nd, jt + 1, w)


def test_split_init_sees_the_whole_field(ab):
    """AEROBULK_INIT decides on GLOBAL statistics: a field whose humidity is 'sh'-like in one shard and would read as
    relative humidity in the other alone must be classified once, for the whole field, exactly as on one GPU; and a
    unit error that lives in one shard only stops the call on every device."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    Ni, Nj = 512, 64
    f = synth.fields(Ni, Nj, humidity="rh")
    hum = f["hum_zt"].copy()
    hum[:, : Nj // 2] = 0.01           # first shard alone: mean 0.01 -> would be detected as specific humidity
    ab.set_devices(1)
    one = ab.aerobulk_model(1, 1, "coare3p6", 2.0, 10.0, f["sst"], f["t_zt"], hum, f["U_zu"], f["V_zu"], f["slp"])
    h1 = ab.humidity_type()
    ab.reset()
    ab.set_devices(2)
    two = ab.aerobulk_model(1, 1, "coare3p6", 2.0, 10.0, f["sst"], f["t_zt"], hum, f["U_zu"], f["V_zu"], f["slp"])
    assert ab.humidity_type() == h1 == "rh"
    for k in one:
        assert np.array_equal(one[k], two[k]), k
    # specific humidity in the first shard, dew points in the second: EACH shard alone is a valid field ('sh', 'dp'), the
    # whole is neither -> the reference's type_of_humidity stops (error 5), and so must the split call
    hum2 = hum.copy()
    hum2[:, Nj // 2:] = 280.0
    codes = []
    for nd in (2, 1):
        ab.set_devices(1)
        ab.reset()
        ab.set_devices(nd)
        with pytest.raises(ab.AerobulkError) as e:
            ab.aerobulk_model(1, 1, "coare3p6", 2.0, 10.0, f["sst"], f["t_zt"], hum2, f["U_zu"], f["V_zu"], f["slp"])
        codes.append((e.value.code, e.value.message))
    assert codes[0][0] == codes[1][0] == 5
    assert codes[0][1] == codes[1][1]
    # and the error ended the session cleanly on every device: the next call works
    ab.set_devices(2)
    ok = ab.aerobulk_model(1, 1, "coare3p6", 2.0, 10.0, f["sst"], f["t_zt"], hum, f["U_zu"], f["V_zu"], f["slp"])
    assert np.array_equal(ok["QL"], one["QL"])


def test_split_reports_wind_stress_error_with_global_indices(ab):
    """tau > 10 N/m2 in the second shard: same error code and the (ji, jj) of the CALLER's grid."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    Ni, Nj = 256, 64
    f = synth.fields(Ni, Nj)
    U = f["U_zu"].copy()
    V = f["V_zu"].copy()
    U[100, 50], V[100, 50] = 48.0, 0.0       # 48 m/s: inside the sanity range (<= 50), tau > 10 N/m2 with COARE 3.6
    msgs = []
    for nd in (1, 2):
        ab.set_devices(1)
        ab.reset()
        ab.set_devices(nd)
        with pytest.raises(ab.AerobulkError) as e:
            ab.aerobulk_model(1, 1, "coare3p6", 2.0, 10.0, f["sst"], f["t_zt"], f["hum_zt"], U, V, f["slp"])
        assert e.value.code == 8
        msgs.append(e.value.message)
    assert "ji, jj = 0101, 0051" in msgs[0], msgs[0]
    assert msgs[0] == msgs[1]


def test_cxx_bridge_symbol_is_split_too(ab):
    """The drop-in symbol the reference's C++ wrapper binds (src/aerobulk.cpp:5-19) goes through the same split."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    L = ab.lib()
    m = 50000
    f = synth.fields(m, 1)
    flat = {k: np.ascontiguousarray(v.ravel()) for k, v in f.items()}
    outs = {}
    for nd in (1, 2):
        ab.set_devices(1)
        ab.reset()
        ab.set_devices(nd)
        o = [np.zeros(m) for _ in range(5)]
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        i = lambda v: C.byref(C.c_int(v))
        d = lambda v: C.byref(C.c_double(v))
        L.aerobulk_cxx_no_skin(i(1), i(1), b"ecmwf", d(2.0), d(10.0), *[p(flat[k]) for k in IN_KEYS], *[p(a) for a in o],
                               i(7), i(5), i(m))
        outs[nd] = o
    for a, b in zip(outs[1], outs[2]):
        assert np.array_equal(a, b)


def test_split_call_with_speculative_init_per_shard(ab):
    """Shards big enough for several pipeline chunks each (>= 400 000 points): every shard runs its chunks on its LOCAL
    running AEROBULK_INIT verdict and recomputes what the GLOBAL verdict (statistics of all devices) overrules.  A
    relative-humidity field whose first rows -- the whole first chunk of shard 0 -- read as specific humidity."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    Ni, Nj = 1440, 720
    f = synth.fields(Ni, Nj, humidity="rh")
    f["hum_zt"][:, :150] = 0.02
    kw = dict(Niter=5, l_use_skin=True, rad_sw=f["rad_sw"], rad_lw=f["rad_lw"])
    ab.set_devices(1)
    ab.reset()
    one = ab.aerobulk_model(1, 1, "ecmwf", 2.0, 10.0, *[f[k] for k in IN_KEYS], **kw)
    assert ab.humidity_type() == "rh"
    ab.reset()
    ab.set_devices(2)
    two = ab.aerobulk_model(1, 1, "ecmwf", 2.0, 10.0, *[f[k] for k in IN_KEYS], **kw)
    assert ab.humidity_type() == "rh"
    for k in one:
        assert np.array_equal(one[k], two[k]), k
