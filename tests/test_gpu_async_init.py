"""Asynchronous AEROBULK_INIT of device-resident sessions: the stats-dependent half of src/mod_aerobulk.f90:100-160 runs
on the device (init_decide_kernel), the kernels read its verdict from device memory, the host catches up -- and raises the
reference's error -- at its next synchronisation.  Same results as the blocking host-array path, bit for bit."""
import numpy as np
import pytest

from aerobulk_b200 import synth

pytestmark = pytest.mark.gpu
IN_KEYS = ("sst", "t_zt", "hum_zt", "U_zu", "V_zu", "slp")
OUT_KEYS = ("QL", "QH", "Tau_x", "Tau_y", "Evap", "T_s")


@pytest.fixture()
def ab():
    import aerobulk_b200 as ab
    ab.lib()
    ab.reset()
    ab.set_verbose(False)
    yield ab
    ab.set_stream(None)
    ab.reset()


def _dev(f):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v.ravel(order="F"))).cuda() for k, v in f.items()}


@pytest.mark.parametrize("hum", ["sh", "rh", "dp"])
def test_async_init_session_equals_blocking_session(ab, hum):
    """3-step COARE 3.6 skin session on device tensors with the banners off: no synchronisation between jt = 1 and jt = Nt,
    humidity type decided on the device; every step equals the host-array session."""
    import torch
    Ni, Nj, Nt = 200, 96, 3
    f = synth.fields(Ni, Nj, humidity=hum)
    host = []
    for jt in range(1, Nt + 1):
        o = ab.aerobulk_model(jt, Nt, "coare3p6", 2.0, 10.0, *[f[k] for k in IN_KEYS], Niter=5, l_use_skin=True,
                              rad_sw=f["rad_sw"], rad_lw=f["rad_lw"])
        host.append({k: v.ravel(order="F").copy() for k, v in o.items()})
    assert ab.humidity_type() == hum
    ab.reset()
    ab.set_verbose(False)
    d = _dev(f)
    st = torch.cuda.Stream()
    ab.set_stream(st.cuda_stream)
    outs = [{k: torch.empty(Ni * Nj, dtype=torch.float64, device="cuda") for k in OUT_KEYS} for _ in range(Nt)]
    for jt in range(1, Nt + 1):
        ab.aerobulk_model_device(jt, Nt, "coare3p6", 2.0, 10.0, *[d[k] for k in IN_KEYS], out=outs[jt - 1], Niter=5,
                                 l_use_skin=True, rad_sw=d["rad_sw"], rad_lw=d["rad_lw"], shape=(Ni, Nj))
        if jt == 1:
            assert ab.humidity_type() == hum       # the device's verdict, read back on demand
    ab.synchronize()
    for jt in range(Nt):
        for k in OUT_KEYS:
            assert np.array_equal(outs[jt][k].cpu().numpy(), host[jt][k]), (hum, jt + 1, k)


def test_async_init_failure_surfaces_at_synchronize(ab):
    """An unidentifiable humidity field: the device-side init flags it, the flux kernel computes nothing, and the host
    raises the reference's type_of_humidity error (code 5) at the next synchronising call; the session is then over."""
    import torch
    Ni, Nj = 128, 32
    f = synth.fields(Ni, Nj)
    f["hum_zt"] = f["hum_zt"] * 1000.0 + 110.0     # neither sh, rh nor dp
    d = _dev(f)
    out = {k: torch.full((Ni * Nj,), -777.0, dtype=torch.float64, device="cuda") for k in OUT_KEYS[:5]}
    ab.aerobulk_model_device(1, 4, "ecmwf", 2.0, 10.0, *[d[k] for k in IN_KEYS], out=out, shape=(Ni, Nj))
    ab.aerobulk_model_device(2, 4, "ecmwf", 2.0, 10.0, *[d[k] for k in IN_KEYS], out=out, shape=(Ni, Nj))
    with pytest.raises(ab.AerobulkError) as e:
        ab.synchronize()
    assert e.value.code == 5 and "un-identified humidity type" in e.value.message
    assert float(out["QL"].min()) == -777.0 == float(out["QL"].max())      # nothing was computed after the failed init
    # all masked (SST in Celsius everywhere) -> error 4, same route
    g = synth.fields(Ni, Nj)
    g["sst"] = g["sst"] - 273.15
    d = _dev(g)
    ab.aerobulk_model_device(1, 2, "ncar", 2.0, 10.0, *[d[k] for k in IN_KEYS], out=out, shape=(Ni, Nj))
    with pytest.raises(ab.AerobulkError) as e:
        ab.synchronize()
    assert e.value.code == 4
    # a clean session afterwards
    d = _dev(synth.fields(Ni, Nj))
    ab.aerobulk_model_device(1, 1, "ncar", 2.0, 10.0, *[d[k] for k in IN_KEYS], out=out, shape=(Ni, Nj))
    ab.synchronize()
    assert np.isfinite(out["QL"].cpu().numpy()).all() and float(out["QL"].max()) != -777.0


def test_gathered_stats_init_on_device(ab):
    """The sharded AEROBULK_INIT without a host round trip: row-block vectors written to device memory, gathered (here:
    concatenated, as one all-gather would), combined and judged on the device.  Each row block then equals the same rows
    of the one-shot call, including a humidity type only the WHOLE field determines."""
    import torch
    from aerobulk_b200 import model as abm
    Ni, Nj = 96, 64
    full = synth.fields(Ni, Nj, humidity="rh")
    full["hum_zt"][:, :20] = 0.01                   # the first block alone would read as specific humidity
    want = ab.aerobulk_model(1, 1, "coare3p0", 2.0, 10.0, *[full[k] for k in IN_KEYS], Niter=5)
    assert ab.humidity_type() == "rh"
    blocks = ((0, 20), (20, 64))
    devs = [_dev({k: np.asfortranarray(v[:, a:b]) for k, v in full.items()}) for a, b in blocks]
    ab.reset()
    ab.set_verbose(False)
    gathered = torch.empty((2, 64), dtype=torch.float64, device="cuda")
    for r, d in enumerate(devs):
        abm.init_local_stats_device(*[d[k] for k in IN_KEYS], rad_lw=None, out=gathered[r])
    host_parts = [abm.init_local_stats(*[d[k] for k in IN_KEYS]) for d in devs]
    assert np.array_equal(gathered.cpu().numpy(), np.stack(host_parts))
    for (a, b), d in zip(blocks, devs):
        n = Ni * (b - a)
        out = {k: torch.empty(n, dtype=torch.float64, device="cuda") for k in OUT_KEYS[:5]}
        abm.init_from_gathered_stats(1, "coare3p0", None, False, gathered, 2)
        ab.aerobulk_model_device(1, 1, "coare3p0", 2.0, 10.0, *[d[k] for k in IN_KEYS], out=out, Niter=5, shape=(Ni, b - a))
        ab.synchronize()
        assert ab.humidity_type() == "rh"
        for k in out:
            assert np.array_equal(out[k].cpu().numpy(), want[k][:, a:b].ravel(order="F")), (a, b, k)


def _device_call(ab, algo, f, Ni, Nj, skin):
    """The non-speculative route to the same numbers: device tensors, asynchronous init over the whole field."""
    import torch
    d = _dev(f)
    keys = OUT_KEYS if skin else OUT_KEYS[:5]
    out = {k: torch.empty(Ni * Nj, dtype=torch.float64, device="cuda") for k in keys}
    kw = dict(l_use_skin=True, rad_sw=d["rad_sw"], rad_lw=d["rad_lw"]) if skin else {}
    ab.aerobulk_model_device(1, 1, algo, 2.0, 10.0, *[d[k] for k in IN_KEYS], out=out, Niter=5, shape=(Ni, Nj), **kw)
    ab.synchronize()
    return {k: v.cpu().numpy().reshape((Ni, Nj), order="F") for k, v in out.items()}


@pytest.mark.parametrize("algo,skin", [("coare3p6", True), ("ecmwf", False)])
def test_speculative_init_of_the_staged_pipeline(ab, algo, skin):
    """jt == 1 of a host-array call with several pipeline chunks: each chunk's flux launch runs on the RUNNING verdict of
    AEROBULK_INIT (statistics of the chunks so far) and is recomputed if the final verdict differs.  Three fields:
    (a) ordinary; (b) relative humidity whose first chunk alone reads as specific humidity (wrong running verdict ->
    recomputed); (c) a first chunk that is entirely masked (running verdict 'whole domain masked' -> nothing computed ->
    recomputed).  All equal the single-verdict device path bit for bit."""
    Ni, Nj = 1440, 720                       # 1 036 800 points: 5 chunks
    assert ab.lib().aerobulk_gpu_chunk_plan(Ni * Nj, 1, (__import__("ctypes").c_longlong * 17)()) >= 2
    cases = {"plain": synth.fields(Ni, Nj)}
    rh = synth.fields(Ni, Nj, humidity="rh")
    rh["hum_zt"][:, :300] = 0.02
    cases["rh with an sh-like first chunk"] = rh
    cold = synth.fields(Ni, Nj)
    cold["sst"][:, :330] -= 273.15           # degrees Celsius: masked out by the sanity range, the rest is fine
    cases["first chunk masked"] = cold
    for name, f in cases.items():
        ab.reset()
        ab.set_verbose(False)
        kw = dict(l_use_skin=True, rad_sw=f["rad_sw"], rad_lw=f["rad_lw"]) if skin else {}
        got = ab.aerobulk_model(1, 1, algo, 2.0, 10.0, *[f[k] for k in IN_KEYS], Niter=5, **kw)
        hum = ab.humidity_type()
        ab.reset()
        ab.set_verbose(False)
        want = _device_call(ab, algo, f, Ni, Nj, skin)
        assert ab.humidity_type() == hum, name
        for k in want:
            assert np.array_equal(got[k], want[k], equal_nan=True), (name, k)


def test_speculative_init_raises_the_final_verdict(ab):
    """A humidity field that only becomes unidentifiable with its LAST chunk: the error of the reference (code 5)."""
    Ni, Nj = 1440, 720
    f = synth.fields(Ni, Nj)
    f["hum_zt"][:, 700:] = 280.0             # dew points in the last rows of a specific-humidity field
    with pytest.raises(ab.AerobulkError) as e:
        ab.aerobulk_model(1, 1, "ncar", 2.0, 10.0, *[f[k] for k in IN_KEYS])
    assert e.value.code == 5
    g = synth.fields(Ni, Nj)
    ok = ab.aerobulk_model(1, 1, "ncar", 2.0, 10.0, *[g[k] for k in IN_KEYS])     # and the next session is clean
    assert np.isfinite(ok["QL"]).all()
