"""GPU parity: the CUDA path (through the C ABI of libaerobulk_gpu.so) against the CPU oracle
on identical seeded inputs, at the sizes BASELINE.json names.

Tolerance (stated by BASELINE.json north_star, metric of SURVEY.md 8d):
    err_f = |gpu - oracle| / (|oracle| + S_f),  S = 10 W/m2 (QL,QH), 1e-2 N/m2 (tau), 1e-5 kg/m2/s (Evap), 1 K (T_s)
    gate: err <= 1e-10 at EVERY point, except points PROVEN to sit on a discontinuity of the reference algorithm:
    the oracle is re-run on each such point with its inputs nudged by 1..64 ulp and must itself move by at least a
    tenth of the GPU's deviation (tests/parity_util.py).  Unproven points fail; proven ones are counted, must stay
    below 2e-5 of the grid and below 1e-3.

The oracle (oracle/aerobulk_oracle.c) is pinned to the reference's own captured output doc/ex_ab.dat
(tests/test_oracle_golden.py); what no reference fixture pins -- warm-layer integration with Rsw > 0, rh/dp
humidity, zt == zu, nb_iter != 50 -- is defined by the restatement alone ("parity unpinned").
"""
import math
import os

import numpy as np
import pytest

import parity_util as pu
from aerobulk_b200 import synth

pytestmark = pytest.mark.gpu

TOL = pu.TOL
IN_KEYS = pu.IN_KEYS
NTHREADS = os.cpu_count() or 8


@pytest.fixture(scope="module")
def ab():
    import aerobulk_b200 as ab
    ab.lib()
    return ab


@pytest.fixture(scope="module")
def oracle():
    from oracle.oracle import OracleSession
    return OracleSession


def _single_call(ab, oracle, tag, algo, zt, zu, f, nb_iter=5, skin=False):
    """One aerobulk_model call (jt = Nt = 1) on the GPU and on the oracle; gate + branch-flip proof."""
    kw = dict(Niter=nb_iter)
    if skin:
        kw.update(l_use_skin=True, rad_sw=f["rad_sw"], rad_lw=f["rad_lw"])
    ab.reset()
    got = ab.aerobulk_model(1, 1, algo, zt, zu, *[f[k] for k in IN_KEYS], **kw)
    ref = oracle(threads=NTHREADS).model(1, 1, algo, zt, zu, *[f[k] for k in IN_KEYS], **kw)
    if skin:
        assert "T_s" in got
    inputs = {k: f[k] for k in IN_KEYS}
    inputs["rad_sw"] = f["rad_sw"] if skin else None
    inputs["rad_lw"] = f["rad_lw"] if skin else None
    rep = pu.assert_parity(tag, pu.worst_per_point(got, ref), inputs,
                           pu.oracle_runner(oracle, algo, zt, zu, nb_iter, skin))
    return got, ref, rep


@pytest.mark.parametrize("algo", ["ncar", "andreas", "coare3p0", "coare3p6", "ecmwf"])
@pytest.mark.parametrize("nb_iter", [5, 10])
def test_noskin_360x180(ab, oracle, algo, nb_iter):
    """BASELINE config 1 (1 deg grid 360x180, bulk SST, one call) for all five algorithms."""
    _single_call(ab, oracle, f"C1 {algo} noskin nb_iter={nb_iter}", algo, 2.0, 10.0, synth.fields(360, 180), nb_iter)


@pytest.mark.parametrize("algo", ["coare3p0", "coare3p6", "ecmwf"])
def test_skin_single_step(ab, oracle, algo):
    _single_call(ab, oracle, f"{algo} skin 360x180", algo, 2.0, 10.0, synth.fields(360, 180), 5, skin=True)


STATE_SCALE = (1.0, 20.0, 1e7, 1e3)   # dT_wl [K], Hz_wl [m], Qnt_ac [J/m2], Tau_ac [N s/m2]


def _session(ab, oracle, tag, algo, nb_iter, Ni, Nj, Nt=24, check_state=True):
    """A state-carrying skin session of Nt hourly steps (diurnal rad_sw): fluxes compared at every step, the
    device-resident warm-layer state after every step; the gate applies to the per-point worst over all steps."""
    f = synth.fields(Ni, Nj)
    n = Ni * Nj
    ab.reset()
    osess = oracle(threads=NTHREADS)
    worst = np.zeros(n)
    worst_state = np.zeros(n)
    rsw_all = []
    nstate = 4 if algo != "ecmwf" else 1
    for jt in range(1, Nt + 1):
        rsw = synth.rad_sw_hour(Ni, Nj, jt)
        rsw_all.append(rsw)
        kw = dict(Niter=nb_iter, l_use_skin=True, rad_sw=rsw, rad_lw=f["rad_lw"])
        got = ab.aerobulk_model(jt, Nt, algo, 2.0, 10.0, *[f[k] for k in IN_KEYS], **kw)
        ref = osess.model(jt, Nt, algo, 2.0, 10.0, *[f[k] for k in IN_KEYS], **kw)
        worst = np.maximum(worst, pu.worst_per_point(got, ref))
        if jt < Nt:
            if check_state:
                for which in range(nstate):
                    sg = ab.get_state(which, n)
                    so = osess.state(which, n)
                    assert sg is not None and so is not None
                    worst_state = np.maximum(worst_state, np.abs(sg - so) / (np.abs(so) + STATE_SCALE[which]))
        else:
            assert ab.get_state(0, n) is None   # *_EXIT at jt == nitend frees the state
    inputs = {k: f[k] for k in IN_KEYS}
    inputs["rad_sw"] = rsw_all
    inputs["rad_lw"] = f["rad_lw"]
    rep = pu.assert_parity(tag, worst, inputs, pu.oracle_runner(oracle, algo, 2.0, 10.0, nb_iter, True, nt=Nt))
    if check_state:
        # the state is compared with the same gate; a point whose FLUXES were proven to sit on a branch may carry a
        # different state, every other point must agree to 1e-9 (state scales above)
        bad_state = np.flatnonzero(worst_state > 1e-9)
        flux_bad = set(np.flatnonzero(worst > TOL).tolist())
        unexplained = [int(i) for i in bad_state if int(i) not in flux_bad]
        print(f"[parity] {tag}: warm-layer state max scaled err {worst_state.max():.3e}, points>1e-9: {bad_state.size}")
        assert not unexplained, (tag, "warm-layer state differs where the fluxes agree", unexplained[:10])
    return rep


@pytest.mark.parametrize("algo,nb_iter", [("coare3p6", 5), ("coare3p6", 6), ("coare3p0", 10), ("ecmwf", 5)])
def test_skin_24_steps_state_carried(ab, oracle, algo, nb_iter):
    """BASELINE config 2 in small (144x72): 24 hourly steps, warm-layer state device-resident between calls.
    nb_iter 5/6/10 exercise the MOD(nb_iter,jit) commit quirk."""
    _session(ab, oracle, f"{algo} nb_iter={nb_iter} 144x72x24", algo, nb_iter, 144, 72)


def test_config2_full_size(ab, oracle):
    """BASELINE config 2 AT FULL SIZE: COARE 3.6 + cool-skin/warm-layer, 1440x720 (1 036 800 points), 24 hourly steps,
    state carried on the device and compared with the oracle's after EVERY step.
    Reference loop body: src/mod_blk_coare3p6.f90:302-383, src/mod_skin_coare.f90:97-250."""
    _session(ab, oracle, "C2 coare3p6 skin 1440x720x24 FULL", "coare3p6", 5, 1440, 720)


def test_config2_full_size_ecmwf(ab, oracle):
    """The same session with the ECMWF skin scheme (prognostic warm layer advanced at every iteration,
    src/mod_skin_ecmwf.f90:113-230): the case with the worst round-1 margin (1.4e-11)."""
    _session(ab, oracle, "C2-like ecmwf skin 1440x720x24 FULL", "ecmwf", 5, 1440, 720)


@pytest.mark.parametrize("algo", ["ncar", "coare3p6", "ecmwf", "andreas", "coare3p0"])
def test_zt_equal_zu(ab, oracle, algo):
    _single_call(ab, oracle, f"{algo} zt==zu", algo, 10.0, 10.0, synth.fields(180, 90))


@pytest.mark.parametrize("algo", ["coare3p6", "ecmwf", "ncar"])
@pytest.mark.parametrize("hum", ["rh", "dp"])
def test_humidity_types(ab, oracle, hum, algo):
    f = synth.fields(180, 90, humidity=hum)
    _single_call(ab, oracle, f"{algo} humidity {hum}", algo, 2.0, 10.0, f)
    assert ab.humidity_type() == hum


def test_golden_ex_ab_on_gpu(ab, golden_dir):
    """The reference's own captured output (doc/ex_ab.dat) straight through the CUDA path."""
    import json, os
    g = json.load(open(os.path.join(golden_dir, "ex_ab.json")))
    inp = g["inputs"]
    rt0 = inp["rt0"]
    ins = [np.array(inp["sst_C"]) + rt0, np.array(inp["t_zt_C"]) + rt0, np.array(inp["q_zt"]),
           np.array(inp["U_zu"]), np.array(inp["V_zu"]), np.array(inp["slp"])]
    for algo in ("coare3p6", "ecmwf", "ncar", "andreas"):   # coare3p0 rows of the file are stale
        ab.reset()
        kw = dict(Niter=50)
        if algo in ("coare3p6", "ecmwf"):
            kw.update(l_use_skin=True, rad_sw=np.array(inp["rad_sw"]), rad_lw=np.array(inp["rad_lw"]))
        o = ab.aerobulk_model(1, 1, algo, inp["zt"], inp["zu"], *ins, **kw)
        ga = g["algos"][algo]
        pairs = [("QH", o["QH"]), ("QL", o["QL"]), ("Evap_mm_day", o["Evap"] * 86400.0), ("Tau_x", o["Tau_x"])]
        if "T_s" in o:
            pairs.append(("SSST_C", o["T_s"] - rt0))
        for key, val in pairs:
            for k in range(2):
                s = ga[key + "_str"][k]
                mant, _, ex = s.upper().partition("E")
                last = 10.0 ** (-len(mant.split(".")[1]) + (int(ex) if ex else 0))
                assert abs(val[k] - float(s)) <= 0.5 * last + abs(float(s)) * 2.0 ** -23, (algo, key, k, val[k], s)


def test_survey_b3_series_on_gpu(ab, golden_dir):
    """one-point 24 h warm-layer series, state compared after selected steps."""
    import json, os
    k = json.load(open(os.path.join(golden_dir, "survey_kat.json")))
    for algo, nb in sorted({(r["algo"], r["nb_iter"]) for r in k["B3"]}):
        ab.reset()
        want = {r["jt"]: r for r in k["B3"] if r["algo"] == algo and r["nb_iter"] == nb}
        one = lambda v: np.array([v], dtype=float)
        for jt in range(1, 25):
            rsw = max(0.0, 900.0 * math.sin(math.pi * (jt - 6) / 12.0))
            o = ab.aerobulk_model(jt, 24, algo, 2.0, 10.0, one(301.15), one(300.15), one(0.018), one(3.0), one(1.0),
                                  one(101000.0), Niter=nb, l_use_skin=True, rad_sw=one(rsw), rad_lw=one(400.0))
            if jt in want:
                r = want[jt]
                for key in ("T_s", "QL", "QH"):
                    assert o[key][0] == pytest.approx(r[key], rel=1e-10), (algo, nb, jt, key)
                if jt < 24 and r["dT_wl"] is not None:
                    assert ab.get_state(0, 1)[0] == pytest.approx(r["dT_wl"], rel=1e-9, abs=1e-13)


def test_partition_invariance(ab):
    """Row-block sharding: any partition of the grid gives bit-identical fluxes."""
    Ni, Nj = 128, 96
    f = synth.fields(Ni, Nj)
    ab.reset()
    full = ab.aerobulk_model(1, 1, "coare3p6", 2.0, 10.0, *[f[k] for k in IN_KEYS], l_use_skin=True,
                             rad_sw=f["rad_sw"], rad_lw=f["rad_lw"])
    for parts in (2, 3):
        edges = np.linspace(0, Nj, parts + 1).astype(int)
        for j0, j1 in zip(edges[:-1], edges[1:]):
            b = synth.fields(Ni, Nj, j0=j0, j1=j1)
            ab.reset()
            blk = ab.aerobulk_model(1, 1, "coare3p6", 2.0, 10.0, *[b[k] for k in IN_KEYS], l_use_skin=True,
                                    rad_sw=b["rad_sw"], rad_lw=b["rad_lw"])
            for key in full:
                assert np.array_equal(blk[key], full[key][:, j0:j1]), (parts, j0, key)


def test_device_api_matches_host_api(ab):
    import torch
    Ni, Nj = 256, 64
    f = synth.fields(Ni, Nj)
    ab.reset()
    kw = dict(Niter=5, l_use_skin=True)
    host = ab.aerobulk_model(1, 2, "ecmwf", 2.0, 10.0, *[f[k] for k in IN_KEYS], rad_sw=f["rad_sw"], rad_lw=f["rad_lw"], **kw)
    host2 = ab.aerobulk_model(2, 2, "ecmwf", 2.0, 10.0, *[f[k] for k in IN_KEYS], rad_sw=f["rad_sw"], rad_lw=f["rad_lw"], **kw)
    ab.reset()
    dev = {k: torch.from_numpy(np.ascontiguousarray(v.ravel(order="F"))).cuda() for k, v in f.items()}
    out = {k: torch.empty(Ni * Nj, dtype=torch.float64, device="cuda") for k in ("QL", "QH", "Tau_x", "Tau_y", "Evap", "T_s")}
    view = lambda d: {k: v.view(-1) for k, v in d.items()}
    ins = [dev[k] for k in IN_KEYS]
    ab.set_stream(torch.cuda.current_stream().cuda_stream)
    try:
        for jt, want in ((1, host), (2, host2)):
            # 1-D device arrays: Ni*Nj points as (m,1)
            ab.aerobulk_model_device(jt, 2, "ecmwf", 2.0, 10.0, *ins, out=out, rad_sw=dev["rad_sw"], rad_lw=dev["rad_lw"], **kw)
            ab.synchronize()
            for key in want:
                assert np.array_equal(out[key].cpu().numpy(), want[key].ravel(order="F")), (jt, key)
    finally:
        ab.set_stream(None)


@pytest.mark.parametrize("algo", ["andreas", "coare3p0"])
def test_config4_nb_iter_sweep(ab, oracle, algo):
    """BASELINE config 4 (ANDREAS and COARE 3.0, nb_iter sweep 5..30) on a 1/10-scale grid against the
    oracle: convergence with nb_iter and tolerance at every count (6 and 10 included: SURVEY quirk 1)."""
    f = synth.fields(432, 216)
    prev = None
    for nb in (5, 6, 8, 10, 15, 20, 30):
        got, _, _ = _single_call(ab, oracle, f"C4 {algo} 432x216 nb_iter={nb}", algo, 2.0, 10.0, f, nb)
        if prev is not None and nb >= 15:
            # the fixed-point iteration has converged for the bulk of the grid
            d = np.abs(got["QL"] - prev["QL"]) / (np.abs(prev["QL"]) + 10.0)
            assert np.median(d) < 1e-4, (algo, nb, float(np.median(d)))
        prev = got


@pytest.fixture(scope="module")
def fields_twelfth_degree():
    return synth.fields(4320, 2160)


@pytest.mark.parametrize("algo,nb_iter", [("andreas", 5), ("andreas", 30), ("coare3p0", 5), ("coare3p0", 30)])
def test_config4_full_size(ab, oracle, fields_twelfth_degree, algo, nb_iter):
    """BASELINE config 4 AT FULL SIZE: 1/12 deg grid 4320x2160 (9 331 200 points), both ends of the nb_iter sweep.
    Reference: src/mod_blk_andreas.f90:66-272, src/mod_blk_coare3p0.f90:54-358."""
    _single_call(ab, oracle, f"C4 {algo} 4320x2160 nb_iter={nb_iter} FULL", algo, 2.0, 10.0, fields_twelfth_degree, nb_iter)


def test_config3_full_size(ab, oracle, fields_twelfth_degree):
    """BASELINE config 3 AT FULL SIZE: ECMWF with the skin scheme on the 1/12 deg grid 4320x2160.
    Reference: src/mod_blk_ecmwf.f90:63-383, src/mod_skin_ecmwf.f90:68-230."""
    _single_call(ab, oracle, "C3 ecmwf skin 4320x2160 FULL", "ecmwf", 2.0, 10.0, fields_twelfth_degree, 5, skin=True)


C5_VARIANTS = [("ncar", False), ("andreas", False), ("coare3p0", False), ("coare3p6", False), ("ecmwf", False),
               ("coare3p6", True), ("ecmwf", True)]


@pytest.mark.parametrize("shard", range(8))
def test_config5_rows_of_every_shard(ab, oracle, shard):
    """BASELINE config 5 (1/36 deg grid 12960x6480, row blocks over 8 GPUs): 64 full-width rows out of EACH of the 8
    shards (8 x 64 = 512 rows, 6.6 M points) for all five algorithms and both skin variants against the oracle.
    The rows are rows of the global grid (synth is counter-based), computed as their own row block -- bit-identical
    to the same rows inside a whole-shard call, which test_config5_one_shard_of_eight asserts."""
    Ni, Nj, G = 12960, 6480, 8
    j0 = shard * Nj // G + 373
    f = synth.fields(Ni, Nj, j0=j0, j1=j0 + 64)
    for algo, skin in C5_VARIANTS:
        _single_call(ab, oracle, f"C5 shard {shard} rows {j0}..{j0 + 64} {algo}{' skin' if skin else ''}", algo, 2.0, 10.0,
                     f, 5, skin=skin)


@pytest.mark.parametrize("algo,skin", [("ncar", False), ("coare3p6", True)])
def test_config5_one_shard_of_eight(ab, oracle, algo, skin):
    """BASELINE config 5: the WHOLE shard of rank 3 (12960x810 = 10.5 M points) in one call;
    (a) a row sub-block of the shard computed alone is bit-identical (sharding invariance at full width);
    (b) those 64 rows against the oracle."""
    Ni, Nj, G, r = 12960, 6480, 8, 3
    j0, j1 = r * Nj // G, (r + 1) * Nj // G
    f = synth.fields(Ni, Nj, j0=j0, j1=j1)
    kw = dict(l_use_skin=True, rad_sw=f["rad_sw"], rad_lw=f["rad_lw"]) if skin else {}
    ab.reset()
    big = ab.aerobulk_model(1, 1, algo, 2.0, 10.0, *[f[k] for k in IN_KEYS], **kw)
    for k, v in big.items():
        assert np.isfinite(v).all(), k
    a, b = 400, 464
    sub = {k: np.asfortranarray(v[:, a:b]) for k, v in f.items()}
    small, _, _ = _single_call(ab, oracle, f"C5 whole shard 3 rows {a}..{b} {algo}", algo, 2.0, 10.0, sub, 5, skin=skin)
    for k in small:
        assert np.array_equal(small[k], big[k][:, a:b]), k
