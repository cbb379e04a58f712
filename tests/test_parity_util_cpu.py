"""CPU checks of the parity gate itself (tests/parity_util.py): the branch-flip proof accepts what ulp-level input
noise does to the oracle and rejects a real mismatch.  No GPU: the "other implementation" here is the oracle fed
with inputs moved by a few ulp, which is the size of perturbation the GPU arithmetic introduces."""
import numpy as np
import pytest

import parity_util as pu
from aerobulk_b200 import synth
from oracle.oracle import OracleSession


def _run(algo, f, nb_iter=5, skin=False):
    kw = dict(Niter=nb_iter)
    if skin:
        kw.update(l_use_skin=True, rad_sw=f["rad_sw"], rad_lw=f["rad_lw"])
    return OracleSession(threads=8).model(1, 1, algo, 2.0, 10.0, *[f[k] for k in pu.IN_KEYS], **kw)


@pytest.mark.parametrize("algo,skin", [("andreas", False), ("ecmwf", True), ("coare3p6", True)])
def test_gate_accepts_ulp_noise(algo, skin):
    f = synth.fields(240, 120)
    ref = _run(algo, f, skin=skin)
    g = {k: pu._nudge(np.ravel(v, order="F"), 3).reshape(v.shape, order="F") for k, v in f.items()}
    got = _run(algo, g, skin=skin)
    inputs = {k: f[k] for k in pu.IN_KEYS}
    inputs["rad_sw"] = f["rad_sw"] if skin else None
    inputs["rad_lw"] = f["rad_lw"] if skin else None
    rep = pu.assert_parity(f"cpu {algo}", pu.worst_per_point(got, ref), inputs,
                           pu.oracle_runner(OracleSession, algo, 2.0, 10.0, 5, skin))
    assert rep["above_tol"] == rep["proven"]


def test_gate_rejects_a_real_mismatch():
    f = synth.fields(96, 48)
    ref = _run("coare3p6", f)
    got = {k: v.copy() for k, v in ref.items()}
    got["QL"][17, 5] *= 1.0 + 3e-9          # a genuine 3e-9 error at a well-conditioned point
    inputs = {k: f[k] for k in pu.IN_KEYS}
    inputs["rad_sw"] = inputs["rad_lw"] = None
    with pytest.raises(AssertionError):
        pu.assert_parity("cpu mismatch", pu.worst_per_point(got, ref), inputs,
                         pu.oracle_runner(OracleSession, "coare3p6", 2.0, 10.0, 5, False))


def test_gate_proves_a_real_discontinuity():
    """NCAR: ChN = 18 (stable) or 32.7 (unstable) with the sign of zeta (src/mod_blk_ncar.f90:210).  These two adjacent
    doubles of t_zt (found by bisection) straddle the switch: QH jumps from 4.73 to 9.13 W/m2.  An implementation that
    rounds differently lands on the other side; the gate must recognise the point as a discontinuity of the reference."""
    t_lo, t_hi = 296.49341479971497, 296.493414799715
    assert np.nextafter(t_lo, np.inf) == t_hi
    one = lambda v: np.array([v])
    f = dict(sst=one(295.15), hum_zt=one(0.012), U_zu=one(5.0), V_zu=one(0.0), slp=one(101000.0))
    run = lambda t: OracleSession(threads=1).model(1, 1, "ncar", 10.0, 10.0, f["sst"], one(t), f["hum_zt"], f["U_zu"],
                                                   f["V_zu"], f["slp"], Niter=5)
    ref, got = run(t_lo), run(t_hi)
    assert abs(ref["QH"][0] - 4.7298178830982485) < 1e-9 and abs(got["QH"][0] - 9.132098972060398) < 1e-9
    inputs = dict(f, t_zt=one(t_lo), rad_sw=None, rad_lw=None)
    rep = pu.assert_parity("cpu ncar jump", pu.worst_per_point(got, ref), inputs,
                           pu.oracle_runner(OracleSession, "ncar", 10.0, 10.0, 5, False))
    assert rep["above_tol"] == 1 and rep["proven"] == 1 and rep["max"] > 0.1


def test_gate_rejects_many_outliers():
    f = synth.fields(96, 48)
    ref = _run("ncar", f)
    got = {k: v.copy() for k, v in ref.items()}
    got["QH"][::7, ::5] += 1e-6
    with pytest.raises(AssertionError):
        pu.assert_parity("cpu many", pu.worst_per_point(got, ref))


def test_nudge_moves_by_ulps():
    a = np.array([1.0, -2.5, 0.0, 1e-300])
    up = pu._nudge(a, 4)
    assert up[2] == 0.0
    assert up[0] == 1.0 + 4 * np.spacing(1.0)
    assert up[1] == -2.5 + 4 * np.spacing(2.5)
    assert np.array_equal(pu._nudge(up, -4), a)
