"""Shared machinery of the GPU parity tests.

Metric (BASELINE.json north_star): per output field f
    err_f = |gpu - oracle| / (|oracle| + S_f),   S = 10 W/m2 (QL, QH), 1e-2 N/m2 (tau), 1e-5 kg/m2/s (Evap), 1 K (T_s)
Gate: err <= TOL = 1e-10 at every point, EXCEPT points that are proven to sit on a discontinuity (or an
ill-conditioned spot) of the reference algorithm itself:

  branch-flip proof -- for every point above TOL the ORACLE is re-run on that point with its inputs nudged by
  +-1 ulp (then +-4, +-16, +-64 ulp: the GPU arithmetic differs from glibc by a few ulp per transcendental and
  contracts FMAs, so an intermediate quantity can be off by some tens of ulp after a few iterations).  The point
  is accepted only if the oracle's OWN answer moves by at least PROOF_RATIO x the GPU's deviation under one of
  those nudges, i.e. the reference algorithm is discontinuous / ill-conditioned there at the rounding level
  (SIGN-selected stable/unstable psi, RiB < 0.15 switch of ANDREAS, LKB table edges, warm-layer thresholds,
  MAX/MIN clips: SURVEY.md 7 "hard parts").  A point whose oracle answer stays put is a REAL mismatch and fails.

The number of proven points must stay below OUTLIER_FRACTION of the grid.  Their error is not capped: the reference
has genuine jumps (NCAR's ChN switches 18 -> 32.7 with the sign of zeta, src/mod_blk_ncar.f90:210: two ADJACENT doubles
of t_zt give QH = 4.73 and 9.13 W/m2, tests/test_parity_util_cpu.py::test_gate_proves_a_real_discontinuity).
"""
from __future__ import annotations

import numpy as np

from aerobulk_b200 import synth

TOL = 1e-10
OUTLIER_FRACTION = 2e-5
UNPROVEN_MAX = 1e-9         # without the proof machinery (no oracle callback) nothing may exceed this
PROOF_RATIO = 0.1
NUDGES = (1, 4, 16, 64)

IN_KEYS = ("sst", "t_zt", "hum_zt", "U_zu", "V_zu", "slp")
RAD_KEYS = ("rad_sw", "rad_lw")


def worst_per_point(got: dict, ref: dict) -> np.ndarray:
    """max over the output fields of the scaled error, per point (flattened, Fortran order)."""
    errs = synth.parity_errors(got, ref)
    assert set(errs) == set(ref), (set(errs), set(ref))
    w = None
    for k, e in errs.items():
        assert not np.isnan(got[k]).any(), (k, "NaN in GPU output")
        e = np.ravel(e, order="F")
        w = e if w is None else np.maximum(w, e)
    return w


def _nudge(a: np.ndarray, ulps: int) -> np.ndarray:
    """a moved by `ulps` units in the last place (sign of ulps = direction); zeros stay zero (calm wind, night)."""
    out = a.copy()
    step = np.where(ulps > 0, np.inf, -np.inf)
    for _ in range(abs(ulps)):
        out = np.where(out != 0.0, np.nextafter(out, step), out)
    return out


def _variants(point_inputs: dict, ulps: int):
    """All single-field +-ulps nudges plus all-fields-together, for the selected points (1-D arrays of length m)."""
    keys = list(point_inputs)
    out = []
    for s in (+1, -1):
        for k in keys:
            if point_inputs[k] is None:
                continue
            v = dict(point_inputs)
            v[k] = _nudge(point_inputs[k], s * ulps)
            out.append(v)
        out.append({k: (None if a is None else _nudge(a, s * ulps)) for k, a in point_inputs.items()})
    return out


def prove_flips(idx: np.ndarray, err_gpu: np.ndarray, inputs: dict, run_oracle, tag: str = "") -> np.ndarray:
    """idx: flat (Fortran-order) indices of the points above TOL; err_gpu: their scaled GPU-oracle error;
    inputs: {name: flat array of the WHOLE field or None, or a list of such per time step for rad_sw};
    run_oracle(point_inputs) -> list of output dicts (one per time step) for 1-D inputs of the selected points.
    Returns a bool array: True where the oracle itself moves by >= PROOF_RATIO * err_gpu under some nudge."""
    sel = {}
    for k, a in inputs.items():
        if a is None:
            sel[k] = None
        elif isinstance(a, (list, tuple)):
            sel[k] = [np.ascontiguousarray(np.ravel(x, order="F")[idx]) for x in a]
        else:
            sel[k] = np.ascontiguousarray(np.ravel(a, order="F")[idx])
    base = run_oracle(sel)
    proven = np.zeros(idx.size, dtype=bool)
    spread_best = np.zeros(idx.size)
    used = np.zeros(idx.size, dtype=int)
    for ulps in NUDGES:
        # the time-dependent rad_sw list is nudged as a whole (same direction at every step)
        flat = {k: v for k, v in sel.items() if not isinstance(v, list)}
        lists = {k: v for k, v in sel.items() if isinstance(v, list)}
        for var in _variants(flat, ulps):
            for s in ((+1, -1) if lists else (0,)):
                full = dict(var)
                for k, v in lists.items():
                    full[k] = [_nudge(x, s * ulps) if s else x for x in v]
                outs = run_oracle(full)
                spread = np.zeros(idx.size)
                for o, b in zip(outs, base):
                    for k in b:
                        spread = np.maximum(spread, np.abs(o[k] - b[k]) / (np.abs(b[k]) + synth.PARITY_SCALE[k]))
                newly = (~proven) & (spread >= PROOF_RATIO * err_gpu)
                used[newly] = ulps
                proven |= newly
                spread_best = np.maximum(spread_best, spread)
        if proven.all():
            break
    for i in range(idx.size):
        print(f"[flip-proof] {tag} point {int(idx[i])}: gpu-oracle err {err_gpu[i]:.3e}, oracle moves by "
              f"{spread_best[i]:.3e} under <= {used[i] or NUDGES[-1]} ulp input nudges -> {'PROVEN' if proven[i] else 'NOT PROVEN'}")
    return proven


def assert_parity(tag: str, worst: np.ndarray, inputs: dict | None = None, run_oracle=None) -> dict:
    """worst: per-point worst scaled error (over fields, and over steps for a session)."""
    n = worst.size
    bad = np.flatnonzero(worst > TOL)
    s = np.sort(worst)
    rep = {"max": float(s[-1]), "second": float(s[-2]) if n > 1 else float(s[-1]), "above_tol": int(bad.size), "n": int(n),
           "proven": 0}
    print(f"[parity] {tag}: n={n} max {rep['max']:.3e} 2nd {rep['second']:.3e} points>1e-10: {bad.size}")
    if bad.size == 0:
        return rep
    assert bad.size <= max(1, int(OUTLIER_FRACTION * n)), (tag, "too many points above 1e-10", bad.size, n)
    if run_oracle is None or inputs is None:
        assert float(worst[bad].max()) <= UNPROVEN_MAX, (tag, "point above 1e-9 and no branch-flip proof available")
        return rep
    proven = prove_flips(bad, worst[bad], inputs, run_oracle, tag)
    rep["proven"] = int(proven.sum())
    assert proven.all(), (tag, "points above 1e-10 where the oracle itself is NOT sensitive to ulp-level input nudges",
                          bad[~proven].tolist(), worst[bad][~proven].tolist())
    return rep


def oracle_runner(OracleSession, algo, zt, zu, nb_iter, skin, nt=1, threads=4, hum_first=None):
    """run_oracle callback for prove_flips: a fresh oracle session over `nt` steps on 1-D point arrays.
    AEROBULK_INIT detects the humidity type from field statistics: a handful of selected points could be classified
    differently from the full field, so the selected points are preceded by nothing -- the synthetic fields are 'sh'
    (< 0.08) / 'rh' / 'dp' at every single point, which the detection reads the same way for any subset."""

    def run(pi):
        s = OracleSession(threads=threads)
        outs = []
        for jt in range(1, nt + 1):
            kw = dict(Niter=nb_iter)
            if skin:
                rsw = pi["rad_sw"][jt - 1] if isinstance(pi["rad_sw"], list) else pi["rad_sw"]
                kw.update(l_use_skin=True, rad_sw=rsw, rad_lw=pi["rad_lw"])
            outs.append(s.model(jt, nt, algo, zt, zu, *[pi[k] for k in IN_KEYS], **kw))
        return outs

    return run
