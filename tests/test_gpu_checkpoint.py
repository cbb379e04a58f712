"""SURVEY.md 8f row 3 (first half): checkpoint / restart of the warm-layer state.  The reference keeps it in module
arrays that cannot be saved (src/mod_skin_coare.f90:31-36); here aerobulk_gpu_get_state / set_state make a 24-step
session restartable: run to step 12, save, "restart the process" (aerobulk_gpu_reset), re-open the session with
jt = 1, restore, continue at step 13 -- bit-identical to the uninterrupted session."""
import numpy as np
import pytest

from aerobulk_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("algo", ["coare3p6", "coare3p0", "ecmwf"])
def test_restart_from_saved_state_is_bit_identical(algo):
    import aerobulk_b200 as ab
    Ni, Nj, Nt, cut = 256, 120, 24, 12
    n = Ni * Nj
    f = synth.fields(Ni, Nj)
    keys = ("sst", "t_zt", "hum_zt", "U_zu", "V_zu", "slp")
    rsw = [synth.rad_sw_hour(Ni, Nj, jt) for jt in range(1, Nt + 1)]

    def step(jt):
        return ab.aerobulk_model(jt, Nt, algo, 2.0, 10.0, *[f[k] for k in keys], Niter=6, l_use_skin=True,
                                 rad_sw=rsw[jt - 1], rad_lw=f["rad_lw"])

    ab.reset()
    full = [step(jt) for jt in range(1, Nt + 1)]

    ab.reset()
    for jt in range(1, cut + 1):
        step(jt)
    saved = {w: ab.get_state(w, n) for w in range(4)}
    assert saved[0] is not None and np.any(saved[0] != 0.0)          # the warm layer is active at noon
    if algo == "ecmwf":
        assert saved[2] is None and saved[3] is None                 # ECMWF keeps dT_wl only ...
        assert np.all(saved[1] == 3.0)                               # ... Hz_wl is the constant 3 m

    ab.reset()                                                       # process restart
    step(1)                                                          # re-opens the session (sticky globals, allocation)
    for w, v in saved.items():
        if v is not None:
            assert ab.set_state(w, v)
    for jt in range(cut + 1, Nt + 1):
        r = step(jt)
        for k in ("QL", "QH", "Tau_x", "Tau_y", "Evap", "T_s"):
            assert np.array_equal(r[k], full[jt - 1][k]), (algo, jt, k)
    assert ab.get_state(0, n) is None                                # released at jt == Nt as usual
    # a state of the wrong size is refused
    ab.reset()
    step(1)
    assert not ab.set_state(0, np.zeros(n - 1))
    ab.reset()
