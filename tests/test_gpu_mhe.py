"""Moving horizon estimator on the GPU through the C ABI: the reference's regression loop
(testing/test_rotating_oscillating_masses_mhe_mpc.py) fed with the measurements of the reference run; state and parameter
estimates, parameter vectors and multipliers against testing/results/results_rotatingMasses.pkl (estimator part)."""
import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_mhe_five_steps_match_the_reference_run():
    import rotating_masses
    g = np.load(GOLDEN / "kkt_rotatingMasses_mhe.npz")
    mhe = rotating_masses.masses_mhe()
    assert (mhe.n_opt_x, mhe.n_opt_lagr, mhe.n_opt_p) == (423, 450, 322)
    mhe.x0 = np.zeros(8)
    mhe.p_est0 = 1e-4
    mhe.set_initial_guess()
    ineq = ~np.isfinite(mhe.nlp_cons_lb)
    for t in range(5):
        x_est = mhe.make_step(g["y"][t].reshape(-1, 1))
        assert mhe.S.stats()["success"], (t, mhe.S.stats())
        # the parameter vector handed to the solver is the reference's (x_prev comes from our previous solution)
        assert np.abs(mhe.opt_p_num.full().ravel() - g["opt_p_num"][t]).max() < 1e-5
        x_ref = mhe._opt_x(g["opt_x_num"][t])
        assert np.abs(x_est - x_ref["_x", -1, -1].full()).max() < 1e-5, t
        assert abs(float(mhe.p_est0.full()[0, 0]) - float(x_ref["_p_est"].full()[0, 0])) < 1e-5 * 1e-4 + 1e-9, t
        lam = mhe.lam_g_num.full().ravel()
        assert np.abs(lam - g["lam_g_num"][t])[~ineq].max() < 1e-4 * max(1.0, np.abs(g["lam_g_num"][t]).max())
    d = mhe.data
    assert d["_x"].shape == (5, 8) and d["_y"].shape == (5, 5) and d["_opt_x_num"].shape == (5, 423) and d["_lam_g_num"].shape == (5, 450)
    assert np.abs(d["_x"][1:] - g["x"][1:]).max() < 1e-5            # stored estimates = the reference's stored estimates
    assert mhe.S.launch_info()["n_launches"] >= 5
