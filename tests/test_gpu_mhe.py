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


def test_mpc_simulator_mhe_loop_like_the_reference_regression_test():
    """testing/test_rotating_oscillating_masses_mhe_mpc.py end to end on the GPU: MPC -> simulator -> MHE, 5 steps, the stored
    trajectories of all three modules against the reference's results_rotatingMasses.pkl (reference tolerance 1e-8 on its own
    re-run; here the stated solver tolerance 1e-5 x scaling, simulator 1e-6)."""
    import rotating_masses
    g = np.load(GOLDEN / "loop_rotatingMasses.npz")
    model = rotating_masses.masses_model()
    mpc = rotating_masses.masses_mpc(model)
    simulator = rotating_masses.masses_simulator(model)
    mhe = rotating_masses.masses_mhe(model)
    np.random.seed(99)
    x0_true = np.random.rand(model.n_x) - 0.5
    x0 = np.zeros(model.n_x)
    mpc.x0 = x0; simulator.x0 = x0_true; mhe.x0 = x0; mhe.p_est0 = 1e-4
    mpc.set_initial_guess(); mhe.set_initial_guess()
    for k in range(5):
        u0 = mpc.make_step(np.asarray(x0).reshape(-1, 1))
        y_next = simulator.make_step(u0)
        x0 = mhe.make_step(y_next)
    assert np.abs(mpc.data["_u"] - g["mpc_u"]).max() < 1e-4 and np.abs(mpc.data["_x"] - g["mpc_x"]).max() < 1e-4
    assert np.abs(simulator.data["_x"] - g["sim_x"]).max() < 1e-4 and np.abs(simulator.data["_u"] - g["sim_u"]).max() < 1e-4
    assert np.abs(mhe.data["_x"] - g["mhe_x"]).max() < 1e-4 and np.abs(mhe.data["_y"] - g["mhe_y"]).max() < 1e-4
