"""SURVEY §8c tier T1 (basin-controlled parity): started at the reference's own solution (golden opt_x_num, lam_g_num, opt_p_num
of testing/results/*.pkl) the solver must stay there.  Cold start of the multipliers (IPOPT default) from the golden primal point,
and — with ipopt.warm_start_init_point — from the full golden primal-dual point, where convergence has to be immediate."""
import numpy as np
import pytest

from conftest import HostSim
from examples_common import EXAMPLES, golden
from oracle import models
from oracle.nlp import FlatNLP

ORACLE_MODELS = {"oscillatingMasses": models.oscillating_masses, "oscillatingMasses_dae": models.oscillating_masses_dae, "batch_reactor": models.batch_reactor, "CSTR": models.cstr_robust,
                 "dip": models.dip}


@pytest.mark.parametrize("name", ["oscillatingMasses", "oscillatingMasses_dae", "batch_reactor", "CSTR", "dip"])
def test_stays_at_the_reference_solution(name):
    make, _ = EXAMPLES[name]
    mpc = make(); mpc.prepare_nlp(); r = mpc.recipe
    g = golden(name)
    sim = HostSim(mpc, name); sim.o.max_iter = 300
    steps = [0, 2, 4]
    X = g["opt_x_num"][steps] / r.opt_x_scaling            # goldens store the unscaled solution
    P = g["opt_p_num"][steps]
    cold = sim.solve(P, x_init=X)
    assert np.all(cold["status"] == 0)
    # robust CSTR: Q_dot is only weakly determined by the objective (flat direction); both solvers stop at tol 1e-8 on different
    # sides of it (same effect as in tests/test_gpu_closed_loop.py)
    assert np.abs(cold["u0_scaled"] - g["u"][steps] / r.u_scaling).max() < (1e-3 if name == "CSTR" else 1e-5)
    # full primal-dual warm start: lam_x from stationarity of the reference point (it is not stored in the pickles)
    nlp = FlatNLP(ORACLE_MODELS[name]())
    LX = np.array([-(nlp.grad_f(X[i], P[i]) + nlp.jac_g(X[i], P[i]).T @ g["lam_g_num"][steps][i]) for i in range(len(steps))])
    # (the usual MPC warm-start settings: tiny pushes so that the point and its multipliers are taken as they are)
    sim.o.warm_start = 1; sim.o.mu_init = 1e-8; sim.o.ws_bound_push = 1e-9; sim.o.ws_bound_frac = 1e-9; sim.o.ws_mult_bound_push = 1e-9
    warm = sim.solve(P, x_init=X, lam_g0=g["lam_g_num"][steps], lam_x0=LX)
    assert np.all(warm["status"] == 0), warm["status"]
    # 1..3 iterations when the point is accepted as is; up to ~10 when variables sitting on their (relaxed) bounds have to be
    # re-centred first (s*z << mu after the push) — far below the 15..70 iterations of the cold starts
    assert warm["iter_count"].max() <= 12 and np.median(warm["iter_count"]) <= 5, warm["iter_count"]
    du = np.abs(warm["u0_scaled"] - g["u"][steps] / r.u_scaling)
    if name == "CSTR":
        # flat direction (Q_dot): the input moves by up to 6e-3 of its scale while the objective agrees to 3e-6 (the barrier bias
        # mu * #bounds of two different final mu) — F is pinned
        assert du[:, 0].max() < 2e-5 and du[:, 1].max() < 1e-2
        assert np.abs(warm["f"] - cold["f"]).max() < 1e-5
        return
    assert du.max() < 1e-5
    used = nlp.used_mask()
    assert np.abs(warm["x"] - X)[:, used].max() < 1e-4 * max(1.0, np.abs(X[:, used]).max())
    assert np.abs(warm["lam_g"] - g["lam_g_num"][steps]).max() < 1e-4 * max(1.0, np.abs(g["lam_g_num"][steps]).max())
