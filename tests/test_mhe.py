"""Moving horizon estimator: the reference's regression example (rotating oscillating masses,
testing/test_rotating_oscillating_masses_mhe_mpc.py) with the reference's own results
(testing/results/results_rotatingMasses.pkl, estimator part -> tests/golden/kkt_rotatingMasses_mhe.npz).
  * the oracle's MHE NLP (oracle/mhe.py) has the golden (opt_p -> opt_x*, lam_g*) as KKT points   [pins the oracle]
  * oracle/ipm.py from the reference's initial guesses reproduces the golden solutions
  * the kernel's algorithm source (CPU emulation) on this package's MHE recipe agrees with both."""
import numpy as np
import pytest

from conftest import GOLDEN, HostSim

import rotating_masses
from oracle import models
from oracle.ipm import IPMSolver, IpmOptions
from oracle.mhe import FlatMHE


@pytest.fixture(scope="module")
def gold():
    d = np.load(GOLDEN / "kkt_rotatingMasses_mhe.npz")
    return {k: d[k] for k in d.files}


@pytest.fixture(scope="module")
def nlp():
    return FlatMHE(models.rotating_masses_mhe())


def start_point(nlp, gold, t):
    """initial guess of step t exactly as the reference run had it: set_initial_guess() broadcast (x0 = 0, p_est0 = 1e-4) for
    the first call, the previous solution afterwards (optimizer.py:741-742)."""
    if t == 0:
        return nlp.initial_guess(np.zeros(8), np.zeros(2), [1e-4])
    return gold["opt_x_num"][t - 1] / nlp.scaling


def test_sizes_and_golden_kkt_points(nlp, gold):
    assert (nlp.n_w, nlp.n_g, nlp.n_p) == (423, 450, 322) == (gold["opt_x_num"].shape[1], gold["lam_g_num"].shape[1], gold["opt_p_num"].shape[1])
    used = nlp.used_mask()
    for t in range(5):
        x = gold["opt_x_num"][t] / nlp.scaling
        p, lam = gold["opt_p_num"][t], gold["lam_g_num"][t]
        g = nlp.g(x, p)
        assert np.maximum(np.maximum(nlp.lbg - g, g - nlp.ubg), 0).max() < 1e-8
        r = nlp.grad_f(x, p) + nlp.jac_g(x, p).T @ lam
        dl, du = x - nlp.lbx, nlp.ubx - x
        assert (np.abs(r) * np.minimum(np.minimum(dl, du), 1.0))[used].max() < 1e-7
        assert lam[np.isinf(nlp.lbg)].min() > -1e-9


def test_oracle_ipm_reproduces_the_reference_estimates(nlp, gold):
    sol = IPMSolver(nlp, IpmOptions(max_iter=300))
    for t in range(5):
        r = sol.solve(start_point(nlp, gold, t), gold["opt_p_num"][t])
        assert r["success"], (t, r["return_status"])
        x_ref = gold["opt_x_num"][t]
        assert np.abs(r["x"] * nlp.scaling - x_ref)[nlp.used_mask()].max() < 1e-6, t


def test_kernel_source_on_the_mhe_recipe(nlp, gold):
    mhe = rotating_masses.masses_mhe(setup=False)
    mhe.prepare_nlp()
    r = mhe.recipe
    assert (r.n_w, r.n_g, mhe.n_opt_p) == (423, 450, 322)
    sim = HostSim(mhe, "mhe_masses")
    sim.o.max_iter = 300
    orc = IPMSolver(nlp, IpmOptions(max_iter=300))
    used = nlp.used_mask()
    for t in range(5):
        x_init = start_point(nlp, gold, t)
        out = sim.solve(mhe.kernel_params(gold["opt_p_num"][t]), x_init=x_init[None, :])
        assert out["status"][0] == 0, (t, out["status"])
        assert np.abs(out["x"][0] * nlp.scaling - gold["opt_x_num"][t])[used].max() < 1e-5, t       # stated tolerance: 1e-5 x scaling
        ro = orc.solve(x_init, gold["opt_p_num"][t])
        # the lifted formulation carries multipliers for the parameter-consistency rows: the least-squares multiplier estimate
        # and the dual scaling s_d differ slightly from the flat NLP's, so iteration counts agree only approximately
        print("MHE step", t, "iterations kernel", int(out["iter_count"][0]), "oracle", ro["iter_count"])
        assert abs(int(out["iter_count"][0]) - ro["iter_count"]) <= 3, (t, out["iter_count"], ro["iter_count"])
        # multipliers in the reference's g layout (lifted-parameter rows have no flat counterpart)
        lam = mhe.reference_multipliers(out["x"][0], out["lam_g"][0], gold["opt_p_num"][t])
        ineq = np.isinf(nlp.lbg)              # (the duplicated inequality rows share their multiplier arbitrarily: compare the equalities)
        assert np.abs(lam - ro["lam_g"])[~ineq].max() < 1e-5 * max(1.0, np.abs(ro["lam_g"]).max())
        assert np.abs(lam - gold["lam_g_num"][t])[~ineq].max() < 1e-5 * max(1.0, np.abs(gold["lam_g_num"][t]).max())
