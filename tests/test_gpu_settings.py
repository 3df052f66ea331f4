"""Controller settings that change the structure handed to the kernel, on the GPU through the C ABI: the single slack of
`nl_cons_single_slack=True` (reference _mpc.py:1120-1123) is a lifted state with a free root — the same kernel mechanism the MHE
uses.  Compared with the oracle IPM on the oracle's flat restatement of the NLP (KKT conditions of the flat NLP at the returned
x, lam_g, lam_x; first input within 1e-5 * u_scaling)."""
import numpy as np
import pytest

from conftest import GOLDEN
from oracle.ipm import IPMSolver, IpmOptions, solve_mpc
from test_hostsim_settings import _build

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["soft_nl_cons", "soft_nl_cons_single_slack"])
def test_soft_constraint_variants_on_gpu(name):
    mpc, nlp = _build(name)
    mpc.create_nlp()
    r = mpc.recipe
    d = np.load(GOLDEN / "cstr_dataset.npz")
    idx = np.where(d["success"])[0][:16]
    out = mpc.make_step_batch(d["x0"][idx], d["u_prev"][idx], return_full=True)
    orc = IPMSolver(nlp, IpmOptions(max_iter=300))
    used = nlp.used_mask()
    eq = nlp.lbg == nlp.ubg
    for b, i in enumerate(idx):
        ro = solve_mpc(nlp, d["x0"][i], d["u_prev"][i], solver=orc)
        assert bool(out["success"][b]) == ro["success"], (name, i)
        if not ro["success"]:
            continue
        assert abs(int(out["iter_count"][b]) - ro["iter_count"]) <= (6 if r.n_lift else 1), (name, i)
        # lifted slack: the inertia regularisation delta_w acts on every copy, so the iterates are not the flat NLP's; both end at
        # KKT points to tol = 1e-8 and the weakly determined cooling input agrees to 2e-4 of its scale (flat NLP: 1e-7)
        assert (np.abs(out["u0"][b] - ro["u0"]) / r.u_scaling).max() < (2e-4 if r.n_lift else 1e-7), (name, i)
        x, lam_g, lam_x, p = out["x"][b], out["lam_g"][b], out["lam_x"][b], out["p"][b]
        g = nlp.g(x, p)
        assert np.abs(g[eq]).max() < 1e-8 and np.all(g[~eq] <= nlp.ubg[~eq] + 2e-8 * np.maximum(1.0, np.abs(nlp.ubg[~eq])))   # bound_relax_factor 1e-8 (relative)
        stat = nlp.grad_f(x, p) + nlp.jac_g(x, p).T @ lam_g + lam_x
        assert np.abs(stat[used]).max() < 1e-7, (name, i)
        assert abs(out["f"][b] - ro["f"]) < 1e-7 * max(1.0, abs(ro["f"]))
    # the single-controller call of the reference API goes through the same library (flat x0 in, lifted copies initialised on the device)
    i0 = idx[0]
    mpc.x0 = d["x0"][i0]; mpc.u0 = d["u_prev"][i0]; mpc.set_initial_guess()
    u = mpc.make_step(d["x0"][i0].reshape(-1, 1))
    assert np.abs(u.ravel() - out["u0"][0]).max() < 1e-6 * r.u_scaling.max()
    eps = mpc.opt_x_num["_eps", 0, 0].full().ravel()
    assert eps.shape == (1,) and eps[0] >= -1e-8


def test_dae_under_a_scenario_tree_on_gpu():
    import oscillating_masses as om
    from oracle import models
    from oracle.nlp import FlatNLP
    mpc = om.osc_dae_robust_mpc(n_horizon=5, n_robust=1, setup=True)
    nlp = FlatNLP(models.oscillating_masses_dae(5, n_robust=1, uncertain=True))
    r = mpc.recipe
    rng = np.random.default_rng(5)
    X0 = rng.uniform(-1.0, 1.0, (32, 4)) * np.array([1.5, 3.0, 1.5, 3.0]); U0 = rng.uniform(-0.3, 0.3, (32, 1))
    out = mpc.make_step_batch(X0, U0, Z0=np.zeros((32, 4)), return_full=True)
    orc = IPMSolver(nlp, IpmOptions(max_iter=300))
    for b in range(0, 32, 4):
        ro = solve_mpc(nlp, X0[b], U0[b], z0=np.zeros(4), solver=orc)
        assert bool(out["success"][b]) == ro["success"]
        if ro["success"]:
            assert abs(int(out["iter_count"][b]) - ro["iter_count"]) <= 1
            assert np.abs(out["u0"][b] - ro["u0"]).max() < 1e-7
            assert np.abs(out["x"][b] - ro["x"])[nlp.used_mask()].max() < 1e-6
