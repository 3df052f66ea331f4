"""Controller settings the reference exposes (do_mpc/controller/_mpcsettings.py): every discretisation / constraint option that
changes the NLP is run through the kernel's algorithm source (CPU emulation) and compared with the oracle IPM on the
oracle's independent restatement of the same NLP — sizes, return status, iteration count, first input."""
import dataclasses

import numpy as np
import pytest

from conftest import GOLDEN, HostSim, cstr_batch_inputs
from oracle import models
from oracle.ipm import IPMSolver, IpmOptions, solve_mpc
from oracle.nlp import FlatNLP

N = 6
VARIANTS = {
    "deg3": dict(collocation_deg=3),
    "deg1": dict(collocation_deg=1),
    "ni2": dict(collocation_ni=2),
    "deg3_ni2": dict(collocation_deg=3, collocation_ni=2),
    "legendre": dict(collocation_type="legendre"),
    "horizon1": dict(n_horizon=1),
    "no_colloc_bounds": dict(cons_check_colloc_points=False),
    "terminal_bounds": dict(use_terminal_bounds=True),
    "hard_nl_cons_colloc": dict(nl_cons_check_colloc_points=True),
    "soft_nl_cons": dict(),
    "soft_nl_cons_single_slack": dict(nl_cons_single_slack=True),     # one slack for the horizon: lifted to a state (recipe.py)
    "open_loop_single_scenario": dict(open_loop=True),        # no effect without a scenario tree (reference _mpc.py:1113-1117)
}


def _build(name):
    import cstr
    kw = dict(VARIANTS[name])
    n_hor = kw.pop("n_horizon", N)
    mpc = cstr.cstr_nominal_mpc(n_horizon=n_hor, setup=False)
    for k, v in kw.items():
        setattr(mpc.settings, k, v)
    if name == "terminal_bounds":
        mpc.terminal_bounds["upper", "C_b"] = 1.2          # reference form (_mpc.py:424-431)
    if name == "hard_nl_cons_colloc":
        mpc.set_nl_cons("T_K_hard", mpc.model.x["T_K"] - 0.5 * mpc.model.x["T_R"], ub=70.0)
    if name.startswith("soft_nl_cons"):
        mpc.set_nl_cons("T_R_soft", mpc.model.x["T_R"], ub=132.0, soft_constraint=True, penalty_term_cons=50.0)
    mpc.prepare_nlp()
    d = dataclasses.replace(models.cstr_nominal(n_hor), **kw)
    if name == "hard_nl_cons_colloc":
        d = dataclasses.replace(d, nl_cons=[models.NlCons("T_K_hard", d.model.x[3] - 0.5 * d.model.x[2], ub=70.0)])
    if name.startswith("soft_nl_cons"):
        d = dataclasses.replace(d, nl_cons=[models.NlCons("T_R_soft", d.model.x[2], ub=132.0, soft=True, penalty=50.0)])
    if name == "terminal_bounds":
        # only C_b gets a terminal upper bound: the other entries stay +inf (the default "terminal = state bounds" applies
        # only while NO terminal bound is set, reference _mpc.py:914-918), the untouched lower bounds default to x_lb
        d = dataclasses.replace(d, x_terminal_ub=np.array([np.inf, 1.2, np.inf, np.inf]), x_terminal_lb=None)
    return mpc, FlatNLP(d)


@pytest.mark.parametrize("name", list(VARIANTS))
def test_setting_variant_matches_oracle(name):
    mpc, nlp = _build(name)
    r = mpc.recipe
    assert (r.n_w, r.n_g, r.n_p) == (nlp.n_w, nlp.n_g, nlp.n_p)
    lbx, ubx = mpc._lb_opt_x.full().ravel(), mpc._ub_opt_x.full().ravel()
    assert np.array_equal(np.nan_to_num(lbx, neginf=-1e30), np.nan_to_num(nlp.lbx, neginf=-1e30))
    assert np.array_equal(np.nan_to_num(ubx, posinf=1e30), np.nan_to_num(nlp.ubx, posinf=1e30))
    d = np.load(GOLDEN / "cstr_dataset.npz")
    idx = np.where(d["success"])[0][[0, 3, 7]]
    P, x0s, u0s = cstr_batch_inputs(r, d["x0"][idx], d["u_prev"][idx])
    sim = HostSim(mpc, "cstr_" + name)
    out = sim.solve(P, x0s, u0s)
    orc = IPMSolver(nlp, IpmOptions(max_iter=300))
    for b, i in enumerate(idx):
        ro = solve_mpc(nlp, d["x0"][i], d["u_prev"][i], solver=orc)
        assert bool(out["success"][b]) == ro["success"], (name, i, out["status"][b], ro["return_status"])
        if ro["success"]:
            # the lifted slack carries its own consistency multipliers: same solution, the path of the iterates differs slightly
            assert abs(int(out["iter_count"][b]) - ro["iter_count"]) <= (6 if r.n_lift else 1), (name, i)
            assert np.abs(out["u0_scaled"][b] - ro["u0"] / r.u_scaling).max() < (1e-5 if r.n_lift else 1e-7), (name, i)
            used = nlp.used_mask()
            if not r.n_lift:
                assert np.abs(out["x"][b] - ro["x"])[used].max() < 1e-6
            else:
                # both are converged to tol = 1e-8 but along different iterates: compare through the KKT conditions of the flat
                # (reference-layout) NLP at the returned x, lam_g, lam_x; weakly determined inputs may differ by more than u0 does
                x, lam_g, lam_x = out["x"][b], out["lam_g"][b], out["lam_x"][b]
                g = nlp.g(x, P[b])
                eq = nlp.lbg == nlp.ubg
                assert np.abs(g[eq]).max() < 1e-8 and np.all(g[~eq] <= nlp.ubg[~eq] + 2e-8 * np.maximum(1.0, np.abs(nlp.ubg[~eq])))   # bound_relax_factor 1e-8 (relative)
                stat = nlp.grad_f(x, P[b]) + nlp.jac_g(x, P[b]).T @ lam_g + lam_x
                assert np.abs(stat[used]).max() < 1e-7, (name, i)
                assert np.all(lam_g[~eq] >= -1e-9)
                assert abs(out["f"][b] - ro["f"]) < 1e-7 * max(1.0, abs(ro["f"]))
                assert np.abs(x - ro["x"])[used].max() < 1e-3


def test_dae_under_a_scenario_tree_matches_oracle():
    """Algebraic states together with n_robust > 0: supported where lterm / nl_cons / aux do not depend on z (the reference's
    `_z[k, s]` indexing quirk, SURVEY App. F.5, has no effect then).  Discrete DAE with an uncertain input gain, 3 scenarios."""
    import oscillating_masses as om
    mpc = om.osc_dae_robust_mpc(n_horizon=5, n_robust=1, setup=False)
    mpc.prepare_nlp()
    nlp = FlatNLP(models.oscillating_masses_dae(5, n_robust=1, uncertain=True))
    r = mpc.recipe
    assert (r.n_w, r.n_g, r.n_p) == (nlp.n_w, nlp.n_g, nlp.n_p)
    rng = np.random.default_rng(5)
    X0 = rng.uniform(-1.0, 1.0, (3, 4)) * np.array([1.5, 3.0, 1.5, 3.0]); U0 = rng.uniform(-0.3, 0.3, (3, 1))
    P, x0s, u0s, z0s = mpc.pack_batch(X0, U0, Z0=np.zeros((3, 4)))
    sim = HostSim(mpc, "osc_dae_robust")
    out = sim.solve(P, x0s, u0s, z0s)
    orc = IPMSolver(nlp, IpmOptions(max_iter=300))
    for b in range(3):
        ro = solve_mpc(nlp, X0[b], U0[b], z0=np.zeros(4), solver=orc)
        assert np.array_equal(nlp.pack_p(X0[b], U0[b]), P[b])
        assert bool(out["success"][b]) == ro["success"] == True
        assert abs(int(out["iter_count"][b]) - ro["iter_count"]) <= 1
        assert np.abs(out["u0_scaled"][b] - ro["u0"] / r.u_scaling).max() < 1e-7
        assert np.abs(out["x"][b] - ro["x"])[nlp.used_mask()].max() < 1e-6


def test_dae_tree_with_z_dependent_stage_cost_is_refused():
    import oscillating_masses as om
    mpc = om.osc_dae_robust_mpc(n_horizon=3, n_robust=1, setup=False)
    mpc.set_objective(mterm=mpc.model.aux["cost"], lterm=mpc.model.aux["cost"] + mpc.model.z["x_next"][0] ** 2)
    with pytest.raises(NotImplementedError):
        mpc.prepare_nlp()


def test_parameter_scaling_reaches_the_model_equations_only():
    """scaling['_p', name] multiplies the parameter inside rhs / alg (reference optimizer.py:811-815): a controller with
    scaling 2 and uncertainty values v / 2 solves the same NLP as the unscaled controller with values v."""
    import cstr
    ref = cstr.cstr_robust_mpc(n_horizon=4, n_robust=1, setup=False); ref.prepare_nlp()
    mpc = cstr.cstr_robust_mpc(n_horizon=4, n_robust=1, setup=False)
    mpc.scaling["_p", "alpha"] = 2.0; mpc.scaling["_p", "beta"] = 0.5
    mpc.set_uncertainty_values(alpha=np.array([1.0, 1.05, 0.95]) / 2.0, beta=np.array([1.0, 1.1, 0.9]) / 0.5)
    mpc.prepare_nlp()
    d = np.load(GOLDEN / "cstr_dataset.npz")
    idx = np.where(d["success"])[0][[0, 3]]
    outs = []
    for m_, tag in ((ref, "cstr_rob4"), (mpc, "cstr_rob4_pscaled")):
        P, x0s, u0s, _ = m_.pack_batch(d["x0"][idx], d["u_prev"][idx])
        outs.append(HostSim(m_, tag).solve(P, x0s, u0s))
    assert np.array_equal(outs[0]["status"], outs[1]["status"]) and (outs[0]["status"] == 0).all()
    assert np.array_equal(outs[0]["iter_count"], outs[1]["iter_count"])
    assert np.abs(outs[0]["x"] - outs[1]["x"]).max() < 1e-9
