"""All model families on the GPU through the C-ABI against the reference's IPOPT results (golden step 0 = cold
start as testing/test_*.py runs it) and against the CPU emulation of the same algorithm."""
import numpy as np
import pytest

from examples_common import EXAMPLES, golden, cold_start_inputs

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("name", ["oscillatingMasses", "oscillatingMasses_dae", "batch_reactor", "CSTR", "dip"])
def test_cold_start_matches_reference_ipopt_on_gpu(name):
    make, sizes = EXAMPLES[name]
    mpc = make(setup=True)
    r = mpc.recipe
    g = golden(name)
    mpc.S.set_options(max_iter=500)
    P, x0s, u0s, z0s = cold_start_inputs(r, g)
    # a batch: the golden instance plus perturbed copies (exercises many CTAs / groups on the device)
    B = 48
    rng = np.random.RandomState(0)
    Pb = np.repeat(P, B, axis=0); xb = np.repeat(x0s, B, axis=0) * (1 + 0.01 * rng.randn(B, r.n_x)); xb[0] = x0s[0]
    Pb[:, r.p_x0:r.p_x0 + r.n_x] = xb * r.x_scaling
    out = mpc.S.solve_batch(Pb, xb, np.repeat(u0s, B, axis=0), None if z0s is None else np.repeat(z0s, B, axis=0),
                            lbx=mpc._lb_opt_x.full().ravel(), ubx=mpc._ub_opt_x.full().ravel(), return_full=True)
    assert out["status"][0] == 0
    assert np.abs(out["u0_scaled"][0] - g["u"][0] / r.u_scaling).max() < TOL
    eq = np.isfinite(mpc.nlp_cons_lb)
    ok = out["success"]
    assert ok.mean() > 0.8
    assert np.abs(out["g"][ok][:, eq]).max() < 1e-7
    # same algorithm source on the CPU: identical iteration counts and solutions up to round-off
    from conftest import HostSim
    sim = HostSim(mpc, name)
    sim.o.max_iter = 500
    k = 4
    ref = sim.solve(Pb[:k], x0s=xb[:k], u0s=np.repeat(u0s, k, axis=0), z0s=None if z0s is None else np.repeat(z0s, k, axis=0))
    assert np.array_equal(ref["success"], out["success"][:k])
    same = (np.abs(ref["iter_count"] - out["iter_count"][:k]) <= 1) & (ref["status"] == out["status"][:k])
    # Path mismatches are counted, not folded into a tolerance (tier T2).  The pendulum swing-up (N = 100) takes
    # 125 - 150 iterations on a non-convex landscape and is chaotic with respect to rounding: on the golden instance the oracle
    # needs 134 iterations, the CPU emulation 125, the GPU (FMA contraction) 137 — all three end at the reference's IPOPT
    # solution.  For that family the requirement is therefore on the solution (same success flags — asserted above —, same
    # basin: u0 within the north-star tolerance), with basin mismatches of perturbed instances counted and bounded; every
    # other family must agree on every instance, iteration for iteration.
    n_path = int((~same).sum())
    du = np.abs(ref["u0_scaled"] - out["u0_scaled"][:k]).max(axis=1)
    print(f"{name}: {n_path} of {k} instances took a different path on the GPU than on the CPU emulation; max |u0 difference| (scaled) {du.max():.2e}")
    if name == "dip":
        n_basin = int((du[ref["success"]] > TOL).sum())
        assert du[0] < TOL and n_basin <= k // 4, (n_basin, du, ref["iter_count"], out["iter_count"][:k])
    else:
        assert n_path == 0, (name, ref["iter_count"], out["iter_count"][:k], ref["status"], out["status"][:k])
        assert du.max() < 1e-7


def test_make_step_loop_like_reference_regression_test():
    """5 controller steps as in testing/test_batch_reactor.py (states fed back from the golden closed loop because the
    plant simulator is out of scope): u, stored solution sizes, aux, warm start."""
    name = "batch_reactor"
    make, _ = EXAMPLES[name]
    mpc = make(setup=True)
    g = golden(name)
    mpc.x0 = g["x"][0]
    mpc.set_initial_guess()
    for t in range(5):
        u0 = mpc.make_step(g["x"][t].reshape(-1, 1))
        assert mpc.S.stats()["success"]
        assert np.abs(u0.ravel() - g["u"][t]).max() / mpc.recipe.u_scaling.max() < TOL
    assert mpc.data["_x"].shape == (5, 4) and mpc.data["_opt_x_num"].shape == (5, 608) and mpc.data["_lam_g_num"].shape == (5, 564)
    assert mpc.data["_aux"].shape == (5, 1) and mpc.data["success"].shape == (5, 1)
    pred = mpc.data.prediction(("_x", "P_s"))
    assert pred.shape == (1, 21, 1)


def test_industrial_poly_matches_reference_on_gpu():
    make, _ = EXAMPLES["industrial_poly"]
    mpc = make(setup=True)
    r = mpc.recipe
    g = golden("industrial_poly")
    mpc.S.set_options(max_iter=300)
    P, x0s, u0s, z0s = cold_start_inputs(r, g)
    out = mpc.S.solve_batch(P, x0s, u0s, z0s, lbx=mpc._lb_opt_x.full().ravel(), ubx=mpc._ub_opt_x.full().ravel(), return_full=True)
    # 0 = Solve_Succeeded; 1 = Solved_To_Acceptable_Level (E0 <= 1e-6 when the line search reaches the round-off floor of
    # this badly scaled problem; CasADi reports success for both, like the reference's stats()['success'])
    assert out["status"][0] in (0, 1) and out["success"][0]
    assert np.abs(out["u0_scaled"][0] - g["u"][0] / r.u_scaling).max() < TOL
    assert np.abs(out["g"][0]).max() < 1e-7


@pytest.mark.parametrize("name", ["deg3_ni2", "legendre", "horizon1", "hard_nl_cons_colloc", "soft_nl_cons", "terminal_bounds"])
def test_setting_variants_on_the_gpu_match_the_oracle(name):
    """The settings sweep of tests/test_hostsim_settings.py through the real kernel + C ABI (each variant is its own generated library)."""
    from test_hostsim_settings import _build
    from conftest import GOLDEN, cstr_batch_inputs
    from oracle.ipm import IPMSolver, IpmOptions, solve_mpc
    mpc, nlp = _build(name)
    mpc.create_nlp()
    r = mpc.recipe
    d = np.load(GOLDEN / "cstr_dataset.npz")
    idx = np.where(d["success"])[0][[0, 3, 7, 11]]
    out = mpc.make_step_batch(d["x0"][idx], d["u_prev"][idx])
    orc = IPMSolver(nlp, IpmOptions(max_iter=300))
    for b, i in enumerate(idx):
        ro = solve_mpc(nlp, d["x0"][i], d["u_prev"][i], solver=orc)
        assert bool(out["success"][b]) == ro["success"], (name, i)
        if ro["success"]:
            assert abs(int(out["iter_count"][b]) - ro["iter_count"]) <= 1
            assert np.abs((out["u0"][b] - ro["u0"]) / r.u_scaling).max() < 1e-7
