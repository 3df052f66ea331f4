"""Parity of the CUDA path (through the C-ABI / ctypes, no CPU fallback) with the reference's IPOPT results and
the oracle.  Tolerances (BASELINE.json north_star / SURVEY.md §8c): own scaled KKT residuals <= 1e-8,
|u0 - u0_ref| <= 1e-5 * u_scaling on instances where the reference succeeded; success flags must agree."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mpc():
    import cstr
    return cstr.cstr_nominal_mpc()


def test_dataset_parity_through_cabi(mpc, dataset):
    out = mpc.make_step_batch(dataset["x0"], dataset["u_prev"])
    ok_ref = dataset["success"].astype(bool)
    assert np.array_equal(out["success"], ok_ref), "success split differs from the reference's 692/308"
    dev = np.abs(out["u0"][ok_ref] - dataset["u0"][ok_ref]) / mpc.recipe.u_scaling
    assert dev.max() < 1e-5
    assert np.abs(out["iter_count"][ok_ref] - dataset["iter_count"][ok_ref]).max() <= 1
    info = mpc.S.launch_info()
    assert info["n_launches"] >= 1 and info["grid"] >= 1


def test_full_solution_kkt_and_oracle(mpc, dataset):
    from oracle import models
    from oracle.ipm import IPMSolver, IpmOptions, solve_mpc
    from oracle.nlp import FlatNLP
    nlp = FlatNLP(models.cstr_nominal())
    used = nlp.used_mask()
    idx = np.where(dataset["success"])[0][:6]
    out = mpc.make_step_batch(dataset["x0"][idx], dataset["u_prev"][idx], return_full=True)
    orc = IPMSolver(nlp, IpmOptions(max_iter=300))
    for b, i in enumerate(idx):
        p = nlp.pack_p(dataset["x0"][i], dataset["u_prev"][i])
        x, lam_g, lam_x = out["x"][b], out["lam_g"][b], out["lam_x"][b]
        assert np.abs(nlp.g(x, p)).max() < 1e-8                              # primal infeasibility
        stat = nlp.grad_f(x, p) + nlp.jac_g(x, p).T @ lam_g + lam_x
        assert np.abs(stat[used]).max() < 1e-7                                  # dual infeasibility
        r = solve_mpc(nlp, dataset["x0"][i], dataset["u_prev"][i], solver=orc)
        assert r["iter_count"] == out["iter_count"][b]
        assert np.abs(r["x"][used] - x[used]).max() < 1e-8
        # unused reference variables keep the initial guess, like set_initial_guess leaves them
        w0 = nlp.initial_guess(dataset["x0"][i], dataset["u_prev"][i])
        assert np.array_equal(x[~used], w0[~used])


def test_nlpsol_call_signature_and_make_step(mpc, dataset):
    i = 0
    mpc.reset_history()
    mpc.x0 = dataset["x0"][i]; mpc.u0 = dataset["u_prev"][i]
    mpc.set_initial_guess()
    u0 = mpc.make_step(dataset["x0"][i].reshape(-1, 1))
    assert u0.shape == (2, 1)
    assert np.abs(u0.ravel() - dataset["u0"][i]).max() / 2000 < 1e-5
    st = mpc.S.stats()
    assert st["success"] is True and st["iter_count"] == dataset["iter_count"][i] and st["return_status"] == "Solve_Succeeded"
    assert mpc.data["_u"].shape == (1, 2) and mpc.data["_opt_x_num"].shape == (1, 376) and mpc.data["_lam_g_num"].shape == (1, 324)
    # second step warm-starts from the previous solution (reference optimizer.py:741-742)
    u1 = mpc.make_step(dataset["x0"][i].reshape(-1, 1))
    assert mpc.S.stats()["success"] and np.isfinite(u1).all()
    # error behaviour: shape mismatch raises, solver failure does not
    with pytest.raises(Exception):
        mpc.make_step(np.zeros((3, 1)))


def test_batch_4096_properties(mpc, dataset):
    """BASELINE config #2 size: independence from batch position, permutation invariance, per-instance KKT stats."""
    rng = np.random.RandomState(42)
    B = 4096
    x0 = rng.uniform([0.1, 0.1, 50, 50], [2, 2, 140, 140], size=(B, 4))
    up = rng.uniform([5, -8500], [100, 0], size=(B, 2))
    x0[:1000], up[:1000] = dataset["x0"], dataset["u_prev"]
    a = mpc.make_step_batch(x0, up)
    assert np.array_equal(a["success"][:1000], dataset["success"].astype(bool))
    perm = rng.permutation(B)
    b = mpc.make_step_batch(x0[perm], up[perm])
    assert np.array_equal(b["status"], a["status"][perm])
    assert np.array_equal(b["iter_count"], a["iter_count"][perm])
    assert np.array_equal(b["u0"], a["u0"][perm]), "results must not depend on the position in the batch"
    ok = a["success"]
    assert 0.55 < ok.mean() < 0.85                        # reference: 69.2 % of random starts are feasible
    st = a["stats"]
    assert st[ok, 1].max() <= 1e-8 and np.all(st[ok, 0] <= 1e-8)      # inf_pr, final mu
    lo, hi = np.array([5, -8500.0]), np.array([100, 0.0])
    assert np.all(a["u0"][ok] >= lo - 1e-4 * np.abs(lo) - 1e-6) and np.all(a["u0"][ok] <= hi + 1e-4)


def test_device_resident_api_matches_host_api(mpc, dataset):
    import torch
    r = mpc.recipe
    n = 256
    P = np.zeros((n, r.n_p)); P[:, :4] = dataset["x0"][:n]; P[:, 4:6] = dataset["u_prev"][:n]
    host = mpc.S.solve_batch(P, dataset["x0"][:n] / r.x_scaling, dataset["u_prev"][:n] / r.u_scaling,
                             lbx=mpc._lb_opt_x.full().ravel(), ubx=mpc._ub_opt_x.full().ravel())
    dev = mpc.S.solve_batch_device(torch.tensor(P, device="cuda"), torch.tensor(dataset["x0"][:n] / r.x_scaling, device="cuda"),
                                   torch.tensor(dataset["u_prev"][:n] / r.u_scaling, device="cuda"))
    torch.cuda.synchronize()
    assert np.array_equal(dev["status"].cpu().numpy(), host["status"])
    assert np.array_equal(dev["u0_scaled"].cpu().numpy(), host["u0_scaled"])


def test_batched_sampler_reproduces_reference_dataset_table(mpc, dataset):
    """do_mpc_b200.sampling: the reference's serial AMPCSampler loop as one batched solve, same table columns."""
    from do_mpc_b200.sampling import sample_open_loop, sampling_plan_box
    plan = [{"x0": dataset["x0"][i].reshape(-1, 1), "u_prev": dataset["u_prev"][i].reshape(-1, 1), "id": f"{i:03d}"} for i in range(200)]
    df = sample_open_loop(mpc, plan)
    assert list(df.columns) == ["x0", "u_prev", "id", "u0", "status", "t_make_step", "t_wall", "iter_count"]
    assert np.array_equal(df["status"].to_numpy(), dataset["success"][:200].astype(bool))
    ok = df["status"].to_numpy()
    u0 = np.array([df.iloc[i]["u0"].ravel() for i in np.where(ok)[0]])
    assert np.abs(u0 - dataset["u0"][:200][ok]).max() / 2000 < 1e-5
    assert np.abs(df["iter_count"].to_numpy()[ok] - dataset["iter_count"][:200][ok]).max() <= 1
    p2 = sampling_plan_box(mpc, 16, seed=42)
    assert len(p2) == 16 and p2[0]["x0"].shape == (4, 1) and np.all(p2[3]["u_prev"].ravel() <= [100, 0])


def test_ampc_sampler_open_and_closed_loop_feed_the_trainer(mpc, tmp_path):
    """reference workflow (documentation/.../approximate_mpc notebooks): AMPCSampler -> data_<name>_opt.pkl -> Trainer -> ApproxMPC."""
    import pandas as pd
    from do_mpc_b200.approximateMPC import AMPCSampler, ApproxMPC, Trainer
    box = dict(lbx=[0.1, 0.1, 50.0, 50.0], ubx=[2.0, 2.0, 140.0, 140.0])
    s = AMPCSampler(mpc)
    s.settings.n_samples, s.settings.dataset_name, s.settings.data_dir, s.settings.seed = 256, "ol", str(tmp_path), 3
    s.settings.lbx, s.settings.ubx = box["lbx"], box["ubx"]
    s.setup()
    df = s.default_sampling()
    assert len(df) == 256 and list(df.columns) == ["x0", "u_prev", "id", "u0", "status", "t_make_step", "t_wall", "iter_count"]
    opt = pd.read_pickle(tmp_path / "ol" / "data_ol_opt.pkl")
    assert 0.4 * 256 < len(opt) < 256 and opt["status"].all()
    c = AMPCSampler(mpc)
    c.settings.n_samples, c.settings.dataset_name, c.settings.data_dir, c.settings.seed = 64, "cl", str(tmp_path), 4
    c.settings.closed_loop_flag, c.settings.trajectory_length = True, 4
    c.settings.lbx, c.settings.ubx = box["lbx"], box["ubx"]
    c.setup()
    dfc = c.default_sampling()
    assert 64 < len(dfc) <= 256                         # trajectories stop at their first failed solve
    # consecutive samples of one trajectory: x0 of step k+1 is the simulated successor, u_prev the applied input
    t0 = dfc[dfc["id"].str.endswith("_0") & dfc["status"]].iloc[0]
    t1 = dfc[dfc["id"] == t0["id"][:-2] + "_1"].iloc[0]
    assert np.allclose(t1["u_prev"], t0["u0"]) and not np.allclose(t1["x0"], t0["x0"])
    a = ApproxMPC(mpc)
    a.settings.n_hidden_layers, a.settings.n_neurons = 2, 24
    a.settings.lbx, a.settings.ubx = box["lbx"], box["ubx"]
    a.setup()
    t = Trainer(a)
    t.settings.dataset_name, t.settings.n_epochs, t.settings.data_dir, t.settings.print_frequency = "ol", 30, str(tmp_path), 1000
    t.setup()
    h = t.default_training()
    assert h["train_loss"][-1] < h["train_loss"][0]
    assert a.make_step(np.array([[0.8], [0.4], [134.0], [130.0]])).shape == (2, 1)


def test_ipopt_warm_start_option_through_the_cabi(dataset):
    """nlpsol_opts {'ipopt.warm_start_init_point': 'yes', 'ipopt.mu_init': 1e-4}: lam_g0 / lam_x0 are taken over (dmb_solve_host_ws);
    same KKT point as the cold solve in fewer iterations, and identical to the CPU emulation of the kernel source."""
    import cstr
    from conftest import HostSim, cstr_batch_inputs
    m = cstr.cstr_nominal_mpc(setup=False)
    m.settings.nlpsol_opts = {"ipopt.warm_start_init_point": "yes", "ipopt.mu_init": 1e-4}
    m.setup()
    assert m.S.options.warm_start == 1 and m.S.options.mu_init == 1e-4
    r = m.recipe
    idx = np.where(dataset["success"])[0][:64]
    P, x0s, u0s = cstr_batch_inputs(r, dataset["x0"][idx], dataset["u_prev"][idx])
    lbx, ubx = m._lb_opt_x.full().ravel(), m._ub_opt_x.full().ravel()
    m.S.set_options(warm_start=0, mu_init=0.1)
    first = m.S.solve_batch(P, x0s, u0s, lbx=lbx, ubx=ubx, return_full=True)
    P2 = P.copy(); P2[:, r.p_x0:r.p_x0 + r.n_x] *= 1.002; P2[:, r.p_uprev:r.p_uprev + r.n_u] = first["u0_scaled"] * r.u_scaling
    cold = m.S.solve_batch(P2, x_init=first["x"], lbx=lbx, ubx=ubx, return_full=True)
    m.S.set_options(warm_start=1, mu_init=1e-4)
    warm = m.S.solve_batch(P2, x_init=first["x"], lam_g0=first["lam_g"], lam_x0=first["lam_x"], lbx=lbx, ubx=ubx, return_full=True)
    ok = cold["success"] & warm["success"]
    assert ok.mean() > 0.9
    du = np.abs(warm["u0_scaled"] - cold["u0_scaled"])[ok].max(axis=1)
    # same KKT point; a few instances have inputs that the objective barely determines (both runs stop at tol 1e-8)
    assert np.median(du) < 1e-6 and np.quantile(du, 0.9) < 1e-4, (np.median(du), np.quantile(du, 0.9), du.max())
    assert warm["iter_count"][ok].mean() < 0.7 * cold["iter_count"][ok].mean()
    sim = HostSim(m, "cstr_nom"); sim.o.warm_start = 1; sim.o.mu_init = 1e-4
    hs = sim.solve(P2[:8], x_init=first["x"][:8], lam_g0=first["lam_g"][:8], lam_x0=first["lam_x"][:8])
    assert np.array_equal(hs["iter_count"], warm["iter_count"][:8]) and np.abs(hs["x"] - warm["x"][:8]).max() < 1e-9
