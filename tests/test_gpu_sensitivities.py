"""Parametric sensitivities on the GPU through the C ABI (dmb_sens_host): against the dense restatement of the reference's
NLPDifferentiator, against the CPU emulation of the same kernel source, and against finite differences of GPU re-solves."""
import numpy as np
import pytest

from conftest import GOLDEN, HostSim

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def solved():
    import cstr
    d = np.load(GOLDEN / "cstr_dataset.npz")
    mpc = cstr.cstr_nominal_mpc()
    idx = np.where(d["success"])[0][:256]
    out = mpc.make_step_batch(d["x0"][idx], d["u_prev"][idx], return_full=True)
    assert np.all(out["success"])
    return mpc, out


def test_matches_reference_algorithm_and_cpu_emulation(solved):
    from oracle import models
    from oracle.nlp import FlatNLP
    from oracle.sens import nlp_sensitivity
    mpc, out = solved
    r = mpc.S.sensitivities(out["x"], out["lam_g"], out["p"], status=out["status"], want_dlam=True)
    assert np.all(r["sens_status"] == 0)
    nlp = FlatNLP(models.cstr_nominal())
    for b in (0, 17, 200):
        dxo, dlo = nlp_sensitivity(nlp, out["x"][b], out["lam_g"][b], out["p"][b], nlp.lbx, nlp.ubx, nlp.lbg, nlp.ubg)
        assert np.abs(r["dxdp"][b] - dxo).max() < 1e-6 * max(1.0, np.abs(dxo).max())
        assert np.abs(r["dlamdp"][b] - dlo).max() < 1e-6 * max(1.0, np.abs(dlo).max())
    sim = HostSim(mpc, "cstr_nom")
    dx, sst = sim.sens(out["x"][:32], out["lam_g"][:32], out["p"][:32])
    assert np.abs(r["dxdp"][:32] - dx.transpose(0, 2, 1)).max() < 1e-7
    # u0-only mode returns the same rows
    iu = np.asarray(mpc._opt_x.f_index(("_u", 0, 0)), dtype=int)
    r1 = mpc.S.sensitivities(out["x"], out["lam_g"], out["p"], status=out["status"], u0_only=True, cols=[0, 1, 2, 3, 4, 5])
    assert np.array_equal(r1["dxdp"], r["dxdp"][:, iu, :])


def test_du0dx0_matches_finite_differences_of_gpu_solves(solved):
    mpc, out = solved
    rec = mpc.recipe
    B = 64
    x, lam, P = out["x"][:B], out["lam_g"][:B], out["p"][:B]
    cols = list(range(rec.p_x0, rec.p_x0 + rec.n_x)) + list(range(rec.p_uprev, rec.p_uprev + rec.n_u))
    du = mpc.S.sensitivities(x, lam, P, cols=cols, u0_only=True)["dxdp"]            # [B, n_u, 6] scaled u
    lbx, ubx = mpc._lb_opt_x.full().ravel(), mpc._ub_opt_x.full().ravel()
    fd = np.zeros_like(du)
    for jc, j in enumerate(cols):
        h = 1e-5 * np.maximum(1.0, np.abs(P[:, j]))
        Pp = P.copy(); Pp[:, j] += h; Pm = P.copy(); Pm[:, j] -= h
        op = mpc.S.solve_batch(Pp, x_init=x, lbx=lbx, ubx=ubx); om = mpc.S.solve_batch(Pm, x_init=x, lbx=lbx, ubx=ubx)
        assert np.all(op["success"]) and np.all(om["success"])
        fd[:, :, jc] = (op["u0_scaled"] - om["u0_scaled"]) / (2 * h[:, None])
    # instances whose later inputs sit 1e-6..1e-3 from a bound are degenerate for the reference's primal active-set test
    dist = np.minimum(np.abs(x - lbx), np.abs(ubx - x))
    regular = np.sum((dist > 1e-6) & (dist < 1e-3), axis=1) <= 4
    err = np.abs(du - fd).max(axis=(1, 2))
    assert regular.sum() >= 20 and np.median(err[regular]) < 1e-4 and np.quantile(err[regular], 0.9) < 5e-3


def test_do_mpc_differentiator_api(solved):
    """reference usage: nlp_diff = DoMPCDifferentiator(mpc); mpc.make_step(x0); dxdp, dlamdp = nlp_diff.differentiate();
    nlp_diff.sens_num['dxdp', indexf['_u',0,0], indexf['_x0']] (do_mpc/differentiator/_nlpdifferentiator.py:700-790)."""
    import do_mpc_b200 as do_mpc
    from do_mpc_b200.differentiator import indexf
    mpc, out = solved
    d = np.load(GOLDEN / "cstr_dataset.npz")
    i = int(np.where(d["success"])[0][0])
    nlp_diff = do_mpc.differentiator.DoMPCDifferentiator(mpc)
    mpc.reset_history()
    mpc.x0 = d["x0"][i]; mpc.u0 = d["u_prev"][i]; mpc.set_initial_guess()
    mpc.make_step(d["x0"][i].reshape(-1, 1))
    dxdp, dlamdp = nlp_diff.differentiate()
    assert dxdp.shape == (mpc.n_opt_x, mpc.n_opt_p) and dlamdp.shape == (mpc.recipe.n_g + mpc.n_opt_x, mpc.n_opt_p)
    assert nlp_diff.status.lse_solved
    du0dx0 = nlp_diff.sens_num["dxdp", indexf["_u", 0, 0], indexf["_x0"]]
    assert du0dx0.shape == (2, 4)
    # unscaled: rows multiplied by u_scaling, same numbers as the batch entry
    rb = nlp_diff.differentiate_batch(out, cols=[0, 1, 2, 3])
    assert np.abs(rb["du0dp"][0] - du0dx0).max() < 1e-6 * max(1.0, np.abs(du0dx0).max())


def test_tvp_and_p_columns_on_the_examples():
    """robust CSTR (p columns, scenario tree) and DIP (tvp columns, hard nl_cons) against the CPU emulation."""
    from examples_common import EXAMPLES, golden, cold_start_inputs
    for name, cols in (("CSTR", [0, 5, 12, 22, 23]), ("dip", [0, 6, 27, 107, 125])):
        make, _ = EXAMPLES[name]
        mpc = make(setup=True); r = mpc.recipe
        mpc.S.set_options(max_iter=500)
        g = golden(name)
        P, x0s, u0s, z0s = cold_start_inputs(r, g)
        out = mpc.S.solve_batch(P, x0s, u0s, z0s, lbx=mpc._lb_opt_x.full().ravel(), ubx=mpc._ub_opt_x.full().ravel(), return_full=True)
        assert out["status"][0] == 0
        rs = mpc.S.sensitivities(out["x"], out["lam_g"], P, cols=cols)
        sim = HostSim(mpc, name)
        dx, sst = sim.sens(out["x"], out["lam_g"], P, cols=cols)
        assert rs["sens_status"][0] == sst[0] == 0
        assert np.abs(rs["dxdp"][0] - dx[0].T).max() < 1e-6 * max(1.0, np.abs(dx).max()), name
