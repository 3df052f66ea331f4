"""The CUDA kernel's algorithm source, compiled for the CPU (tests/hostsim), against the reference's dataset and
the oracle.  This validates the generated code + structured KKT solve without a GPU; the GPU tests repeat the
comparison through the real C-ABI."""
import numpy as np
import pytest

import cstr
from conftest import HostSim, cstr_batch_inputs
from oracle import models
from oracle.ipm import IPMSolver, IpmOptions, solve_mpc
from oracle.nlp import FlatNLP


@pytest.fixture(scope="module")
def sim():
    mpc = cstr.cstr_nominal_mpc(setup=False)
    mpc.prepare_nlp()
    return HostSim(mpc, "cstr_nominal")


def test_recipe_sizes_match_reference_formulas(sim):
    r = sim.r
    assert (r.n_w, r.n_g, r.n_p) == (376, 324, 6)
    assert r.n_y == 12 and len(r.edge_k) == 20 and len(r.node_k) == 21


def test_dataset_parity_full(sim, dataset):
    """All 1000 shipped IPOPT solves: identical success split, u0 within 1e-5*scaling, iteration counts within 1."""
    P, x0s, u0s = cstr_batch_inputs(sim.r, dataset["x0"], dataset["u_prev"])
    sim.o.max_iter = 300
    out = sim.solve(P, x0s, u0s)
    ok_ref = dataset["success"].astype(bool)
    assert np.array_equal(out["success"], ok_ref)
    dev = np.abs(out["u0_scaled"][ok_ref] - dataset["u0"][ok_ref] / sim.r.u_scaling)
    assert dev.max() < 1e-5
    assert np.abs(out["iter_count"][ok_ref] - dataset["iter_count"][ok_ref]).max() <= 1
    assert (out["iter_count"][ok_ref] == dataset["iter_count"][ok_ref]).mean() > 0.99


def test_full_solution_is_kkt_point_of_oracle_nlp(sim, dataset):
    """x, lam_g, lam_x returned in the reference's flat layout satisfy the oracle NLP's KKT conditions."""
    nlp = FlatNLP(models.cstr_nominal())
    idx = np.where(dataset["success"])[0][:8]
    P, x0s, u0s = cstr_batch_inputs(sim.r, dataset["x0"][idx], dataset["u_prev"][idx])
    out = sim.solve(P, x0s, u0s)
    used = nlp.used_mask()
    for b in range(len(idx)):
        x, lam_g, lam_x, p = out["x"][b], out["lam_g"][b], out["lam_x"][b], P[b]
        g = nlp.g(x, p)
        assert np.abs(g).max() < 1e-8
        assert np.abs(g - out["g"][b]).max() < 1e-9
        stat = nlp.grad_f(x, p) + nlp.jac_g(x, p).T @ lam_g + lam_x
        assert np.abs(stat[used]).max() < 1e-7
        assert abs(nlp.f(x, p) - out["f"][b]) < 1e-9 * max(1, abs(out["f"][b]))
        assert np.all(x[used] >= nlp.lbx[used] - 1e-7) and np.all(x[used] <= nlp.ubx[used] + 1e-7)


def test_matches_oracle_iterates(sim, dataset):
    """Same algorithm, different linear algebra (stage elimination vs dense LDL^T): solutions agree to 1e-9."""
    nlp = FlatNLP(models.cstr_nominal())
    orc = IPMSolver(nlp, IpmOptions(max_iter=300))
    idx = [0, 1, 3, 7, 11]
    P, x0s, u0s = cstr_batch_inputs(sim.r, dataset["x0"][idx], dataset["u_prev"][idx])
    out = sim.solve(P, x0s, u0s)
    for b, i in enumerate(idx):
        r = solve_mpc(nlp, dataset["x0"][i], dataset["u_prev"][i], solver=orc)
        assert r["success"] == bool(out["success"][b])
        assert r["iter_count"] == out["iter_count"][b]
        if r["success"]:
            used = nlp.used_mask()
            assert np.abs(r["x"][used] - out["x"][b][used]).max() < 1e-8
            assert np.abs(r["lam_g"] - out["lam_g"][b]).max() < 1e-6 * max(1.0, np.abs(r["lam_g"]).max())


def test_pathological_inputs_terminate_with_a_failure_status():
    """NaN / Inf / absurd initial states must end the solve at once with a failure status (the reference does not raise either,
    optimizer.py:770-778: only `success` is False) — on the device a solve that never ends would take the whole batch with it."""
    import cstr
    from conftest import HostSim
    mpc = cstr.cstr_nominal_mpc(setup=False); mpc.prepare_nlp()
    sim = HostSim(mpc, "cstr_nom"); sim.o.max_iter = 300
    x0 = np.array([0.8, 0.5, 134.14, 130.0]); up = np.array([5.0, -1000.0])
    X = np.stack([np.where(np.arange(4) == 0, np.nan, x0), np.where(np.arange(4) == 1, np.inf, x0), x0 * 1e12, -10 * x0, 0 * x0, x0])
    P, x0s, u0s, z0s = mpc.pack_batch(X, np.repeat(up[None, :], len(X), axis=0), None)
    out = sim.solve(P, x0s=x0s, u0s=u0s, z0s=z0s)
    assert np.all(out["iter_count"] <= 300)
    assert not out["success"][:5].any() and out["success"][5]                 # the sane instance in the same batch still converges
    assert np.all(np.isin(out["status"][:2], (-3, -13)))                      # NaN / Inf: step computation fails / invalid number
