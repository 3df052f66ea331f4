"""Pin the CPU oracle against the reference's own golden data.  CPU only."""
import numpy as np
import pytest

from conftest import GOLDEN
from oracle import models
from oracle.collocation import collocation_matrices
from oracle.ipm import IPMSolver, IpmOptions, solve_mpc
from oracle.nlp import FlatNLP

CASES = {
    "oscillatingMasses": (models.oscillating_masses, (39, 32, 5)),
    "oscillatingMasses_dae": (models.oscillating_masses_dae, (67, 60, 5)),
    "batch_reactor": (models.batch_reactor, (608, 564, 23)),
    "CSTR": (models.cstr_robust, (3564, 3064, 24)),
    "industrial_poly": (models.industrial_poly, (8100, 7210, 31)),
    "dip": (models.dip, (4330, 4506, 126)),
}


def test_collocation_constants_match_reference_values():
    # SURVEY.md App. C (values printed from casadi.collocation_points in the reference's documentation run)
    tau, C, D = collocation_matrices(2, "radau")
    assert np.allclose(tau, [0, 1 / 3, 1], atol=1e-15)
    assert np.allclose(C, [[-4, -2, 2], [4.5, 1.5, -4.5], [-0.5, 0.5, 2.5]], atol=1e-13)
    assert np.allclose(D, [0, 0, 1], atol=1e-13)
    tau, C, D = collocation_matrices(3, "radau")
    assert np.allclose(tau, [0, 0.155051025721682, 0.644948974278318, 1], atol=1e-14)
    assert np.allclose(C[0], [-9, -4.139387691339832, 1.739387691339823, -3], atol=1e-12)


@pytest.mark.parametrize("name", list(CASES))
def test_golden_points_are_kkt_points_of_the_restated_nlp(name):
    """Every stored (p -> x*, lam_g*) of the reference's regression pickles satisfies the oracle NLP's KKT
    conditions: same sizes, primal feasibility, stationarity weighted by the distance to the bounds
    (IPOPT terminates with mu ~ 1e-9 barrier residue), multiplier signs."""
    make, sizes = CASES[name]
    nlp = FlatNLP(make())
    assert (nlp.n_w, nlp.n_g, nlp.n_p) == sizes
    d = np.load(GOLDEN / f"kkt_{name}.npz")
    used = nlp.used_mask()
    for t in range(d["opt_x_num"].shape[0]):
        x = d["opt_x_num"][t] / nlp.scaling
        p, lam = d["opt_p_num"][t], d["lam_g_num"][t]
        g = nlp.g(x, p)
        viol = np.maximum(np.maximum(nlp.lbg - g, g - nlp.ubg), 0).max()
        assert viol < 2e-8
        r = nlp.grad_f(x, p) + nlp.jac_g(x, p).T @ lam
        dl, du = x - nlp.lbx, nlp.ubx - x
        comp = np.abs(r) * np.minimum(np.minimum(dl, du), 1.0)
        assert comp[used].max() < 1e-7
        assert np.all(r[(dl < 1e-6) & used] > -1e-6) and np.all(r[(du < 1e-6) & used] < 1e-6)
        assert min(dl.min(), du.min()) > -1e-5            # bound_relax_factor
        ineq = np.isinf(nlp.lbg)
        if ineq.any():
            assert lam[ineq].min() > -1e-9


def test_oracle_ipm_reproduces_shipped_cstr_dataset(dataset):
    """Cold starts as the reference's sampler does them: same success flags, same u0 (|du| <= 1e-5*scaling),
    same IPOPT iteration counts (+-1) on a deterministic subset of the 1000 shipped samples."""
    nlp = FlatNLP(models.cstr_nominal())
    sol = IPMSolver(nlp, IpmOptions(max_iter=300))
    idx = np.arange(0, 1000, 37)
    n_ok = 0
    for i in idx:
        r = solve_mpc(nlp, dataset["x0"][i], dataset["u_prev"][i], solver=sol)
        assert r["success"] == bool(dataset["success"][i])
        if r["success"]:
            n_ok += 1
            assert np.abs((r["u0"] - dataset["u0"][i]) / nlp.d.u_scaling).max() < 1e-5
            assert abs(r["iter_count"] - dataset["iter_count"][i]) <= 1
    assert n_ok >= 10


def test_oracle_ipm_restarts_at_golden_solution():
    """T1 (basin-controlled): started at a golden primal point the oracle IPM stays at that point."""
    nlp = FlatNLP(models.batch_reactor())
    d = np.load(GOLDEN / "kkt_batch_reactor.npz")
    sol = IPMSolver(nlp, IpmOptions(max_iter=100))
    x = d["opt_x_num"][0] / nlp.scaling
    r = sol.solve(x, d["opt_p_num"][0])
    assert r["success"]
    u0 = r["x"][nlp.ix_u[0, 0]] * nlp.d.u_scaling
    assert abs(u0[0] - d["u"][0][0]) < 1e-5
