"""ipopt.warm_start_init_point = yes (with mu_init): the multipliers of the previous solution that optimizer.solve() passes as
lam_g0 / lam_x0 (do_mpc/optimizer.py:760-768) are taken over instead of being ignored.  Kernel source (CPU emulation) against
the oracle IPM with the same warm-start rule, and against the cold solve of the same problem (same KKT point, fewer iterations).
No reference golden exists for this option (no example sets it): the warm-start constants follow IPOPT's documented defaults."""
import numpy as np
import pytest

from conftest import GOLDEN, HostSim, cstr_batch_inputs
from oracle import models
from oracle.ipm import IPMSolver, IpmOptions
from oracle.nlp import FlatNLP


def test_warm_started_second_step_matches_oracle_and_cold_solution():
    import cstr
    d = np.load(GOLDEN / "cstr_dataset.npz")
    mpc = cstr.cstr_nominal_mpc(setup=False); mpc.prepare_nlp(); r = mpc.recipe
    sim = HostSim(mpc, "cstr_nom")
    idx = np.where(d["success"])[0][:4]
    P, x0s, u0s = cstr_batch_inputs(r, d["x0"][idx], d["u_prev"][idx])
    first = sim.solve(P, x0s, u0s)
    assert np.all(first["status"] == 0)
    # "next sampling instant": the state moved a little, u_prev is the input just applied
    P2 = P.copy()
    P2[:, r.p_x0:r.p_x0 + r.n_x] *= 1.002
    P2[:, r.p_uprev:r.p_uprev + r.n_u] = first["u0_scaled"] * r.u_scaling
    cold = sim.solve(P2, x_init=first["x"])
    sim.o.warm_start = 1; sim.o.mu_init = 1e-4
    warm = sim.solve(P2, x_init=first["x"], lam_g0=first["lam_g"], lam_x0=first["lam_x"])
    assert np.all(cold["status"] == 0) and np.all(warm["status"] == 0)
    assert np.abs(warm["u0_scaled"] - cold["u0_scaled"]).max() < 1e-6           # same KKT point
    assert warm["iter_count"].mean() < 0.6 * cold["iter_count"].mean(), (warm["iter_count"], cold["iter_count"])
    nlp = FlatNLP(models.cstr_nominal())
    orc = IPMSolver(nlp, IpmOptions(max_iter=300, warm_start_init_point=True, mu_init=1e-4))
    for b in range(4):
        ro = orc.solve(first["x"][b], P2[b], lam_g0=first["lam_g"][b], lam_x0=first["lam_x"][b])
        assert ro["success"] and abs(ro["iter_count"] - int(warm["iter_count"][b])) <= 1
        assert np.abs(ro["x"][nlp.ix_u[0, 0]] - warm["u0_scaled"][b]).max() < 1e-7
    # multipliers are ignored without the option, like IPOPT
    sim.o.warm_start = 0; sim.o.mu_init = 0.1
    again = sim.solve(P2, x_init=first["x"], lam_g0=first["lam_g"], lam_x0=first["lam_x"])
    assert np.array_equal(again["iter_count"], cold["iter_count"]) and np.array_equal(again["x"], cold["x"])
