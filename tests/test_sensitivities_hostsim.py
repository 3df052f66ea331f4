"""Parametric sensitivities (reference do_mpc/differentiator): the kernel's algorithm source (CPU emulation of
csrc/ipm_sens.cuh) against (a) the dense restatement of the reference's NLPDifferentiator (oracle/sens.py) and
(b) finite differences of the solution map itself — an oracle-independent check."""
import numpy as np
import pytest

from conftest import GOLDEN, HostSim, cstr_batch_inputs
from examples_common import EXAMPLES, golden, cold_start_inputs
from oracle import models
from oracle.nlp import FlatNLP
from oracle.sens import nlp_sensitivity

ORACLE_MODELS = {"oscillatingMasses": models.oscillating_masses, "oscillatingMasses_dae": models.oscillating_masses_dae, "batch_reactor": models.batch_reactor,
                 "CSTR": models.cstr_robust, "dip": models.dip, "industrial_poly": models.industrial_poly}


@pytest.fixture(scope="module")
def cstr_solved():
    import cstr
    d = np.load(GOLDEN / "cstr_dataset.npz")
    mpc = cstr.cstr_nominal_mpc(setup=False); mpc.prepare_nlp()
    sim = HostSim(mpc, "cstr_nom")
    idx = np.where(d["success"])[0][:5]
    P, x0s, u0s = cstr_batch_inputs(mpc.recipe, d["x0"][idx], d["u_prev"][idx])
    out = sim.solve(P, x0s, u0s)
    assert np.all(out["status"] == 0)
    return mpc, sim, P, out


def test_cstr_matches_reference_algorithm(cstr_solved):
    mpc, sim, P, out = cstr_solved
    dx, dl, sst = sim.sens(out["x"], out["lam_g"], P, want_dlam=True)
    assert np.all(sst == 0)
    nlp = FlatNLP(models.cstr_nominal())
    for b in range(3):
        dxo, dlo = nlp_sensitivity(nlp, out["x"][b], out["lam_g"][b], P[b], nlp.lbx, nlp.ubx, nlp.lbg, nlp.ubg)
        assert np.abs(dx[b].T - dxo).max() < 1e-6 * max(1.0, np.abs(dxo).max())
        assert np.abs(dl[b].T - dlo).max() < 1e-6 * max(1.0, np.abs(dlo).max())


def test_cstr_matches_finite_differences_of_the_solution_map(cstr_solved):
    """du0/dx0 and du0/du_prev (what approximate-MPC training consumes) against central differences of re-solved problems.
    Only the first input is compared: far down the horizon some inputs sit 1e-6..1e-4 away from a bound with multipliers
    ~1e-4 (no strict complementarity) — the reference's primal test calls them inactive while the barrier keeps them
    where they are; their effect on u0 is below the tolerance for these instances."""
    mpc, sim, P, out = cstr_solved
    r = mpc.recipe
    cols = list(range(r.p_x0, r.p_x0 + r.n_x)) + list(range(r.p_uprev, r.p_uprev + r.n_u))
    dx, sst = sim.sens(out["x"], out["lam_g"], P, cols=cols)
    iu = np.asarray(mpc._opt_x.f_index(("_u", 0, 0)), dtype=int)
    seen = 0.0
    for b in (0, 1, 4):
        for jc, j in enumerate(cols):
            h = 1e-5 * max(1.0, abs(P[b, j]))
            Pp = P[b:b + 1].copy(); Pp[0, j] += h
            Pm = P[b:b + 1].copy(); Pm[0, j] -= h
            op = sim.solve(Pp, x_init=out["x"][b:b + 1]); om = sim.solve(Pm, x_init=out["x"][b:b + 1])
            assert op["status"][0] == 0 and om["status"][0] == 0
            fd = (op["x"][0] - om["x"][0]) / (2 * h)
            assert np.abs(dx[b, jc][iu] - fd[iu]).max() < 2e-4, (b, j)
            seen = max(seen, np.abs(fd[iu]).max())
    assert seen > 0.1          # the compared derivatives are not trivially zero


@pytest.mark.parametrize("name,cols", [("oscillatingMasses", None), ("oscillatingMasses_dae", None), ("batch_reactor", None), ("CSTR", [0, 2, 5, 12, 22, 23]),
                                        ("dip", [0, 4, 6, 27, 106, 107, 110, 125])])
def test_model_families_match_reference_algorithm(name, cols):
    """discrete / multi-element collocation / scenario tree + soft constraints / DAE + hard nl_cons + tvp:
    x0, u_prev, tvp and p columns.  (industrial_poly is not in this list: with the reference's active-set rule its reduced
    KKT matrix is singular — 9 'active' rows, LICQ violated — and the reference's sparse solve fails; see the next test.)"""
    make, _ = EXAMPLES[name]
    mpc = make(); mpc.prepare_nlp(); r = mpc.recipe
    g = golden(name)
    sim = HostSim(mpc, name); sim.o.max_iter = 500
    P, x0s, u0s, z0s = cold_start_inputs(r, g)
    out = sim.solve(P, x0s=x0s, u0s=u0s, z0s=z0s)
    assert out["status"][0] == 0
    cols = list(range(r.n_p)) if cols is None else cols
    dx, dl, sst = sim.sens(out["x"], out["lam_g"], P, cols=cols, want_dlam=True)
    assert sst[0] == 0
    nlp = FlatNLP(ORACLE_MODELS[name]())
    dxo, dlo = nlp_sensitivity(nlp, out["x"][0], out["lam_g"][0], P[0], nlp.lbx, nlp.ubx, nlp.lbg, nlp.ubg, cols=cols)
    assert np.abs(dx[0].T - dxo).max() < 1e-6 * max(1.0, np.abs(dxo).max())
    assert np.abs(dl[0].T - dlo).max() < 1e-6 * max(1.0, np.abs(dlo).max())
    for k in range(len(cols)):          # column-wise: small columns must be small, large ones accurate
        assert np.abs(dx[0, k] - dxo[:, k]).max() < 1e-6 * max(1e-2, np.abs(dxo[:, k]).max()), (name, cols[k])


def test_industrial_poly_is_degenerate_for_the_reference_rule_too():
    """industrial_poly has unscaled temperature bounds ~360: IPOPT's bound_relax_factor moves them by 3.6e-6 > active_set_tol,
    so the reference's primal test keeps only 9 of the ~530 bounds the solver converged onto.  The reference's reduced KKT
    matrix is then singular (its sparse solve fails); the structured solve reports an indefinite reduced Hessian (status 1).
    The relaxed-bounds rule (extension) finds the full active set; the KKT point itself is not a strict minimiser
    (IPOPT converged with delta_w > 0), so no finite-difference claim is made for this model."""
    make, _ = EXAMPLES["industrial_poly"]
    mpc = make(); mpc.prepare_nlp(); r = mpc.recipe
    g = golden("industrial_poly")
    sim = HostSim(mpc, "industrial_poly"); sim.o.max_iter = 500
    P, x0s, u0s, z0s = cold_start_inputs(r, g)
    out = sim.solve(P, x0s=x0s, u0s=u0s, z0s=z0s)
    d_ref, dl_ref, s_ref = sim.sens(out["x"], out["lam_g"], P, cols=[0], relaxed_rule=0, want_dlam=True)
    d_rel, dl_rel, s_rel = sim.sens(out["x"], out["lam_g"], P, cols=[0], relaxed_rule=1, want_dlam=True)
    assert s_ref[0] == 1 and np.all(np.isfinite(d_ref))
    n_act = lambda dl: int(np.sum(dl[0, 0][r.n_g:] != 0.0))
    assert n_act(dl_ref) < 20 and n_act(dl_rel) > 400
    nlp = FlatNLP(models.industrial_poly())
    with pytest.raises(np.linalg.LinAlgError):
        nlp_sensitivity(nlp, out["x"][0], out["lam_g"][0], P[0], nlp.lbx, nlp.ubx, nlp.lbg, nlp.ubg, cols=[0])


def test_both_active_set_rules_agree_when_bounds_are_order_one(cstr_solved):
    mpc, sim, P, out = cstr_solved
    a, _ = sim.sens(out["x"], out["lam_g"], P, cols=[0, 5], relaxed_rule=0)
    b, _ = sim.sens(out["x"], out["lam_g"], P, cols=[0, 5], relaxed_rule=1)
    assert np.abs(a - b).max() < 1e-7


def test_model_parameter_columns_match_finite_differences():
    """Oracle-independent check of the generated d/dtheta right-hand sides: the uncertain model parameters of the batch reactor
    (Y_x, S_in of the nominal scenario) against central differences of re-solved problems, first input.  Agreement to a few per
    cent only: inputs that sit 1e-6..1e-3 from a bound further down the horizon are 'inactive' for the reference's primal test but
    held by the barrier in the re-solves (see the CSTR test above); the exact comparison is the one against oracle/sens.py."""
    name, cols = "batch_reactor", [4, 5]
    make, _ = EXAMPLES[name]
    mpc = make(); mpc.prepare_nlp(); r = mpc.recipe
    g = golden(name)
    sim = HostSim(mpc, name); sim.o.max_iter = 500
    P, x0s, u0s, z0s = cold_start_inputs(r, g)
    out = sim.solve(P, x0s=x0s, u0s=u0s, z0s=z0s)
    assert out["status"][0] == 0
    dx, sst = sim.sens(out["x"], out["lam_g"], P, cols=cols)
    iu = np.asarray(mpc._opt_x.f_index(("_u", 0, 0)), dtype=int)
    for jc, j in enumerate(cols):
        h = 1e-4 * max(1.0, abs(P[0, j]))
        Pp = P.copy(); Pp[0, j] += h; Pm = P.copy(); Pm[0, j] -= h
        op = sim.solve(Pp, x_init=out["x"]); om = sim.solve(Pm, x_init=out["x"])
        assert op["status"][0] == 0 and om["status"][0] == 0
        fd = (op["x"][0] - om["x"][0]) / (2 * h)
        scale = np.abs(fd[iu]).max()
        assert scale > 1e-5 and np.abs(dx[0, jc][iu] - fd[iu]).max() < 5e-2 * scale, (j, dx[0, jc][iu], fd[iu])


@pytest.mark.parametrize("name", ["kite", "kinematic_bicycle", "dynamic_bicycle"])
def test_example_families_without_golden(name):
    """examples/kite (objective through a parameter-dependent aux expression: the columns of the uncertain parameters are not
    trivial; soft constraint with a bounded slack), kinematic / dynamic bicycle: every opt_p column of dx*/dp against the
    restated reference algorithm and, for the first input, against central differences of re-solved problems.  The multiplier
    sensitivities are compared where they are unique (the dynamic bicycle ends with the yaw-rate bound active on node and
    collocation copies of the same state: LICQ fails, the split of the bound multipliers is arbitrary while dx is not)."""
    import bicycle, kite
    make, mo, x0, up = {"kite": (lambda: kite.kite_mpc(setup=False), models.kite, np.array([0.5, 0.3, 0.2]), np.zeros(1)),
                        "kinematic_bicycle": (lambda: bicycle.kinematic_mpc(setup=False), models.kinematic_bicycle, np.array([0.2, -0.3, 0.1, 0.5]), np.zeros(2)),
                        "dynamic_bicycle": (lambda: bicycle.dynamic_mpc(setup=False), models.dynamic_bicycle, np.array([0.0, 0.2, 0.05, 1.0, 0.0, 0.0]), np.zeros(2))}[name]
    mpc = make(); mpc.prepare_nlp(); r = mpc.recipe
    sim = HostSim(mpc, "sens_" + name); sim.o.max_iter = 300
    P, x0s, u0s, z0s = mpc.pack_batch(x0[None], up[None], None)
    out = sim.solve(P, x0s=x0s, u0s=u0s, z0s=z0s)
    assert out["status"][0] == 0
    dx, dl, sst = sim.sens(out["x"], out["lam_g"], P, want_dlam=True)
    assert sst[0] == 0
    nlp = FlatNLP(mo())
    dxo, dlo = nlp_sensitivity(nlp, out["x"][0], out["lam_g"][0], P[0], nlp.lbx, nlp.ubx, nlp.lbg, nlp.ubg)
    assert np.abs(dx[0].T - dxo).max() < 1e-6 * max(1.0, np.abs(dxo).max())
    if name != "dynamic_bicycle":
        assert np.abs(dl[0].T - dlo).max() < 1e-6 * max(1.0, np.abs(dlo).max())
    iu = np.asarray(mpc._opt_x.f_index(("_u", 0, 0)), dtype=int)
    seen = 0.0
    for j in range(r.n_p):
        h = 1e-5 * max(1.0, abs(P[0, j]))
        Pp = P.copy(); Pp[0, j] += h
        Pm = P.copy(); Pm[0, j] -= h
        op = sim.solve(Pp, x_init=out["x"]); om = sim.solve(Pm, x_init=out["x"])
        assert op["status"][0] == 0 and om["status"][0] == 0
        fd = (op["x"][0] - om["x"][0]) / (2 * h)
        assert np.abs(dx[0, j][iu] - fd[iu]).max() < 5e-4 * max(1.0, np.abs(fd[iu]).max()), (name, j)
        seen = max(seen, np.abs(fd[iu]).max())
    assert seen > 0.05
