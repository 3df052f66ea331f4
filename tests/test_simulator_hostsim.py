"""Plant simulator (first 'next' row): the integrator source compiled for the CPU against the reference's
closed-loop goldens — golden `_x[t+1]` is the reference simulator (CVODES/IDAS, tol 1e-10) applied to (`_x[t]`, `_u[t]`)
with the nominal plant parameters of the example's template_simulator.py."""
import ctypes
import hashlib
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from examples_common import golden

import batch_reactor, cstr, dip, industrial_poly, oscillating_masses

CASES = {
    "CSTR": (lambda: cstr.cstr_model(True), 0.005, [1.0, 1.0], None),
    "batch_reactor": (batch_reactor.reactor_model, 1.0, [0.5, 200.0], None),
    "industrial_poly": (industrial_poly.poly_model, 50.0 / 3600.0, [950.0, 7.0], None),
    "dip": (dip.dip_model, 0.04, [0.2, 0.2], [0.0]),
    "oscillatingMasses": (oscillating_masses.osc_model, 0.5, None, None),
    "oscillatingMasses_dae": (oscillating_masses.osc_dae_model, 0.5, None, None),
}


def build_simhost(model, tag):
    from do_mpc_b200.solver.sim_codegen import generate_sim
    src = generate_sim(model)
    csrc = ROOT / "do_mpc_b200" / "csrc"
    h = hashlib.sha256(src.encode() + (csrc / "sim_core.cuh").read_bytes() + (ROOT / "tests/hostsim/simhost.cpp").read_bytes()).hexdigest()[:12]
    d = ROOT / "tests" / "hostsim" / "_build" / f"sim_{tag}_{h}"
    so = d / "simhost.so"
    if not so.exists():
        d.mkdir(parents=True, exist_ok=True)
        (d / "sim_gen.h").write_text(src)
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", f"-I{d}", f"-I{csrc}", str(ROOT / "tests/hostsim/simhost.cpp"), "-o", str(so)], check=True)
    return ctypes.CDLL(str(so))


def sim_step(lib, model, x, u, z, tvp, p, t_step, n_sub=20, mode=0):
    B = x.shape[0]
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    x, u, z, tvp, p = c(x), c(u), c(z), c(tvp), c(p)
    xn = np.zeros((B, model.n_x)); zn = np.zeros((B, max(model.n_z, 1))); st = np.zeros(B, np.int32)
    ptr = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
    lib.simhost_step(B, ptr(x), ptr(u), ptr(z), ptr(tvp), ptr(p), ctypes.c_double(t_step), n_sub, ctypes.c_double(1e-13), 25, mode,
                     ptr(xn), ptr(zn), ptr(st))
    return xn, zn[:, :model.n_z], st


@pytest.mark.parametrize("name", list(CASES))
def test_one_step_reproduces_reference_simulator(name):
    make, t_step, p, tvp = CASES[name]
    model = make()
    lib = build_simhost(model, name)
    g = golden(name)
    T = g["x"].shape[0]
    X, U = g["x"][:T - 1], g["u"][:T - 1]
    B = X.shape[0]
    P = None if p is None else np.tile(p, (B, 1))
    TV = None if tvp is None else np.tile(tvp, (B, 1))
    Z = None
    if model.n_z:
        _, Z, st = sim_step(lib, model, X, U, np.zeros((B, model.n_z)), TV, P, t_step, mode=1)
        assert np.all(st == 0)
    xn, zn, st = sim_step(lib, model, X, U, Z, TV, P, t_step)
    assert np.all(st == 0)
    err = np.abs(xn - g["x"][1:]) / (1.0 + np.abs(g["x"][1:]))
    assert err.max() < 5e-8, (name, err.max())       # the reference integrates with abstol = reltol = 1e-10 (simulator.py:90-93)


def test_integrator_order_and_substeps():
    """Halving the step of the order-5 method must shrink the error by >> 2 until round-off; 20 substeps are converged."""
    model = cstr.cstr_model(True)
    lib = build_simhost(model, "CSTR")
    x = np.array([[0.8, 0.5, 134.14, 130.0]]); u = np.array([[10.0, -1000.0]]); p = np.array([[1.0, 1.0]])
    ref, _, _ = sim_step(lib, model, x, u, None, None, p, 0.005, n_sub=200)
    e1 = np.abs(sim_step(lib, model, x, u, None, None, p, 0.005, n_sub=1)[0] - ref).max()
    e2 = np.abs(sim_step(lib, model, x, u, None, None, p, 0.005, n_sub=2)[0] - ref).max()
    e20 = np.abs(sim_step(lib, model, x, u, None, None, p, 0.005, n_sub=20)[0] - ref).max()
    assert e2 < e1 / 8 and e20 < 1e-10


@pytest.mark.parametrize("name", ["kite", "kinematic_bicycle", "dynamic_bicycle"])
def test_example_families_without_golden_against_an_independent_integrator(name):
    """examples/kite, kinematic_bicycle_model, dynamic_bicycle_model ship no closed-loop golden: the generated integrator code
    (tan / atan / atan2 / sqrt paths of the code generator) against scipy's DOP853 (rtol = atol = 1e-12) on the oracle's sympy
    restatement of the same right-hand side, from 16 random states and inputs inside the examples' bounds."""
    import sympy as sy
    from scipy.integrate import solve_ivp
    import bicycle, kite
    from oracle import models
    make, d, t_step, p = {"kite": (kite.kite_model, models.kite(), 0.15, [6.0, 10.0]),
                          "kinematic_bicycle": (bicycle.kinematic_model, models.kinematic_bicycle(), 0.05, None),
                          "dynamic_bicycle": (bicycle.dynamic_model, models.dynamic_bicycle(), 0.1, None)}[name]
    model = make()
    lib = build_simhost(model, name)
    m = d.model
    f = sy.lambdify([m.x, m.u, m.p], m.rhs, "numpy")
    rng = np.random.RandomState(1)
    lo = np.where(np.isfinite(d.x_lb), np.maximum(d.x_lb, -2.0), -2.0); hi = np.where(np.isfinite(d.x_ub), np.minimum(d.x_ub, 2.0), 2.0)
    X = lo + (0.2 + 0.6 * rng.uniform(size=(16, m.n_x))) * (hi - lo)                  # inside 60 % of the (clipped) state box
    U = d.u_lb + (0.3 + 0.4 * rng.uniform(size=(16, m.n_u))) * (d.u_ub - d.u_lb)
    P = None if p is None else np.tile(p, (16, 1))
    xn, _, st = sim_step(lib, model, X, U, None, None, P, t_step)
    assert np.all(st == 0)
    for b in range(16):
        pv = [] if p is None else list(P[b])
        sol = solve_ivp(lambda t, x: np.array(f(list(x), list(U[b]), pv), dtype=float), (0.0, t_step), X[b], method="DOP853", rtol=1e-12, atol=1e-12)
        err = np.abs(xn[b] - sol.y[:, -1]) / (1.0 + np.abs(sol.y[:, -1]))
        assert err.max() < 5e-8, (name, b, err.max())          # the bar of the golden families above
