"""Two oscillating masses, discrete-time linear model (same problem as the reference's
examples/oscillating_masses_discrete; BASELINE config #1 uses N=10)."""
import numpy as np

import do_mpc_b200 as do_mpc
from do_mpc_b200.casadi_api import sum1


def osc_model(symvar_type="SX"):
    model = do_mpc.model.Model("discrete", symvar_type)
    x = model.set_variable("_x", "x", shape=(4, 1))
    u = model.set_variable("_u", "u", shape=(1, 1))
    model.set_expression("cost", sum1(x ** 2))
    A = np.array([[0.763, 0.460, 0.115, 0.020], [-0.899, 0.763, 0.420, 0.115],
                  [0.115, 0.020, 0.763, 0.460], [0.420, 0.115, -0.899, 0.763]])
    B = np.array([[0.014], [0.063], [0.221], [0.367]])
    model.set_rhs("x", A @ x + B @ u)
    model.setup()
    return model


def osc_dae_model(symvar_type="SX", uncertain=False):
    """The same plant as a discrete DAE (reference examples/oscillating_masses_discrete_dae): x+ = z, 0 = z - A x - B u.
    uncertain=True: the input gain carries an uncertain factor kB (test variant for DAE + scenario tree)."""
    model = do_mpc.model.Model("discrete", symvar_type)
    x = model.set_variable("_x", "x", shape=(4, 1))
    u = model.set_variable("_u", "u", shape=(1, 1))
    model.set_expression("cost", sum1(x ** 2))
    A = np.array([[0.763, 0.460, 0.115, 0.020], [-0.899, 0.763, 0.420, 0.115],
                  [0.115, 0.020, 0.763, 0.460], [0.420, 0.115, -0.899, 0.763]])
    B = np.array([[0.014], [0.063], [0.221], [0.367]])
    x_next = model.set_variable("_z", "x_next", shape=(4, 1))
    kB = model.set_variable("_p", "kB") if uncertain else 1.0
    model.set_rhs("x", x_next)
    model.set_alg("x_next", x_next - A @ x - kB * (B @ u))
    model.setup()
    return model


def osc_dae_mpc(n_horizon=7, setup=True):
    return osc_mpc(osc_dae_model(), n_horizon=n_horizon, setup=setup)


def osc_dae_robust_mpc(n_horizon=5, n_robust=1, setup=True):
    return osc_mpc(osc_dae_model(uncertain=True), n_horizon=n_horizon, setup=setup, n_robust=n_robust)


def osc_mpc(model=None, n_horizon=7, setup=True, n_robust=0):
    model = model or osc_model()
    mpc = do_mpc.controller.MPC(model)
    mpc.settings.n_horizon, mpc.settings.n_robust, mpc.settings.open_loop, mpc.settings.t_step = n_horizon, n_robust, 0, 0.5
    mpc.settings.state_discretization, mpc.settings.store_full_solution = "discrete", True
    cost = model.aux["cost"]
    mpc.set_objective(mterm=cost, lterm=cost)
    mpc.set_rterm(u=1e-4)
    max_x = np.array([[4.0], [10.0], [4.0], [10.0]])
    mpc.bounds["lower", "_x", "x"] = -max_x; mpc.bounds["upper", "_x", "x"] = max_x
    mpc.bounds["lower", "_u", "u"] = -0.5; mpc.bounds["upper", "_u", "u"] = 0.5
    if model.n_p:
        mpc.set_uncertainty_values(kB=np.array([1.0, 1.3, 0.7]))
    if setup:
        mpc.setup()
    return mpc
