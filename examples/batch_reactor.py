"""Fed-batch bioreactor (same problem as the reference's examples/batch_reactor; BASELINE config #3 is the
robust multi-stage variant with n_robust=1, N=30)."""
import numpy as np

import do_mpc_b200 as do_mpc


def reactor_model(symvar_type="SX"):
    model = do_mpc.model.Model("continuous", symvar_type)
    mu_m, K_m, K_i, v_par, Y_p = 0.02, 0.05, 5.0, 0.004, 1.2
    X_s = model.set_variable("_x", "X_s"); S_s = model.set_variable("_x", "S_s")
    P_s = model.set_variable("_x", "P_s"); V_s = model.set_variable("_x", "V_s")
    inp = model.set_variable("_u", "inp")
    Y_x = model.set_variable("_p", "Y_x"); S_in = model.set_variable("_p", "S_in")
    mu_S = mu_m * S_s / (K_m + S_s + (S_s ** 2 / K_i))
    model.set_rhs("X_s", mu_S * X_s - inp / V_s * X_s)
    model.set_rhs("S_s", -mu_S * X_s / Y_x - v_par * X_s / Y_p + inp / V_s * (S_in - S_s))
    model.set_rhs("P_s", v_par * X_s - inp / V_s * P_s)
    model.set_rhs("V_s", inp)
    model.setup()
    return model


def reactor_mpc(model=None, n_horizon=20, n_robust=0, setup=True):
    model = model or reactor_model()
    mpc = do_mpc.controller.MPC(model)
    s = mpc.settings
    s.n_horizon, s.n_robust, s.open_loop, s.t_step = n_horizon, n_robust, 0, 1.0
    s.state_discretization, s.collocation_type, s.collocation_deg, s.collocation_ni = "collocation", "radau", 2, 2
    s.store_full_solution = True
    mpc.set_objective(mterm=-model.x["P_s"], lterm=-model.x["P_s"])
    mpc.set_rterm(inp=1.0)
    for n, v in (("X_s", 0.0), ("S_s", -0.01), ("P_s", 0.0), ("V_s", 0.0)):
        mpc.bounds["lower", "_x", n] = v
    mpc.bounds["upper", "_x", "X_s"] = 3.7; mpc.bounds["upper", "_x", "P_s"] = 3.0
    mpc.bounds["lower", "_u", "inp"] = 0.0; mpc.bounds["upper", "_u", "inp"] = 0.2
    mpc.set_uncertainty_values(Y_x=np.array([0.5, 0.4, 0.3]), S_in=np.array([200.0, 220.0, 180.0]))
    if setup:
        mpc.setup()
    return mpc
