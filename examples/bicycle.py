"""Bicycle vehicle models (same problems as the reference's examples/kinematic_bicycle_model and
examples/dynamic_bicycle_model: template_model.py:33-90 / template_mpc.py:33-85 of each)."""
import numpy as np

import do_mpc_b200 as do_mpc
from do_mpc_b200.casadi_api import atan, atan2, cos, sin, sqrt, tan


def kinematic_model(symvar_type="SX"):
    model = do_mpc.model.Model("continuous", symvar_type)
    lf, lr = 0.3, 0.3
    X_p = model.set_variable("_x", "X_p"); Y_p = model.set_variable("_x", "Y_p")
    Psi = model.set_variable("_x", "Psi"); V = model.set_variable("_x", "V")
    Delta = model.set_variable("_u", "Delta"); Acc = model.set_variable("_u", "Acc")
    Beta = atan((lr / (lr + lf)) * tan(Delta))
    model.set_rhs("X_p", V * cos(Psi + Beta))
    model.set_rhs("Y_p", V * sin(Psi + Beta))
    model.set_rhs("Psi", (V / lr) * sin(Beta))
    model.set_rhs("V", Acc)
    model.setup()
    return model


def kinematic_mpc(model=None, n_horizon=10, setup=True):
    model = model or kinematic_model()
    mpc = do_mpc.controller.MPC(model)
    s = mpc.settings
    s.n_horizon, s.n_robust, s.open_loop, s.t_step = n_horizon, 0, 0, 0.05
    s.state_discretization, s.collocation_type, s.collocation_deg, s.collocation_ni = "collocation", "radau", 2, 1
    s.store_full_solution = True
    mterm = (model.x["Y_p"] - 2) ** 2 + (model.x["X_p"] - 3) ** 2 + (model.x["Psi"] - 0) ** 2
    lterm = (model.x["Y_p"] - 1) ** 2
    mpc.set_objective(mterm=mterm, lterm=lterm * 0)
    mpc.set_rterm(Delta=1.0, Acc=1e-3)
    for n, v in (("X_p", 50), ("Y_p", 50), ("Psi", np.pi / 2), ("V", 5)):
        mpc.bounds["lower", "_x", n] = -v; mpc.bounds["upper", "_x", n] = v
    for n in ("Delta", "Acc"):
        mpc.bounds["lower", "_u", n] = -5; mpc.bounds["upper", "_u", n] = 5
    if setup:
        mpc.setup()
    return mpc


def dynamic_model(symvar_type="SX"):
    model = do_mpc.model.Model("continuous", symvar_type)
    m, I_z, lf, lr = 5.692, 0.204, 0.178, 0.147
    D_f, D_r, C_f, C_r, B_f, B_r = 134.585, 159.919, 0.085, 0.133, 9.242, 17.716
    c_m1, c_m2, c_m3, c_m4 = 20, 6.92 * 1e-7, 3.99, 0.67
    X_p = model.set_variable("_x", "X_p"); Y_p = model.set_variable("_x", "Y_p"); Psi = model.set_variable("_x", "Psi")
    V_x = model.set_variable("_x", "V_x"); V_y = model.set_variable("_x", "V_y"); W = model.set_variable("_x", "W")
    Delta = model.set_variable("_u", "Delta"); d = model.set_variable("_u", "d")
    model.set_expression("Vel", sqrt(V_x ** 2 + V_y ** 2))
    alpha_f = -atan2(W * lf + V_y, V_x) + Delta
    alpha_r = atan2((W * lr - V_y), V_x)
    F_f_y = D_f * sin(C_f * atan(B_f * alpha_f))
    F_r_y = D_r * sin(C_r * atan(B_r * alpha_r))
    F_x = (c_m1 - c_m2 * V_x) * d - c_m4 * V_x ** 2 - c_m3
    model.set_rhs("X_p", V_x * cos(Psi) - V_y * sin(Psi))
    model.set_rhs("Y_p", V_x * sin(Psi) + V_y * cos(Psi))
    model.set_rhs("Psi", W)
    model.set_rhs("V_x", (1 / m) * (F_x - F_f_y * sin(Delta) + m * V_y * W))
    model.set_rhs("V_y", (1 / m) * (F_r_y + F_f_y * cos(Delta) - m * V_x * W))
    model.set_rhs("W", (1 / I_z) * (F_f_y * lf * cos(Delta) - lf * F_x * sin(Delta) - lr * F_r_y))
    model.setup()
    return model


def dynamic_mpc(model=None, n_horizon=10, setup=True):
    model = model or dynamic_model()
    mpc = do_mpc.controller.MPC(model)
    s = mpc.settings
    s.n_horizon, s.n_robust, s.open_loop, s.t_step = n_horizon, 0, 0, 0.1
    s.state_discretization, s.collocation_type, s.collocation_deg, s.collocation_ni = "collocation", "radau", 2, 1
    s.store_full_solution = True
    cost = (model.x["Y_p"] - 1) ** 2
    mpc.set_objective(mterm=cost, lterm=cost)
    mpc.set_rterm(Delta=1e-3, d=1e-3)
    for n, lo, hi in (("X_p", -50000, 50000), ("Y_p", -2, 2), ("Psi", -0.78, 0.78), ("V_x", 0.1, 5), ("V_y", -1, 1), ("W", -0.2, 0.2)):
        mpc.bounds["lower", "_x", n] = lo; mpc.bounds["upper", "_x", n] = hi
    mpc.bounds["lower", "_u", "Delta"] = -2; mpc.bounds["upper", "_u", "Delta"] = 2
    mpc.bounds["lower", "_u", "d"] = 0; mpc.bounds["upper", "_u", "d"] = 1
    if setup:
        mpc.setup()
    return mpc
