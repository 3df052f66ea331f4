"""Continuous stirred tank reactor (van de Vusse kinetics) — model and controllers.

Same optimal-control problems as the reference's examples/CSTR (robust multi-stage) and
examples/CSTR_approximate_mpc (nominal, BASELINE config #2), written against this package's API.
"""
import numpy as np

import do_mpc_b200 as do_mpc
from do_mpc_b200.casadi_api import exp


def cstr_model(uncertain: bool = True, symvar_type: str = "SX"):
    model = do_mpc.model.Model("continuous", symvar_type)
    k0 = dict(ab=1.287e12, bc=1.287e12, ad=9.043e9)          # pre-exponential factors [1/h], [l/(mol h)]
    ea = dict(ab=9758.3, bc=9758.3, ad=8560.0)                # activation energies / R  [K]
    dh = dict(ab=4.2, bc=-11.0, ad=-41.85)                    # reaction enthalpies [kJ/mol]
    rho, cp, cp_k = 0.9342, 3.01, 2.0                         # density, heat capacities
    area, volume, m_k, t_in, k_w = 0.215, 10.01, 5.0, 130.0, 4032.0
    c_a0 = (5.7 + 4.5) / 2.0

    C_a = model.set_variable("_x", "C_a"); C_b = model.set_variable("_x", "C_b")
    T_R = model.set_variable("_x", "T_R"); T_K = model.set_variable("_x", "T_K")
    F = model.set_variable("_u", "F"); Q_dot = model.set_variable("_u", "Q_dot")
    if uncertain:
        alpha = model.set_variable("_p", "alpha"); beta = model.set_variable("_p", "beta")
    else:
        alpha = beta = 1
    T_dif = model.set_expression("T_dif", T_R - T_K)
    absT = T_R + 273.15
    K_1 = beta * k0["ab"] * exp(-ea["ab"] / absT)
    K_2 = k0["bc"] * exp(-ea["bc"] / absT)
    K_3 = k0["ad"] * exp(-alpha * ea["ad"] / absT)

    model.set_rhs("C_a", F * (c_a0 - C_a) - K_1 * C_a - K_3 * C_a ** 2)
    model.set_rhs("C_b", -F * C_b + K_1 * C_a - K_2 * C_b)
    heat = (K_1 * C_a * dh["ab"] + K_2 * C_b * dh["bc"] + K_3 * C_a ** 2 * dh["ad"]) / (-rho * cp)
    model.set_rhs("T_R", heat + F * (t_in - T_R) + (k_w * area * (-T_dif)) / (rho * cp * volume))
    model.set_rhs("T_K", (Q_dot + k_w * area * T_dif) / (m_k * cp_k))
    model.setup()
    return model


def _common(mpc, n_horizon, n_robust):
    s = mpc.settings
    s.n_horizon, s.n_robust, s.open_loop, s.t_step = n_horizon, n_robust, 0, 0.005
    s.state_discretization, s.collocation_type, s.collocation_deg, s.collocation_ni = "collocation", "radau", 2, 1
    s.store_full_solution = True
    mpc.scaling["_x", "T_R"] = 100; mpc.scaling["_x", "T_K"] = 100
    mpc.scaling["_u", "Q_dot"] = 2000; mpc.scaling["_u", "F"] = 100
    mpc.set_rterm(F=0.1, Q_dot=1e-3)
    for name, lo in (("C_a", 0.1), ("C_b", 0.1), ("T_R", 50), ("T_K", 50)):
        mpc.bounds["lower", "_x", name] = lo
    mpc.bounds["upper", "_x", "C_a"] = 2; mpc.bounds["upper", "_x", "C_b"] = 2; mpc.bounds["upper", "_x", "T_K"] = 140
    mpc.bounds["lower", "_u", "F"] = 5; mpc.bounds["lower", "_u", "Q_dot"] = -8500
    mpc.bounds["upper", "_u", "F"] = 100; mpc.bounds["upper", "_u", "Q_dot"] = 0.0


def cstr_nominal_mpc(model=None, n_horizon=20, setup=True):
    """BASELINE config #2: nominal controller, N=20, collocation degree 2, hard T_R <= 140."""
    model = model or cstr_model(uncertain=False)
    mpc = do_mpc.controller.MPC(model)
    _common(mpc, n_horizon, 0)
    cost = (model.x["C_b"] - 0.6) ** 2 + (model.x["C_a"] - 0.7) ** 2
    mpc.set_objective(mterm=cost, lterm=cost)
    mpc.bounds["upper", "_x", "T_R"] = 140
    if setup:
        mpc.setup()
    return mpc


def cstr_robust_mpc(model=None, n_horizon=20, n_robust=1, setup=True):
    """Multi-stage controller with 3x3 uncertainty realisations and a soft T_R <= 140 constraint."""
    model = model or cstr_model(uncertain=True)
    mpc = do_mpc.controller.MPC(model)
    _common(mpc, n_horizon, n_robust)
    cost = (model.x["C_b"] - 0.6) ** 2
    mpc.set_objective(mterm=cost, lterm=cost)
    mpc.set_nl_cons("T_R", model.x["T_R"], ub=140, soft_constraint=True, penalty_term_cons=1e2)
    mpc.set_uncertainty_values(alpha=np.array([1.0, 1.05, 0.95]), beta=np.array([1.0, 1.1, 0.9]))
    if setup:
        mpc.setup()
    return mpc
