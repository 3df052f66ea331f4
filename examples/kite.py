"""Towing kite (same problem as the reference's examples/kite: template_model.py:33-85, template_mpc.py:33-90):
3 states, 1 input, 2 uncertain parameters (glide ratio E_0, wind speed v_0), thrust maximisation through an
aux expression that reads the parameters, soft height constraint on an aux expression, N = 80."""
import numpy as np

import do_mpc_b200 as do_mpc
from do_mpc_b200.casadi_api import DM, cos, sin, sqrt, tan


def kite_model(symvar_type="SX"):
    model = do_mpc.model.Model("continuous", symvar_type)
    L_tether, A, rho, beta, c_tilde = 400.0, 300.0, 1.0, 0, 0.028
    theta = model.set_variable("_x", "theta"); phi = model.set_variable("_x", "phi"); psi = model.set_variable("_x", "psi")
    u_tilde = model.set_variable("_u", "u_tilde")
    E_0 = model.set_variable("_p", "E_0"); v_0 = model.set_variable("_p", "v_0")
    model.set_expression("E_0", E_0); model.set_expression("v_0", v_0)
    E = E_0 - c_tilde * u_tilde ** 2
    v_a = v_0 * E * cos(theta)
    P_D = (rho * v_0 ** 2) / 2.0
    T_F = (P_D * A * cos(theta) ** 2 * (E + 1.0) * sqrt(E ** 2 + 1.0)) * (cos(theta) * np.cos(beta) + sin(theta) * np.sin(beta) * sin(phi))
    model.set_expression("T_F", T_F)
    model.set_expression("height_kite", L_tether * sin(theta) * cos(phi))
    dphi = -v_a / (L_tether * sin(theta)) * sin(psi)
    model.set_rhs("theta", v_a / L_tether * (cos(psi) - tan(theta) / E))
    model.set_rhs("phi", dphi)
    model.set_rhs("psi", v_a / L_tether * u_tilde + dphi * (cos(theta)))
    model.setup()
    return model


def kite_mpc(model=None, w_ref=10.0, E_0=6.0, h_min=100.0, n_horizon=80, setup=True):
    model = model or kite_model()
    mpc = do_mpc.controller.MPC(model)
    s = mpc.settings
    s.n_horizon, s.n_robust, s.open_loop, s.t_step, s.store_full_solution = n_horizon, 0, 0, 0.15, True
    mpc.set_objective(mterm=DM(0), lterm=-model.aux["T_F"] / 1e4)
    mpc.set_rterm(u_tilde=0.5)
    mpc.bounds["lower", "_x", "theta"] = 0.0; mpc.bounds["upper", "_x", "theta"] = 0.5 * np.pi
    mpc.bounds["lower", "_x", "phi"] = -0.5 * np.pi; mpc.bounds["upper", "_x", "phi"] = 0.5 * np.pi
    mpc.bounds["lower", "_x", "psi"] = -1.0 * np.pi; mpc.bounds["upper", "_x", "psi"] = 1.0 * np.pi
    mpc.bounds["lower", "_u", "u_tilde"] = -10; mpc.bounds["upper", "_u", "u_tilde"] = 10
    mpc.set_nl_cons("height_kite", -model.aux["height_kite"], ub=-h_min, soft_constraint=True, penalty_term_cons=1e3,
                    maximum_violation=10)
    mpc.set_uncertainty_values(E_0=np.array([E_0]), v_0=np.array([w_ref, w_ref * 0.8, w_ref * 1.2]))
    if setup:
        mpc.setup()
    return mpc
