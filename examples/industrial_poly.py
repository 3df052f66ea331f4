"""Industrial polymerisation reactor, 10 states / 3 inputs, robust multi-stage (same problem as the reference's
examples/industrial_poly; BASELINE config #5)."""
import numpy as np

import do_mpc_b200 as do_mpc
from do_mpc_b200.casadi_api import exp


def poly_model(symvar_type="SX"):
    model = do_mpc.model.Model("continuous", symvar_type)
    R, T_F, E_a, A_tank = 8.314, 25 + 273.15, 8500.0, 65.0
    k_U2, k_U1, w_WF, w_AF = 32.0, 4.0, .333, .667
    m_M_KW, fm_M_KW, m_AWT_KW, fm_AWT_KW, m_AWT, fm_AWT, m_S = 5000.0, 300000.0, 1000.0, 100000.0, 200.0, 20000.0, 39000.0
    c_pW, c_pS, c_pF, c_pR = 4.2, .47, 3.0, 5.0
    k_WS, k_AS, k_PS = 17280.0, 3600.0, 360.0
    alfa, p_1 = 5 * 20e4 * 3.6, 1.0
    names = ["m_W", "m_A", "m_P", "T_R", "T_S", "Tout_M", "T_EK", "Tout_AWT", "accum_monom", "T_adiab"]
    m_W, m_A, m_P, T_R, T_S, Tout_M, T_EK, Tout_AWT, accum_monom, T_adiab = (model.set_variable("_x", n) for n in names)
    m_dot_f = model.set_variable("_u", "m_dot_f"); T_in_M = model.set_variable("_u", "T_in_M"); T_in_EK = model.set_variable("_u", "T_in_EK")
    delH_R = model.set_variable("_p", "delH_R"); k_0 = model.set_variable("_p", "k_0")
    U_m = m_P / (m_A + m_P)
    m_ges = m_W + m_A + m_P
    k_R1 = k_0 * exp(-E_a / (R * T_R)) * ((k_U1 * (1 - U_m)) + (k_U2 * U_m))
    k_R2 = k_0 * exp(-E_a / (R * T_EK)) * ((k_U1 * (1 - U_m)) + (k_U2 * U_m))
    k_K = ((m_W / m_ges) * k_WS) + ((m_A / m_ges) * k_AS) + ((m_P / m_ges) * k_PS)
    react = k_R1 * (m_A - ((m_A * m_AWT) / (m_W + m_A + m_P)))
    ehe = p_1 * k_R2 * (m_A / m_ges) * m_AWT
    dot_m_W = m_dot_f * w_WF
    dot_m_A = (m_dot_f * w_AF) - react - ehe
    dot_m_P = react + ehe
    dot_T_R = 1. / (c_pR * m_ges) * ((m_dot_f * c_pF * (T_F - T_R)) - (k_K * A_tank * (T_R - T_S))
                                     - (fm_AWT * c_pR * (T_R - T_EK)) + (delH_R * react))
    model.set_rhs("m_W", dot_m_W); model.set_rhs("m_A", dot_m_A); model.set_rhs("m_P", dot_m_P); model.set_rhs("T_R", dot_T_R)
    model.set_rhs("T_S", 1. / (c_pS * m_S) * ((k_K * A_tank * (T_R - T_S)) - (k_K * A_tank * (T_S - Tout_M))))
    model.set_rhs("Tout_M", 1. / (c_pW * m_M_KW) * ((fm_M_KW * c_pW * (T_in_M - Tout_M)) + (k_K * A_tank * (T_S - Tout_M))))
    model.set_rhs("T_EK", 1. / (c_pR * m_AWT) * ((fm_AWT * c_pR * (T_R - T_EK)) - (alfa * (T_EK - Tout_AWT)) + (ehe * delH_R)))
    model.set_rhs("Tout_AWT", 1. / (c_pW * m_AWT_KW) * ((fm_AWT_KW * c_pW * (T_in_EK - Tout_AWT)) - (alfa * (Tout_AWT - T_EK))))
    model.set_rhs("accum_monom", m_dot_f)
    model.set_rhs("T_adiab", delH_R / (m_ges * c_pR) * dot_m_A - (dot_m_A + dot_m_W + dot_m_P) * (m_A * delH_R / (m_ges * m_ges * c_pR)) + dot_T_R)
    model.setup()
    return model


def poly_mpc(model=None, n_horizon=20, n_robust=1, setup=True):
    model = model or poly_model()
    mpc = do_mpc.controller.MPC(model)
    s = mpc.settings
    s.n_horizon, s.n_robust, s.open_loop, s.t_step = n_horizon, n_robust, 0, 50.0 / 3600.0
    s.state_discretization, s.store_full_solution = "collocation", True
    mpc.set_objective(mterm=-model.x["m_P"], lterm=-model.x["m_P"])
    mpc.set_rterm(m_dot_f=0.002, T_in_M=0.004, T_in_EK=0.002)
    tr = 2.0
    lo = dict(m_W=0.0, m_A=0.0, m_P=26.0, T_R=363.15 - tr, T_S=298.0, Tout_M=298.0, T_EK=288.0, Tout_AWT=288.0, accum_monom=0.0)
    hi = dict(T_R=363.15 + tr, T_S=400.0, Tout_M=400.0, T_EK=400.0, Tout_AWT=400.0, accum_monom=30000.0, T_adiab=382.15)
    for k, v in lo.items(): mpc.bounds["lower", "_x", k] = v
    for k, v in hi.items(): mpc.bounds["upper", "_x", k] = v
    mpc.bounds["lower", "_u", "m_dot_f"] = 0.0; mpc.bounds["lower", "_u", "T_in_M"] = 333.15; mpc.bounds["lower", "_u", "T_in_EK"] = 333.15
    mpc.bounds["upper", "_u", "m_dot_f"] = 3.0e4; mpc.bounds["upper", "_u", "T_in_M"] = 373.15; mpc.bounds["upper", "_u", "T_in_EK"] = 373.15
    for n in ("m_W", "m_A", "m_P", "accum_monom"): mpc.scaling["_x", n] = 10
    mpc.scaling["_u", "m_dot_f"] = 100
    if n_robust == 0:
        mpc.set_nl_cons("T_R_UB", model.x["T_R"], ub=363.15 + tr, soft_constraint=True, penalty_term_cons=1e4)
    mpc.set_uncertainty_values(delH_R=np.array([950.0, 950.0 * 1.30, 950.0 * 0.70]), k_0=np.array([7.0, 7.0 * 1.30, 7.0 * 0.70]))
    if setup:
        mpc.setup()
    return mpc
