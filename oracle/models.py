"""sympy restatements of the reference's example problems (TEST INFRASTRUCTURE, see oracle/README.md).

Each function returns an oracle.nlp.MPCDef that describes the same optimal-control problem the cited
reference template files configure.  Only the mathematics is restated; nothing is imported from
/root/reference.

  cstr_nominal()        examples/CSTR_approximate_mpc/template_model.py:33-105, template_mpc.py:33-91  (BASELINE cfg #2)
  cstr_robust()         examples/CSTR/template_model.py:33-100,  template_mpc.py:33-110
  batch_reactor()       examples/batch_reactor/template_model.py:33-75, template_mpc.py:33-86
  oscillating_masses()  examples/oscillating_masses_discrete/template_model.py:33-75, template_mpc.py:33-70
  oscillating_masses_dae()  examples/oscillating_masses_discrete_dae/template_model.py:33-75, template_mpc.py:33-70
  dip()                 examples/double_inverted_pendulum/template_model.py:33-150, template_mpc.py:33-95
  industrial_poly()     examples/industrial_poly/template_model.py:33-125, template_mpc.py:33-105
  kite()                examples/kite/template_model.py:33-85, template_mpc.py:33-90
  kinematic_bicycle()   examples/kinematic_bicycle_model/template_model.py:33-70, template_mpc.py:33-85
  dynamic_bicycle()     examples/dynamic_bicycle_model/template_model.py:33-95, template_mpc.py:33-90
The last three have no golden in the reference's testing/results: their restatements are pinned only through the shared
NLP assembly (oracle/nlp.py) that the six golden families pin.
"""
import itertools

import numpy as np
import sympy as sy

from .nlp import ModelDef, MPCDef, NlCons


def _scenarios(*value_lists):
    """set_uncertainty_values: all combinations, first = nominal (reference _mpc.py:860-870)."""
    return np.array(list(itertools.product(*value_lists)), dtype=float)


# ------------------------------------------------------------------------------------------------
def _cstr_model(uncertain):
    C_a, C_b, T_R, T_K = sy.symbols("C_a C_b T_R T_K")
    F, Q_dot = sy.symbols("F Q_dot")
    if uncertain:
        alpha, beta = sy.symbols("alpha beta")
        p = [alpha, beta]
    else:
        alpha, beta, p = 1, 1, []
    K0_ab, K0_bc, K0_ad = 1.287e12, 1.287e12, 9.043e9
    E_A_ab, E_A_bc, E_A_ad = 9758.3, 9758.3, 8560.0
    H_R_ab, H_R_bc, H_R_ad = 4.2, -11.0, -41.85
    Rou, Cp, Cp_k, A_R, V_R, m_k, T_in, K_w = 0.9342, 3.01, 2.0, 0.215, 10.01, 5.0, 130.0, 4032.0
    C_A0 = (5.7 + 4.5) / 2.0
    T_dif = T_R - T_K
    K_1 = beta * K0_ab * sy.exp(-E_A_ab / (T_R + 273.15))
    K_2 = K0_bc * sy.exp(-E_A_bc / (T_R + 273.15))
    K_3 = K0_ad * sy.exp(-alpha * E_A_ad / (T_R + 273.15))
    rhs = [
        F * (C_A0 - C_a) - K_1 * C_a - K_3 * C_a ** 2,
        -F * C_b + K_1 * C_a - K_2 * C_b,
        (K_1 * C_a * H_R_ab + K_2 * C_b * H_R_bc + K_3 * C_a ** 2 * H_R_ad) / (-Rou * Cp)
        + F * (T_in - T_R) + (K_w * A_R * (-T_dif)) / (Rou * Cp * V_R),
        (Q_dot + K_w * A_R * T_dif) / (m_k * Cp_k),
    ]
    return ModelDef("continuous", x=[C_a, C_b, T_R, T_K], u=[F, Q_dot], p=p, rhs=rhs, aux={"T_dif": T_dif})


def cstr_nominal(n_horizon=20):
    m = _cstr_model(False)
    C_a, C_b = m.x[0], m.x[1]
    cost = (C_b - 0.6) ** 2 + (C_a - 0.7) ** 2
    return MPCDef(m, n_horizon=n_horizon, t_step=0.005, n_robust=0, collocation_deg=2, collocation_ni=1,
                  lterm=cost, mterm=cost, rterm_factor=[0.1, 1e-3],
                  x_scaling=[1, 1, 100, 100], u_scaling=[100, 2000],
                  x_lb=[0.1, 0.1, 50, 50], x_ub=[2, 2, 140, 140], u_lb=[5, -8500], u_ub=[100, 0.0])


def cstr_robust(n_horizon=20, n_robust=1):
    m = _cstr_model(True)
    C_b, T_R = m.x[1], m.x[2]
    cost = (C_b - 0.6) ** 2
    return MPCDef(m, n_horizon=n_horizon, t_step=0.005, n_robust=n_robust, collocation_deg=2, collocation_ni=1,
                  lterm=cost, mterm=cost, rterm_factor=[0.1, 1e-3],
                  x_scaling=[1, 1, 100, 100], u_scaling=[100, 2000],
                  x_lb=[0.1, 0.1, 50, 50], x_ub=[2, 2, np.inf, 140], u_lb=[5, -8500], u_ub=[100, 0.0],
                  nl_cons=[NlCons("T_R", T_R, ub=140, soft=True, penalty=1e2)],
                  p_scenarios=_scenarios([1.0, 1.05, 0.95], [1.0, 1.1, 0.9]))


# ------------------------------------------------------------------------------------------------
def batch_reactor(n_horizon=20, n_robust=0):
    X_s, S_s, P_s, V_s, inp, Y_x, S_in = sy.symbols("X_s S_s P_s V_s inp Y_x S_in")
    mu_m, K_m, K_i, v_par, Y_p = 0.02, 0.05, 5.0, 0.004, 1.2
    mu_S = mu_m * S_s / (K_m + S_s + (S_s ** 2 / K_i))
    rhs = [mu_S * X_s - inp / V_s * X_s,
           -mu_S * X_s / Y_x - v_par * X_s / Y_p + inp / V_s * (S_in - S_s),
           v_par * X_s - inp / V_s * P_s,
           inp]
    m = ModelDef("continuous", x=[X_s, S_s, P_s, V_s], u=[inp], p=[Y_x, S_in], rhs=rhs)
    return MPCDef(m, n_horizon=n_horizon, t_step=1.0, n_robust=n_robust, collocation_deg=2, collocation_ni=2,
                  lterm=-P_s, mterm=-P_s, rterm_factor=[1.0],
                  x_lb=[0.0, -0.01, 0.0, 0.0], x_ub=[3.7, np.inf, 3.0, np.inf], u_lb=[0.0], u_ub=[0.2],
                  p_scenarios=_scenarios([0.5, 0.4, 0.3], [200.0, 220.0, 180.0]))


# ------------------------------------------------------------------------------------------------
def oscillating_masses(n_horizon=7):
    x = list(sy.symbols("x0:4")); u = [sy.Symbol("u")]
    A = np.array([[0.763, 0.460, 0.115, 0.020], [-0.899, 0.763, 0.420, 0.115],
                  [0.115, 0.020, 0.763, 0.460], [0.420, 0.115, -0.899, 0.763]])
    B = np.array([0.014, 0.063, 0.221, 0.367])
    rhs = [sum(A[i, j] * x[j] for j in range(4)) + B[i] * u[0] for i in range(4)]
    cost = sum(xi ** 2 for xi in x)
    m = ModelDef("discrete", x=x, u=u, rhs=rhs, aux={"cost": cost})
    mx = np.array([4.0, 10.0, 4.0, 10.0])
    return MPCDef(m, n_horizon=n_horizon, t_step=0.5, n_robust=0, lterm=cost, mterm=cost, rterm_factor=[1e-4],
                  x_lb=-mx, x_ub=mx, u_lb=[-0.5], u_ub=[0.5])


def oscillating_masses_dae(n_horizon=7, n_robust=0, uncertain=False):
    """examples/oscillating_masses_discrete_dae/template_model.py:33-75: the same plant written as a discrete DAE —
    x+ = z, 0 = z - A x - B u (algebraic state z = x_next).  uncertain=True (test variant, not a reference example): the input
    gain is scaled by an uncertain parameter kB so that the DAE can be put under a scenario tree."""
    x = list(sy.symbols("x0:4")); u = [sy.Symbol("u")]; z = list(sy.symbols("z0:4"))
    kB = sy.Symbol("kB") if uncertain else 1
    A = np.array([[0.763, 0.460, 0.115, 0.020], [-0.899, 0.763, 0.420, 0.115],
                  [0.115, 0.020, 0.763, 0.460], [0.420, 0.115, -0.899, 0.763]])
    B = np.array([0.014, 0.063, 0.221, 0.367])
    alg = [z[i] - sum(A[i, j] * x[j] for j in range(4)) - kB * B[i] * u[0] for i in range(4)]
    cost = sum(xi ** 2 for xi in x)
    m = ModelDef("discrete", x=x, u=u, z=z, p=[kB] if uncertain else [], rhs=list(z), alg=alg, aux={"cost": cost})
    mx = np.array([4.0, 10.0, 4.0, 10.0])
    kw = dict(p_scenarios=_scenarios([1.0, 1.3, 0.7])) if uncertain else {}
    return MPCDef(m, n_horizon=n_horizon, t_step=0.5, n_robust=n_robust, lterm=cost, mterm=cost, rterm_factor=[1e-4],
                  x_lb=-mx, x_ub=mx, u_lb=[-0.5], u_ub=[0.5], **kw)


# ------------------------------------------------------------------------------------------------
def dip(n_horizon=100, obstacles=({"x": 0.0, "y": 0.6, "r": 0.3},)):
    m0, L1, L2 = 0.6, 0.5, 0.5
    l1, l2 = L1 / 2, L2 / 2
    J1 = (0.2 * l1 ** 2) / 3
    J2 = (0.2 * l2 ** 2) / 3
    m1, m2 = sy.symbols("m1 m2")
    g = 9.80665
    h1 = m0 + m1 + m2
    h2 = m1 * l1 + m2 * L1
    h3 = m2 * l2
    h4 = m1 * l1 ** 2 + m2 * L1 ** 2 + J1
    h5 = m2 * l2 * L1
    h6 = m2 * l2 ** 2 + J2
    h7 = (m1 * l1 + m2 * L1) * g
    h8 = m2 * l2 * g
    pos_set = sy.Symbol("pos_set")
    pos, th0, th1, dpos, dth0, dth1 = sy.symbols("pos th0 th1 dpos dth0 dth1")
    ddpos, ddth0, ddth1 = sy.symbols("ddpos ddth0 ddth1")
    u = sy.Symbol("force")
    rhs = [dpos, dth0, dth1, ddpos, ddth0, ddth1]
    alg = [
        h1 * ddpos + h2 * ddth0 * sy.cos(th0) + h3 * ddth1 * sy.cos(th1)
        - (h2 * dth0 ** 2 * sy.sin(th0) + h3 * dth1 ** 2 * sy.sin(th1) + u),
        h2 * sy.cos(th0) * ddpos + h4 * ddth0 + h5 * sy.cos(th0 - th1) * ddth1
        - (h7 * sy.sin(th0) - h5 * dth1 ** 2 * sy.sin(th0 - th1)),
        h3 * sy.cos(th1) * ddpos + h5 * sy.cos(th0 - th1) * ddth0 + h6 * ddth1
        - (h5 * dth0 ** 2 * sy.sin(th0 - th1) + h8 * sy.sin(th1)),
    ]
    half = sy.Rational(1, 2)
    E_kin_cart = half * m0 * dpos ** 2
    E_kin_p1 = half * m1 * ((dpos + l1 * dth0 * sy.cos(th0)) ** 2 + (l1 * dth0 * sy.sin(th0)) ** 2) + half * J1 * dth0 ** 2
    E_kin_p2 = half * m2 * ((dpos + L1 * dth0 * sy.cos(th0) + l2 * dth1 * sy.cos(th1)) ** 2
                            + (L1 * dth0 * sy.sin(th0) + l2 * dth1 * sy.sin(th1)) ** 2) + half * J2 * dth0 ** 2
    E_kin = E_kin_cart + E_kin_p1 + E_kin_p2
    E_pot = m1 * g * l1 * sy.cos(th0) + m2 * g * (L1 * sy.cos(th0) + l2 * sy.cos(th1))
    n0x, n0y = pos, 0
    n1x, n1y = n0x + L1 * sy.sin(th0), n0y + L1 * sy.cos(th0)
    n2x, n2y = n1x + L2 * sy.sin(th1), n1y + L2 * sy.cos(th1)
    dist = []
    for ob in obstacles:
        for (nx_, ny_) in ((n0x, n0y), (n1x, n1y), (n2x, n2y)):
            dist.append(sy.sqrt((nx_ - ob["x"]) ** 2 + (ny_ - ob["y"]) ** 2) - ob["r"] * 1.05)
    aux = {"E_kin": E_kin, "E_pot": E_pot}
    for i, d_ in enumerate(dist):
        aux[f"obstacle_distance_{i}"] = d_
    aux["tvp"] = pos_set
    m = ModelDef("continuous", x=[pos, th0, th1, dpos, dth0, dth1], u=[u], z=[ddpos, ddth0, ddth1],
                 tvp=[pos_set], p=[m1, m2], rhs=rhs, alg=alg, aux=aux)
    return MPCDef(m, n_horizon=n_horizon, t_step=0.04, n_robust=0, collocation_deg=3, collocation_ni=1,
                  lterm=-E_pot + 10 * (pos - pos_set) ** 2, mterm=E_kin - E_pot, rterm_factor=[0.1],
                  u_lb=[-4], u_ub=[4],
                  nl_cons=[NlCons(f"obstacles_{i}", -d_, ub=0.0) for i, d_ in enumerate(dist)],
                  p_scenarios=_scenarios(0.2 * np.array([1, 0.95, 1.05]), 0.2 * np.array([1, 0.95, 1.05])))


# ------------------------------------------------------------------------------------------------
def industrial_poly(n_horizon=20, n_robust=1):
    R, T_F, E_a, A_tank = 8.314, 25 + 273.15, 8500.0, 65.0
    k_U2, k_U1, w_WF, w_AF = 32.0, 4.0, .333, .667
    m_M_KW, fm_M_KW, m_AWT_KW, fm_AWT_KW, m_AWT, fm_AWT, m_S = 5000.0, 300000.0, 1000.0, 100000.0, 200.0, 20000.0, 39000.0
    c_pW, c_pS, c_pF, c_pR = 4.2, .47, 3.0, 5.0
    k_WS, k_AS, k_PS = 17280.0, 3600.0, 360.0
    alfa = 5 * 20e4 * 3.6
    p_1 = 1.0
    m_W, m_A, m_P, T_R, T_S, Tout_M, T_EK, Tout_AWT, accum_monom, T_adiab = sy.symbols(
        "m_W m_A m_P T_R T_S Tout_M T_EK Tout_AWT accum_monom T_adiab")
    m_dot_f, T_in_M, T_in_EK = sy.symbols("m_dot_f T_in_M T_in_EK")
    delH_R, k_0 = sy.symbols("delH_R k_0")
    U_m = m_P / (m_A + m_P)
    m_ges = m_W + m_A + m_P
    k_R1 = k_0 * sy.exp(-E_a / (R * T_R)) * ((k_U1 * (1 - U_m)) + (k_U2 * U_m))
    k_R2 = k_0 * sy.exp(-E_a / (R * T_EK)) * ((k_U1 * (1 - U_m)) + (k_U2 * U_m))
    k_K = ((m_W / m_ges) * k_WS) + ((m_A / m_ges) * k_AS) + ((m_P / m_ges) * k_PS)
    dot_m_W = m_dot_f * w_WF
    react = k_R1 * (m_A - ((m_A * m_AWT) / (m_W + m_A + m_P)))
    ehe = p_1 * k_R2 * (m_A / m_ges) * m_AWT
    dot_m_A = (m_dot_f * w_AF) - react - ehe
    dot_m_P = react + ehe
    dot_T_R = 1. / (c_pR * m_ges) * ((m_dot_f * c_pF * (T_F - T_R)) - (k_K * A_tank * (T_R - T_S))
                                     - (fm_AWT * c_pR * (T_R - T_EK)) + (delH_R * react))
    rhs = [
        dot_m_W, dot_m_A, dot_m_P, dot_T_R,
        1. / (c_pS * m_S) * ((k_K * A_tank * (T_R - T_S)) - (k_K * A_tank * (T_S - Tout_M))),
        1. / (c_pW * m_M_KW) * ((fm_M_KW * c_pW * (T_in_M - Tout_M)) + (k_K * A_tank * (T_S - Tout_M))),
        1. / (c_pR * m_AWT) * ((fm_AWT * c_pR * (T_R - T_EK)) - (alfa * (T_EK - Tout_AWT)) + (ehe * delH_R)),
        1. / (c_pW * m_AWT_KW) * ((fm_AWT_KW * c_pW * (T_in_EK - Tout_AWT)) - (alfa * (Tout_AWT - T_EK))),
        m_dot_f,
        delH_R / (m_ges * c_pR) * dot_m_A - (dot_m_A + dot_m_W + dot_m_P) * (m_A * delH_R / (m_ges * m_ges * c_pR)) + dot_T_R,
    ]
    m = ModelDef("continuous", x=[m_W, m_A, m_P, T_R, T_S, Tout_M, T_EK, Tout_AWT, accum_monom, T_adiab],
                 u=[m_dot_f, T_in_M, T_in_EK], p=[delH_R, k_0], rhs=rhs)
    tr = 2.0
    inf = np.inf
    nl = []
    x_ub = [inf, inf, inf, 363.15 + tr, 400.0, 400.0, 400.0, 400.0, 30000.0, 382.15]
    if n_robust == 0:
        nl = [NlCons("T_R_UB", T_R, ub=363.15 + tr, soft=True, penalty=1e4)]
    return MPCDef(m, n_horizon=n_horizon, t_step=50.0 / 3600.0, n_robust=n_robust, collocation_deg=2, collocation_ni=1,
                  lterm=-m_P, mterm=-m_P, rterm_factor=[0.002, 0.004, 0.002],
                  x_scaling=[10, 10, 10, 1, 1, 1, 1, 1, 10, 1], u_scaling=[100, 1, 1],
                  x_lb=[0.0, 0.0, 26.0, 363.15 - tr, 298.0, 298.0, 288.0, 288.0, 0.0, -inf], x_ub=x_ub,
                  u_lb=[0.0, 333.15, 333.15], u_ub=[3.0e4, 373.15, 373.15], nl_cons=nl,
                  p_scenarios=_scenarios([950.0, 950.0 * 1.30, 950.0 * 0.70], [7.0, 7.0 * 1.30, 7.0 * 0.70]))


# ------------------------------------------------------------------------------------------------
def kite(w_ref=10.0, E_0_nom=6.0, h_min=100.0, n_horizon=80):
    """Towing kite: thrust maximisation through an aux expression that reads the uncertain parameters, soft height
    constraint; n_robust = 0 (only the nominal combination of the three wind speeds enters the NLP)."""
    L_tether, A, rho, beta, c_tilde = 400.0, 300.0, 1.0, 0.0, 0.028
    theta, phi, psi, u_tilde, E_0, v_0 = sy.symbols("theta phi psi u_tilde E_0 v_0")
    E = E_0 - c_tilde * u_tilde ** 2
    v_a = v_0 * E * sy.cos(theta)
    P_D = (rho * v_0 ** 2) / 2.0
    T_F = (P_D * A * sy.cos(theta) ** 2 * (E + 1.0) * sy.sqrt(E ** 2 + 1.0)) * (sy.cos(theta) * np.cos(beta) + sy.sin(theta) * np.sin(beta) * sy.sin(phi))
    height = L_tether * sy.sin(theta) * sy.cos(phi)
    dphi = -v_a / (L_tether * sy.sin(theta)) * sy.sin(psi)
    rhs = [v_a / L_tether * (sy.cos(psi) - sy.tan(theta) / E), dphi, v_a / L_tether * u_tilde + dphi * sy.cos(theta)]
    m = ModelDef("continuous", x=[theta, phi, psi], u=[u_tilde], p=[E_0, v_0], rhs=rhs,
                 aux={"E_0": E_0, "v_0": v_0, "T_F": T_F, "height_kite": height})
    return MPCDef(m, n_horizon=n_horizon, t_step=0.15, n_robust=0, collocation_deg=2, collocation_ni=1,
                  lterm=-T_F / 1e4, mterm=sy.Integer(0), rterm_factor=[0.5],
                  x_lb=[0.0, -0.5 * np.pi, -np.pi], x_ub=[0.5 * np.pi, 0.5 * np.pi, np.pi], u_lb=[-10], u_ub=[10],
                  nl_cons=[NlCons("height_kite", -height, ub=-h_min, soft=True, penalty=1e3, max_violation=10)],
                  p_scenarios=_scenarios([E_0_nom], [w_ref, w_ref * 0.8, w_ref * 1.2]))


def kinematic_bicycle(n_horizon=10):
    lf, lr = 0.3, 0.3
    X_p, Y_p, Psi, V, Delta, Acc = sy.symbols("X_p Y_p Psi V Delta Acc")
    Beta = sy.atan((lr / (lr + lf)) * sy.tan(Delta))
    rhs = [V * sy.cos(Psi + Beta), V * sy.sin(Psi + Beta), (V / lr) * sy.sin(Beta), Acc]
    m = ModelDef("continuous", x=[X_p, Y_p, Psi, V], u=[Delta, Acc], rhs=rhs)
    return MPCDef(m, n_horizon=n_horizon, t_step=0.05, n_robust=0, collocation_deg=2, collocation_ni=1,
                  lterm=sy.Integer(0), mterm=(Y_p - 2) ** 2 + (X_p - 3) ** 2 + Psi ** 2, rterm_factor=[1.0, 1e-3],
                  x_lb=[-50, -50, -np.pi / 2, -5], x_ub=[50, 50, np.pi / 2, 5], u_lb=[-5, -5], u_ub=[5, 5])


def dynamic_bicycle(n_horizon=10):
    mass, I_z, lf, lr = 5.692, 0.204, 0.178, 0.147
    D_f, D_r, C_f, C_r, B_f, B_r = 134.585, 159.919, 0.085, 0.133, 9.242, 17.716
    c_m1, c_m2, c_m3, c_m4 = 20, 6.92 * 1e-7, 3.99, 0.67
    X_p, Y_p, Psi, V_x, V_y, W, Delta, d = sy.symbols("X_p Y_p Psi V_x V_y W Delta d")
    alpha_f = -sy.atan2(W * lf + V_y, V_x) + Delta
    alpha_r = sy.atan2(W * lr - V_y, V_x)
    F_f_y = D_f * sy.sin(C_f * sy.atan(B_f * alpha_f))
    F_r_y = D_r * sy.sin(C_r * sy.atan(B_r * alpha_r))
    F_x = (c_m1 - c_m2 * V_x) * d - c_m4 * V_x ** 2 - c_m3
    rhs = [V_x * sy.cos(Psi) - V_y * sy.sin(Psi), V_x * sy.sin(Psi) + V_y * sy.cos(Psi), W,
           (1 / mass) * (F_x - F_f_y * sy.sin(Delta) + mass * V_y * W),
           (1 / mass) * (F_r_y + F_f_y * sy.cos(Delta) - mass * V_x * W),
           (1 / I_z) * (F_f_y * lf * sy.cos(Delta) - lf * F_x * sy.sin(Delta) - lr * F_r_y)]
    m = ModelDef("continuous", x=[X_p, Y_p, Psi, V_x, V_y, W], u=[Delta, d], rhs=rhs, aux={"Vel": sy.sqrt(V_x ** 2 + V_y ** 2)})
    cost = (Y_p - 1) ** 2
    return MPCDef(m, n_horizon=n_horizon, t_step=0.1, n_robust=0, collocation_deg=2, collocation_ni=1,
                  lterm=cost, mterm=cost, rterm_factor=[1e-3, 1e-3],
                  x_lb=[-50000, -2, -0.78, 0.1, -1, -0.2], x_ub=[50000, 2, 0.78, 5, 1, 0.2], u_lb=[-2, 0], u_ub=[2, 1])


# ------------------------------------------------------------------------------------------------
def rotating_masses_mhe(n_horizon=10):
    """examples/rotating_oscillating_masses_mhe_mpc/template_{model,mhe}.py: 8 states, 2 inputs, 5 noisy measurements,
    Theta_1 estimated, default objective with P_x = 1e-4 I, P_v = tvp, P_p = p, two hard constraints on Theta_1 at all
    collocation points."""
    from .mhe import MHEDef
    phi = sy.symbols("phi_1 phi_2 phi_3")
    dphi = sy.symbols("dphi_0 dphi_1 dphi_2")
    phi_m = sy.symbols("phi_m_0 phi_m_1")
    u = sy.symbols("phi_m_set_0 phi_m_set_1")
    phi_2_set = sy.Symbol("phi_2_set")
    P_v = sy.symbols("P_v_0:25")                      # 5 x 5, column-major like CasADi's vec (symmetric here)
    P_p = sy.Symbol("P_p")
    T1, T2, T3 = sy.symbols("Theta_1 Theta_2 Theta_3")
    c = np.array([2.697, 2.66, 3.05, 2.86]) * 1e-3
    d = np.array([6.78, 8.01, 8.82]) * 1e-5
    tau = 1e-2
    rhs = [dphi[0], dphi[1], dphi[2],
           -c[0] / T1 * (phi[0] - phi_m[0]) - c[1] / T1 * (phi[0] - phi[1]) - d[0] / T1 * dphi[0],
           -c[1] / T2 * (phi[1] - phi[0]) - c[2] / T2 * (phi[1] - phi[2]) - d[1] / T2 * dphi[1],
           -c[2] / T3 * (phi[2] - phi[1]) - c[3] / T3 * (phi[2] - phi_m[1]) - d[2] / T3 * dphi[2],
           1 / tau * (u[0] - phi_m[0]), 1 / tau * (u[1] - phi_m[1])]
    x = list(phi) + list(dphi) + list(phi_m)
    m = ModelDef("continuous", x=x, u=list(u), tvp=[phi_2_set] + list(P_v), p=[P_p, T1, T2, T3], rhs=rhs)
    v = list(sy.symbols("v_0:5"))
    x_prev = list(sy.symbols("xprev_0:8")); pe_prev = [sy.Symbol("Theta_1_prev")]
    Pv = sy.Matrix(5, 5, lambda i, j: P_v[j * 5 + i])
    vv = sy.Matrix(v)
    stage = (vv.T * Pv * vv)[0, 0]
    dx = sy.Matrix([a - b for a, b in zip(x, x_prev)])
    arrival = 1e-4 * (dx.T * dx)[0, 0] + (T1 - pe_prev[0]) * P_p * (T1 - pe_prev[0])
    inf = np.inf
    return MHEDef(m, n_horizon=n_horizon, t_step=0.1, p_est=[T1], meas=list(phi) + list(u), stage_cost=stage, arrival_cost=arrival,
                  v=v, x_prev=x_prev, p_est_prev=pe_prev, nl_cons_check_colloc_points=True,
                  nl_cons=[NlCons("p_est_lb", -T1 + 1e-5, ub=0.0), NlCons("p_est_ub", T1 - 1e-3, ub=0.0)],
                  x_lb=[-inf, -inf, -inf, -6, -6, -6, -inf, -inf], x_ub=[inf, inf, inf, 6, 6, 6, inf, inf],
                  u_lb=[-5, -5], u_ub=[5, 5])
