"""Double inverted pendulum on a cart, DAE formulation with obstacle-avoidance constraints (same problem as the
reference's examples/double_inverted_pendulum; BASELINE config #4 uses N=50)."""
import numpy as np

import do_mpc_b200 as do_mpc
from do_mpc_b200.casadi_api import sin, cos, sqrt, vertcat


def dip_model(obstacles=({"x": 0.0, "y": 0.6, "r": 0.3},), symvar_type="SX"):
    model = do_mpc.model.Model("continuous", symvar_type)
    m0, L1, L2, g = 0.6, 0.5, 0.5, 9.80665
    l1, l2 = L1 / 2, L2 / 2
    J1, J2 = (0.2 * l1 ** 2) / 3, (0.2 * l2 ** 2) / 3
    m1 = model.set_variable("_p", "m1"); m2 = model.set_variable("_p", "m2")
    h1 = m0 + m1 + m2; h2 = m1 * l1 + m2 * L1; h3 = m2 * l2
    h4 = m1 * l1 ** 2 + m2 * L1 ** 2 + J1; h5 = m2 * l2 * L1; h6 = m2 * l2 ** 2 + J2
    h7 = (m1 * l1 + m2 * L1) * g; h8 = m2 * l2 * g
    pos_set = model.set_variable("_tvp", "pos_set")
    pos = model.set_variable("_x", "pos"); theta = model.set_variable("_x", "theta", (2, 1))
    dpos = model.set_variable("_x", "dpos"); dtheta = model.set_variable("_x", "dtheta", (2, 1))
    ddpos = model.set_variable("_z", "ddpos"); ddtheta = model.set_variable("_z", "ddtheta", (2, 1))
    u = model.set_variable("_u", "force")
    model.set_rhs("pos", dpos); model.set_rhs("theta", dtheta); model.set_rhs("dpos", ddpos); model.set_rhs("dtheta", ddtheta)
    model.set_alg("euler_lagrange", vertcat(
        h1 * ddpos + h2 * ddtheta[0] * cos(theta[0]) + h3 * ddtheta[1] * cos(theta[1])
        - (h2 * dtheta[0] ** 2 * sin(theta[0]) + h3 * dtheta[1] ** 2 * sin(theta[1]) + u),
        h2 * cos(theta[0]) * ddpos + h4 * ddtheta[0] + h5 * cos(theta[0] - theta[1]) * ddtheta[1]
        - (h7 * sin(theta[0]) - h5 * dtheta[1] ** 2 * sin(theta[0] - theta[1])),
        h3 * cos(theta[1]) * ddpos + h5 * cos(theta[0] - theta[1]) * ddtheta[0] + h6 * ddtheta[1]
        - (h5 * dtheta[0] ** 2 * sin(theta[0] - theta[1]) + h8 * sin(theta[1]))))
    E_kin = (0.5 * m0 * dpos ** 2
             + 0.5 * m1 * ((dpos + l1 * dtheta[0] * cos(theta[0])) ** 2 + (l1 * dtheta[0] * sin(theta[0])) ** 2) + 0.5 * J1 * dtheta[0] ** 2
             + 0.5 * m2 * ((dpos + L1 * dtheta[0] * cos(theta[0]) + l2 * dtheta[1] * cos(theta[1])) ** 2
                           + (L1 * dtheta[0] * sin(theta[0]) + l2 * dtheta[1] * sin(theta[1])) ** 2) + 0.5 * J2 * dtheta[0] ** 2)
    E_pot = m1 * g * l1 * cos(theta[0]) + m2 * g * (L1 * cos(theta[0]) + l2 * cos(theta[1]))
    model.set_expression("E_kin", E_kin); model.set_expression("E_pot", E_pot)
    n0x, n0y = pos, 0.0
    n1x, n1y = n0x + L1 * sin(theta[0]), n0y + L1 * cos(theta[0])
    n2x, n2y = n1x + L2 * sin(theta[1]), n1y + L2 * cos(theta[1])
    dist = []
    for ob in obstacles:
        for nx_, ny_ in ((n0x, n0y), (n1x, n1y), (n2x, n2y)):
            dist.append(sqrt((nx_ - ob["x"]) ** 2 + (ny_ - ob["y"]) ** 2) - ob["r"] * 1.05)
    model.set_expression("obstacle_distance", vertcat(*dist))
    model.set_expression("tvp", pos_set)
    model.setup()
    return model


def dip_mpc(model=None, n_horizon=100, setup=True):
    model = model or dip_model()
    mpc = do_mpc.controller.MPC(model)
    s = mpc.settings
    s.n_horizon, s.n_robust, s.open_loop, s.t_step = n_horizon, 0, 0, 0.04
    s.state_discretization, s.collocation_type, s.collocation_deg, s.collocation_ni = "collocation", "radau", 3, 1
    s.store_full_solution = True
    mpc.set_objective(mterm=model.aux["E_kin"] - model.aux["E_pot"],
                      lterm=-model.aux["E_pot"] + 10 * (model.x["pos"] - model.tvp["pos_set"]) ** 2)
    mpc.set_rterm(force=0.1)
    mpc.bounds["lower", "_u", "force"] = -4; mpc.bounds["upper", "_u", "force"] = 4
    mpc.set_nl_cons("obstacles", -model.aux["obstacle_distance"], 0)
    mpc.set_uncertainty_values(m1=0.2 * np.array([1, 0.95, 1.05]), m2=0.2 * np.array([1, 0.95, 1.05]))
    tvp_template = mpc.get_tvp_template()
    ind_switch = 4 // s.t_step

    def tvp_fun(t_ind):
        ind = t_ind // s.t_step
        tvp_template["_tvp", :, "pos_set"] = -0.8 if ind <= ind_switch else 0.8
        return tvp_template
    mpc.set_tvp_fun(tvp_fun)
    if setup:
        mpc.setup()
    return mpc
