"""Rotating oscillating masses with a moving-horizon estimator (same problem as the reference's
examples/rotating_oscillating_masses_mhe_mpc: template_model.py, template_mhe.py)."""
import numpy as np

import do_mpc_b200 as do_mpc
from do_mpc_b200.casadi_api import vertcat


def masses_model(symvar_type="SX"):
    model = do_mpc.model.Model("continuous", symvar_type)
    phi_1 = model.set_variable(var_type="_x", var_name="phi_1")
    phi_2 = model.set_variable(var_type="_x", var_name="phi_2")
    phi_3 = model.set_variable(var_type="_x", var_name="phi_3")
    phi = vertcat(phi_1, phi_2, phi_3)
    dphi = model.set_variable(var_type="_x", var_name="dphi", shape=(3, 1))
    phi_m_set = model.set_variable(var_type="_u", var_name="phi_m_set", shape=(2, 1))
    phi_m = model.set_variable(var_type="_x", var_name="phi_m", shape=(2, 1))
    model.set_variable(var_type="_tvp", var_name="phi_2_set")
    model.set_variable(var_type="_p", var_name="P_p")
    model.set_variable(var_type="_tvp", var_name="P_v", shape=(5, 5))
    model.set_meas("phi_1_meas", phi)
    model.set_meas("phi_m_set_meas", phi_m_set)
    Theta_1 = model.set_variable("parameter", "Theta_1")
    Theta_2 = model.set_variable("parameter", "Theta_2")
    Theta_3 = model.set_variable("parameter", "Theta_3")
    c = np.array([2.697, 2.66, 3.05, 2.86]) * 1e-3
    d = np.array([6.78, 8.01, 8.82]) * 1e-5
    model.set_rhs("phi_1", dphi[0]); model.set_rhs("phi_2", dphi[1]); model.set_rhs("phi_3", dphi[2])
    dphi_next = vertcat(
        -c[0] / Theta_1 * (phi[0] - phi_m[0]) - c[1] / Theta_1 * (phi[0] - phi[1]) - d[0] / Theta_1 * dphi[0],
        -c[1] / Theta_2 * (phi[1] - phi[0]) - c[2] / Theta_2 * (phi[1] - phi[2]) - d[1] / Theta_2 * dphi[1],
        -c[2] / Theta_3 * (phi[2] - phi[1]) - c[3] / Theta_3 * (phi[2] - phi_m[1]) - d[2] / Theta_3 * dphi[2])
    model.set_rhs("dphi", dphi_next)
    tau = 1e-2
    model.set_rhs("phi_m", 1 / tau * (phi_m_set - phi_m))
    model.setup()
    return model


def masses_mhe(model=None, setup=True):
    model = model or masses_model()
    mhe = do_mpc.estimator.MHE(model, ["Theta_1"])
    s = mhe.settings
    s.n_horizon, s.t_step, s.store_full_solution, s.nl_cons_check_colloc_points = 10, 0.1, True, True
    mhe.set_default_objective(1e-4 * np.eye(8), model.tvp["P_v"], model.p["P_p"])
    tvp_template = mhe.get_tvp_template()
    tvp_template["_tvp", :, "P_v"] = np.diag(np.array([1, 1, 1, 20, 20]))
    mhe.set_tvp_fun(lambda t_now: tvp_template)
    p_template = mhe.get_p_template()

    def p_fun(t_now):
        p_template["Theta_2"] = 2.25e-4
        p_template["Theta_3"] = 2.25e-4
        p_template["P_p"] = np.eye(1)
        return p_template
    mhe.set_p_fun(p_fun)
    y_template = mhe.get_y_template()

    def y_fun(t_now):
        n_steps = min(mhe.data._y.shape[0], mhe.settings.n_horizon)
        for k in range(-n_steps, 0):
            y_template["y_meas", k] = mhe.data._y[k]
        return y_template
    mhe.set_y_fun(y_fun)
    mhe.bounds["lower", "_u", "phi_m_set"] = -5
    mhe.bounds["upper", "_u", "phi_m_set"] = 5
    mhe.bounds["lower", "_x", "dphi"] = -6
    mhe.bounds["upper", "_x", "dphi"] = 6
    mhe.set_nl_cons("p_est_lb", -mhe._p_est["Theta_1"] + 1e-5, 0)
    mhe.set_nl_cons("p_est_ub", mhe._p_est["Theta_1"] - 1e-3, 0)
    if setup:
        mhe.setup()
    return mhe


def masses_mpc(model=None, setup=True):
    """examples/rotating_oscillating_masses_mhe_mpc/template_mpc.py"""
    from do_mpc_b200.casadi_api import DM, vertsplit
    model = model or masses_model()
    mpc = do_mpc.controller.MPC(model)
    s = mpc.settings
    s.n_robust, s.n_horizon, s.t_step, s.store_full_solution = 0, 20, 0.1, True
    _x, _tvp = model["x", "tvp"]
    mpc.set_objective(mterm=DM(1), lterm=(_x["phi_2"] - _tvp["phi_2_set"]) ** 2)
    mpc.set_rterm(phi_m_set=1e-2)

    def random_setpoint(tvp0):
        tvp_next = (0.5 - np.random.rand()) * np.pi
        switch = np.random.rand() >= 0.95
        return (1 - switch) * tvp0 + switch * tvp_next
    np.random.seed(999)
    tvp_traj = [np.array([0])]
    for i in range(400):
        tvp_traj.append(random_setpoint(tvp_traj[i]))
    tvp_traj = np.concatenate(tvp_traj)
    tvp_template = mpc.get_tvp_template()

    def tvp_fun(t_now):
        ind = int(t_now / mpc.settings.t_step)
        tvp_template["_tvp", :-1] = vertsplit(tvp_traj[ind:ind + mpc.settings.n_horizon])
        return tvp_template
    mpc.set_tvp_fun(tvp_fun)
    mpc.set_uncertainty_values(Theta_1=2.25 * 1e-4 * np.array([1., 1.1]), Theta_2=2.25 * 1e-4 * np.array([1.]), Theta_3=2.25 * 1e-4 * np.array([1.]))
    mpc.bounds["lower", "_u", "phi_m_set"] = -5
    mpc.bounds["upper", "_u", "phi_m_set"] = 5
    if setup:
        mpc.setup()
    return mpc


def masses_simulator(model=None):
    """examples/rotating_oscillating_masses_mhe_mpc/template_simulator.py"""
    model = model or masses_model()
    simulator = do_mpc.simulator.Simulator(model)
    simulator.set_param(t_step=0.1)
    p_template = simulator.get_p_template()

    def p_fun(t_now):
        p_template["Theta_1"] = 2.25e-4
        p_template["Theta_2"] = 2.25e-4
        p_template["Theta_3"] = 2.25e-4
        return p_template
    simulator.set_p_fun(p_fun)
    tvp_template = simulator.get_tvp_template()
    simulator.set_tvp_fun(lambda t_now: tvp_template)
    simulator.setup()
    return simulator
