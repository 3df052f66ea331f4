"""ctypes mirror of `struct DmbOptions` (do_mpc_b200/csrc/ipm_core.cuh) and mapping of ipopt.* option names."""
import ctypes

_D = ["tol", "dual_inf_tol", "constr_viol_tol", "compl_inf_tol",
      "acceptable_tol", "acceptable_dual_inf_tol", "acceptable_constr_viol_tol", "acceptable_compl_inf_tol",
      "s_max", "mu_init", "mu_min", "kappa_mu", "theta_mu", "kappa_eps", "tau_min",
      "bound_push", "bound_frac", "bound_relax_factor", "constr_mult_init_max", "bound_mult_init_val",
      "kappa_sigma", "kappa_d",
      "alpha_red_factor", "eta_phi", "theta_max_fact", "theta_min_fact", "s_phi", "s_theta", "gamma_phi", "gamma_theta",
      "alpha_min_frac", "delta", "kappa_soc",
      "delta_w_first", "delta_w_min", "delta_w_max", "kappa_w_plus_first", "kappa_w_plus", "kappa_w_minus",
      "nlp_scaling_max_gradient", "nlp_scaling_min_value", "kappa_resto", "residual_ratio_max",
      "ws_bound_push", "ws_bound_frac", "ws_mult_bound_push", "ws_mult_init_max"]
_I = ["max_iter", "acceptable_iter", "max_soc", "mimic_dummy_reg", "resto_max_iter", "warm_start", "max_refine", "pad_"]


class DmbOptions(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in _D] + [(n, ctypes.c_int) for n in _I]


# IPOPT option name -> DmbOptions field (reference passes `settings.nlpsol_opts` straight to IPOPT, _mpc.py:1328)
IPOPT_MAP = {
    "ipopt.tol": "tol", "ipopt.max_iter": "max_iter", "ipopt.dual_inf_tol": "dual_inf_tol",
    "ipopt.constr_viol_tol": "constr_viol_tol", "ipopt.compl_inf_tol": "compl_inf_tol",
    "ipopt.acceptable_tol": "acceptable_tol", "ipopt.acceptable_iter": "acceptable_iter",
    "ipopt.mu_init": "mu_init", "ipopt.mu_min": "mu_min", "ipopt.bound_push": "bound_push",
    "ipopt.bound_frac": "bound_frac", "ipopt.bound_relax_factor": "bound_relax_factor",
    "ipopt.max_soc": "max_soc", "ipopt.nlp_scaling_max_gradient": "nlp_scaling_max_gradient",
    "ipopt.kappa_d": "kappa_d", "ipopt.tau_min": "tau_min", "ipopt.s_max": "s_max",
    "ipopt.warm_start_bound_push": "ws_bound_push", "ipopt.warm_start_bound_frac": "ws_bound_frac",
    "ipopt.warm_start_slack_bound_push": "ws_bound_push", "ipopt.warm_start_slack_bound_frac": "ws_bound_frac",
    "ipopt.warm_start_mult_bound_push": "ws_mult_bound_push", "ipopt.warm_start_mult_init_max": "ws_mult_init_max",
}
IGNORED_PREFIXES = ("ipopt.print_level", "ipopt.sb", "print_time", "ipopt.linear_solver", "expand", "ipopt.hsllib",
                    "ipopt.print_", "ipopt.output_file", "ipopt.file_print_level", "verbose",
                    "ipopt.ma", "ipopt.mumps", "record_time", "ipopt.timing_statistics", "ipopt.print_timing_statistics")


def apply_nlpsol_opts(opts_struct: DmbOptions, nlpsol_opts: dict):
    import warnings
    for k, v in (nlpsol_opts or {}).items():
        if k == "ipopt.warm_start_init_point":
            opts_struct.warm_start = int(str(v).lower() in ("yes", "true", "1"))
        elif k in IPOPT_MAP:
            f = IPOPT_MAP[k]
            setattr(opts_struct, f, int(v) if f in _I else float(v))
        elif k.startswith("dmb."):
            f = k[4:]
            setattr(opts_struct, f, int(v) if f in _I else float(v))
        elif k.startswith(IGNORED_PREFIXES):
            continue
        else:
            warnings.warn(f"solver option '{k}' has no counterpart in the B200 interior-point solver and is ignored")
    return opts_struct
