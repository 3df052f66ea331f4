"""Moving horizon estimator with the do_mpc.estimator.MHE API on top of the batched CUDA NLP solver.

API mirror of reference do_mpc/estimator/_mhe.py (class MHE :37; __init__ :150, p_est0 :235, set_objective :489,
set_default_objective :602, get_p_template :717, set_p_fun :733, get_y_template :745, set_y_fun :791, set_initial_guess :864,
setup :885, make_step :896, _update_bounds :995, _prepare_nlp :1030, _create_nlp :1213) and of the shared base
do_mpc/optimizer.py (bounds, scaling, set_nl_cons, tvp functions, solve).

Same boundary as the controller: `self.S` replaces `castools.nlpsol('S', 'ipopt', nlp, opts)` (_mhe.py:1250-1251); `opt_x`,
`opt_p`, `g` keep the reference's flat layouts.  The NLP differs from the controller's in three ways, all expressed through the
stage recipe (do_mpc_b200/recipe.py) rather than new kernel code paths:
  * the root state is FREE and carries the arrival cost (kernel: DMB_NFREE0 / gen_arrival*, csrc/ipm_core.cuh);
  * the estimated parameters `_p_est` are one global variable: they are lifted to extra state components with the trivial
    dynamics p+ = p (linear consistency constraints do not change the Newton iterates of a point that satisfies them);
  * the measurement equations  h(x_{k+1}, u_k, z_k, tvp_k, p, v_k) = y_meas_k  join the square collocation block of the stage,
    with the measurement noise v_k as its unknowns (requires noise on every measurement: dh/dv = I keeps the block regular).
"""
from __future__ import annotations

import warnings
from dataclasses import asdict, dataclass, field
from typing import Callable, List

import numpy as np

from .controller import _Indexed
from .data import Data
from .model import Model
from .recipe import Recipe, collocation_constants
from .sym import expr as E
from .sym.struct import NumStruct, SymStruct, entry
from .sym.sx import DM, SX, Function, sum1, vec, vertcat


@dataclass
class MHESettings:
    """reference do_mpc/estimator/_estimatorsettings.py (MHESettings)."""
    n_horizon: int = None
    t_step: float = None
    meas_from_data: bool = True
    state_discretization: str = "collocation"
    collocation_type: str = "radau"
    collocation_deg: int = 2
    collocation_ni: int = 1
    nl_cons_check_colloc_points: bool = False
    nl_cons_single_slack: bool = False
    cons_check_colloc_points: bool = True
    store_full_solution: bool = False
    store_lagr_multiplier: bool = True
    store_solver_stats: List[str] = field(default_factory=lambda: ["success", "t_wall_total"])
    nlpsol_opts: dict = field(default_factory=dict)

    def check_for_mandatory_settings(self):
        if self.n_horizon is None:
            raise ValueError("n_horizon must be set")
        if self.t_step is None:
            raise ValueError("t_step must be set")

    def supress_ipopt_output(self):
        self.nlpsol_opts.update({"ipopt.print_level": 0, "ipopt.sb": "yes", "print_time": 0})

    def set_linear_solver(self, solver_name: str = "MA27"):
        self.nlpsol_opts.update({"ipopt.linear_solver": solver_name})


class MHE:
    def __init__(self, model: Model, p_est_list: list = []):
        assert model.flags["setup"] == True, "Model for MHE was not setup. After the complete model creation call model.setup()."
        self.model = m = model
        self.settings = self._settings = MHESettings()
        self.data = Data(model)
        self.data.dtype = "MHE"
        # estimated / set parameters (reference :163-193): the union of both are the model parameters, model order kept
        names = list(m._p.keys())
        self._p_est = SymStruct([entry("default", shape=(0, 1))] + [entry(n, shape=m._p[n].shape) for n in names if n in p_est_list])
        self._p_set = SymStruct([entry(n, shape=m._p[n].shape) for n in names if n not in p_est_list])
        subs = [self._p_est[n] if n in p_est_list else self._p_set[n] for n in names]
        p_cat = vertcat(*[vec(s_) for s_ in subs]) if subs else SX.zeros(0, 1)
        self._p_cat_fun = Function("p_cat_fun", [self._p_est.cat, self._p_set.cat], [p_cat])
        self.n_p_est, self.n_p_set = self._p_est.size, self._p_set.size
        self._p_est_scaling, self._p_set_scaling = self._p_est(1.0), self._p_set(1.0)
        self._p_est_lb, self._p_est_ub = self._p_est(-np.inf), self._p_est(np.inf)
        self._p_est0 = self._p_est(0.0)
        self._y_meas = m._y
        self._x = m._x
        self._x_prev = SymStruct([entry(n, shape=m._x[n].shape) for n in m._x.keys()])      # fresh symbols with the same layout
        self._p_est_prev = SymStruct([entry("default", shape=(0, 1))] + [entry(n, shape=m._p[n].shape) for n in names if n in p_est_list])
        self._p = m._p
        self._w, self._v = m._w, m._v
        self.flags = {"setup": False, "prepare_nlp": False, "set_tvp_fun": False, "set_p_fun": False, "set_y_fun": False,
                      "set_objective": False, "set_initial_guess": False, "initial_run": False, "MINLP": bool(m.integer)}
        if self.flags["MINLP"]:
            raise NotImplementedError("integer inputs (MINLP / bonmin) are outside the GPU interior-point path")
        self._x_lb, self._x_ub = m._x(-np.inf), m._x(np.inf)
        self._u_lb, self._u_ub = m._u(-np.inf), m._u(np.inf)
        self._z_lb, self._z_ub = m._z(-np.inf), m._z(np.inf)
        self._x_scaling, self._u_scaling, self._z_scaling = m._x(1.0), m._u(1.0), m._z(1.0)
        self._x0, self._u0, self._z0, self._t0 = m._x(0.0), m._u(0.0), m._z(0.0), np.array([0.0])
        self.nl_cons_list, self.slack_vars_list = [], []
        self.bounds = _Indexed(self._get_bounds, self._set_bounds)
        self.scaling = _Indexed(self._get_scaling, self._set_scaling)
        self.S = None
        self.solver_stats = {}

    # ---- legacy set_param ------------------------------------------------------------------------------------
    def set_param(self, **kwargs):
        assert self.flags["setup"] == False, "Setting parameters after setup is prohibited."
        for k, v in kwargs.items():
            if hasattr(self.settings, k):
                setattr(self.settings, k, v)
            else:
                print(f"Warning: Key {k} does not exist for MHE.")

    # ---- bounds / scaling (optimizer.py:268-420 with the additional var_type '_p_est') ---------------------------------------
    def _bstruct(self, ind):
        assert isinstance(ind, tuple) and len(ind) >= 3, "Power index must include bound_type, var_type, var_name (as a tuple)."
        bt, vt = ind[0], ind[1]
        assert bt in ("lower", "upper"), f"Invalid power index {bt} for bound_type. Must be from (lower, upper)."
        assert vt in ("_x", "_u", "_z", "_p_est"), f"Invalid power index {vt} for var_type. Must be from (_x, _u, _z, _p_est)."
        s = getattr(self, f"{vt}_{'lb' if bt == 'lower' else 'ub'}")
        assert ind[2] in s.keys(), f"Calling .bounds with {ind} is not valid. Possible keys are {s.keys()}."
        return s, ind[2:]

    def _get_bounds(self, ind):
        s, k = self._bstruct(ind); return s[k]

    def _set_bounds(self, ind, val):
        s, k = self._bstruct(ind); s[k] = val
        if self.flags["prepare_nlp"]:
            self._update_bounds()

    def _sstruct(self, ind):
        assert isinstance(ind, tuple) and len(ind) >= 2, "Power index must include var_type, var_name (as a tuple)."
        vt = ind[0]
        assert vt in ("_x", "_u", "_z", "_p_est"), f"Invalid power index {vt} for var_type."
        return getattr(self, f"{vt}_scaling"), ind[1:]

    def _get_scaling(self, ind):
        s, k = self._sstruct(ind); return s[k]

    def _set_scaling(self, ind, val):
        assert self.flags["setup"] == False, "Scaling can only be set before setup (kernels are generated with it)."
        s, k = self._sstruct(ind); s[k] = val

    # ---- objective (reference :489-716) --------------------------------------------------------------------------------------
    def set_objective(self, stage_cost, arrival_cost) -> None:
        stage_cost = stage_cost if isinstance(stage_cost, SX) else SX(stage_cost)
        arrival_cost = arrival_cost if isinstance(arrival_cost, SX) else SX(arrival_cost)
        assert stage_cost.shape == (1, 1), f"stage_cost must have shape=(1,1). You have {stage_cost.shape}"
        assert arrival_cost.shape == (1, 1), f"arrival_cost must have shape=(1,1). You have {arrival_cost.shape}"
        assert self.flags["setup"] == False, "Cannot call .set_objective after .setup."
        m = self.model
        from .sym.sx import substitute
        p_cat = self._p_cat_fun(self._p_est.cat, self._p_set.cat)
        arrival_cost = substitute(arrival_cost, m._p.cat, p_cat)
        names_e = [n for n in self._p_est.keys() if n != "default"]
        if names_e:
            stage_cost = substitute(stage_cost, vertcat(*[vec(self._p_est[n]) for n in names_e]),
                                    vertcat(*[vec(m._p[n]) for n in names_e]))
        names_s = list(self._p_set.keys())
        if names_s:
            stage_cost = substitute(stage_cost, vertcat(*[vec(self._p_set[n]) for n in names_s]),
                                    vertcat(*[vec(m._p[n]) for n in names_s]))
        try:
            self.stage_cost_fun = Function("stage_cost_fun", [self._w.cat, self._v.cat, m._tvp.cat, m._p.cat], [stage_cost])
        except Exception:
            raise Exception("objective cost equation must be solely depending on w, v, p and tvp.")
        try:
            self.arrival_cost_fun = Function("arrival_cost_fun", [self._x.cat, self._x_prev.cat, self._p_est.cat, self._p_est_prev.cat,
                                                                  self._p_set.cat], [arrival_cost])
        except Exception:
            raise Exception("Arrival cost equation must be solely depending on x_0, x_prev, p_0, p_prev, p_set")
        self.flags["set_objective"] = True

    def set_default_objective(self, P_x, P_v=None, P_p=None, P_w=None) -> None:
        m = self.model
        ok = (np.ndarray, SX, DM)
        assert isinstance(P_x, ok), f"P_x must be of type {ok}, you have {type(P_x)}"
        n_x, n_w, n_v, n_p = m.n_x, m.n_w, m.n_v, self.n_p_est
        mat = lambda P: P if isinstance(P, SX) else SX(np.asarray(P.full() if isinstance(P, DM) else P, dtype=float))
        assert tuple(P_x.shape) == (n_x, n_x), f"P_x has wrong shape:{P_x.shape}, must be {(n_x, n_x)}"
        stage_cost = SX(0.0)
        if P_v is None:
            assert n_v == 0, "Must pass weighting factor P_v, since you have measurement noise on some measurements (configured in model)."
        else:
            assert tuple(P_v.shape) == (n_v, n_v), f"P_v has wrong shape:{P_v.shape}, must be {(n_v, n_v)}"
            v = self._v.cat
            stage_cost = stage_cost + v.T @ mat(P_v) @ v
        if P_w is None:
            assert n_w == 0, "Must pass weighting factor P_w, since you have process noise on some states (configured in model)."
        else:
            assert tuple(P_w.shape) == (n_w, n_w), f"P_w has wrong shape:{P_w.shape}, must be {(n_w, n_w)}"
            w = self._w.cat
            stage_cost = stage_cost + w.T @ mat(P_w) @ w
        dx = self._x.cat - self._x_prev.cat
        arrival_cost = dx.T @ mat(P_x) @ dx
        if P_p is None:
            assert n_p == 0, "Must pass weighting factor P_p, since you are trying to estimate parameters."
        else:
            assert tuple(P_p.shape) == (n_p, n_p), f"P_p has wrong shape:{P_p.shape}, must be {(n_p, n_p)}"
            dp = self._p_est.cat - self._p_est_prev.cat
            arrival_cost = arrival_cost + dp.T @ mat(P_p) @ dp
        self.set_objective(stage_cost, arrival_cost)

    # ---- nonlinear constraints (optimizer.py:483-585; inputs x, u, z, tvp, p_est, p_set, _mhe.py:1034-1036) ----------------------
    def set_nl_cons(self, expr_name, expr, ub=np.inf, soft_constraint=False, penalty_term_cons=1, maximum_violation=np.inf):
        assert self.flags["setup"] == False, "Cannot call .set_expression after .setup()."
        assert isinstance(expr_name, str), f"expr_name must be str, you have: {type(expr_name)}"
        assert isinstance(expr, SX), f"expr must be a symbolic expression, you have: {type(expr)}"
        if soft_constraint:
            self.slack_vars_list.append({"slack_name": expr_name, "shape": expr.shape, "ub": maximum_violation, "penalty": penalty_term_cons})
        self.nl_cons_list.append({"expr_name": expr_name, "expr": expr, "ub": ub})
        return expr

    def _setup_nl_cons(self):
        m = self.model
        self._eps = SymStruct([entry(s["slack_name"], shape=s["shape"]) for s in self.slack_vars_list])
        self.n_eps = self._eps.size
        self._eps_lb, self._eps_ub = self._eps(0.0), self._eps(np.inf)
        exprs = {c["expr_name"]: c["expr"] for c in self.nl_cons_list}
        cost = SX(0.0)
        for s in self.slack_vars_list:
            self._eps_ub[s["slack_name"]] = s["ub"]
            exprs[s["slack_name"]] = exprs[s["slack_name"]] - self._eps[s["slack_name"]]
            cost = cost + sum1(s["penalty"] * self._eps[s["slack_name"]])
        self._nl_cons = SymStruct([entry(c["expr_name"], expr=exprs[c["expr_name"]]) for c in self.nl_cons_list], symbolic=False)
        self.n_nl = self._nl_cons.size
        self.epsterm_fun = Function("epsterm", [self._eps.cat], [cost])
        self._nl_cons_fun = Function("nl_cons_fun", [m._x.cat, m._u.cat, m._z.cat, m._tvp.cat, self._p_est.cat, self._p_set.cat, self._eps.cat],
                                     [self._nl_cons.cat])
        self._nl_cons_ub = self._nl_cons(np.inf)
        self._nl_cons_lb = self._nl_cons(-np.inf)
        for c in self.nl_cons_list:
            self._nl_cons_ub[c["expr_name"]] = c["ub"]

    # ---- parameter / measurement functions (reference :717-801) ----------------------------------------------------------------------
    def get_p_template(self) -> NumStruct: return self._p_set(0)

    def set_p_fun(self, p_fun: Callable) -> None:
        assert self.get_p_template().labels() == p_fun(0).labels(), "Incorrect output of p_fun. Use get_p_template to obtain the required structure."
        self.p_fun = p_fun; self.flags["set_p_fun"] = True

    def get_tvp_template(self) -> NumStruct:
        return SymStruct([entry("_tvp", repeat=self.settings.n_horizon, struct=self.model._tvp)])(0)

    def set_tvp_fun(self, tvp_fun: Callable) -> None:
        assert self.get_tvp_template().labels() == tvp_fun(0).labels(), "Incorrect output of tvp_fun. Use get_tvp_template to obtain the required structure."
        self.tvp_fun = tvp_fun; self.flags["set_tvp_fun"] = True

    def get_y_template(self) -> NumStruct:
        return SymStruct([entry("y_meas", repeat=self.settings.n_horizon, struct=self.model._y)])(0)

    def set_y_fun(self, y_fun: Callable) -> None:
        assert self.get_y_template().labels() == y_fun(0).labels(), "Incorrect output of y_fun. Use get_y_template to obtain the required structure."
        self.y_fun = y_fun; self.flags["set_y_fun"] = True

    def _check_validity(self):
        if self.flags["set_objective"] == False:
            raise Exception("Objective is undefined. Please call .set_objective() or .set_default_objective() prior to .setup().")
        if self.flags["set_tvp_fun"] == False and self.model.n_tvp > 0:
            raise Exception("You have not supplied a function to obtain the time-varying parameters defined in model. Use .set_tvp_fun() prior to setup.")
        if self.flags["set_p_fun"] == False and self._p_set.size > 0:
            raise Exception("You have not supplied a function to obtain the parameters defined in model. Use .set_p_fun() prior to setup.")
        if not hasattr(self, "tvp_fun"):
            _tvp = self.get_tvp_template(); self.set_tvp_fun(lambda t: _tvp)
        if not hasattr(self, "p_fun"):
            _p = self.get_p_template(); self.set_p_fun(lambda t: _p)
        if not hasattr(self, "y_fun"):
            # default: the most recent measurements from the data object (reference :843-862)
            y_template = self.get_y_template()

            def y_fun(t_now):
                n_steps = min(self.data._y.shape[0], self.settings.n_horizon)
                for k in range(-n_steps, 0):
                    y_template["y_meas", k] = self.data._y[k]
                try:
                    for k in range(self.settings.n_horizon - n_steps):
                        y_template["y_meas", k] = self.data._y[-n_steps]
                except Exception:
                    pass
                return y_template
            self.set_y_fun(y_fun)

    # ---- state ------------------------------------------------------------------------------------------------------------------------------
    def _num_prop(name):
        def g(self): return getattr(self, name)
        def s(self, val):
            cur = getattr(self, name)
            if isinstance(val, NumStruct): val = val.full()
            a = np.asarray(val.full() if isinstance(val, DM) else val, dtype=float).ravel()
            if a.size == 1 and cur.size > 1: a = np.full(cur.size, a[0])
            assert a.size == cur.size, f"Variable cannot be set because the supplied vector has the wrong size. You have {a.size} and the model is setup for {cur.size}"
            setattr(self, name, cur(a))
        return property(g, s)
    x0 = _num_prop("_x0"); u0 = _num_prop("_u0"); z0 = _num_prop("_z0"); p_est0 = _num_prop("_p_est0")

    @property
    def t0(self): return self._t0
    @t0.setter
    def t0(self, v): self._t0 = np.asarray(v, dtype=float).reshape(1)

    @property
    def opt_x_num(self): return self._opt_x_num
    @property
    def opt_p_num(self): return self._opt_p_num
    @property
    def lb_opt_x(self): return self._lb_opt_x
    @property
    def ub_opt_x(self): return self._ub_opt_x

    def reset_history(self) -> None:
        self.data.init_storage(); self._t0 = np.array([0.0])

    def set_initial_guess(self) -> None:
        """reference :864-883: x0, u0, z0 and p_est0 (scaled) into every entry of opt_x_num."""
        assert self.flags["setup"] == True, "mhe was not setup yet. Please call mhe.setup()."
        self.opt_x_num["_x"] = self._x0.full() / self._x_scaling.full()
        self.opt_x_num["_u"] = self._u0.full() / self._u_scaling.full()
        if self.model.n_z:
            self.opt_x_num["_z"] = self._z0.full() / self._z_scaling.full()
        if self.n_p_est:
            self.opt_x_num["_p_est"] = self._p_est0.full() / self._p_est_scaling.full()
        self.flags["set_initial_guess"] = True

    # ---- setup ----------------------------------------------------------------------------------------------------------------------------
    def prepare_nlp(self) -> None:
        st, m = self.settings, self.model
        st.check_for_mandatory_settings()
        self._setup_nl_cons()
        self._check_validity()
        self.recipe = r = build_mhe_recipe(self)
        N, P, Pz = r.N, r.P, r.Pz
        n_eps_rep = 1 if st.nl_cons_single_slack else N
        self._opt_x = SymStruct([
            entry("_x", repeat=[N + 1, 1 + P], struct=m._x),
            entry("_z", repeat=[N, Pz], struct=m._z),
            entry("_u", repeat=[N], struct=m._u),
            entry("_w", repeat=[N], struct=m._w),
            entry("_v", repeat=[N], struct=m._v),
            entry("_eps", repeat=[n_eps_rep], struct=self._eps),
            entry("_p_est", struct=self._p_est),
        ])
        self._opt_p = SymStruct([
            entry("_x_prev", struct=m._x),
            entry("_p_est_prev", struct=self._p_est_prev),
            entry("_p_set", struct=self._p_set),
            entry("_tvp", repeat=N, struct=m._tvp),
            entry("_y_meas", repeat=N, struct=m._y),
        ])
        self.aux_struct = SymStruct([entry("_aux", repeat=[N], struct=m._aux_expression)])
        self.n_opt_x, self.n_opt_p, self.n_opt_aux = self._opt_x.size, self._opt_p.size, self.aux_struct.size
        assert self.n_opt_x == r.n_w, (self.n_opt_x, r.n_w)
        self.n_opt_lagr = r.n_g
        osc = self._opt_x(1.0)
        osc["_x"] = self._x_scaling.cat; osc["_u"] = self._u_scaling.cat
        if m.n_z: osc["_z"] = self._z_scaling.cat
        self.opt_x_scaling = osc
        self._lb_opt_x, self._ub_opt_x = self._opt_x(-np.inf), self._opt_x(np.inf)
        self._update_bounds()
        self.nlp_cons_lb = np.zeros(r.n_g); self.nlp_cons_ub = np.zeros(r.n_g)
        for k in range(N):
            o = self._g_off_nl[k]
            self.nlp_cons_lb[o:o + r.nl_rows_per_edge] = -np.inf
            self.nlp_cons_ub[o:o + r.nl_rows_per_edge] = r.nl_ub
        self._opt_x_num = self._opt_x(0.0)
        self.opt_x_num_unscaled = self._opt_x(0.0)
        self._opt_p_num = self._opt_p(0.0)
        self.opt_aux_num = self.aux_struct(0.0)
        self.opt_g_num = DM(np.zeros((r.n_g, 1)))
        self.lam_g_num = DM(np.zeros((r.n_g, 1)))
        self.lam_x_num = DM(np.zeros((r.n_w, 1)))
        self.flags["prepare_nlp"] = True

    def create_nlp(self) -> None:
        from .solver import BatchedNlpSolver
        self.S = BatchedNlpSolver(self.recipe, self.settings.nlpsol_opts)
        self.S.bounds_source = lambda: (self._lb_opt_x.full().ravel(), self._ub_opt_x.full().ravel())
        self.data.set_meta(**{k: getattr(self.settings, k) for k in asdict(self.settings).keys()})
        self._prepare_data()
        self.flags["setup"] = True

    def _prepare_data(self):
        d = self.data
        d.data_fields.update({"_y_meas": self.model.n_y, "opt_p_num": self.n_opt_p})
        d.opt_p = self._opt_p
        if self.settings.store_full_solution:
            d.data_fields.update({"_opt_x_num": self.n_opt_x, "_opt_aux_num": self.n_opt_aux})
            d.opt_x = self._opt_x
        if self.settings.store_lagr_multiplier:
            d.data_fields.update({"_lam_g_num": self.n_opt_lagr})
        for k in self.settings.store_solver_stats:
            d.data_fields.update({k: 1})
        d.init_storage()

    def setup(self) -> None:
        self.prepare_nlp()
        self.create_nlp()

    def _update_bounds(self):
        """reference :995-1028 (note: unlike the controller, the state bounds apply to ALL k including 0 and N)."""
        N = self.settings.n_horizon
        lb, ub = self._opt_x(-np.inf), self._opt_x(np.inf)
        if self.settings.cons_check_colloc_points:
            lb["_x"] = self._x_lb.cat; ub["_x"] = self._x_ub.cat
            lb["_z"] = self._z_lb.cat; ub["_z"] = self._z_ub.cat
        else:
            lb["_x", 1:N, -1] = self._x_lb.cat; ub["_x", 1:N, -1] = self._x_ub.cat
            lb["_z", :, 0] = self._z_lb.cat; ub["_z", :, 0] = self._z_ub.cat
        lb["_u"] = self._u_lb.cat; ub["_u"] = self._u_ub.cat
        if self.n_eps:
            lb["_eps"] = self._eps_lb.cat; ub["_eps"] = self._eps_ub.cat
        if self.n_p_est:
            lb["_p_est"] = self._p_est_lb.cat; ub["_p_est"] = self._p_est_ub.cat
        sc = self.opt_x_scaling.full().ravel()
        self._lb_opt_x = self._opt_x(lb.full().ravel() / sc)
        self._ub_opt_x = self._opt_x(ub.full().ravel() / sc)

    # ---- runtime ------------------------------------------------------------------------------------------------------------------------------
    def kernel_params(self, opt_p_rows) -> np.ndarray:
        """opt_p in the reference order [_x_prev | _p_est_prev | _p_set | _tvp (N) | _y_meas (N)] -> parameter rows of the kernel
        [x0 (unused) | per stage (tvp_k, y_meas_k), N + 1 slots | (p_set, x_prev, p_est_prev) | u_prev (unused)]."""
        r, m = self.recipe, self.model
        op = np.asarray(opt_p_rows, dtype=float).reshape(-1, self.n_opt_p)
        B = op.shape[0]
        N, ntv, ny = r.N, m.n_tvp, m.n_y
        o = 0
        xprev = op[:, o:o + m.n_x]; o += m.n_x
        pprev = op[:, o:o + self.n_p_est]; o += self.n_p_est
        pset = op[:, o:o + self.n_p_set]; o += self.n_p_set
        tvp = op[:, o:o + N * ntv].reshape(B, N, ntv); o += N * ntv
        ym = op[:, o:o + N * ny].reshape(B, N, ny)
        P = np.zeros((B, r.n_p))
        st = np.concatenate([tvp, ym], axis=2)                                # (B, N, ntv + ny)
        P[:, r.p_tvp:r.p_tvp + N * r.n_tvp] = st.reshape(B, -1)
        P[:, r.p_p:r.p_p + r.n_pm] = np.concatenate([pset, xprev, pprev], axis=1)
        return P

    def reference_multipliers(self, x_flat, lam_g, opt_p_row) -> np.ndarray:
        """lam_g of the reference's NLP from the solver's.  The stage recipe evaluates the measurement function at the
        END-OF-INTERVAL collocation state (an unknown of the stage block), the reference at the node state `_x[k+1,-1]`
        (_mhe.py:1140); both are equal through the continuity row, so primal solutions coincide and the continuity multipliers
        differ by exactly  (dh/dx)' lam_meas : added here (host, n_horizon small products)."""
        r, m = self.recipe, self.model
        lam = np.array(lam_g, dtype=float).ravel().copy()
        if m.n_y == 0:
            return lam
        if not hasattr(self, "_meas_jac_fun"):
            xs = SX.sym("xs_scaled", m.n_x)
            sx = SX(np.asarray(self._x_scaling.full()).reshape(-1, 1))
            h = m._meas_fun(xs * sx, m._u.cat, m._z.cat, m._tvp.cat, m._p.cat, m._v.cat)
            from .sym.sx import jacobian
            self._meas_jac_fun = Function("meas_jac", [xs, m._u.cat, m._z.cat, m._tvp.cat, m._p.cat, m._v.cat], [jacobian(h, xs)])
        X = self._opt_x(np.asarray(x_flat, dtype=float).ravel())
        OP = self._opt_p(np.asarray(opt_p_row, dtype=float).ravel())
        p_all = self._p_cat_fun(X["_p_est"], OP["_p_set"]).full()
        su, sz = self._u_scaling.full(), self._z_scaling.full()
        n_x, n_y, P = m.n_x, m.n_y, r.P
        rows_edge = (P * (n_x + m.n_z) if not r.discrete else m.n_z) + n_x + n_y + r.nl_rows_per_edge
        for k in range(r.N):
            zk = X["_z", k, -1].full() * sz if m.n_z else np.zeros((0, 1))
            J = np.asarray(self._meas_jac_fun(X["_x", k + 1, -1].full(), X["_u", k].full() * su, zk, OP["_tvp", k].full() if m.n_tvp else np.zeros((0, 1)),
                                              p_all, X["_v", k].full()).full()).reshape(n_y, n_x)
            g0 = k * rows_edge + (P * (n_x + m.n_z) if not r.discrete else m.n_z)
            lam[g0:g0 + n_x] += J.T @ lam[g0 + n_x:g0 + n_x + n_y]
        return lam

    def solve(self) -> None:
        assert self.flags["setup"] == True, "optimizer was not setup yet. Please call optimizer.setup()."
        kw = dict(x0=self.opt_x_num, lbx=self._lb_opt_x, ubx=self._ub_opt_x, lbg=self.nlp_cons_lb, ubg=self.nlp_cons_ub,
                  p=self.kernel_params(self.opt_p_num.full().ravel())[0])
        if self.flags["initial_run"]:
            kw.update(lam_x0=self.lam_x_num, lam_g0=self.lam_g_num)
        r = self.S(**kw)
        self.opt_x_num.master = r["x"]
        self.opt_x_num_unscaled.master = r["x"].full() * self.opt_x_scaling.full()
        self.opt_g_num = r["g"]
        self.lam_g_num = DM(self.reference_multipliers(r["x"].full().ravel(), r["lam_g"].full().ravel(), self.opt_p_num.full().ravel()))
        self.lam_x_num = r["lam_x"]
        self.solver_stats = self.S.stats()
        # aux expressions at every stage (reference :1188-1190), evaluated on the host (n_aux * N numbers)
        m = self.model
        if m.n_aux:
            xu = self.opt_x_num_unscaled
            p0 = self._p_cat_fun(xu["_p_est"], self.opt_p_num["_p_set"]).full()
            for k in range(self.settings.n_horizon):
                zk = xu["_z", k, -1] if m.n_z else np.zeros((0, 1))
                self.opt_aux_num["_aux", k] = m._aux_expression_fun(xu["_x", k, -1], xu["_u", k], zk, self.opt_p_num["_tvp", k], p0).full()
        self.flags["initial_run"] = True

    def make_step(self, y0) -> np.ndarray:
        assert self.flags["setup"] == True, "optimizer was not setup yet. Please call optimizer.setup()."
        if isinstance(y0, NumStruct):
            y0 = y0.cat
        if not isinstance(y0, (np.ndarray, DM)):
            raise Exception(f"Invalid type {type(y0)} for y0. Must be {(np.ndarray, DM, NumStruct)}")
        y0 = np.asarray(y0.full() if isinstance(y0, DM) else y0, dtype=float)
        assert y0.size == self.model.n_y, f"Wrong input with shape {y0.size}. Expected vector with {self.model.n_y} elements"
        if not self.flags["set_initial_guess"]:
            warnings.warn("Intial guess for the optimizer was not set. The solver call is likely to fail.")
            self.flags["set_initial_guess"] = True
        self.data.update(_y=y0.reshape(-1, 1))
        p_est0, x0, t0 = self._p_est0, self._x0, self._t0
        tvp0, p_set0, y_traj = self.tvp_fun(t0), self.p_fun(t0), self.y_fun(t0)
        self.opt_p_num["_x_prev"] = self.opt_x_num["_x", 1, -1].full() * self._x_scaling.full()
        if self.n_p_est:
            self.opt_p_num["_p_est_prev"] = p_est0.full()
        if self.n_p_set:
            self.opt_p_num["_p_set"] = p_set0.full()
        if self.model.n_tvp:
            self.opt_p_num["_tvp"] = np.concatenate([np.asarray(v.full()) for v in tvp0["_tvp"]])
        self.opt_p_num["_y_meas"] = np.concatenate([np.asarray(v.full()) for v in y_traj["y_meas"]])
        self.solve()
        m = self.model
        x_next = self.opt_x_num["_x", -1, -1].full() * self._x_scaling.full()
        p_est_next = self.opt_x_num["_p_est"].full() * self._p_est_scaling.full() if self.n_p_est else np.zeros((0, 1))
        u0 = self.opt_x_num["_u", -1].full() * self._u_scaling.full()
        z0 = self.opt_x_num["_z", -1, -1].full() * self._z_scaling.full() if m.n_z else np.zeros((0, 1))
        aux0 = self.opt_aux_num["_aux", -1].full() if m.n_aux else np.zeros((0, 1))
        p0 = self._p_cat_fun(p_est0.cat, p_set0.cat).full()
        d = self.data
        d.update(_x=x0.full(), _u=u0, _time=t0, _aux=aux0)
        if m.n_z: d.update(_z=z0)
        if m.n_p: d.update(_p=p0)
        if m.n_tvp: d.update(_tvp=tvp0["_tvp", -1].full())
        d.update(opt_p_num=self.opt_p_num.full())
        if self.settings.store_full_solution:
            d.update(_opt_x_num=self.opt_x_num_unscaled.full(), _opt_aux_num=self.opt_aux_num.full())
        if self.settings.store_lagr_multiplier:
            d.update(_lam_g_num=self.lam_g_num.full())
        for k in self.settings.store_solver_stats:
            if k in self.solver_stats:
                d.update(**{k: self.solver_stats[k]})
        self._t0 = self._t0 + self.settings.t_step
        self._x0 = m._x(x_next); self._u0 = m._u(u0)
        if self.n_p_est: self._p_est0 = self._p_est(p_est_next)
        if m.n_z: self._z0 = m._z(z0)
        return x_next


# ---------------------------------------------------------------------------------------------------------------------------------------
def build_mhe_recipe(mhe: MHE) -> Recipe:
    """The MHE NLP of reference _mhe.py:1030-1195 as a stage recipe (chain, no scenario tree).

    stage state   x~ = (x, p_est)            scaled by (x_scaling, 1)           [p_est lifted, trivial dynamics]
    stage inputs  u~ = (u, w)                scaled by (u_scaling, 1)
    edge unknowns y  = (collocation points of x~ ..., algebraic states ..., v)   [v: measurement noise of the stage]
    edge block    G  = (collocation / algebraic equations, measurement equations h(.) + ... - y_meas)
    parameters    per stage (tvp_k, y_meas_k); global (p_set, x_prev, p_est_prev)
    """
    m, st = mhe.model, mhe.settings
    n_x, n_u, n_z, n_tvp, n_w, n_v, n_y = m.n_x, m.n_u, m.n_z, m.n_tvp, m.n_w, m.n_v, m.n_y
    n_pe, n_ps = mhe.n_p_est, mhe.n_p_set
    N = st.n_horizon
    if n_v != n_y:
        raise NotImplementedError("MHE on the batched solver needs measurement noise on every measurement (set_meas(..., meas_noise=True)): "
                                  "the measurement equations join the square collocation block with v as their unknowns")
    if st.nl_cons_single_slack and mhe.n_eps > 0:
        raise NotImplementedError("nl_cons_single_slack in the MHE")
    if np.any(mhe._p_est_scaling.full() != 1.0):
        raise NotImplementedError("scaling of the estimated parameters")
    discrete = (m.model_type == "discrete")
    deg, ni = st.collocation_deg, st.collocation_ni
    P = 0 if discrete else (deg + 1) * ni
    Pz = max(P, 1)
    n_eps, n_nl = mhe.n_eps, mhe.n_nl
    NX, NU = n_x + n_pe, n_u + n_w
    sx, su, sz = (np.asarray(v.full()).ravel() for v in (mhe._x_scaling.cat, mhe._u_scaling.cat, mhe._z_scaling.cat))
    sxa = np.concatenate([sx, np.ones(n_pe)]); sua = np.concatenate([su, np.ones(n_w)])

    Xn = SX.sym("Xn", NX); Un = SX.sym("Un", NU); EPS = SX.sym("EPS", n_eps) if n_eps else SX.zeros(0, 1)
    ntv_k = n_tvp + n_y
    TV = SX.sym("TV", ntv_k) if ntv_k else SX.zeros(0, 1)
    n_pm = n_ps + n_x + n_pe
    PP = SX.sym("PP", n_pm) if n_pm else SX.zeros(0, 1)
    UPREV = SX.sym("UPREV", NU)
    tvp_k, ymeas_k = TV[0:n_tvp] if n_tvp else SX.zeros(0, 1), TV[n_tvp:ntv_k]
    p_set, x_prev, pe_prev = PP[0:n_ps] if n_ps else SX.zeros(0, 1), PP[n_ps:n_ps + n_x], (PP[n_ps + n_x:n_pm] if n_pe else SX.zeros(0, 1))
    sxS, suS = SX(sx.reshape(-1, 1)), SX(su.reshape(-1, 1))
    szS = SX(sz.reshape(-1, 1)) if n_z else SX.zeros(0, 1)
    x_n, pe_n = Xn[0:n_x], (Xn[n_x:NX] if n_pe else SX.zeros(0, 1))         # parent node: states and (lifted) parameters
    u_n, w_n = Un[0:n_u], (Un[n_u:NU] if n_w else SX.zeros(0, 1))
    p_all = mhe._p_cat_fun(pe_n, p_set)                                        # model parameter vector at this stage

    def rhs_scaled(x, z):            # optimizer.py:807-818 (x, z scaled symbols)
        return m._rhs_fun(x * sxS, u_n * suS, z * szS if n_z else z, tvp_k, p_all, w_n) / sxS

    def alg(x, z):
        return m._alg_fun(x * sxS, u_n * suS, z * szS if n_z else z, tvp_k, p_all, w_n)

    if discrete:
        n_y_blk = n_z + NX + n_v
        Y = SX.sym("Y", n_y_blk)
        zk = Y[0:n_z] if n_z else SX.zeros(0, 1)
        xf = Y[n_z:n_z + NX]
        vk = Y[n_z + NX:n_y_blk]
        rows = []
        if n_z:
            rows.append(alg(x_n, zk))
        rows.append(rhs_scaled(x_n, zk) - xf[0:n_x])
        if n_pe:
            rows.append(pe_n - xf[n_x:NX])
        x_end, z_end = xf[0:n_x], zk
        xf_sel = np.arange(n_z, n_z + NX)
        yx, yz = [], [zk]
        z_first = zk
    else:
        n_y_blk = P * (NX + n_z) + n_v
        Y = SX.sym("Y", n_y_blk)
        _, C, D = collocation_constants(deg, st.collocation_type)
        C[np.abs(C) < 1e-14] = 0.0; D[np.abs(D) < 1e-14] = 0.0
        h = st.t_step / ni
        yx = [Y[i * NX:(i + 1) * NX] for i in range(P)]
        yz = [Y[P * NX + i * n_z:P * NX + (i + 1) * n_z] if n_z else SX.zeros(0, 1) for i in range(P)]
        vk = Y[P * (NX + n_z):n_y_blk]
        xs = [[None] * (deg + 1) for _ in range(ni)]; zs = [[None] * (deg + 1) for _ in range(ni)]
        xs[0][0] = Xn
        off, offz = 0, 0
        zs[0][0] = yz[offz]; offz += 1
        first_j = 1
        for i in range(ni):
            for j in range(first_j, deg + 1):
                xs[i][j] = yx[off]; off += 1
                zs[i][j] = yz[offz]; offz += 1
            first_j = 0
        xkf = yx[off]; off += 1
        rows = []
        for i in range(ni):
            if n_z:
                rows.append(alg(xs[i][0][0:n_x], zs[i][0]))
            for j in range(1, deg + 1):
                xp = SX.zeros(NX, 1)
                for r_ in range(deg + 1):
                    if C[r_, j] != 0.0:
                        xp = xp + float(C[r_, j]) * xs[i][r_]
                f_j = vertcat(h * rhs_scaled(xs[i][j][0:n_x], zs[i][j]), SX.zeros(n_pe, 1)) if n_pe else h * rhs_scaled(xs[i][j][0:n_x], zs[i][j])
                rows.append(f_j - xp)
                if n_z:
                    rows.append(alg(xs[i][j][0:n_x], zs[i][j]))
            xf_i = SX.zeros(NX, 1)
            for r_ in range(deg + 1):
                if D[r_] != 0.0:
                    xf_i = xf_i + float(D[r_]) * xs[i][r_]
            x_next = xs[i + 1][0] if i + 1 < ni else xkf
            rows.append(x_next - xf_i)
        x_end, z_end = xkf[0:n_x], yz[P - 1]
        xf_sel = np.arange((P - 1) * NX, P * NX)
        z_first = yz[0]
    # measurement equations (reference :1140-1161): end-of-interval state, this stage's input, last algebraic state
    meas = m._meas_fun(x_end * sxS, u_n * suS, z_end * szS if n_z else z_end, tvp_k, p_all, vk) - ymeas_k
    rows.append(meas)
    G = vertcat(*rows)
    assert G.shape[0] == n_y_blk, (G.shape, n_y_blk)

    q = vertcat(Xn, Un, EPS, Y)
    lterm = mhe.stage_cost_fun(w_n, vk, tvp_k, p_all)
    epscost = mhe.epsterm_fun(EPS) if n_eps else SX(0.0)
    nl_rows = []
    if n_nl:
        if st.nl_cons_check_colloc_points and not discrete:
            for i in range(P):
                nl_rows.append(mhe._nl_cons_fun(yx[i][0:n_x] * sxS, u_n * suS, yz[i] * szS if n_z else yz[i], tvp_k, pe_n, p_set, EPS))
        else:
            nl_rows.append(mhe._nl_cons_fun(x_n * sxS, u_n * suS, z_first * szS if n_z else z_first, tvp_k, pe_n, p_set, EPS))
        nl_rows.append(nl_rows[-1])                   # the reference appends the last block a second time (_mhe.py:1180-1182)
    nl = vertcat(*nl_rows) if nl_rows else SX.zeros(0, 1)
    nl_ub = np.tile(np.asarray(mhe._nl_cons_ub.cat.full()).ravel(), len(nl_rows)) if nl_rows else np.zeros(0)
    X0s = SX.sym("X0r", NX)
    arrival = mhe.arrival_cost_fun(X0s[0:n_x] * sxS, x_prev, X0s[n_x:NX] if n_pe else SX.zeros(0, 1), pe_prev, p_set)
    XN = SX.sym("XN", NX)
    aux = SX.zeros(0, 1)

    # ---- chain tables --------------------------------------------------------------------------------------------------------------
    n_node, n_edge = N + 1, N
    node_k = np.arange(n_node); node_s = np.zeros(n_node, int)
    node_first_edge = np.concatenate([np.arange(N), [-1]]); node_n_edges = np.concatenate([np.ones(N, int), [0]])
    node_parent_edge = np.arange(n_node) - 1
    # ---- flat layouts (reference _mhe.py:1051-1059 opt_x, :1082-1088 opt_p, :1128-1185 g) ------------------------------------------
    o = 0
    ix_x = o + np.arange((N + 1) * (P + 1) * n_x).reshape(N + 1, P + 1, n_x); o += ix_x.size
    ix_z = o + np.arange(N * Pz * n_z).reshape(N, Pz, n_z); o += ix_z.size
    ix_u = o + np.arange(N * n_u).reshape(N, n_u); o += ix_u.size
    ix_w = o + np.arange(N * n_w).reshape(N, n_w); o += ix_w.size
    ix_v = o + np.arange(N * n_v).reshape(N, n_v); o += ix_v.size
    ix_eps = o + np.arange(N * n_eps).reshape(N, n_eps); o += ix_eps.size
    ix_pe = o + np.arange(n_pe); o += n_pe
    n_w_flat = o
    nvp = NX + NU + n_eps
    w2f = []
    for k in range(n_node):
        w2f += list(ix_x[k, -1]) + (list(ix_pe) if k == 0 else [-1] * n_pe)
        w2f += (list(ix_u[k]) + list(ix_w[k]) + list(ix_eps[k])) if k < N else [-1] * (NU + n_eps)
    node0_pe = [n_x + i for i in range(n_pe)]                                # internal index of the root's parameter copy
    init_from = [-1] * (n_node * nvp)
    for k in range(1, n_node):
        for i in range(n_pe):
            init_from[k * nvp + n_x + i] = node0_pe[i]
    for k in range(n_edge):
        if discrete:
            w2f += list(ix_z[k, 0]) + list(ix_x[k + 1, -1]) + [-1] * n_pe + list(ix_v[k])
            # the auxiliary end-of-interval state is a copy of the child's state: handled by load_initial_guess (discrete branch)
            init_from += [-1] * n_z + [-1] * n_x + node0_pe + [-1] * n_v
        else:
            for j in range(P):
                w2f += list(ix_x[k + 1, j]) + [-1] * n_pe
                init_from += [-1] * n_x + node0_pe
            for j in range(P):
                w2f += list(ix_z[k, j])
                init_from += [-1] * n_z
            w2f += list(ix_v[k]); init_from += [-1] * n_v
    # multiplier rows: [init (all free: no flat row) | per edge: G rows, continuity, nl rows]
    l2f = [-1] * NX
    g_row = 0
    g_off_nl = []
    n_g_edge_ref = (P * (n_x + n_z) if not discrete else n_z) + n_x + n_y + nl.shape[0]
    for k in range(n_edge):
        g0 = g_row
        ref = g0
        if discrete:
            l2f += [ref + i for i in range(n_z)]; ref += n_z
            # reference: rhs/x_scaling - x_next is the continuity row; here: aux row (-1) + continuity row
            l2f += [-1] * n_x + [-1] * n_pe
        else:
            for i in range(ni):
                if n_z:
                    l2f += [ref + a for a in range(n_z)]; ref += n_z
                for j in range(1, deg + 1):
                    l2f += [ref + a for a in range(n_x)] + [-1] * n_pe; ref += n_x
                    if n_z:
                        l2f += [ref + a for a in range(n_z)]; ref += n_z
                l2f += [ref + a for a in range(n_x)] + [-1] * n_pe; ref += n_x
        cont = ref; ref += n_x
        meas_row = ref; ref += n_y
        l2f += [meas_row + a for a in range(n_y)]                            # last rows of G: the measurement equations
        l2f += [cont + a for a in range(n_x)] + [-1] * n_pe                  # continuity rows (state part; parameter part internal)
        g_off_nl.append(ref)
        l2f += [ref + a for a in range(nl.shape[0])]; ref += nl.shape[0]
        assert ref - g0 == n_g_edge_ref
        g_row = ref
    mhe._g_off_nl = g_off_nl
    n_p_int = NX + (N + 1) * ntv_k + n_pm + NU
    osc = np.ones(n_w_flat)
    osc[ix_x] = sx
    if n_z: osc[ix_z] = sz
    osc[ix_u] = su
    return Recipe(
        n_x=NX, n_u=NU, n_z=n_z, n_tvp=ntv_k, n_pm=n_pm, n_comb=1, n_eps=n_eps, n_nl=n_nl, N=N, n_robust=0, P=P, Pz=Pz,
        discrete=discrete, n_y=n_y_blk, nl_rows_per_edge=nl.shape[0], n_branches=[1] * N, n_scenarios=[1] * (N + 1),
        node_k=node_k, node_s=node_s, node_first_edge=node_first_edge, node_n_edges=node_n_edges, node_parent_edge=node_parent_edge,
        edge_k=np.arange(N), edge_parent=np.arange(N), edge_child=np.arange(N) + 1, edge_pidx=np.zeros(N, int), edge_omega=np.ones(N),
        n_w=n_w_flat, n_g=g_row, n_p=n_p_int,
        node_x_flat=np.array([int(ix_x[k, -1, 0]) for k in range(n_node)]), node_u_flat=-np.ones(n_node, int), node_eps_flat=-np.ones(n_node, int),
        edge_yx_flat=-np.ones(n_edge, int), edge_yz_flat=-np.ones(n_edge, int), edge_g_flat=np.zeros(n_edge, int),
        p_x0=0, p_tvp=NX, p_p=NX + (N + 1) * ntv_k, p_uprev=NX + (N + 1) * ntv_k + n_pm,
        q=q, tvp=TV, pp=PP, uprev=UPREV, xN=XN, G=G, lterm=lterm, epscost=epscost, nl=nl, nl_ub=nl_ub,
        rterm=SX(0.0), mterm=SX(0.0), aux=aux, xf_sel=xf_sel, x_scaling=sxa, u_scaling=sua, z_scaling=sz, opt_x_scaling=osc,
        n_free0=NX, arrival=arrival, x0sym=X0s, w2flat=np.array(w2f), lam2flat=np.array(l2f), w_init_from=np.array(init_from),
        has_dummies=(P > 0),
    )
