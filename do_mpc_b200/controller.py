"""Model predictive controller with the do_mpc.controller.MPC API on top of the batched CUDA NLP solver.

API mirror of reference do_mpc/controller/_mpc.py (class MPC :37; set_objective :525, set_rterm :593,
set_uncertainty_values :819, setup :933, set_initial_guess :955, make_step :975) and of the shared base
do_mpc/optimizer.py (bounds :268, scaling :356, set_nl_cons :483, get/set_tvp_fun :588-676, solve :731).
Settings mirror do_mpc/controller/_controllersettings.py:64-175.

Difference by design: `setup()` does not build a flat symbolic NLP for CasADi/IPOPT; it builds a stage
recipe (do_mpc_b200/recipe.py), generates + compiles sm_100a kernels (do_mpc_b200/solver) and installs
`self.S`, an object with the call signature of the CasADi nlpsol Function (reference optimizer.py:754-778).
`make_step_batch` is the batched entry point that replaces the reference's serial sampling loop
(do_mpc/approximateMPC/_ampc_sampler.py:312-320).
"""
from __future__ import annotations

import itertools
import time
import warnings
from dataclasses import dataclass, field, asdict
from typing import Callable, List, Optional

import numpy as np

from .data import MPCData
from .model import Model
from .sym.sx import SX, DM, Function, vertcat, sum1
from .sym.struct import SymStruct, NumStruct, entry


@dataclass
class MPCSettings:
    """Settings of the MPC (reference _controllersettings.py:64-175); unknown ipopt.* options are accepted."""
    n_horizon: int = None
    n_robust: int = 0
    open_loop: bool = False
    t_step: float = None
    use_terminal_bounds: bool = False
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
    gauss_newton: bool = False

    def check_for_mandatory_settings(self):
        if self.n_horizon is None:
            raise ValueError("n_horizon must be set")
        if self.t_step is None:
            raise ValueError("t_step must be set")

    def supress_ipopt_output(self):
        self.nlpsol_opts.update({"ipopt.print_level": 0, "ipopt.sb": "yes", "print_time": 0})

    def set_linear_solver(self, solver_name: str = "MA27"):
        self.nlpsol_opts.update({"ipopt.linear_solver": solver_name})


class _Indexed:
    """`obj.bounds['lower','_x','C_a'] = v` style access (reference tools/_indexedproperty.py:3)."""

    def __init__(self, getter, setter):
        self._g, self._s = getter, setter

    def __getitem__(self, ind): return self._g(ind)
    def __setitem__(self, ind, val): self._s(ind, val)


class MPC:
    def __init__(self, model: Model, settings: Optional[MPCSettings] = None):
        assert model.flags["setup"] == True, "Model for MPC was not setup. After the complete model creation call model.setup()."
        self.model = model
        self.settings = self._settings = settings or MPCSettings()
        self.data = MPCData(model)
        self.flags = {"setup": False, "prepare_nlp": False, "set_objective": False, "set_rterm": False,
                      "rterm_fun": False, "set_tvp_fun": False, "set_p_fun": False, "set_initial_guess": False,
                      "initial_run": False, "MINLP": bool(model.integer)}
        if self.flags["MINLP"]:
            raise NotImplementedError("integer inputs (MINLP / bonmin) are outside the GPU interior-point path")
        m = model
        self._x_lb, self._x_ub = m._x(-np.inf), m._x(np.inf)
        self._x_terminal_lb, self._x_terminal_ub = m._x(-np.inf), m._x(np.inf)
        self._u_lb, self._u_ub = m._u(-np.inf), m._u(np.inf)
        self._z_lb, self._z_ub = m._z(-np.inf), m._z(np.inf)
        self._x_scaling, self._u_scaling, self._z_scaling, self._p_scaling = m._x(1.0), m._u(1.0), m._z(1.0), m._p(1.0)
        self.rterm_factor = m._u(0.0)
        self.u_prev = SX.sym("u_prev", m.n_u)
        self._x0, self._u0, self._z0, self._t0 = m._x(0.0), m._u(0.0), m._z(0.0), np.array([0.0])
        self.nl_cons_list, self.slack_vars_list = [], []
        self.slack_cost = SX(0.0)
        self.n_combinations = 1
        self.bounds = _Indexed(self._get_bounds, self._set_bounds)
        self.terminal_bounds = _Indexed(self._get_tbounds, self._set_tbounds)
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
                print(f"Warning: Key {k} does not exist for MPC.")

    # ---- bounds / scaling ----------------------------------------------------------------------------------------
    def _bstruct(self, ind):
        assert isinstance(ind, tuple) and len(ind) >= 3, "Power index must include bound_type, var_type, var_name (as a tuple)."
        bt, vt = ind[0], ind[1]
        assert bt in ("lower", "upper"), f"Invalid power index {bt} for bound_type. Must be from (lower, upper)."
        assert vt in ("_x", "_u", "_z"), f"Invalid power index {vt} for var_type. Must be from (_x, _u, _z)."
        name = f"{vt}_{'lb' if bt == 'lower' else 'ub'}"
        s = getattr(self, name)
        assert ind[2] in s.keys(), f"Calling .bounds with {ind} is not valid. Possible keys are {s.keys()}."
        return s, ind[2:]

    def _get_bounds(self, ind):
        s, k = self._bstruct(ind); return s[k]

    def _set_bounds(self, ind, val):
        s, k = self._bstruct(ind); s[k] = val
        if self.flags["prepare_nlp"]:
            self._update_bounds()

    def _tbstruct(self, ind):
        """terminal_bounds index = (bound_type, var_name[, ...]) as in the reference (_mpc.py:408-478: always the states);
        the 3-tuple form with an explicit '_x' var_type is tolerated."""
        assert isinstance(ind, tuple), "Power index must include bound_type, var_name (as a tuple)."
        assert len(ind) >= 2, "Power index must include bound_type, var_type, var_name (as a tuple)."
        bt = ind[0]
        assert bt in ("lower", "upper"), f"Invalid power index {bt} for bound_type. Must be from (lower, upper)."
        name = ind[1:]
        if name[0] == "_x" and len(name) >= 2:
            name = name[1:]
        s = self._x_terminal_lb if bt == "lower" else self._x_terminal_ub
        assert name[0] in s.keys(), f"Calling .bounds with {ind} is not valid. Possible keys are {s.keys()}."
        return s, name

    def _get_tbounds(self, ind):
        s, k = self._tbstruct(ind); return s[k]

    def _set_tbounds(self, ind, val):
        s, k = self._tbstruct(ind); s[k] = val
        if self.flags["prepare_nlp"]:
            self._update_bounds()

    def _sstruct(self, ind):
        assert isinstance(ind, tuple) and len(ind) >= 2, "Power index must include var_type, var_name (as a tuple)."
        vt = ind[0]
        assert vt in ("_x", "_u", "_z", "_p"), f"Invalid power index {vt} for var_type."
        return getattr(self, f"{vt}_scaling"), ind[1:]

    def _get_scaling(self, ind):
        s, k = self._sstruct(ind); return s[k]

    def _set_scaling(self, ind, val):
        assert self.flags["setup"] == False, "Scaling can only be set before setup (kernels are generated with it)."
        s, k = self._sstruct(ind); s[k] = val

    # ---- objective -----------------------------------------------------------------------------------------------------
    def set_objective(self, mterm=None, lterm=None) -> None:
        assert self.flags["setup"] == False, "Cannot call .set_objective after .setup()."
        m = self.model
        mterm = SX(0.0) if mterm is None else (mterm if isinstance(mterm, SX) else SX(mterm))
        lterm = SX(0.0) if lterm is None else (lterm if isinstance(lterm, SX) else SX(lterm))
        assert mterm.shape == (1, 1), f"mterm must have shape=(1,1). You have {mterm.shape}"
        assert lterm.shape == (1, 1), f"lterm must have shape=(1,1). You have {lterm.shape}"
        self.mterm, self.lterm = mterm, lterm
        try:
            self.mterm_fun = Function("mterm", [m._x.cat, m._tvp.cat, m._p.cat], [mterm])
        except Exception:
            raise Exception("mterm contains invalid symbolic variables as inputs. Must contain only: _x, _tvp, _p")
        try:
            self.lterm_fun = Function("lterm", [m._x.cat, m._u.cat, m._z.cat, m._tvp.cat, m._p.cat], [lterm])
        except Exception:
            raise Exception("lterm contains invalid symbolic variables as inputs. Must contain only: _x, _u, _z, _tvp, _p")
        self.flags["set_objective"] = True

    def set_rterm(self, rterm=None, **kwargs) -> None:
        assert self.flags["setup"] == False, "Cannot call .set_rterm after .setup()."
        m = self.model
        if rterm is not None:
            rterm = rterm if isinstance(rterm, SX) else SX(rterm)
            assert rterm.shape == (1, 1), f"rterm must have shape=(1,1). You have {rterm.shape}"
            self.rterm_fun = Function("rterm", [m._x.cat, m._u.cat, self.u_prev, m._z.cat, m._tvp.cat, m._p.cat], [rterm])
            self.flags["rterm_fun"] = True
        else:
            for key, val in kwargs.items():
                assert key in m._u.keys(), \
                    f"Must pass keywords that refer to input names defined in model. Valid is: {m._u.keys()}. You have: {key}"
                assert isinstance(val, (int, float, np.ndarray)), f"Value for {key} must be int, float or numpy.ndarray."
                self.rterm_factor[key] = val
        self.flags["set_rterm"] = True

    # ---- nonlinear constraints (optimizer.py:483-585) ----------------------------------------------------------------
    def set_nl_cons(self, expr_name, expr, ub=np.inf, soft_constraint=False, penalty_term_cons=1, maximum_violation=np.inf):
        assert self.flags["setup"] == False, "Cannot call .set_expression after .setup()."
        assert isinstance(expr_name, str), f"expr_name must be str, you have: {type(expr_name)}"
        assert isinstance(expr, SX), f"expr must be a symbolic expression, you have: {type(expr)}"
        assert isinstance(ub, (int, float, np.ndarray)), f"ub must be float, int or numpy.ndarray, you have: {type(ub)}"
        assert isinstance(soft_constraint, bool), f"soft_constraint must be boolean, you have: {type(soft_constraint)}"
        if soft_constraint:
            self.slack_vars_list.append({"slack_name": expr_name, "shape": expr.shape, "ub": maximum_violation,
                                         "penalty": penalty_term_cons})
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
        self.slack_cost = cost
        self._nl_cons = SymStruct([entry(c["expr_name"], expr=exprs[c["expr_name"]]) for c in self.nl_cons_list], symbolic=False)
        self.n_nl = self._nl_cons.size
        self.epsterm_fun = Function("epsterm", [self._eps.cat], [cost])
        self._nl_cons_fun = Function("nl_cons_fun", [m._x.cat, m._u.cat, m._z.cat, m._tvp.cat, m._p.cat, self._eps.cat],
                                     [self._nl_cons.cat])
        self._nl_cons_ub = self._nl_cons(np.inf)
        self._nl_cons_lb = self._nl_cons(-np.inf)
        for c in self.nl_cons_list:
            self._nl_cons_ub[c["expr_name"]] = c["ub"]

    # ---- parameters ----------------------------------------------------------------------------------------------------------
    def get_p_template(self, n_combinations: int) -> NumStruct:
        self.n_combinations = n_combinations
        return SymStruct([entry("_p", repeat=n_combinations, struct=self.model._p)])(0)

    def set_p_fun(self, p_fun: Callable) -> None:
        assert self.get_p_template(self.n_combinations).labels() == p_fun(0).labels(), \
            "Incorrect output of p_fun. Use get_p_template to obtain the required structure."
        self.flags["set_p_fun"] = True
        self.p_fun = p_fun

    def set_uncertainty_values(self, **kwargs) -> None:
        if not kwargs:
            return None
        names = list(kwargs.keys())
        valid = self.model.p.keys()
        assert set(names).issubset(set(valid)), \
            f"You passed keywords {names}. Valid keywords are: {valid} (refering to user-defined parameter names)."
        p_scenario = list(itertools.product(*kwargs.values()))
        p_template = self.get_p_template(len(p_scenario))
        p_template["_p", :, names] = np.array(p_scenario, dtype=float)
        self.set_p_fun(lambda t_now: p_template)

    def get_tvp_template(self) -> NumStruct:
        return SymStruct([entry("_tvp", repeat=self.settings.n_horizon + 1, struct=self.model._tvp)])(0)

    def set_tvp_fun(self, tvp_fun: Callable) -> None:
        assert self.get_tvp_template().labels() == tvp_fun(0).labels(), \
            "Incorrect output of tvp_fun. Use get_tvp_template to obtain the required structure."
        self.flags["set_tvp_fun"] = True
        self.tvp_fun = tvp_fun

    # ---- setup --------------------------------------------------------------------------------------------------------------------
    def _check_validity(self):
        if self.flags["set_objective"] == False:
            raise Exception("Objective is undefined. Please call .set_objective() prior to .setup().")
        if self.flags["set_rterm"] == False:
            warnings.warn("rterm was not set and defaults to zero. Changes in the control inputs are not penalized. "
                          "Can lead to oscillatory behavior.")
        if self.flags["set_tvp_fun"] == False and self.model.n_tvp > 0:
            raise Exception("You have not supplied a function to obtain the time-varying parameters defined in model. "
                            "Use .set_tvp_fun() prior to setup.")
        if self.flags["set_p_fun"] == False and self.model.n_p > 0:
            raise Exception("You have not supplied a function to obtain the parameters defined in model. Use .set_p_fun() "
                            "(low-level API) or .set_uncertainty_values() (high-level API) prior to setup.")
        for lb, ub in zip([self._x_lb, self._u_lb, self._z_lb], [self._x_ub, self._u_ub, self._z_ub]):
            bad = lb.full().ravel() > ub.full().ravel()
            if np.any(bad):
                fail = [l for i, l in enumerate(lb.labels()) if bad[i]]
                raise Exception(f"Your bounds are inconsistent. For {fail} you have lower bound > upper bound.")
        if np.all(self._x_terminal_ub.full() == np.inf) and self.settings.use_terminal_bounds:
            self._x_terminal_ub = self._x_ub
        if np.all(self._x_terminal_lb.full() == -np.inf) and self.settings.use_terminal_bounds:
            self._x_terminal_lb = self._x_lb
        if not hasattr(self, "tvp_fun"):
            _tvp = self.get_tvp_template()
            self.set_tvp_fun(lambda t: _tvp)
        if not hasattr(self, "p_fun"):
            _p = self.get_p_template(1)
            self.set_p_fun(lambda t: _p)

    def prepare_nlp(self) -> None:
        from .recipe import build_recipe
        st, m = self.settings, self.model
        st.check_for_mandatory_settings()
        self._setup_nl_cons()
        self._check_validity()
        self.recipe = r = build_recipe(self)
        N, S, P, Pz = r.N, r.n_comb ** r.n_robust, r.P, r.Pz
        self._opt_x = SymStruct([
            entry("_x", repeat=[N + 1, S, 1 + P], struct=m._x),
            entry("_z", repeat=[N, S, Pz], struct=m._z),
            entry("_u", repeat=[N, S], struct=m._u),
            entry("_eps", repeat=[1 if st.nl_cons_single_slack else N, S], struct=self._eps),     # reference _mpc.py:1120-1133
        ])
        self._opt_p = SymStruct([
            entry("_x0", struct=m._x),
            entry("_tvp", repeat=N + 1, struct=m._tvp),
            entry("_p", repeat=self.n_combinations, struct=m._p),
            entry("_u_prev", struct=m._u),
        ])
        self.aux_struct = SymStruct([entry("_aux", repeat=[N, S], struct=m._aux_expression)])
        self.n_opt_x, self.n_opt_p, self.n_opt_aux = self._opt_x.size, self._opt_p.size, self.aux_struct.size
        assert self.n_opt_x == r.n_w and self.n_opt_p == r.n_p
        self.n_opt_lagr = r.n_g
        self.opt_x_scaling = self._opt_x(r.opt_x_scaling)
        self._lb_opt_x, self._ub_opt_x = self._opt_x(-np.inf), self._opt_x(np.inf)
        self._update_bounds()
        self.nlp_cons_lb = np.zeros(r.n_g); self.nlp_cons_ub = np.zeros(r.n_g)
        for e in range(len(r.edge_k)):
            o = r.edge_g_flat[e] + (r.n_z + r.n_x - r.n_lift if r.discrete else r.n_y + r.n_x - 2 * r.n_lift)
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
        meta = {k: getattr(self.settings, k) for k in asdict(self.settings).keys()}
        self.data.set_meta(**meta)
        self.data.prepare(self)
        self.flags["setup"] = True

    def setup(self) -> None:
        self.prepare_nlp()
        self.create_nlp()

    def _update_bounds(self):
        """reference _mpc.py:1061-1095; values stored scaled (optimizer.py:233-265)."""
        N = self.settings.n_horizon
        lb, ub = self._opt_x(-np.inf), self._opt_x(np.inf)
        if self.settings.cons_check_colloc_points:
            lb["_x", 1:N] = self._x_lb.cat; ub["_x", 1:N] = self._x_ub.cat
            lb["_z"] = self._z_lb.cat; ub["_z"] = self._z_ub.cat
        else:
            lb["_x", 1:N, :, -1] = self._x_lb.cat; ub["_x", 1:N, :, -1] = self._x_ub.cat
            lb["_z", :, :, 0] = self._z_lb.cat; ub["_z", :, :, 0] = self._z_ub.cat
        lb["_x", N, :, -1] = self._x_terminal_lb.cat; ub["_x", N, :, -1] = self._x_terminal_ub.cat
        lb["_u"] = self._u_lb.cat; ub["_u"] = self._u_ub.cat
        if self.n_eps:
            lb["_eps"] = self._eps_lb.cat; ub["_eps"] = self._eps_ub.cat
        sc = self.opt_x_scaling.full().ravel()
        self._lb_opt_x = self._opt_x(lb.full().ravel() / sc)
        self._ub_opt_x = self._opt_x(ub.full().ravel() / sc)

    # ---- properties -------------------------------------------------------------------------------------------------------------------------
    @property
    def opt_x_num(self): return self._opt_x_num
    @property
    def opt_p_num(self): return self._opt_p_num
    @property
    def lb_opt_x(self): return self._lb_opt_x
    @property
    def ub_opt_x(self): return self._ub_opt_x

    def _num_prop(name):
        def g(self): return getattr(self, name)
        def s(self, val):
            cur = getattr(self, name)
            if isinstance(val, NumStruct): val = val.full()
            a = np.asarray(val.full() if isinstance(val, DM) else val, dtype=float).ravel()
            assert a.size == cur.size, f"Variable cannot be set because the supplied vector has the wrong size. You have {a.size} and the model is setup for {cur.size}"
            setattr(self, name, cur(a))
        return property(g, s)
    x0 = _num_prop("_x0"); u0 = _num_prop("_u0"); z0 = _num_prop("_z0")

    @property
    def t0(self): return self._t0
    @t0.setter
    def t0(self, v): self._t0 = np.asarray(v, dtype=float).reshape(1)

    def reset_history(self) -> None:
        self.data.init_storage()
        self._t0 = np.array([0.0])

    def set_initial_guess(self) -> None:
        assert self.flags["setup"] == True, "MPC was not setup yet. Please call MPC.setup()."
        self.opt_x_num["_x"] = self._x0.full() / self._x_scaling.full()
        self.opt_x_num["_u"] = self._u0.full() / self._u_scaling.full()
        if self.model.n_z:
            self.opt_x_num["_z"] = self._z0.full() / self._z_scaling.full()
        self.flags["set_initial_guess"] = True

    # ---- runtime ------------------------------------------------------------------------------------------------------------------------------
    def solve(self) -> None:
        assert self.flags["setup"] == True, "optimizer was not setup yet. Please call optimizer.setup()."
        kw = dict(x0=self.opt_x_num, lbx=self._lb_opt_x, ubx=self._ub_opt_x, lbg=self.nlp_cons_lb,
                  ubg=self.nlp_cons_ub, p=self.opt_p_num)
        if self.flags["initial_run"]:
            kw.update(lam_x0=self.lam_x_num, lam_g0=self.lam_g_num)
        r = self.S(**kw)
        self.opt_x_num.master = r["x"]
        self.opt_x_num_unscaled.master = r["x"].full() * self.opt_x_scaling.full()
        self.opt_g_num = r["g"]
        self.lam_g_num = r["lam_g"]
        self.lam_x_num = r["lam_x"]
        self.solver_stats = self.S.stats()
        self.opt_aux_num.master = self.S.aux(r["x"], self.opt_p_num)
        self.flags["initial_run"] = True

    def make_step(self, x0) -> np.ndarray:
        assert self.flags["setup"] == True, "MPC was not setup yet. Please call MPC.setup()."
        if isinstance(x0, NumStruct):
            x0 = x0.cat
        if not isinstance(x0, (np.ndarray, DM)):
            raise Exception(f"Invalid type {type(x0)} for x0. Must be {(np.ndarray, DM, NumStruct)}")
        x0 = np.asarray(x0.full() if isinstance(x0, DM) else x0, dtype=float)
        assert x0.size == self.model.n_x, f"Wrong input with shape {x0.size}. Expected vector with {self.model.n_x} elements"
        x0 = x0.reshape(-1, 1)
        if not self.flags["set_initial_guess"]:
            warnings.warn("Intial guess for the MPC was not set. The solver call is likely to fail.")
            self.flags["set_initial_guess"] = True
        u_prev = self._u0
        tvp0 = self.tvp_fun(self._t0)
        p0 = self.p_fun(self._t0)
        t0 = self._t0
        self.opt_p_num["_x0"] = x0
        self.opt_p_num["_u_prev"] = u_prev.full()
        if self.model.n_tvp:
            self.opt_p_num["_tvp"] = tvp0["_tvp"] if False else np.concatenate([np.asarray(v.full()) for v in tvp0["_tvp"]])
        if self.model.n_p:
            self.opt_p_num["_p"] = np.concatenate([np.asarray(v.full()) for v in p0["_p"]])
        self.solve()
        u0 = self.opt_x_num["_u", 0, 0].full() * self._u_scaling.full()
        z0 = self.opt_x_num["_z", 0, 0, 0].full() * self._z_scaling.full() if self.model.n_z else np.zeros((0, 1))
        aux0 = self.opt_aux_num["_aux", 0, 0].full()
        d = self.data
        d.update(_x=x0, _u=u0, _z=z0, _time=t0, _aux=aux0)
        if self.model.n_tvp: d.update(_tvp=tvp0["_tvp", 0].full())
        if self.model.n_p: d.update(_p=p0["_p", 0].full())
        d.update(opt_p_num=self.opt_p_num.full())
        if self.settings.store_full_solution:
            d.update(_opt_x_num=self.opt_x_num_unscaled.full(), _opt_aux_num=self.opt_aux_num.full())
        if self.settings.store_lagr_multiplier:
            d.update(_lam_g_num=self.lam_g_num.full())
        for k in self.settings.store_solver_stats:
            if k in self.solver_stats:
                d.update(**{k: self.solver_stats[k]})
        self._t0 = self._t0 + self.settings.t_step
        self._x0 = self.model._x(x0); self._u0 = self.model._u(u0)
        if self.model.n_z: self._z0 = self.model._z(z0)
        return u0

    def pack_batch(self, X0, U_prev=None, Z0=None, tvp=None, p=None):
        """Inputs of B independent controller steps in the solver's layout: opt_p rows P [B, n_p] (reference order
        [_x0 | _tvp(N+1) | _p(n_comb) | _u_prev], _mpc.py:1160-1165) and the scaled broadcast initial guess (x0s, u0s, z0s) that
        set_initial_guess() would produce per instance (_mpc.py:969-971).  tvp / p default to tvp_fun(t0) / p_fun(t0)."""
        assert self.flags["prepare_nlp"] == True, "MPC was not prepared yet. Please call MPC.setup() or MPC.prepare_nlp()."
        X0 = np.ascontiguousarray(np.asarray(X0, dtype=float).reshape(-1, self.model.n_x))
        B = X0.shape[0]
        U_prev = np.tile(self._u0.full().ravel(), (B, 1)) if U_prev is None else \
            np.ascontiguousarray(np.asarray(U_prev, dtype=float).reshape(B, self.model.n_u))
        r = self.recipe
        P_ = np.zeros((B, r.n_p))
        P_[:, r.p_x0:r.p_x0 + self.model.n_x] = X0
        P_[:, r.p_uprev:r.p_uprev + r.n_u] = U_prev
        if r.n_tvp:
            t = self.tvp_fun(self._t0)["_tvp"] if tvp is None else None
            tv = np.concatenate([np.asarray(v.full()).ravel() for v in t]) if tvp is None else np.asarray(tvp, dtype=float)
            P_[:, r.p_tvp:r.p_tvp + (r.N + 1) * r.n_tvp] = tv.reshape(-1, (r.N + 1) * r.n_tvp)
        if r.n_pm:
            pv = np.concatenate([np.asarray(v.full()).ravel() for v in self.p_fun(self._t0)["_p"]]) if p is None \
                else np.asarray(p, dtype=float)
            P_[:, r.p_p:r.p_p + r.n_comb * r.n_pm] = pv.reshape(-1, r.n_comb * r.n_pm)
        Zs = None if not r.n_z else (np.tile(self._z0.full().ravel(), (B, 1)) if Z0 is None else np.asarray(Z0, dtype=float).reshape(B, r.n_z))
        x0s = X0 / r.x_scaling[:self.model.n_x]
        if r.n_lift:                                   # lifted slack copies start at 0 like `_eps` after set_initial_guess()
            x0s = np.concatenate([x0s, np.zeros((B, r.n_lift))], axis=1)
        return P_, x0s, U_prev / r.u_scaling, None if Zs is None else Zs / r.z_scaling

    # ---- batched entry point ---------------------------------------------------------------------------------------------------------------------
    def make_step_batch(self, X0, U_prev=None, Z0=None, tvp=None, p=None, return_full=False, warm_from=None):
        """Solve B independent controller steps at once (cold start per instance, as the reference's sampler
        does one by one: reset_history(); x0=..; u0=..; set_initial_guess(); make_step(x0),
        do_mpc/approximateMPC/_ampc_sampler.py:312-320).

        X0: (B, n_x); U_prev: (B, n_u) (default: current u0); returns dict with 'u0' (B, n_u), 'success' (B,),
        'iter_count' (B,), 'status' (B,) and, if return_full, the primal/dual solution arrays.
        warm_from: the dict returned by a previous call with return_full=True — its primal solution becomes the initial guess
        (what make_step() does from one sampling instant to the next, optimizer.py:741-742) and, with
        nlpsol_opts['ipopt.warm_start_init_point']='yes', its multipliers are taken over as well.
        """
        assert self.flags["setup"] == True, "MPC was not setup yet. Please call MPC.setup()."
        P_, x0s, u0s, z0s = self.pack_batch(X0, U_prev, Z0, tvp, p)
        r = self.recipe
        ws = {}
        if warm_from is not None:
            if "x" not in warm_from:
                raise ValueError("warm_from must come from make_step_batch(..., return_full=True)")
            ws = dict(x_init=warm_from["x"], lam_g0=warm_from.get("lam_g"), lam_x0=warm_from.get("lam_x"))
        out = self.S.solve_batch(P_, x0s, u0s, z0s,
                                 self._lb_opt_x.full().ravel(), self._ub_opt_x.full().ravel(), return_full=return_full, **ws)
        out["u0"] = out["u0_scaled"] * r.u_scaling
        out["p"] = P_
        return out
