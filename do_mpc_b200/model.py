"""Symbolic model container with the do_mpc.model.Model API.

Mirrors the user-facing behaviour of reference do_mpc/model/_model.py (class Model :31;
set_variable :537, set_expression :623, set_meas :670, set_rhs :749, set_alg :811, setup :937;
properties x/u/z/p/tvp/y/aux/w/v :218-533; `model['x','u']` :164).  It is an independent
implementation on this package's own expression graph (do_mpc_b200.sym) — the per-stage atoms
(rhs, alg, aux, meas) stay small graphs that the solver code-generator differentiates and emits as
CUDA; nothing here evaluates numerics on the hot path.
"""
from __future__ import annotations

from typing import Dict, List, Tuple, Union

import numpy as np

from .sym.sx import SX, DM, Function, vertcat, jacobian
from .sym.struct import SymStruct, NumStruct, entry

_VAR_TYPES = ["_x", "_u", "_z", "_p", "_tvp", "_w", "_v"]


class Model:
    """Dynamic model  dx/dt = f(x,u,z,tvp,p,w) (continuous) or x+ = f(...) (discrete), 0 = g(...), y = h(...)."""

    def __init__(self, model_type: str = None, symvar_type: str = "SX"):
        assert isinstance(model_type, str), f"model_type must be string, you have: {type(model_type)}"
        assert model_type in ["discrete", "continuous"], \
            f"model_type must be either discrete or continuous, you have: {model_type}"
        assert symvar_type in ["SX", "MX"], f"symvar_type must be either SX or MX, you have: {symvar_type}"
        self.model_type = model_type
        self.symvar_type = symvar_type          # MX requests are served by the same SX-style graph
        self._vars: Dict[str, List[Tuple[str, SX]]] = {k: [] for k in _VAR_TYPES}
        self._aux_list: List[Tuple[str, SX]] = [("default", SX(0.0))]
        self._y_list: List[Tuple[str, SX]] = []
        self.rhs_list: List[dict] = []
        self.alg_list: List[Tuple[str, SX]] = []
        self.integer: List[str] = []
        self.flags = {"setup": False}

    # -- declaration ---------------------------------------------------------------------------------
    def set_variable(self, var_type: str, var_name: str, shape: Union[int, Tuple] = (1, 1),
                     input_type_integer: bool = False) -> SX:
        assert self.flags["setup"] == False, "Cannot call .set_variable after setup."
        assert isinstance(var_type, str), f"var_type must be str, you have: {type(var_type)}"
        assert isinstance(var_name, str), f"var_name must be str, you have: {type(var_name)}"
        assert isinstance(shape, (tuple, int)), f"shape must be tuple or int, you have: {type(shape)}"
        if isinstance(shape, int):
            shape = (shape, 1)
        if len(shape) == 1:
            shape = (shape[0], 1)
        # long names are accepted like the reference does (model/_model.py:598-604)
        for long_name, short in (("states", "_x"), ("inputs", "_u"), ("algebraic", "_z"), ("parameter", "_p"), ("timevarying_parameter", "_tvp")):
            if var_type == long_name:
                var_type = short
        if not var_type.startswith("_"):
            var_type = "_" + var_type
        if var_type not in _VAR_TYPES:
            raise ValueError(f"Trying to set non-existing variable var_type: {var_type} with var_name {var_name}")
        if var_name in [n for n, _ in self._vars[var_type]]:
            raise ValueError(f"The variable name {var_name} for type {var_type} already exists.")
        var = SX.sym(var_name, shape[0], shape[1])
        self._vars[var_type].append((var_name, var))
        if input_type_integer:
            assert var_type == "_u", "only inputs can be integer"
            self.integer.append(var_name)
        return var

    def set_expression(self, expr_name: str, expr) -> SX:
        assert self.flags["setup"] == False, "Cannot call .set_expression after setup."
        assert isinstance(expr_name, str), f"expr_name must be str, you have: {type(expr_name)}"
        expr = expr if isinstance(expr, SX) else SX(expr)
        self._aux_list.append((expr_name, expr))
        return expr

    def set_meas(self, meas_name: str, expr, meas_noise: bool = True) -> SX:
        assert self.flags["setup"] == False, "Cannot call .set_meas after setup."
        expr = expr if isinstance(expr, SX) else SX(expr)
        if meas_noise:
            v = self.set_variable("_v", meas_name + "_noise", expr.shape)
            expr = expr + v
        self._y_list.append((meas_name, expr))
        return expr

    def set_rhs(self, var_name: str, expr, process_noise: bool = False) -> None:
        assert self.flags["setup"] == False, "Cannot call .set_rhs after .setup."
        assert isinstance(var_name, str), f"var_name must be str, you have: {type(var_name)}"
        names = [n for n, _ in self._vars["_x"]]
        if var_name not in names:
            raise ValueError(f"{var_name} does not exist as a state. Existing states: {names}")
        expr = expr if isinstance(expr, SX) else SX(expr)
        ref = dict(self._vars["_x"])[var_name]
        assert expr.numel() == ref.numel(), \
            f"rhs for {var_name} has {expr.shape}, state has {ref.shape}"
        if process_noise:
            w = self.set_variable("_w", var_name + "_noise", expr.shape)
            expr = expr + w
        self.rhs_list.append({"var_name": var_name, "expr": expr})

    def set_alg(self, expr_name: str, expr) -> None:
        assert self.flags["setup"] == False, "Cannot call .set_alg after .setup."
        expr = expr if isinstance(expr, SX) else SX(expr)
        self.alg_list.append((expr_name, expr))

    # -- setup -----------------------------------------------------------------------------------------------
    def setup(self) -> None:
        """Freeze the declarations: build the structures and the stage functions (reference :937-1060)."""
        assert self.flags["setup"] == False, "setup was already called"
        S = {}
        for vt in _VAR_TYPES:
            S[vt] = SymStruct([entry(n, sym=v) for n, v in self._vars[vt]])
        self._x, self._u, self._z, self._p, self._tvp, self._w, self._v = (S[v] for v in _VAR_TYPES)
        rhs_names = [r["var_name"] for r in self.rhs_list]
        x_names = [n for n, _ in self._vars["_x"]]
        for n in x_names:
            if n not in rhs_names:
                raise Exception(f"No right-hand side was set for state '{n}' (use model.set_rhs).")
        by_name = {r["var_name"]: r["expr"] for r in self.rhs_list}
        self._rhs = SymStruct([entry(n, expr=by_name[n]) for n in x_names], symbolic=False)
        self._alg = SymStruct([entry(n, expr=e) for n, e in self.alg_list], symbolic=False)
        self._aux_expression = SymStruct([entry(n, expr=e) for n, e in self._aux_list], symbolic=False)
        if not self._y_list:
            # state feedback default: y = x (reference :980-986)
            for n, v in self._vars["_x"]:
                self._y_list.append((n, v))
        self._y_expression = SymStruct([entry(n, expr=e) for n, e in self._y_list], symbolic=False)
        self._y = self._y_expression

        self.n_x, self.n_u, self.n_z = self._x.size, self._u.size, self._z.size
        self.n_p, self.n_tvp = self._p.size, self._tvp.size
        self.n_w, self.n_v = self._w.size, self._v.size
        self.n_y, self.n_aux = self._y_expression.size, self._aux_expression.size
        if self._alg.size != self.n_z:
            raise Exception(f"Number of algebraic equations ({self._alg.size}) must equal number of "
                            f"algebraic variables ({self.n_z}).")

        ins = [self._x.cat, self._u.cat, self._z.cat, self._tvp.cat, self._p.cat]
        names = ["_x", "_u", "_z", "_tvp", "_p"]
        self._rhs_fun = Function("rhs_fun", ins + [self._w.cat], [self._rhs.cat], names + ["_w"], ["_rhs"])
        self._alg_fun = Function("alg_fun", ins + [self._w.cat], [self._alg.cat], names + ["_w"], ["_alg"])
        self._aux_expression_fun = Function("aux_expression_fun", ins, [self._aux_expression.cat], names, ["_aux"])
        self._meas_fun = Function("meas_fun", ins + [self._v.cat], [self._y_expression.cat], names + ["_v"], ["_y"])
        self.flags["setup"] = True

    # -- access ------------------------------------------------------------------------------------------------
    def _struct(self, name):
        assert self.flags["setup"], "model.setup() must be called first"
        return getattr(self, name)

    def __getitem__(self, ind):
        if not isinstance(ind, tuple):
            ind = (ind,)
        # the structs themselves, like the reference (model/_model.py:164-191): `_x, _tvp = model['x', 'tvp']; _x['phi_2']`
        out = []
        for i in ind:
            name = i[1:] if i.startswith("_") else i
            assert name in ("x", "u", "z", "p", "tvp", "y", "aux", "w", "v"), f"The queried variable {i} is not valid."
            out.append(getattr(self, name))
        return out[0] if len(out) == 1 else out

    def _pre(self, vt):
        if self.flags["setup"]:
            return getattr(self, vt)
        return SymStruct([entry(n, sym=v) for n, v in self._vars[vt]])

    @property
    def x(self): return self._pre("_x")
    @property
    def u(self): return self._pre("_u")
    @property
    def z(self): return self._pre("_z")
    @property
    def p(self): return self._pre("_p")
    @property
    def tvp(self): return self._pre("_tvp")
    @property
    def w(self): return self._pre("_w")
    @property
    def v(self): return self._pre("_v")
    @property
    def y(self):
        return self._y_expression if self.flags["setup"] else \
            SymStruct([entry(n, expr=e) for n, e in self._y_list], symbolic=False)
    @property
    def aux(self):
        return self._aux_expression if self.flags["setup"] else \
            SymStruct([entry(n, expr=e) for n, e in self._aux_list], symbolic=False)

    def get_linear_system_matrices(self, xss=None, uss=None, z=None, tvp=None, p=None):
        """A = d rhs/dx, B = d rhs/du, C = d y/dx, D = d y/du (reference :1090-1144), optionally evaluated."""
        A = jacobian(self._rhs.cat, self._x.cat); B = jacobian(self._rhs.cat, self._u.cat)
        C = jacobian(self._y_expression.cat, self._x.cat); D = jacobian(self._y_expression.cat, self._u.cat)
        if xss is None and uss is None:
            return A, B, C, D
        ins = [self._x.cat, self._u.cat, self._z.cat, self._tvp.cat, self._p.cat, self._w.cat, self._v.cat]
        f = Function("lin", ins, [A, B, C, D])
        z = np.zeros(self.n_z) if z is None else z
        tvp = np.zeros(self.n_tvp) if tvp is None else tvp
        p = np.zeros(self.n_p) if p is None else p
        r = f(xss, uss, z, tvp, p, np.zeros(self.n_w), np.zeros(self.n_v))
        return tuple(m.full() for m in r)
