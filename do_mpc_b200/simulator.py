"""Batched plant simulator with the do_mpc.simulator.Simulator API (first "next" row).

API mirror of reference do_mpc/simulator.py (class Simulator :106; set_param :426, get/set_tvp_fun :459-522,
get/set_p_fun :524-601, setup :321, init_algebraic_variables :621, make_step :757).  The reference integrates with
SUNDIALS CVODES / IDAS (abstol = reltol = 1e-10); here one CUDA thread integrates one instance with a fixed-step
3-stage Radau IIA collocation method (order 5, `n_substeps` steps per t_step, full Newton with partial pivoting) on
generated model code (csrc/sim_core.cuh).  `make_step_batch` advances B plants at once.  No CPU fallback.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from .data import Data
from .model import Model
from .sym.struct import NumStruct, SymStruct, entry


@dataclass
class SimulatorSettings:
    t_step: float = None
    integration_tool: str = "radau5_fixed"        # accepted for compatibility: 'cvodes' / 'idas' select the same integrator
    abstol: float = 1e-10
    reltol: float = 1e-10
    n_substeps: int = 20
    newton_tol: float = 1e-13
    newton_max_iter: int = 25

    def check_for_mandatory_settings(self):
        if self.t_step is None:
            raise ValueError("t_step must be set")


def _load(path):
    lib = ctypes.CDLL(str(path))
    lib.dms_create.restype = ctypes.c_void_p
    lib.dms_create.argtypes = [ctypes.c_int]
    lib.dms_destroy.argtypes = [ctypes.c_void_p]
    lib.dms_last_error.restype = ctypes.c_char_p
    lib.dms_last_error.argtypes = [ctypes.c_void_p]
    lib.dms_launch_count.restype = ctypes.c_longlong
    lib.dms_launch_count.argtypes = [ctypes.c_void_p]
    lib.dms_step_host.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 5 + \
        [ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4
    lib.dms_step_device.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 5 + \
        [ctypes.c_double, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 5
    return lib


class Simulator:
    def __init__(self, model: Model):
        assert model.flags["setup"] == True, "Model for simulator was not setup. After the complete model creation call model.setup()."
        self.model = model
        self.settings = SimulatorSettings()
        self.data = Data(model)
        self.flags = {"setup": False, "set_tvp_fun": False, "set_p_fun": False}
        self._x0, self._u0, self._z0, self._t0 = model._x(0.0), model._u(0.0), model._z(0.0), np.array([0.0])
        self._h = None

    def set_param(self, **kwargs):
        assert self.flags["setup"] == False, "Setting parameters after setup is prohibited."
        for k, v in kwargs.items():
            if hasattr(self.settings, k):
                setattr(self.settings, k, v)
            else:
                print(f"Warning: Key {k} does not exist for Simulator.")

    # ---- parameters -------------------------------------------------------------------------------------------------------
    def get_tvp_template(self) -> NumStruct: return self.model._tvp(0.0)
    def get_p_template(self) -> NumStruct: return self.model._p(0.0)

    def set_tvp_fun(self, tvp_fun: Callable):
        assert self.get_tvp_template().labels() == tvp_fun(0).labels(), "Incorrect output of tvp_fun. Use get_tvp_template."
        self.tvp_fun = tvp_fun; self.flags["set_tvp_fun"] = True

    def set_p_fun(self, p_fun: Callable):
        assert self.get_p_template().labels() == p_fun(0).labels(), "Incorrect output of p_fun. Use get_p_template."
        self.p_fun = p_fun; self.flags["set_p_fun"] = True

    def setup(self) -> None:
        from .solver.build import build_sim_library
        self.settings.check_for_mandatory_settings()
        if self.model.n_tvp > 0 and not self.flags["set_tvp_fun"]:
            raise Exception("You have not supplied a function to obtain the time-varying parameters defined in model. Use .set_tvp_fun() prior to setup.")
        if self.model.n_p > 0 and not self.flags["set_p_fun"]:
            raise Exception("You have not supplied a function to obtain the parameters defined in model. Use .set_p_fun() prior to setup.")
        if not hasattr(self, "tvp_fun"):
            t = self.get_tvp_template(); self.tvp_fun = lambda t_now: t
        if not hasattr(self, "p_fun"):
            p = self.get_p_template(); self.p_fun = lambda t_now: p
        self.lib_path = build_sim_library(self.model)
        self.lib = _load(self.lib_path)
        import os
        self._h = self.lib.dms_create(int(os.environ.get("DMB_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
        if not self._h:
            raise RuntimeError("do_mpc_b200: could not create the CUDA simulator (no usable GPU). There is no CPU fallback.")
        self.flags["setup"] = True

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try: self.lib.dms_destroy(h)
            except Exception: pass

    def _num_prop(name):
        def g(self): return getattr(self, name)
        def s(self, val):
            cur = getattr(self, name)
            if isinstance(val, NumStruct): val = val.full()
            a = np.asarray(val.full() if hasattr(val, "full") else val, dtype=float).ravel()
            assert a.size == cur.size, f"wrong size {a.size}, expected {cur.size}"
            setattr(self, name, cur(a))
        return property(g, s)
    x0 = _num_prop("_x0"); u0 = _num_prop("_u0"); z0 = _num_prop("_z0")

    @property
    def t0(self): return self._t0
    @t0.setter
    def t0(self, v): self._t0 = np.asarray(v, dtype=float).reshape(1)

    def reset_history(self): self.data.init_storage(); self._t0 = np.array([0.0])
    def set_initial_guess(self): pass

    # ---- compute --------------------------------------------------------------------------------------------------------------
    def _pw(self, PP, W, B):
        """parameter row of the C ABI: [p | w] (process noise is constant over the step like a parameter)."""
        m = self.model
        if m.n_p + m.n_w == 0:
            return None
        out = np.zeros((B, m.n_p + m.n_w))
        if m.n_p:
            out[:, :m.n_p] = np.asarray(PP, dtype=np.float64).reshape(-1, m.n_p)
        if m.n_w and W is not None:
            out[:, m.n_p:] = np.asarray(W, dtype=np.float64).reshape(-1, m.n_w)
        return out

    def _call(self, X, U, Z, TV, PP, mode, W=None):
        m, s = self.model, self.settings
        c = lambda a, n: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, n)) if n else None
        X = c(X, m.n_x); B = X.shape[0]
        U, Z, TV = c(U, m.n_u), c(Z, m.n_z), c(TV, m.n_tvp)
        PP = self._pw(PP, W, B)
        xn = np.empty((B, m.n_x)); zn = np.empty((B, max(m.n_z, 1))); aux = np.empty((B, max(m.n_aux, 1))); st = np.empty(B, dtype=np.int32)
        p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        rc = self.lib.dms_step_host(self._h, B, p(X), p(U), p(Z), p(TV), p(PP), float(s.t_step), int(s.n_substeps),
                                    float(s.newton_tol), int(s.newton_max_iter), int(mode), p(xn), p(zn), p(aux), p(st))
        if rc != 0:
            raise RuntimeError(f"do_mpc_b200 simulator error {rc}: {self.lib.dms_last_error(self._h).decode()}")
        return xn, zn[:, :m.n_z], aux[:, :m.n_aux], st

    def make_step_batch(self, X, U, Z=None, tvp=None, p=None, W=None):
        """Advance B plants by one t_step: X [B,n_x], U [B,n_u] (physical units), optional process noise W [B,n_w]
        -> (X_next, Z_next, status)."""
        m = self.model
        B = np.asarray(X).reshape(-1, m.n_x).shape[0]
        Z = np.tile(self._z0.full().ravel(), (B, 1)) if (Z is None and m.n_z) else Z
        tv = np.tile(self.tvp_fun(self._t0).full().ravel(), (B, 1)) if (tvp is None and m.n_tvp) else tvp
        pp = np.tile(self.p_fun(self._t0).full().ravel(), (B, 1)) if (p is None and m.n_p) else p
        xn, zn, _, st = self._call(X, U, Z, tv, pp, 0, W)
        return xn, zn, st

    def make_step_batch_device(self, X, U, Z=None, tvp=None, p=None, out=None, stream=None, W=None):
        """Device-resident step: CUDA float64 torch tensors in, dict(x, z, status) of CUDA tensors out; queued on `stream`
        (default: torch's current stream), no host synchronisation."""
        import torch
        m, s = self.model, self.settings
        B, dev = X.shape[0], X.device
        dp = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
        if m.n_tvp and tvp is None:
            tvp = torch.as_tensor(self.tvp_fun(self._t0).full().ravel(), dtype=torch.float64, device=dev).repeat(B, 1)
        if m.n_p and p is None:
            p = torch.as_tensor(self.p_fun(self._t0).full().ravel(), dtype=torch.float64, device=dev).repeat(B, 1)
        if m.n_z and Z is None:
            Z = torch.as_tensor(self._z0.full().ravel(), dtype=torch.float64, device=dev).repeat(B, 1)
        if m.n_w:                                      # C ABI parameter row = [p | w]
            W = torch.zeros(B, m.n_w, dtype=torch.float64, device=dev) if W is None else W
            p = torch.cat([p, W], dim=1).contiguous() if m.n_p else W.contiguous()
        if out is None:
            out = {"x": torch.empty(B, m.n_x, dtype=torch.float64, device=dev), "z": torch.empty(B, max(m.n_z, 1), dtype=torch.float64, device=dev),
                   "status": torch.empty(B, dtype=torch.int32, device=dev)}
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        rc = self.lib.dms_step_device(self._h, B, dp(X), dp(U), dp(Z), dp(tvp), dp(p), float(s.t_step), int(s.n_substeps), float(s.newton_tol),
                                      int(s.newton_max_iter), 0, dp(out["x"]), dp(out["z"]), None, dp(out["status"]), ctypes.c_void_p(st.cuda_stream))
        if rc != 0:
            raise RuntimeError(f"do_mpc_b200 simulator error {rc}: {self.lib.dms_last_error(self._h).decode()}")
        return out

    def init_algebraic_variables(self) -> NumStruct:
        if self.model.n_z == 0:
            return self._z0
        xn, zn, _, st = self._call(self._x0.full().ravel(), self._u0.full().ravel(), self._z0.full().ravel(),
                                   self.tvp_fun(self._t0).full().ravel(), self.p_fun(self._t0).full().ravel(), 1)
        if st[0] != 0:
            raise RuntimeError("init_algebraic_variables: Newton iteration on the algebraic equations failed")
        self._z0 = self.model._z(zn[0])
        return self._z0

    def make_step(self, u0=None, v0=None, w0=None) -> np.ndarray:
        assert self.flags["setup"] == True, "Simulator is not setup. Call simulator.setup() first."
        m = self.model
        if u0 is None:
            assert m.n_u == 0, "No input u0 provided. Please provide an input u0."
            u0 = np.zeros((0, 1))
        u0 = np.asarray(u0.full() if hasattr(u0, "full") else u0, dtype=float)
        assert u0.size == m.n_u, f"u0 has incorrect shape. You have: {u0.shape}, expected: {(m.n_u, 1)}"
        tvp0, p0, t0 = self.tvp_fun(self._t0), self.p_fun(self._t0), self._t0
        x0 = self._x0.full().ravel()
        w = None
        if w0 is not None:
            w = np.asarray(w0.full() if hasattr(w0, "full") else w0, dtype=float).ravel()
            assert w.size == m.n_w, f"w0 has incorrect shape. You have: {w.shape}, expected: {(m.n_w, 1)}"
        xn, zn, aux, st = self._call(x0, u0.ravel(), self._z0.full().ravel(), tvp0.full().ravel(), p0.full().ravel(), 0, w)
        if st[0] != 0:
            raise RuntimeError("simulator.make_step: Newton iteration of the integrator failed")
        v = np.zeros(m.n_v) if v0 is None else np.asarray(v0, dtype=float).ravel()
        y = np.asarray(m._meas_fun(xn[0], u0.ravel(), zn[0] if m.n_z else np.zeros(0), tvp0.full().ravel(), p0.full().ravel(), v).full())
        d = self.data
        d.update(_x=x0, _u=u0, _time=t0, _y=y, _aux=aux[0])
        if m.n_z: d.update(_z=self._z0.full())
        if m.n_tvp: d.update(_tvp=tvp0.full())
        if m.n_p: d.update(_p=p0.full())
        self._x0 = m._x(xn[0]); self._u0 = m._u(u0.ravel())
        if m.n_z: self._z0 = m._z(zn[0])
        self._t0 = self._t0 + self.settings.t_step
        return y
