"""BatchedNlpSolver — the object installed as `mpc.S` in place of CasADi's nlpsol Function.

Reference boundary: `self.S = castools.nlpsol('S','ipopt',nlp,opts)` (do_mpc/controller/_mpc.py:1328),
called as `r = self.S(x0=,lbx=,ubx=,lbg=,ubg=,p=[,lam_x0=,lam_g0=])` and `self.S.stats()` (do_mpc/optimizer.py:754-778).
This class keeps that call signature and return keys, and adds the batch entry points.  All numerics run in
the generated CUDA library through the C ABI of include/dompc_b200.h (ctypes); there is NO CPU fallback —
constructing the solver without a visible B200 raises.
"""
from __future__ import annotations

import ctypes
import time
from typing import Optional

import numpy as np

from ..sym.sx import DM
from .build import build_library, latency_threads
from .options import DmbOptions, apply_nlpsol_opts

_STATUS_NAMES = {0: "Solve_Succeeded", 1: "Solved_To_Acceptable_Level", 2: "Infeasible_Problem_Detected",
                 -1: "Maximum_Iterations_Exceeded", -2: "Restoration_Failed", -3: "Error_In_Step_Computation",
                 -13: "Invalid_Number_Detected", -100: "Not_Run"}
(DIM_NW, DIM_NG, DIM_NP, DIM_NX, DIM_NU, DIM_NZ, DIM_NAUX, DIM_NHOR, DIM_NSCEN, DIM_NEDGE, DIM_NNODE, DIM_OPTSIZE,
 DIM_THREADS, DIM_SMEM, DIM_WS, DIM_COUNT) = range(16)
NSTAT = 26
import os as _os
if "DMB_NSTAT" in _os.environ:          # A/B experiments against an older kernel source (21 statistics per instance)
    NSTAT = int(_os.environ["DMB_NSTAT"])
STAT_N_RESTO, STAT_N_RESTO_IT, STAT_N_SOC, STAT_N_REFINE, STAT_N_LS = 21, 22, 23, 24, 25
PHASE_NAMES = ['derivs', 'kkt+mu', 'factor', 'riccati', 'forward+recover', 'stepsize', 'linesearch', 'accept', 'init']

_c_dp = ctypes.c_void_p


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _as2d(a, n):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1, n))


def load_library(path):
    lib = ctypes.CDLL(str(path))
    lib.dmb_abi_version.restype = ctypes.c_int
    lib.dmb_create.restype = ctypes.c_void_p
    lib.dmb_create.argtypes = [ctypes.c_int]
    lib.dmb_destroy.argtypes = [ctypes.c_void_p]
    lib.dmb_last_error.restype = ctypes.c_char_p
    lib.dmb_last_error.argtypes = [ctypes.c_void_p]
    lib.dmb_set_options.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    lib.dmb_set_bounds.argtypes = [ctypes.c_void_p, _c_dp, _c_dp]
    lib.dmb_solve_host.argtypes = [ctypes.c_void_p, ctypes.c_int] + [_c_dp] * 14
    lib.dmb_solve_device.argtypes = [ctypes.c_void_p, ctypes.c_int] + [_c_dp] * 15
    lib.dmb_solve_host_ws.argtypes = [ctypes.c_void_p, ctypes.c_int] + [_c_dp] * 16
    lib.dmb_solve_device_ws.argtypes = [ctypes.c_void_p, ctypes.c_int] + [_c_dp] * 17
    lib.dmb_aux_host.argtypes = [ctypes.c_void_p, ctypes.c_int, _c_dp, _c_dp, _c_dp]
    lib.dmb_get_stats.argtypes = [ctypes.c_void_p] + [_c_dp] * 5
    lib.dmb_set_sens_options.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_int]
    lib.dmb_sens_host.argtypes = [ctypes.c_void_p, ctypes.c_int, _c_dp, _c_dp, _c_dp, _c_dp, ctypes.c_int, _c_dp, ctypes.c_int,
                                  _c_dp, _c_dp, _c_dp]
    lib.dmb_sens_device.argtypes = [ctypes.c_void_p, ctypes.c_int, _c_dp, _c_dp, _c_dp, _c_dp, ctypes.c_int, _c_dp, ctypes.c_int,
                                    _c_dp, _c_dp, _c_dp, _c_dp]
    return lib


class BatchedNlpSolver:
    """One generated library = one CTA size.  The solver owns the *throughput* build (several instances per SM) and, where a
    larger CTA shortens the edge loops of an instance (build.latency_threads), creates a *latency* build on first use: batches of
    at most `latency_max_batch` instances — the reference's real-time `make_step(x0)` is B = 1 — are solved by it with one
    instance per SM.  Same kernel source, same options and bounds; `launch_info()` reports the build used last."""

    def __init__(self, recipe, nlpsol_opts=None, device: Optional[int] = None, _threads: Optional[int] = None, _src: Optional[str] = None):
        import os
        from .codegen import generate
        self.recipe = recipe
        src = self._src = generate(recipe) if _src is None else _src
        self.lib_path = build_library(recipe, src=src, threads=_threads)
        self._lat: Optional["BatchedNlpSolver"] = None          # the latency build, created lazily
        self._lat_threads = None
        self.latency_max_batch = 0
        if _threads is None:
            self.latency_max_batch = int(os.environ.get("DMB_LATENCY_MAX_BATCH", "16"))
            if self.latency_max_batch > 0:
                self._lat_threads = latency_threads(recipe, src)
        self._last = self
        self.lib = load_library(self.lib_path)
        dims = (ctypes.c_int * DIM_COUNT)()
        self.lib.dmb_get_dims(dims)
        self.dims = list(dims)
        self.n_w, self.n_g, self.n_p = self.dims[DIM_NW], self.dims[DIM_NG], self.dims[DIM_NP]
        self.n_x, self.n_u, self.n_z, self.n_aux = (self.dims[i] for i in (DIM_NX, DIM_NU, DIM_NZ, DIM_NAUX))
        assert (self.n_w, self.n_g, self.n_p) == (recipe.n_w, recipe.n_g, recipe.n_p)
        assert self.dims[DIM_OPTSIZE] == ctypes.sizeof(DmbOptions), "DmbOptions layout mismatch between Python and CUDA"
        self.options = DmbOptions()
        self.lib.dmb_default_options(ctypes.byref(self.options))
        apply_nlpsol_opts(self.options, nlpsol_opts)
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) if "DMB_DEVICE" not in os.environ else int(os.environ["DMB_DEVICE"])
        self.device = device
        self._h = self.lib.dmb_create(device)
        if not self._h:
            raise RuntimeError(
                "do_mpc_b200: could not create the CUDA solver (no usable GPU / sm_100a kernel image cannot be loaded). "
                "This package has no CPU fallback for the NLP solve.")
        rc = self.lib.dmb_set_options(self._h, ctypes.byref(self.options), ctypes.sizeof(DmbOptions))
        self._check(rc)
        self._bounds_key = None
        self.bounds_source = None        # callable -> (lbx, ubx); set by MPC.create_nlp so that the device entry points see `mpc.bounds[...] = v`
        self._stats = {"success": False, "return_status": "Not_Run", "iter_count": 0, "t_wall_total": 0.0}
        self.n_launches_timed = 0

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self.lib.dmb_destroy(h)
            except Exception:
                pass
            self._h = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"do_mpc_b200 CUDA solver error {rc}: {self.lib.dmb_last_error(self._h).decode()}")

    def set_options(self, **kw):
        for k, v in kw.items():
            setattr(self.options, k, v)
        self._check(self.lib.dmb_set_options(self._h, ctypes.byref(self.options), ctypes.sizeof(DmbOptions)))
        self._bounds_key = None
        if self._lat is not None:
            self._lat.set_options(**kw)

    def _for_batch(self, B):
        """The build that solves a batch of B instances: the latency build for small batches (if the problem has one)."""
        if self._lat_threads is None or B > self.latency_max_batch:
            self._last = self
            return self
        if self._lat is None:
            self._lat = BatchedNlpSolver(self.recipe, None, self.device, _threads=self._lat_threads, _src=self._src)
            ctypes.memmove(ctypes.byref(self._lat.options), ctypes.byref(self.options), ctypes.sizeof(DmbOptions))
            self._lat.set_options()
            self._lat.bounds_source = self.bounds_source
        if self._bounds_key is not None and self._lat._bounds_key != self._bounds_key:
            self._lat._set_bounds(np.frombuffer(self._bounds_key[0]), np.frombuffer(self._bounds_key[1]))
        self._last = self._lat
        return self._lat

    def _set_bounds(self, lbx, ubx):
        lbx = np.ascontiguousarray(np.asarray(lbx, dtype=np.float64).ravel())
        ubx = np.ascontiguousarray(np.asarray(ubx, dtype=np.float64).ravel())
        key = (lbx.tobytes(), ubx.tobytes())
        if key != self._bounds_key:
            assert lbx.size == self.n_w and ubx.size == self.n_w, "bounds must have n_w entries"
            self._check(self.lib.dmb_set_bounds(self._h, _ptr(lbx), _ptr(ubx)))
            self._bounds_key = key

    def _refresh_bounds(self, lbx=None, ubx=None):
        """Bounds for the device entry points: explicit arguments, else the owner's current bounds (`bounds_source`).
        One launch per handle may be in flight at a time (workspace slots, work counter and staging are per handle)."""
        if lbx is None and self.bounds_source is not None:
            lbx, ubx = self.bounds_source()
        if lbx is not None:
            self._set_bounds(lbx, ubx)
        elif self._bounds_key is None:
            raise RuntimeError("do_mpc_b200: no bounds known to the solver — pass lbx/ubx or call _set_bounds first")

    # ---- batched host entry point --------------------------------------------------------------------------------
    def solve_batch(self, P, x0s=None, u0s=None, z0s=None, lbx=None, ubx=None, x_init=None, return_full=False, lam_g0=None, lam_x0=None):
        """P [B,n_p]; initial guess = x_init [B,n_w] (scaled) or broadcast of x0s,u0s,z0s (scaled). Host arrays.
        lam_g0 / lam_x0: multipliers of a previous solution, used only with the option warm_start (ipopt.warm_start_init_point)."""
        P = _as2d(P, self.n_p)
        B = P.shape[0]
        if lbx is not None:
            self._set_bounds(lbx, ubx)
        S = self._for_batch(B)
        if S is not self:
            return S.solve_batch(P, x0s, u0s, z0s, None, None, x_init, return_full, lam_g0, lam_x0)
        x_init = None if x_init is None else _as2d(x_init, self.n_w)
        x0s = None if x0s is None else _as2d(x0s, self.n_x)
        u0s = None if u0s is None else _as2d(u0s, self.n_u)
        z0s = None if (z0s is None or self.n_z == 0) else _as2d(z0s, self.n_z)
        use_ws = bool(self.options.warm_start) and lam_g0 is not None and lam_x0 is not None and x_init is not None
        lam_g0 = _as2d(lam_g0, self.n_g) if use_ws else None
        lam_x0 = _as2d(lam_x0, self.n_w) if use_ws else None
        f = np.empty(B); status = np.empty(B, dtype=np.int32); iters = np.empty(B, dtype=np.int32)
        u0 = np.empty((B, self.n_u)); stats = np.empty((B, NSTAT))
        x = lam_g = lam_x = g = None
        if return_full:
            x = np.empty((B, self.n_w)); lam_g = np.empty((B, self.n_g)); lam_x = np.empty((B, self.n_w)); g = np.empty((B, self.n_g))
        t0 = time.perf_counter()
        rc = self.lib.dmb_solve_host_ws(self._h, B, _ptr(x_init), _ptr(lam_g0), _ptr(lam_x0), _ptr(x0s), _ptr(u0s), _ptr(z0s), _ptr(P), _ptr(x),
                                     _ptr(lam_g), _ptr(lam_x), _ptr(g), _ptr(f), _ptr(status), _ptr(iters), _ptr(u0), _ptr(stats))
        self._check(rc)
        t = time.perf_counter() - t0
        out = {"u0_scaled": u0, "f": f, "status": status, "iter_count": iters, "success": (status == 0) | (status == 1),
               "stats": stats, "t_wall_total": t}
        if return_full:
            out.update(x=x, lam_g=lam_g, lam_x=lam_x, g=g)
        return out

    # ---- device-resident entry point (torch tensors as memory containers) ----------------------------------------------
    def solve_batch_device(self, P, x0s=None, u0s=None, z0s=None, x_init=None, out=None, stream=None, want_full=False, lam_g0=None, lam_x0=None,
                           lbx=None, ubx=None):
        """All arguments are CUDA float64 torch tensors on this solver's device; returns dict of CUDA tensors.
        Asynchronous with respect to the host (kernel queued on `stream` or the current torch stream).  Bounds (host arrays,
        shared by the batch): `lbx/ubx`, else the owning MPC's current bounds; uploaded only when they changed."""
        import torch
        self._refresh_bounds(lbx, ubx)
        B = P.shape[0]
        S = self._for_batch(B)
        if S is not self:
            return S.solve_batch_device(P, x0s, u0s, z0s, x_init, out, stream, want_full, lam_g0, lam_x0)
        dev = P.device
        def dp(t): return None if t is None else ctypes.c_void_p(t.data_ptr())
        if out is None:
            out = {"f": torch.empty(B, dtype=torch.float64, device=dev), "status": torch.empty(B, dtype=torch.int32, device=dev),
                   "iter_count": torch.empty(B, dtype=torch.int32, device=dev),
                   "u0_scaled": torch.empty(B, self.n_u, dtype=torch.float64, device=dev),
                   "stats": torch.empty(B, NSTAT, dtype=torch.float64, device=dev)}
            if want_full:
                out.update(x=torch.empty(B, self.n_w, dtype=torch.float64, device=dev),
                           lam_g=torch.empty(B, self.n_g, dtype=torch.float64, device=dev),
                           lam_x=torch.empty(B, self.n_w, dtype=torch.float64, device=dev),
                           g=torch.empty(B, self.n_g, dtype=torch.float64, device=dev))
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        if not (self.options.warm_start and lam_g0 is not None and lam_x0 is not None and x_init is not None):
            lam_g0 = lam_x0 = None
        rc = self.lib.dmb_solve_device_ws(self._h, B, dp(x_init), dp(lam_g0), dp(lam_x0), dp(x0s), dp(u0s), dp(z0s), dp(P), dp(out.get("x")),
                                       dp(out.get("lam_g")), dp(out.get("lam_x")), dp(out.get("g")), dp(out["f"]),
                                       dp(out["status"]), dp(out["iter_count"]), dp(out["u0_scaled"]), dp(out["stats"]),
                                       ctypes.c_void_p(st.cuda_stream))
        self._check(rc)
        return out

    # ---- nlpsol-Function look-alike ------------------------------------------------------------------------------------------------
    def __call__(self, x0=None, lbx=None, ubx=None, lbg=None, ubg=None, p=None, lam_x0=None, lam_g0=None):
        """Single-instance call with the keyword signature of the CasADi nlpsol Function (optimizer.py:754-770).
        lam_x0/lam_g0 are used with ipopt.warm_start_init_point=yes and ignored otherwise, like IPOPT."""
        def arr(v):
            return np.asarray(v.full() if hasattr(v, "full") else v, dtype=np.float64).ravel()
        x0, p = arr(x0), arr(p)
        if x0.size != self.n_w or p.size != self.n_p:
            raise ValueError(f"nlpsol call: x0 has {x0.size} (expected {self.n_w}), p has {p.size} (expected {self.n_p})")
        ws = bool(self.options.warm_start) and lam_x0 is not None and lam_g0 is not None
        r = self.solve_batch(p[None, :], x_init=x0[None, :], lbx=arr(lbx), ubx=arr(ubx), return_full=True,
                             lam_g0=arr(lam_g0)[None, :] if ws else None, lam_x0=arr(lam_x0)[None, :] if ws else None)
        st = int(r["status"][0])
        self._stats = {"success": bool(r["success"][0]), "return_status": _STATUS_NAMES.get(st, str(st)),
                       "iter_count": int(r["iter_count"][0]), "t_wall_total": r["t_wall_total"],
                       "t_wall_S": r["t_wall_total"], "unified_return_status": "SOLVER_RET_SUCCESS" if r["success"][0] else "SOLVER_RET_UNKNOWN"}
        return {"x": DM(r["x"][0]), "f": DM(r["f"][0]), "g": DM(r["g"][0]), "lam_g": DM(r["lam_g"][0]),
                "lam_x": DM(r["lam_x"][0]), "lam_p": DM(np.zeros(self.n_p))}

    def stats(self):
        return dict(self._stats)

    def aux(self, x, p):
        x = _as2d(x.full() if hasattr(x, "full") else x, self.n_w); p = _as2d(p.full() if hasattr(p, "full") else p, self.n_p)
        B = x.shape[0]
        out = np.zeros((B, self.dims[DIM_NHOR] * self.dims[DIM_NSCEN] * self.n_aux))
        self._check(self.lib.dmb_aux_host(self._h, B, _ptr(x), _ptr(p), _ptr(out)))
        return out[0] if B == 1 else out

    # ---- parametric sensitivities of solved instances (reference: NLPDifferentiator.differentiate) --------------------------
    def sensitivities(self, x, lam_g, P, status=None, cols=None, u0_only=False, want_dlam=False, active_set_tol=1e-6,
                      active_weight=1e10, relaxed_bounds_rule=False):
        """d opt_x (scaled) / d opt_p[cols] at solved instances: x [B,n_w], lam_g [B,n_g], P [B,n_p] as returned by /
        passed to solve_batch.  Returns dict(dxdp [B, n_w or n_u, ncols], sens_status [B] [, dlamdp [B, n_g+n_w, ncols]])."""
        x, lam_g, P = _as2d(x, self.n_w), _as2d(lam_g, self.n_g), _as2d(P, self.n_p)
        B = P.shape[0]
        cols_a = None if cols is None else np.ascontiguousarray(np.asarray(cols, dtype=np.int32).ravel())
        nc = self.n_p if cols_a is None else cols_a.size
        nrow = self.n_u if u0_only else self.n_w
        dx = np.empty((B, nc, nrow)); sst = np.empty(B, dtype=np.int32)
        dl = np.empty((B, nc, self.n_g + self.n_w)) if (want_dlam and not u0_only) else None
        st = None if status is None else np.ascontiguousarray(np.asarray(status, dtype=np.int32).ravel())
        self._check(self.lib.dmb_set_sens_options(self._h, float(active_set_tol), float(active_weight), int(bool(relaxed_bounds_rule))))
        self._check(self.lib.dmb_sens_host(self._h, B, _ptr(x), _ptr(lam_g), _ptr(P), _ptr(st), 1 if u0_only else 0, _ptr(cols_a),
                                           nc, _ptr(dx), _ptr(dl), _ptr(sst)))
        out = {"dxdp": dx.transpose(0, 2, 1), "sens_status": sst}
        if dl is not None:
            out["dlamdp"] = dl.transpose(0, 2, 1)
        return out

    def sensitivities_device(self, x, lam_g, P, status=None, cols=None, u0_only=True, stream=None):
        """Device-resident variant (CUDA float64 torch tensors); returns dxdp as [B, ncols, n_u or n_w] (kernel layout)."""
        import torch
        self._refresh_bounds()
        B, dev = P.shape[0], P.device
        def dp(t): return None if t is None else ctypes.c_void_p(t.data_ptr())
        cols_a = None if cols is None else np.ascontiguousarray(np.asarray(cols, dtype=np.int32).ravel())
        nc = self.n_p if cols_a is None else cols_a.size
        dx = torch.empty(B, nc, self.n_u if u0_only else self.n_w, dtype=torch.float64, device=dev)
        sst = torch.empty(B, dtype=torch.int32, device=dev)
        st = stream if stream is not None else torch.cuda.current_stream(dev)
        self._check(self.lib.dmb_sens_device(self._h, B, dp(x), dp(lam_g), dp(P), dp(status), 1 if u0_only else 0, _ptr(cols_a), nc,
                                             dp(dx), None, dp(sst), ctypes.c_void_p(st.cuda_stream)))
        return {"dxdp_t": dx, "sens_status": sst}

    def launch_info(self):
        if self._last is not self:
            return self._last.launch_info()
        grid = ctypes.c_int(); block = ctypes.c_int(); smem = ctypes.c_int(); n = ctypes.c_longlong(); ms = ctypes.c_float()
        self.lib.dmb_get_stats(self._h, ctypes.byref(grid), ctypes.byref(block), ctypes.byref(smem), ctypes.byref(n), ctypes.byref(ms))
        return {"grid": grid.value, "block": block.value, "smem_bytes": smem.value, "n_launches": n.value, "last_kernel_ms": ms.value}
