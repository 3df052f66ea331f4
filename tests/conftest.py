import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "examples"))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def dataset():
    d = np.load(GOLDEN / "cstr_dataset.npz")
    return {k: d[k] for k in d.files}


def build_hostsim(recipe, tag):
    """Compile the CPU emulation of the CUDA kernel source for a recipe (test tool; see tests/hostsim/hostsim.cpp)."""
    import ctypes
    from do_mpc_b200.solver.codegen import generate
    import hashlib
    src = generate(recipe)
    csrc = ROOT / "do_mpc_b200" / "csrc"
    h = hashlib.sha256(src.encode())
    for f in ("ipm_core.cuh", "ipm_solve.cuh", "ipm_refine.cuh", "ipm_sens.cuh", "dmb_host_util.h"):
        h.update((csrc / f).read_bytes())
    h.update((ROOT / "tests" / "hostsim" / "hostsim.cpp").read_bytes())
    extra = os.environ.get("DMB_HOSTSIM_FLAGS", "").split()          # experiments only (e.g. -DDMB_EMU_FP32_PROJ)
    h.update(" ".join(extra).encode())
    d = ROOT / "tests" / "hostsim" / "_build" / f"{tag}_{h.hexdigest()[:12]}"
    so = d / "hostsim.so"
    if not so.exists():
        d.mkdir(parents=True, exist_ok=True)
        (d / "problem_gen.h").write_text(src)
        tmp = d / f"hostsim.{os.getpid()}.tmp"
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", *extra, f"-I{d}", f"-I{csrc}",
                        str(ROOT / "tests" / "hostsim" / "hostsim.cpp"), "-o", str(tmp)], check=True)
        os.replace(tmp, so)                          # atomic: concurrent test processes may build the same emulation
    return ctypes.CDLL(str(so))


class HostSim:
    """ctypes front of tests/hostsim (same call shape as dmb_solve_host)."""

    def __init__(self, mpc, tag):
        import ctypes
        from do_mpc_b200.solver.options import DmbOptions
        self.mpc, self.r = mpc, mpc.recipe
        self.lib = build_hostsim(self.r, tag)
        dims = (ctypes.c_int * 16)()
        self.lib.hostsim_dims(dims)
        assert dims[9] == ctypes.sizeof(DmbOptions)
        self.o = DmbOptions()
        self.lib.hostsim_default_options(ctypes.byref(self.o))
        self.lbx = np.ascontiguousarray(mpc._lb_opt_x.full().ravel())
        self.ubx = np.ascontiguousarray(mpc._ub_opt_x.full().ravel())

    def solve(self, P, x0s=None, u0s=None, z0s=None, x_init=None, lam_g0=None, lam_x0=None):
        import ctypes
        r = self.r
        P = np.ascontiguousarray(P, dtype=float); B = P.shape[0]
        c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=float)
        x_init, x0s, u0s, z0s, lam_g0, lam_x0 = c(x_init), c(x0s), c(u0s), c(z0s), c(lam_g0), c(lam_x0)
        out = dict(x=np.zeros((B, r.n_w)), lam_g=np.zeros((B, r.n_g)), lam_x=np.zeros((B, r.n_w)), g=np.zeros((B, r.n_g)),
                   f=np.zeros(B), status=np.zeros(B, np.int32), iter_count=np.zeros(B, np.int32), u0_scaled=np.zeros((B, r.n_u)))
        p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        self.lib.hostsim_solve_ws(B, p(x_init), p(lam_g0), p(lam_x0), p(x0s), p(u0s), p(z0s), p(P), p(self.lbx), p(self.ubx), ctypes.byref(self.o),
                               p(out["x"]), p(out["lam_g"]), p(out["lam_x"]), p(out["g"]), p(out["f"]), p(out["status"]),
                               p(out["iter_count"]), p(out["u0_scaled"]))
        if x_init is None:
            # unused reference variables keep the broadcast initial guess
            xg = np.zeros((B, r.n_w))
            from do_mpc_b200.sym.struct import NumStruct
            for b in range(B):
                s = self.mpc._opt_x(0.0)
                s["_x"] = x0s[b][:r.n_x - r.n_lift]; s["_u"] = u0s[b]
                if r.n_z: s["_z"] = z0s[b]
                xg[b] = s.full().ravel()
            used = np.zeros(r.n_w, bool)
            w2f = self.w2flat()
            used[w2f[w2f >= 0]] = True
            out["x"][:, ~used] = xg[:, ~used]
        out["success"] = (out["status"] == 0) | (out["status"] == 1)
        return out

    def sens(self, x, lam_g, P, cols=None, tol=1e-6, big=1e10, want_dlam=False, relaxed_rule=0):
        """dx [B, ncols, n_w] (scaled opt_x) / d opt_p[col] at solved instances (same phases as dmb_sens_kernel)."""
        import ctypes
        r = self.r
        x = np.ascontiguousarray(x, dtype=float); lam_g = np.ascontiguousarray(lam_g, dtype=float)
        P = np.ascontiguousarray(P, dtype=float); B = P.shape[0]
        cols = np.arange(r.n_p, dtype=np.int32) if cols is None else np.ascontiguousarray(cols, dtype=np.int32)
        dx = np.zeros((B, len(cols), r.n_w)); sst = np.zeros(B, np.int32)
        dl = np.zeros((B, len(cols), r.n_g + r.n_w)) if want_dlam else None
        p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        self.lib.hostsim_sens(B, p(x), p(lam_g), p(P), p(self.lbx), p(self.ubx), ctypes.byref(self.o), ctypes.c_double(tol),
                              ctypes.c_double(big), int(relaxed_rule), p(cols), len(cols), p(dx), p(dl), p(sst))
        return (dx, dl, sst) if want_dlam else (dx, sst)

    def w2flat(self):
        import re
        from do_mpc_b200.solver.codegen import generate
        m = re.search(r"#define DMB_W2FLAT_INIT \{([^}]*)\}", generate(self.r))
        return np.array([int(v) for v in m.group(1).split(",")])


def cstr_batch_inputs(recipe, x0, u_prev):
    B = x0.shape[0]
    P = np.zeros((B, recipe.n_p))
    n_xm = recipe.n_x - recipe.n_lift
    P[:, recipe.p_x0:recipe.p_x0 + n_xm] = x0
    P[:, recipe.p_uprev:recipe.p_uprev + recipe.n_u] = u_prev
    x0s = np.concatenate([x0 / recipe.x_scaling[:n_xm], np.zeros((B, recipe.n_lift))], axis=1)
    return P, x0s, u_prev / recipe.u_scaling
