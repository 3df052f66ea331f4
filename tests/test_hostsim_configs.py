"""BASELINE configurations at their stated sizes: the kernel's algorithm source (CPU emulation, tests/hostsim) against the
oracle fixtures (oracle/ipm.py with second-order correction and restoration ON).  A few instances per config here — the CPU
emulation is single threaded; tests/test_gpu_configs.py runs all 64 per config through the C ABI on the GPU."""
import numpy as np
import pytest

from conftest import HostSim
from config_cases import MAX_ITER, compare_with_oracle, config_mpc, fixture

CASES = {1: np.arange(64), 3: np.array([0, 1]), 4: np.array([0, 1, 2]), 5: np.array([0, 1, 2, 3]),
         "kite": np.arange(32), "kinbike": np.arange(32), "dynbike": np.arange(32),
         "cstrrob": np.arange(16), "reactor": np.arange(16), "polynom": np.arange(8), "oscdae": np.arange(32)}


@pytest.mark.parametrize("cfg", [1, 3, 4, 5, "kite", "kinbike", "dynbike", "cstrrob", "reactor", "polynom", "oscdae"])
def test_config_matches_oracle_fixture(cfg):
    mpc = config_mpc(cfg, setup=False)
    fx = fixture(cfg)
    idx = CASES[cfg]
    sim = HostSim(mpc, f"cfg{cfg}")
    sim.o.max_iter = MAX_ITER
    P, x0s, u0s, z0s = mpc.pack_batch(fx["x0"][idx], fx["u_prev"][idx], fx["z0"][idx] if "z0" in fx else None, tvp=fx.get("tvp"))
    out = sim.solve(P, x0s=x0s, u0s=u0s, z0s=z0s)
    r = mpc.recipe
    res = compare_with_oracle(cfg, fx, idx, out["status"], out["iter_count"], out["u0_scaled"] * r.u_scaling, r.u_scaling)
    assert res["basin"] == 0, res
    assert res["iter_diff_max"] <= 1, res
