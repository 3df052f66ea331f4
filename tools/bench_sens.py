"""Timing of the sensitivity kernel on the bench workload (4096 CSTR instances): dmb_sens_device with device-resident
solutions, CUDA events around the launches.  Not part of bench.py's contract line; numbers go to profiles/."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "examples"))


def main():
    import torch
    import cstr
    B = 4096
    rng = np.random.RandomState(0)
    mpc = cstr.cstr_nominal_mpc()
    r = mpc.recipe
    lb = np.array([0.1, 0.1, 50.0, 50.0]); ub = np.array([2.0, 2.0, 140.0, 140.0])
    X0 = rng.uniform(lb, ub, size=(B, 4)); UP = rng.uniform([5.0, -8500.0], [100.0, 0.0], size=(B, 2))
    out = mpc.make_step_batch(X0, UP, return_full=True)
    dev = torch.device("cuda", 0)
    t = lambda a, dt=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=dev)
    x, lam, P, st = t(out["x"]), t(out["lam_g"]), t(out["p"]), t(out["status"], torch.int32)
    res = {}
    for label, cols, u0_only in (("du0_dx0_uprev", list(range(4)) + [r.p_uprev, r.p_uprev + 1], True), ("full_dxdp_all", None, False)):
        for _ in range(2):
            mpc.S.sensitivities_device(x, lam, P, status=st, cols=cols, u0_only=u0_only)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            o = mpc.S.sensitivities_device(x, lam, P, status=st, cols=cols, u0_only=u0_only)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        ok = int((o["sens_status"] <= 1).sum())
        res[label] = {"ms_per_batch": ms, "instances": B, "differentiated": ok, "columns": 6 if cols is None else len(cols),
                      "instances_per_s": ok / ms * 1e3}
    res["solve_success"] = int(out["success"].sum())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
