#!/usr/bin/env python
"""Device-resident timing of chosen BASELINE configs (kernel development loop): python tools/quick_bench.py 2 3 [--steps 3]
Prints one JSON line per config: ms per launch (CUDA events around the kernel), solves/s, iterations, phase shares if built
with -DDMB_PROFILE (DMB_EXTRA_NVCC_FLAGS)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "examples"))


def cases():
    import batch_reactor, cstr, dip, industrial_poly
    sys.path.insert(0, ROOT)
    from bench import make_inputs, other_config_cases
    x0, up = make_inputs(4096, 0)
    out = {"2": ("cfg2_cstr_nominal_N20_B4096", lambda: cstr.cstr_nominal_mpc(), x0, up, None, None)}
    for k, c in zip(("3", "4", "5"), other_config_cases()):
        out[k] = c
    return out


def main():
    import torch
    argv = sys.argv[1:]
    steps = 3
    if "--steps" in argv:
        i = argv.index("--steps"); steps = int(argv[i + 1]); del argv[i:i + 2]
    which = [a for a in argv if not a.startswith("--")] or ["2"]
    dev = torch.device("cuda", 0)
    cs = cases()
    for k in which:
        name, make, X0, Up, Z0, tvp = cs[k]
        mpc = make(); S = mpc.S
        S.set_options(max_iter=300)
        P, x0s, u0s, z0s = mpc.pack_batch(X0, Up, Z0, tvp=tvp)
        dP, dx, du = (torch.tensor(a, device=dev) for a in (P, x0s, u0s))
        dz = None if z0s is None else torch.tensor(z0s, device=dev)
        res = S.solve_batch_device(dP, dx, du, dz); torch.cuda.synchronize()
        ms = []
        for _ in range(steps):
            res = S.solve_batch_device(dP, dx, du, dz, out=res); torch.cuda.synchronize()
            ms.append(S.launch_info()["last_kernel_ms"])
        it = res["iter_count"].cpu().numpy(); st = res["status"].cpu().numpy()
        stt = res["stats"].cpu().numpy()
        ph = stt[:, 4:13].sum(axis=0); sub = stt[:, 13:21].sum(axis=0)
        line = {"config": name, "ms": float(np.min(ms)), "solves_per_s": X0.shape[0] / (np.min(ms) * 1e-3), "success": float(((st == 0) | (st == 1)).mean()),
                "iters_mean": float(it.mean()), "iters_max": int(it.max()), "launch": S.launch_info(),
                "per_solve": {n: round(float(stt[:, 21 + j].mean()), 3) for j, n in enumerate(["restos", "resto_iters", "soc_solves", "refine_solves", "ls_trials"])} if stt.shape[1] > 21 else {}}
        if ph.sum() > 0:
            line["cycles_per_instance"] = float(ph.sum() / X0.shape[0])
            line["phases"] = {n: round(float(v / ph.sum()), 4) for n, v in zip(
                ["derivs", "kkt_mu", "factor", "riccati", "forward_recover", "stepsize", "linesearch", "accept", "init"], ph)}
            line["sub"] = {n: round(float(v / ph.sum()), 4) for n, v in zip(
                ["f_assemble", "f_lu", "f_backsub", "f_mn", "f_reduce", "ric_sweep", "fwd_sweep", "-"], sub)}
        print(json.dumps(line), flush=True)
        del mpc, S


if __name__ == "__main__":
    main()
