#!/usr/bin/env python
"""Single-instance / small-batch latency of make_step (the reference's real-time use case): kernel time of a cold-start solve for
B in {1, 16} on chosen problems.  The CTA size comes from the library build (DMB_FORCE_THREADS overrides the throughput rule):
    python tools/latency_bench.py cstr_robust reactor_tree cstr_nominal
Prints one JSON line per (problem, B): kernel ms (best of 3), iterations, ms per iteration, launch configuration."""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "examples"))


def problems():
    import batch_reactor, cstr
    g = lambda n: np.load(os.path.join(ROOT, "tests", "golden", f"kkt_{n}.npz"))
    return {"cstr_robust": (lambda: cstr.cstr_robust_mpc(), g("CSTR")["x"][0], np.zeros(2)),
            "reactor_tree": (lambda: batch_reactor.reactor_mpc(n_horizon=30, n_robust=1), np.array([1.0, 0.5, 0.0, 120.0]), np.zeros(1)),
            "cstr_nominal": (lambda: cstr.cstr_nominal_mpc(), np.array([0.8, 0.5, 134.14, 130.0]), np.array([5.0, -1000.0]))}


def main():
    which = sys.argv[1:] or ["cstr_robust"]
    P = problems()
    for name in which:
        make, x0, up = P[name]
        mpc = make()
        mpc.S.set_options(max_iter=300)
        for B in (1, 16):
            X0 = np.repeat(x0[None, :], B, axis=0) * (1 + 0.002 * np.arange(B)[:, None]); Up = np.repeat(up[None, :], B, axis=0)
            ms = []
            for _ in range(4):
                out = mpc.make_step_batch(X0, Up)
                ms.append(mpc.S.launch_info()["last_kernel_ms"])
            it = out["iter_count"]
            print(json.dumps({"problem": name, "B": B, "kernel_ms": float(np.min(ms[1:])), "iters": it.tolist()[:4], "ms_per_iter": float(np.min(ms[1:]) / (it.max() + 1)),
                              "status": out["status"].tolist()[:4], "launch": mpc.S.launch_info(), "threads_env": os.environ.get("DMB_FORCE_THREADS", "rule")}), flush=True)


if __name__ == "__main__":
    main()
