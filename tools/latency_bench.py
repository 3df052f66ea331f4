#!/usr/bin/env python
"""Single-instance / small-batch latency of make_step (the reference's real-time use case): kernel time of a cold-start solve for
B in {1, 16} on chosen problems, with the throughput build (latency_max_batch = 0) and with the latency build the solver picks
for small batches (do_mpc_b200/solver/build.py::latency_threads):
    python tools/latency_bench.py cstr_robust reactor_tree cstr_nominal dip kite
Prints one JSON line per (problem, build, B): kernel ms (best of 3), iterations, ms per iteration, launch configuration."""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "examples"))


def problems():
    import batch_reactor, cstr, dip, kite
    g = lambda n: np.load(os.path.join(ROOT, "tests", "golden", f"kkt_{n}.npz"))
    return {"cstr_robust": (lambda: cstr.cstr_robust_mpc(), g("CSTR")["x"][0], np.zeros(2), None),
            "reactor_tree": (lambda: batch_reactor.reactor_mpc(n_horizon=30, n_robust=1), np.array([1.0, 0.5, 0.0, 120.0]), np.zeros(1), None),
            "cstr_nominal": (lambda: cstr.cstr_nominal_mpc(), np.array([0.8, 0.5, 134.14, 130.0]), np.array([5.0, -1000.0]), None),
            "dip": (lambda: dip.dip_mpc(n_horizon=50), np.array([0.0, 0.9 * np.pi, 0.9 * np.pi, 0.0, 0.0, 0.0]), np.zeros(1), np.full((51, 1), -0.8)),
            "kite": (lambda: kite.kite_mpc(), np.array([0.5, 0.3, 0.2]), np.zeros(1), None)}


def main():
    which = sys.argv[1:] or ["cstr_robust"]
    P = problems()
    for name in which:
        make, x0, up, tvp = P[name]
        mpc = make()
        mpc.S.set_options(max_iter=300)
        lat_batch = mpc.S.latency_max_batch
        for build in ("throughput", "latency"):
            if build == "latency" and mpc.S._lat_threads is None:
                print(json.dumps({"problem": name, "build": build, "note": "no latency build: a larger CTA does not shorten the edge loops"}), flush=True)
                continue
            mpc.S.latency_max_batch = 0 if build == "throughput" else max(lat_batch, 16)
            for B in (1, 16):
                X0 = np.repeat(x0[None, :], B, axis=0) * (1 + 0.002 * np.arange(B)[:, None]); Up = np.repeat(up[None, :], B, axis=0)
                Z0 = np.zeros((B, mpc.recipe.n_z)) if mpc.recipe.n_z else None
                ms = []
                for _ in range(4):
                    out = mpc.make_step_batch(X0, Up, Z0, tvp=tvp) if tvp is not None else mpc.make_step_batch(X0, Up, Z0)
                    ms.append(mpc.S.launch_info()["last_kernel_ms"])
                it = out["iter_count"]
                print(json.dumps({"problem": name, "build": build, "B": B, "kernel_ms": float(np.min(ms[1:])), "iters": it.tolist()[:4],
                                  "ms_per_iter": float(np.min(ms[1:]) / (it.max() + 1)), "status": out["status"].tolist()[:4], "launch": mpc.S.launch_info()}), flush=True)


if __name__ == "__main__":
    main()
