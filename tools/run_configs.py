#!/usr/bin/env python
"""Run the BASELINE.json configurations at full size on one GPU and print one JSON line per config
(success rate, iterations, solves/s through the host API).  Inputs as SURVEY.md §8d."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(__file__), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "examples")); sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(name, mpc, X0, U_prev=None, Z0=None, reps=2, max_iter=500):
    mpc.S.set_options(max_iter=max_iter)
    out = None
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = mpc.make_step_batch(X0, U_prev, Z0)
        ts.append(time.perf_counter() - t0)
    ok = out["success"]
    st, cnt = np.unique(out["status"], return_counts=True)
    r = mpc.recipe
    line = {"config": name, "batch": int(X0.shape[0]), "n_w": r.n_w, "n_g": r.n_g, "n_y": r.n_y, "edges": len(r.edge_k),
            "success_rate": float(ok.mean()), "status_hist": {int(a): int(b) for a, b in zip(st, cnt)},
            "iters_mean": float(out["iter_count"].mean()), "iters_max": int(out["iter_count"].max()),
            "seconds": min(ts), "solves_per_s": X0.shape[0] / min(ts), "launch": mpc.S.launch_info()}
    print(json.dumps(line), flush=True)
    return out


def main():
    which = sys.argv[1:] or ["1", "2", "3", "4", "5"]
    import batch_reactor, cstr, dip, industrial_poly, oscillating_masses
    rng = np.random.RandomState(0)
    if "1" in which:
        np.random.seed(99)
        x0 = np.random.rand(4) - 0.5
        mpc = oscillating_masses.osc_mpc(n_horizon=10)
        out = run("1 oscillating_masses_discrete N=10 B=1", mpc, x0[None, :])
        from oracle import models
        from oracle.ipm import IPMSolver, IpmOptions, solve_mpc
        from oracle.nlp import FlatNLP
        nlp = FlatNLP(models.oscillating_masses(10))
        ro = solve_mpc(nlp, x0, np.zeros(1), solver=IPMSolver(nlp, IpmOptions()))
        print(json.dumps({"config": "1 oracle check", "u0_gpu": out["u0"][0].tolist(), "u0_oracle": ro["u0"].tolist(),
                          "iters_gpu": int(out["iter_count"][0]), "iters_oracle": ro["iter_count"]}))
    if "2" in which:
        d = np.load(os.path.join(ROOT, "tests", "golden", "cstr_dataset.npz"))
        r42 = np.random.RandomState(42)
        B = 4096
        x0 = r42.uniform([0.1, 0.1, 50, 50], [2, 2, 140, 140], size=(B, 4)); up = r42.uniform([5, -8500], [100, 0], size=(B, 2))
        x0[:1000], up[:1000] = d["x0"], d["u_prev"]
        run("2 CSTR nominal N=20 B=4096", cstr.cstr_nominal_mpc(), x0, up)
    if "3" in which:
        B = 1024
        x0 = np.array([1.0, 0.5, 0.0, 120.0]) * rng.uniform(0.9, 1.1, size=(B, 4))
        run("3 batch_reactor robust n_robust=1 N=30 B=1024", batch_reactor.reactor_mpc(n_horizon=30, n_robust=1), x0, np.zeros((B, 1)))
    if "4" in which:
        B = 8192
        x0 = np.zeros((B, 6)); x0[:, 1:3] = 0.9 * np.pi * rng.uniform(0.9, 1.1, size=(B, 2))
        run("4 DIP swing-up N=50 B=8192", dip.dip_mpc(n_horizon=50), x0, np.zeros((B, 1)), np.zeros((B, 3)), reps=1, max_iter=300)
    if "5" in which:
        B = 2048
        g = np.load(os.path.join(ROOT, "tests", "golden", "kkt_industrial_poly.npz"))
        x0 = g["x"][0] * rng.uniform(0.98, 1.02, size=(B, 10))
        run("5 industrial_poly robust N=20 B=2048", industrial_poly.poly_mpc(), x0, np.zeros((B, 3)), max_iter=300)


if __name__ == "__main__":
    main()
