#!/usr/bin/env python
"""Oracle fixtures at the BASELINE.json configuration sizes ("no golden — parity unpinned by the reference
at those exact sizes; pin them by ... our CPU IPM").

For configs 1, 3, 4, 5 this draws the first 64 instances of the §8d input distribution (same seeds as tools/run_configs.py
and bench.py's other_configs leg), solves each with oracle/ipm.py — second-order correction and the restoration substitute
ON, block LDL^T with exact inertia — and stores inputs + results in tests/golden/cfg<k>_oracle.npz.  The GPU tests
(tests/test_gpu_configs.py) solve the same inputs through the C ABI and compare success flags, u0 and iteration counts.

    python tests/golden/make_config_fixtures.py [1 3 4 5] [--n 64] [--procs 8]
"""
import argparse
import os
import sys
import time

for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
    os.environ[_v] = "1"                      # one BLAS thread per worker process (set before numpy is imported)
import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
MAX_ITER = {1: 300, 3: 300, 4: 300, 5: 300, "kite": 300, "kinbike": 300, "dynbike": 300, "cstrrob": 300, "reactor": 300, "oscdae": 300,
            "polynom": 300}
# example families of the reference without a golden in testing/results (examples/kite, kinematic_bicycle_model,
# dynamic_bicycle_model): fixtures "cfgkite", "cfgkinbike", "cfgdynbike", initial states drawn around main.py's


def config_inputs(cfg, n):
    """(x0 [n,n_x], u_prev [n,n_u], z0 or None, tvp or None) — SURVEY.md §8d, identical to tools/run_configs.py."""
    rng = np.random.RandomState(0)
    if cfg == 1:
        np.random.seed(99)
        x0 = np.random.rand(n, 4) - 0.5            # row 0 = the reference test's x0 (test_oscillating_masses_discrete.py:82-83)
        return x0, np.zeros((n, 1)), None, None
    if cfg == 3:
        x0 = np.array([1.0, 0.5, 0.0, 120.0]) * rng.uniform(0.9, 1.1, size=(1024, 4))
        return x0[:n], np.zeros((n, 1)), None, None
    if cfg == 4:
        x0 = np.zeros((8192, 6)); x0[:, 1:3] = 0.9 * np.pi * rng.uniform(0.9, 1.1, size=(8192, 2))
        tvp = np.full((51, 1), -0.8)
        return x0[:n], np.zeros((n, 1)), np.zeros((n, 3)), tvp
    if cfg == 5:
        g = np.load(os.path.join(ROOT, "tests", "golden", "kkt_industrial_poly.npz"))
        x0 = g["x"][0] * rng.uniform(0.98, 1.02, size=(2048, 10))
        return x0[:n], np.zeros((n, 3)), None, None
    if cfg == "kite":
        # examples/kite/main.py:44-56: each state uniformly inside 60 % of its bounds; w_ref / E_0 / h_min fixed to 10 / 6 / 100
        lo, hi = np.array([0.0, -0.5 * np.pi, -np.pi]), np.array([0.5 * np.pi, 0.5 * np.pi, np.pi])
        m_, r_ = (hi + lo) / 2, (hi - lo) / 2
        x0 = m_ - 0.6 * r_ + 2 * 0.6 * r_ * rng.uniform(size=(64, 3))
        return x0[:n], np.zeros((n, 1)), None, None
    if cfg == "kinbike":
        # examples/kinematic_bicycle_model/main.py:45-50: x0 = (0, 0, 0, 0.1), perturbed
        x0 = np.array([0.0, 0.0, 0.0, 0.1]) + rng.uniform([-1, -1, -0.5, 0.0], [1, 1, 0.5, 2.0], size=(64, 4))
        return x0[:n], np.zeros((n, 2)), None, None
    if cfg == "dynbike":
        # examples/dynamic_bicycle_model/main.py:57-64: x0 = (0, 0, 0, 0.1, 0, 0), perturbed inside the state bounds
        x0 = np.array([0.0, 0.0, 0.0, 0.1, 0.0, 0.0]) + rng.uniform([-1, -0.5, -0.3, 0.0, -0.2, -0.1], [1, 0.5, 0.3, 2.0, 0.2, 0.1], size=(64, 6))
        return x0[:n], np.zeros((n, 2)), None, None
    # golden families away from their golden initial state (random starts; CPU-emulation tests only, tests/test_hostsim_configs.py)
    if cfg == "cstrrob":       # shipped robust CSTR (tree + soft constraint), x0 / u_prev from cfg #2's sampling box: 4 of 16 infeasible
        r5 = np.random.RandomState(5)
        return r5.uniform([0.1, 0.1, 50, 50], [2, 2, 140, 140], size=(16, 4))[:n], r5.uniform([5, -8500], [100, 0], size=(16, 2))[:n], None, None
    if cfg in ("reactor", "polynom", "oscdae"):
        r11 = np.random.RandomState(11)
        x_r = np.array([1.0, 0.5, 0.0, 120.0]) * r11.uniform(0.7, 1.3, size=(16, 4)); u_r = r11.uniform(0, 0.2, size=(16, 1))
        g = np.load(os.path.join(ROOT, "tests", "golden", "kkt_industrial_poly.npz"))
        x_p = g["x"][0] * r11.uniform(0.995, 1.005, size=(8, 10))
        x_o = r11.uniform(-1.5, 1.5, size=(32, 4)); u_o = r11.uniform(-1, 1, size=(32, 1))
        if cfg == "reactor":   # shipped batch reactor (N=20, two finite elements), +-30 % around the golden start
            return x_r[:n], u_r[:n], None, None
        if cfg == "polynom":   # industrial_poly without the tree (n_robust = 0: soft T_R constraint), +-0.5 % around the golden start
            return x_p[:n], np.zeros((min(n, 8), 3)), None, None
        return x_o[:n], u_o[:n], np.zeros((min(n, 32), 4)), None          # discrete DAE oscillating masses
    raise ValueError(cfg)


def config_nlp(cfg):
    from oracle import models
    from oracle.nlp import FlatNLP
    d = {1: lambda: models.oscillating_masses(10), 3: lambda: models.batch_reactor(30, 1), 4: lambda: models.dip(50),
         5: lambda: models.industrial_poly(20, 1), "kite": models.kite, "kinbike": models.kinematic_bicycle,
         "dynbike": models.dynamic_bicycle, "cstrrob": models.cstr_robust, "reactor": lambda: models.batch_reactor(20, 0),
         "polynom": lambda: models.industrial_poly(20, 0), "oscdae": lambda: models.oscillating_masses_dae(7)}[cfg]()
    return FlatNLP(d)


_W = {}
STATUS_CODE = {"Solve_Succeeded": 0, "Solved_To_Acceptable_Level": 1, "Infeasible_Problem_Detected": 2,
               "Maximum_Iterations_Exceeded": -1, "Restoration_Failed": -2, "Error_In_Step_Computation": -3,
               "Invalid_Number_Detected": -13}


def _work(args):
    cfg, i, x0, up, z0, tvp = args
    from oracle.ipm import IPMSolver, IpmOptions, solve_mpc
    if cfg not in _W:
        nlp = config_nlp(cfg)
        _W[cfg] = (nlp, IPMSolver(nlp, IpmOptions(max_iter=MAX_ITER[cfg])))
    nlp, sol = _W[cfg]
    t = time.perf_counter()
    r = solve_mpc(nlp, x0, up, z0=z0, tvp=tvp, solver=sol)
    return (i, STATUS_CODE[r["return_status"]], r["iter_count"], r["restorations"], r["u0"], r["f"], time.perf_counter() - t)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cfgs", nargs="*", type=lambda v: int(v) if v.isdigit() else v, default=[1, 3, 4, 5])
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    import multiprocessing as mp
    for cfg in a.cfgs:
        x0, up, z0, tvp = config_inputs(cfg, a.n)
        a_n = len(x0)                                         # some families define fewer instances than --n
        jobs = [(cfg, i, x0[i], up[i], None if z0 is None else z0[i], tvp) for i in range(a_n)]
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(a.procs) as pool:
            res = pool.map(_work, jobs, chunksize=1)
        res.sort()
        out = dict(x0=x0, u_prev=up, status=np.array([r[1] for r in res], dtype=np.int32),
                   iter_count=np.array([r[2] for r in res], dtype=np.int32), restorations=np.array([r[3] for r in res], dtype=np.int32),
                   u0=np.array([r[4] for r in res]), f=np.array([r[5] for r in res]), seconds=np.array([r[6] for r in res]))
        if z0 is not None: out["z0"] = z0
        if tvp is not None: out["tvp"] = tvp
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"cfg{cfg}_oracle.npz"), **out)
        ok = (out["status"] == 0) | (out["status"] == 1)
        print(f"cfg{cfg}: {a_n} instances in {time.perf_counter() - t0:.1f} s wall; success {ok.mean():.3f}; iterations mean "
              f"{out['iter_count'].mean():.1f} max {out['iter_count'].max()}; restorations {out['restorations'].sum()}; "
              f"status hist {dict(zip(*np.unique(out['status'], return_counts=True)))}", flush=True)


if __name__ == "__main__":
    main()
