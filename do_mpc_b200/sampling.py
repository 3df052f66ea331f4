"""Batched data generation for approximate MPC: what the reference's AMPCSampler does with a serial Python loop
(do_mpc/approximateMPC/_ampc_sampler.py:216-330, 387-526: random (x0, u_prev) in the bound box -> reset_history,
set_initial_guess, make_step -> one pickle per sample -> pandas table) as ONE batched solve.

Output table has the reference dataset's columns (`x0, u_prev, id, u0, status, t_make_step, t_wall, iter_count`,
see documentation/.../sampling_notebook/my_dataset/data_my_dataset_all.pkl)."""
from __future__ import annotations

import time

import numpy as np


def sampling_plan_box(mpc, n_samples: int, seed=None, lbx=None, ubx=None, lbu=None, ubu=None):
    """x0 ~ U[lbx, ubx], u_prev ~ U[lbu, ubu] (reference approx_mpc_sampling_plan_box, _ampc_sampler.py:236-283)."""
    rng = np.random.RandomState(seed) if seed is not None else np.random
    lbx = mpc._x_lb.full().ravel() if lbx is None else np.asarray(lbx, dtype=float)
    ubx = mpc._x_ub.full().ravel() if ubx is None else np.asarray(ubx, dtype=float)
    lbu = mpc._u_lb.full().ravel() if lbu is None else np.asarray(lbu, dtype=float)
    ubu = mpc._u_ub.full().ravel() if ubu is None else np.asarray(ubu, dtype=float)
    for name, a in (("state", lbx), ("state", ubx), ("input", lbu), ("input", ubu)):
        if not np.all(np.isfinite(a)):
            raise ValueError(f"box sampling needs finite {name} bounds (pass lbx/ubx/lbu/ubu explicitly)")
    x0 = rng.uniform(lbx, ubx, size=(n_samples, lbx.size))
    up = rng.uniform(lbu, ubu, size=(n_samples, lbu.size))
    return [{"x0": x0[i].reshape(-1, 1), "u_prev": up[i].reshape(-1, 1), "id": f"{i:0{len(str(max(n_samples - 1, 1)))}d}"}
            for i in range(n_samples)]


def sample_open_loop(mpc, plan, as_dataframe: bool = True):
    """Solve every entry of a sampling plan in one batch (cold start per sample, like the reference's sampler)."""
    X0 = np.array([np.asarray(p["x0"], dtype=float).ravel() for p in plan])
    UP = np.array([np.asarray(p["u_prev"], dtype=float).ravel() for p in plan])
    t0 = time.perf_counter()
    out = mpc.make_step_batch(X0, UP)
    dt = time.perf_counter() - t0
    rows = []
    for i, p in enumerate(plan):
        ok = bool(out["success"][i])
        rows.append({"x0": X0[i:i + 1], "u_prev": UP[i:i + 1], "id": p.get("id", str(i)),
                     "u0": out["u0"][i:i + 1] if ok else np.zeros((0,)), "status": ok,
                     "t_make_step": dt / len(plan), "t_wall": float("nan"), "iter_count": int(out["iter_count"][i])})
    if not as_dataframe:
        return rows
    import pandas as pd
    return pd.DataFrame(rows)
