"""Device-resident closed loops: B controllers and B plants advance together, nothing returns to the host between steps.

The reference's closed loop (`u0 = mpc.make_step(x0); y = simulator.make_step(u0); x0 = estimator.make_step(y)`,
e.g. examples/CSTR/main.py, and the closed-loop branch of AMPCSampler, _ampc_sampler.py:420-470) is a Python loop around one
NLP solve and one integrator call.  Here one step is: `dmb_solve_device` (persistent IPM kernel) -> `dms_step_device`
(integrator kernel) -> two small tensor ops that write the next parameter rows — all queued on one CUDA stream; the
trajectories are copied back once at the end.  State feedback only (the estimator returns the simulated state).
"""
from __future__ import annotations

import numpy as np


def closed_loop_device(mpc, simulator, X0, n_steps: int, U_prev=None, warm_start: bool = True, return_device: bool = False):
    """Run `n_steps` closed-loop steps for B initial states X0 [B, n_x] (physical units).

    Returns dict(x [n_steps+1, B, n_x], u [n_steps, B, n_u], status [n_steps, B], iter_count [n_steps, B]).
    Instances keep running after a failed solve (the reference applies the last iterate as well, optimizer.py:772-777);
    `status` tells which steps did not converge.  warm_start: the previous solution is the next initial guess, like
    make_step() does (optimizer.py:741-742); the first step starts from the broadcast of (x0, u_prev).
    """
    import torch
    r, S = mpc.recipe, mpc.S
    assert mpc.flags["setup"] and simulator.flags["setup"]
    if r.n_z or r.n_tvp or r.n_pm:
        raise NotImplementedError("closed_loop_device: models with algebraic states / tvp / parameters use the host loop for now")
    dev = torch.device("cuda", S.device)
    f64 = torch.float64
    X0 = np.ascontiguousarray(np.asarray(X0, dtype=float).reshape(-1, r.n_x)); B = X0.shape[0]
    U0 = np.tile(mpc._u0.full().ravel(), (B, 1)) if U_prev is None else np.asarray(U_prev, dtype=float).reshape(B, r.n_u)
    S._set_bounds(mpc._lb_opt_x.full().ravel(), mpc._ub_opt_x.full().ravel())
    x = torch.as_tensor(X0, dtype=f64, device=dev); up = torch.as_tensor(U0, dtype=f64, device=dev)
    xs = torch.as_tensor(1.0 / r.x_scaling, dtype=f64, device=dev); us = torch.as_tensor(r.u_scaling, dtype=f64, device=dev)
    P = torch.zeros(B, r.n_p, dtype=f64, device=dev)
    traj_x = torch.empty(n_steps + 1, B, r.n_x, dtype=f64, device=dev); traj_u = torch.empty(n_steps, B, r.n_u, dtype=f64, device=dev)
    stat = torch.empty(n_steps, B, dtype=torch.int32, device=dev); iters = torch.empty(n_steps, B, dtype=torch.int32, device=dev)
    traj_x[0] = x
    out = None
    for k in range(n_steps):
        P[:, r.p_x0:r.p_x0 + r.n_x] = x
        P[:, r.p_uprev:r.p_uprev + r.n_u] = up
        if warm_start and out is not None:
            # with ipopt.warm_start_init_point=yes the previous multipliers are taken over as well (copies: no aliasing with the outputs)
            ws = bool(S.options.warm_start)
            out = S.solve_batch_device(P, x_init=out["x"].clone(), out=out, want_full=True,
                                       lam_g0=out["lam_g"].clone() if ws else None, lam_x0=out["lam_x"].clone() if ws else None)
        else:
            out = S.solve_batch_device(P, x0s=x * xs, u0s=up / us, want_full=True)
        u = out["u0_scaled"] * us
        sim = simulator.make_step_batch_device(x, u)
        x, up = sim["x"], u
        traj_x[k + 1] = x; traj_u[k] = u; stat[k] = out["status"]; iters[k] = out["iter_count"]
    res = {"x": traj_x, "u": traj_u, "status": stat, "iter_count": iters}
    if return_device:
        return res
    torch.cuda.synchronize(dev)
    return {k: v.cpu().numpy() for k, v in res.items()}
