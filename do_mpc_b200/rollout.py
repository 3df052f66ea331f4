"""Device-resident closed loops: B controllers and B plants advance together, nothing returns to the host between steps.

The reference's closed loop (`u0 = mpc.make_step(x0); y = simulator.make_step(u0); x0 = estimator.make_step(y)`,
e.g. examples/CSTR/main.py, and the closed-loop branch of AMPCSampler, _ampc_sampler.py:420-470) is a Python loop around one
NLP solve and one integrator call.  Here one step is: `dmb_solve_device` (persistent IPM kernel) -> `dms_step_device`
(integrator kernel) -> two small tensor ops that write the next parameter rows — all queued on one CUDA stream; the
trajectories are copied back once at the end.  State feedback only (the estimator returns the simulated state).  Models with
algebraic states, time-varying parameters and parameters are covered: their tvp / p trajectories are evaluated on the host for
the sampling instants of the run and uploaded once.
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
    dev = torch.device("cuda", S.device)
    f64 = torch.float64
    m = mpc.model
    dt = mpc.settings.t_step
    # time-varying parameters / parameters of the controller and of the plant: evaluated on the host for the n_steps sampling
    # instants (they are the same for all instances, reference _mpc.py:1004-1006 / simulator.py:775-777) and uploaded once
    t_start = float(np.asarray(mpc._t0).ravel()[0])
    tvp_ctl = p_ctl = tvp_sim = p_sim = None
    if r.n_tvp:
        tvp_ctl = torch.as_tensor(np.stack([np.concatenate([np.asarray(v.full()).ravel() for v in mpc.tvp_fun(np.array([t_start + k * dt]))["_tvp"]])
                                            for k in range(n_steps)]), dtype=f64, device=dev)
    if r.n_pm:
        p_ctl = torch.as_tensor(np.stack([np.concatenate([np.asarray(v.full()).ravel() for v in mpc.p_fun(np.array([t_start + k * dt]))["_p"]])
                                          for k in range(n_steps)]), dtype=f64, device=dev)
    ts0 = float(np.asarray(simulator._t0).ravel()[0])
    if m.n_tvp:
        tvp_sim = torch.as_tensor(np.stack([simulator.tvp_fun(np.array([ts0 + k * dt])).full().ravel() for k in range(n_steps)]), dtype=f64, device=dev)
    if m.n_p:
        p_sim = torch.as_tensor(np.stack([simulator.p_fun(np.array([ts0 + k * dt])).full().ravel() for k in range(n_steps)]), dtype=f64, device=dev)
    if r.n_lift:
        raise NotImplementedError("closed_loop_device with nl_cons_single_slack")
    X0 = np.ascontiguousarray(np.asarray(X0, dtype=float).reshape(-1, r.n_x)); B = X0.shape[0]
    U0 = np.tile(mpc._u0.full().ravel(), (B, 1)) if U_prev is None else np.asarray(U_prev, dtype=float).reshape(B, r.n_u)
    x = torch.as_tensor(X0, dtype=f64, device=dev); up = torch.as_tensor(U0, dtype=f64, device=dev)
    z = zs = None
    if r.n_z:                                           # algebraic states: the controller's initial guess and the plant's starting value
        z = torch.as_tensor(np.tile(simulator._z0.full().ravel(), (B, 1)), dtype=f64, device=dev)
        zs = torch.as_tensor(1.0 / r.z_scaling, dtype=f64, device=dev)
    xs = torch.as_tensor(1.0 / r.x_scaling, dtype=f64, device=dev); us = torch.as_tensor(r.u_scaling, dtype=f64, device=dev)
    P = torch.zeros(B, r.n_p, dtype=f64, device=dev)
    traj_x = torch.empty(n_steps + 1, B, r.n_x, dtype=f64, device=dev); traj_u = torch.empty(n_steps, B, r.n_u, dtype=f64, device=dev)
    stat = torch.empty(n_steps, B, dtype=torch.int32, device=dev); iters = torch.empty(n_steps, B, dtype=torch.int32, device=dev)
    traj_x[0] = x
    out = None
    for k in range(n_steps):
        P[:, r.p_x0:r.p_x0 + r.n_x] = x
        P[:, r.p_uprev:r.p_uprev + r.n_u] = up
        if tvp_ctl is not None:
            P[:, r.p_tvp:r.p_tvp + (r.N + 1) * r.n_tvp] = tvp_ctl[k]
        if p_ctl is not None:
            P[:, r.p_p:r.p_p + r.n_comb * r.n_pm] = p_ctl[k]
        if warm_start and out is not None:
            # with ipopt.warm_start_init_point=yes the previous multipliers are taken over as well (copies: no aliasing with the outputs)
            ws = bool(S.options.warm_start)
            out = S.solve_batch_device(P, x_init=out["x"].clone(), out=out, want_full=True,
                                       lam_g0=out["lam_g"].clone() if ws else None, lam_x0=out["lam_x"].clone() if ws else None)
        else:
            out = S.solve_batch_device(P, x0s=x * xs, u0s=up / us, z0s=None if z is None else z * zs, want_full=True)
        u = out["u0_scaled"] * us
        sim = simulator.make_step_batch_device(x, u, Z=z, tvp=None if tvp_sim is None else tvp_sim[k].expand(B, -1).contiguous(),
                                               p=None if p_sim is None else p_sim[k].expand(B, -1).contiguous())
        x, up = sim["x"], u
        if z is not None:
            z = sim["z"][:, :r.n_z]
        traj_x[k + 1] = x; traj_u[k] = u; stat[k] = out["status"]; iters[k] = out["iter_count"]
    res = {"x": traj_x, "u": traj_u, "status": stat, "iter_count": iters}
    if return_device:
        return res
    torch.cuda.synchronize(dev)
    return {k: v.cpu().numpy() for k, v in res.items()}
