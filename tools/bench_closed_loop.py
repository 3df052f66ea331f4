"""Closed-loop throughput on the bench workload: 4096 CSTR controllers + plants, T steps, device-resident loop
(do_mpc_b200.rollout.closed_loop_device) against the same loop through the host APIs (make_step_batch + simulator.make_step_batch).
Not part of bench.py's contract line; numbers go to profiles/."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "examples"))


def main():
    import torch
    import cstr
    import do_mpc_b200 as do_mpc
    from do_mpc_b200.rollout import closed_loop_device
    B, T = 4096, 10
    d = np.load(ROOT / "tests" / "golden" / "cstr_dataset.npz")
    ok = np.where(d["success"])[0]
    idx = ok[np.arange(B) % ok.size]
    rng = np.random.RandomState(1)
    X0 = d["x0"][idx] * rng.uniform(0.999, 1.001, size=(B, 4)); UP = d["u_prev"][idx]
    mpc = cstr.cstr_nominal_mpc()
    sim = do_mpc.simulator.Simulator(mpc.model); sim.set_param(t_step=mpc.settings.t_step); sim.setup()
    closed_loop_device(mpc, sim, X0[:64], 2, U_prev=UP[:64])          # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter(); res = closed_loop_device(mpc, sim, X0, T, U_prev=UP); t_dev = time.perf_counter() - t0
    # host-API loop: cold start every step (make_step_batch has no warm start), so compare with a cold device loop as well
    t0 = time.perf_counter(); res_c = closed_loop_device(mpc, sim, X0, T, U_prev=UP, warm_start=False); t_dev_cold = time.perf_counter() - t0
    # IPOPT-style warm start of the interior-point method itself (user option of the reference: nlpsol_opts
    # {'ipopt.warm_start_init_point': 'yes', 'ipopt.mu_init': 1e-4}); first step cold, the rest warm
    mpc.S.set_options(warm_start=1, mu_init=1e-4)
    mpc.S._set_bounds(mpc._lb_opt_x.full().ravel(), mpc._ub_opt_x.full().ravel())
    closed_loop_device(mpc, sim, X0[:64], 2, U_prev=UP[:64])
    t0 = time.perf_counter(); res_w = closed_loop_device(mpc, sim, X0, T, U_prev=UP); t_ws = time.perf_counter() - t0
    mpc.S.set_options(warm_start=0, mu_init=0.1)
    x, up = X0.copy(), UP.copy()
    t0 = time.perf_counter()
    for k in range(T):
        out = mpc.make_step_batch(x, up)
        x = sim.make_step_batch(x, out["u0"])[0]; up = out["u0"]
    t_host = time.perf_counter() - t0
    print(json.dumps({"instances": B, "steps": T,
                      "device_loop_warm": {"seconds": t_dev, "controller_steps_per_s": B * T / t_dev, "converged": float((res["status"] == 0).mean()),
                                           "mean_iters": float(res["iter_count"].mean())},
                      "device_loop_ipm_warm_start_mu1e-4": {"seconds": t_ws, "controller_steps_per_s": B * T / t_ws, "converged": float((res_w["status"] == 0).mean()),
                                                            "mean_iters": float(res_w["iter_count"].mean()), "mean_iters_after_first_step": float(res_w["iter_count"][1:].mean()),
                                                            "max_state_diff_vs_cold_start_loop": float(np.abs(res_w["x"][-1] - res["x"][-1]).max())},
                      "device_loop_cold": {"seconds": t_dev_cold, "controller_steps_per_s": B * T / t_dev_cold, "mean_iters": float(res_c["iter_count"].mean())},
                      "host_api_loop_cold": {"seconds": t_host, "controller_steps_per_s": B * T / t_host},
                      "max_state_diff_cold_device_vs_host": float(np.abs(res_c["x"][-1] - x).max())}))


if __name__ == "__main__":
    main()
