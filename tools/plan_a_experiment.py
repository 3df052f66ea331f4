#!/usr/bin/env python
"""north_star plan A, accuracy half of the keep/reject decision (no GPU needed).

Plan A = per-stage projection products  Y'(H_yy + Sigma)Y  on tcgen05 (TF32 operands split three ways, FP32 accumulation in
TMEM) with the FP64 iterative refinement loop of the solver absorbing the error.  A TF32x3 product is at best as accurate as an
FP32 product with FP32 accumulation, so this script runs the kernel's algorithm source on the CPU (tests/hostsim) with exactly
those products in FP32 (-DDMB_EMU_FP32_PROJ) — everything else, including the residual evaluation and the refinement, in FP64 —
on the reference's 1000-sample CSTR dataset, and compares success flags, iteration counts, u0 and refinement solves with the
all-FP64 run.  Output: one JSON line (committed as profiles/r2_plan_a_accuracy.json)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "examples"))


def run(flags, n):
    os.environ["DMB_HOSTSIM_FLAGS"] = flags
    import cstr
    from conftest import HostSim, cstr_batch_inputs
    d = np.load(os.path.join(ROOT, "tests", "golden", "cstr_dataset.npz"))
    mpc = cstr.cstr_nominal_mpc(setup=False); mpc.prepare_nlp()
    sim = HostSim(mpc, "plan_a" + ("_fp32" if flags else "_fp64"))
    sim.o.max_iter = 300
    P, x0s, u0s = cstr_batch_inputs(sim.r, d["x0"][:n], d["u_prev"][:n])
    out = sim.solve(P, x0s, u0s)
    return d, out, sim.r


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    d, ref, r = run("", n)
    _, emu, _ = run("-DDMB_EMU_FP32_PROJ", n)
    ok_ref = d["success"][:n].astype(bool)
    both = ref["success"] & emu["success"]
    line = {
        "instances": n, "reference_success": int(ok_ref.sum()),
        "fp64": {"success": int(ref["success"].sum()), "iters_mean_on_success": float(ref["iter_count"][ref["success"]].mean())},
        "fp32_projection": {"success": int(emu["success"].sum()), "flags_equal_to_fp64": bool(np.array_equal(emu["success"], ref["success"])),
                            "lost_successes": int((ref["success"] & ~emu["success"]).sum()),
                            "iters_mean_on_success": float(emu["iter_count"][emu["success"]].mean()) if emu["success"].any() else None,
                            "iters_mean_where_both_succeed": [float(ref["iter_count"][both].mean()), float(emu["iter_count"][both].mean())] if both.any() else None,
                            "max_u0_deviation_scaled_where_both_succeed": float(np.abs(ref["u0_scaled"][both] - emu["u0_scaled"][both]).max()) if both.any() else None,
                            "status_hist": {int(a): int(b) for a, b in zip(*np.unique(emu["status"], return_counts=True))}},
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
