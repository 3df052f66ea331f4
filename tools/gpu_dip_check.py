import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "examples")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
from examples_common import EXAMPLES, golden, cold_start_inputs
name = sys.argv[1] if len(sys.argv) > 1 else "dip"
max_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 500
make, _ = EXAMPLES[name]
mpc = make(setup=True); r = mpc.recipe; g = golden(name)
mpc.S.set_options(max_iter=max_iter, **({'max_refine': int(sys.argv[3])} if len(sys.argv) > 3 else {}))
P, x0s, u0s, z0s = cold_start_inputs(r, g)
out = mpc.S.solve_batch(P, x0s, u0s, z0s, lbx=mpc._lb_opt_x.full().ravel(), ubx=mpc._ub_opt_x.full().ravel(), return_full=True)
print(name, "status", out["status"], "iters", out["iter_count"], "u0", out["u0_scaled"][0] * r.u_scaling, "golden", g["u"][0], "stats", out["stats"][0][:4])
