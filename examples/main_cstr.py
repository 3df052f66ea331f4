"""The reference's examples/CSTR/main.py on this package: robust multi-stage NMPC of the CSTR in closed loop with the plant
simulator and state feedback (graphics omitted).  Run on a machine with a B200:  python examples/main_cstr.py [n_steps]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))

import do_mpc_b200 as do_mpc
from do_mpc_b200.tools import Timer
import cstr


def main(n_steps=50):
    mpc = cstr.cstr_robust_mpc()
    model = mpc.model
    simulator = do_mpc.simulator.Simulator(model)
    simulator.set_param(t_step=mpc.settings.t_step)
    p_num = simulator.get_p_template()
    p_num["alpha"], p_num["beta"] = 1.0, 1.0
    simulator.set_p_fun(lambda t_now: p_num)
    simulator.setup()
    estimator = do_mpc.estimator.StateFeedback(model)
    x0 = np.array([0.8, 0.5, 134.14, 130.0]).reshape(-1, 1)
    mpc.x0 = x0; simulator.x0 = x0; estimator.x0 = x0
    mpc.set_initial_guess()
    timer = Timer()
    for k in range(n_steps):
        timer.tic()
        u0 = mpc.make_step(x0)
        timer.toc()
        y_next = simulator.make_step(u0)
        x0 = estimator.make_step(y_next)
    timer.info()
    print("final state", x0.ravel(), "converged steps", int(mpc.data["success"].sum()), "of", n_steps)
    print("C_b trajectory (every 10th step):", mpc.data["_x", "C_b"].ravel()[::10])


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 50)
