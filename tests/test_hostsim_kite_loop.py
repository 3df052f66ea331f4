"""Closed loop of a reference example that ships no golden (examples/kite: main.py's loop MPC -> simulator -> state feedback
under a varying wind speed): every controller step — the cold start and the steps warm-started from the previous solution
like the reference (optimizer.py:741-742) — solved by the kernel's algorithm source (CPU emulation) and by the oracle from
their own previous solutions; the plant is the integrator source of the batched simulator.  Required: equal flags, u0 within the
north-star tolerance and iteration counts within 1 at every step (the long first solve exercises the inertia correction on
variables and slacks, second-order corrections and the soft-constraint slack)."""
import numpy as np

from conftest import HostSim
from test_simulator_hostsim import build_simhost, sim_step

import kite
from oracle import models
from oracle.ipm import IPMSolver, IpmOptions
from oracle.nlp import FlatNLP

N_STEPS = 8


def test_kite_closed_loop_matches_oracle_step_by_step():
    nlp = FlatNLP(models.kite())
    orc = IPMSolver(nlp, IpmOptions(max_iter=300))
    mpc = kite.kite_mpc(setup=False); mpc.prepare_nlp()
    r = mpc.recipe
    sim = HostSim(mpc, "kite_loop"); sim.o.max_iter = 300
    model = kite.kite_model()
    plant = build_simhost(model, "kite")
    x, up = np.array([0.5, 0.3, 0.2]), np.zeros(1)
    rng = np.random.RandomState(0)
    w_h = w_o = None
    iters = []
    for t in range(N_STEPS):
        P, x0s, u0s, z0s = mpc.pack_batch(x[None], up[None], None)
        p_o = nlp.pack_p(x, up)
        assert np.allclose(P[0], p_o, rtol=0, atol=0)
        if t == 0:
            out = sim.solve(P, x0s=x0s, u0s=u0s, z0s=z0s); ro = orc.solve(nlp.initial_guess(x, up, None), p_o)
        else:
            out = sim.solve(P, x_init=w_h[None, :]); ro = orc.solve(w_o, p_o)
        w_h, w_o = out["x"][0], ro["x"]
        assert out["status"][0] == 0 and ro["success"], (t, out["status"], ro["return_status"])
        u_h, u_o = out["u0_scaled"][0] * r.u_scaling, ro["x"][nlp.ix_u[0, 0]] * nlp.d.u_scaling
        assert np.abs((u_h - u_o) / r.u_scaling).max() < 1e-5, (t, u_h, u_o)
        assert abs(int(out["iter_count"][0]) - ro["iter_count"]) <= 1, (t, out["iter_count"], ro["iter_count"])
        iters.append(int(out["iter_count"][0]))
        p_plant = np.array([[6.0, 10.0 * (1 + 0.1 * rng.randn())]])          # glide ratio nominal, wind speed disturbed
        xn, _, st = sim_step(plant, model, x[None], u_h[None], None, None, p_plant, 0.15)
        assert st[0] == 0
        x, up = xn[0], u_h
    assert iters[0] > 3 * max(iters[1:])                                      # the warm-started steps are the short ones
