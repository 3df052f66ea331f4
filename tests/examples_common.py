"""Shared helpers for the example-parity tests: golden fixtures of the reference's regression pickles
(tests/golden/kkt_<name>.npz) against this repo's example controllers."""
import numpy as np

from conftest import GOLDEN

import batch_reactor
import cstr
import dip
import industrial_poly
import oscillating_masses

EXAMPLES = {
    # name: (factory(setup), golden sizes (n_w, n_g, n_p))
    "oscillatingMasses": (lambda setup=False: oscillating_masses.osc_mpc(setup=setup), (39, 32, 5)),
    "oscillatingMasses_dae": (lambda setup=False: oscillating_masses.osc_dae_mpc(setup=setup), (67, 60, 5)),
    "batch_reactor": (lambda setup=False: batch_reactor.reactor_mpc(setup=setup), (608, 564, 23)),
    "CSTR": (lambda setup=False: cstr.cstr_robust_mpc(setup=setup), (3564, 3064, 24)),
    "dip": (lambda setup=False: dip.dip_mpc(setup=setup), (4330, 4506, 126)),
    "industrial_poly": (lambda setup=False: industrial_poly.poly_mpc(setup=setup), (8100, 7210, 31)),
}


def golden(name):
    d = np.load(GOLDEN / f"kkt_{name}.npz")
    return {k: d[k] for k in d.files}


def cold_start_inputs(r, g, t=0):
    """Inputs of the reference regression test at step t=0: x0 from the golden closed loop, u_prev = 0,
    z0 = 0, parameters exactly as stored (opt_p_num row)."""
    P = g["opt_p_num"][t][None, :]
    x0s = (g["x"][t] / r.x_scaling)[None, :]
    u0s = np.zeros((1, r.n_u))
    z0s = np.zeros((1, r.n_z)) if r.n_z else None
    return P, x0s, u0s, z0s
