"""BASELINE.json configurations 1, 3, 4, 5 as controllers of this package + the oracle fixtures
tests/golden/cfg<k>_oracle.npz (tests/golden/make_config_fixtures.py: oracle/ipm.py with SOC and restoration ON)."""
import numpy as np

from conftest import GOLDEN

import batch_reactor
import bicycle
import cstr
import dip
import industrial_poly
import kite
import oscillating_masses

CONFIGS = {
    1: ("oscillating_masses_discrete N=10", lambda setup: oscillating_masses.osc_mpc(n_horizon=10, setup=setup), (54, 44, 5)),
    3: ("batch_reactor robust n_robust=1 N=30", lambda setup: batch_reactor.reactor_mpc(n_horizon=30, n_robust=1, setup=setup), (8082, 7564, 23)),
    4: ("DIP swing-up N=50", lambda setup: dip.dip_mpc(n_horizon=50, setup=setup), (2180, 2256, 76)),
    5: ("industrial_poly robust N=20", lambda setup: industrial_poly.poly_mpc(setup=setup), (8100, 7210, 31)),
    # example families of the reference that ship no golden (sizes from running the reference's templates through compat/,
    # tests/test_compat_templates.py): 32 initial states each around the example's main.py
    "kite": ("towing kite N=80 (aux objective reading parameters, soft aux constraint)", lambda setup: kite.kite_mpc(setup=setup), (1132, 1043, 10)),
    "kinbike": ("kinematic bicycle N=10", lambda setup: bicycle.kinematic_mpc(setup=setup), (196, 164, 6)),
    "dynbike": ("dynamic bicycle N=10 (Pacejka tyre forces, atan2)", lambda setup: bicycle.dynamic_mpc(setup=setup), (284, 246, 8)),
    # golden families away from their golden initial state: random starts (CPU-emulation tests only)
    "cstrrob": ("robust CSTR (tree + soft constraint), random starts", lambda setup: cstr.cstr_robust_mpc(setup=setup), (3564, 3064, 24)),
    "reactor": ("batch reactor N=20, random starts", lambda setup: batch_reactor.reactor_mpc(setup=setup), (608, 564, 23)),
    "polynom": ("industrial_poly without the tree (soft T_R constraint), random starts", lambda setup: industrial_poly.poly_mpc(n_robust=0, setup=setup), None),
    "oscdae": ("discrete DAE oscillating masses, random starts", lambda setup: oscillating_masses.osc_dae_mpc(setup=setup), (67, 60, 5)),
}
EXAMPLE_FAMILIES = ("kite", "kinbike", "dynbike")
RANDOM_START_FAMILIES = ("cstrrob", "reactor", "polynom", "oscdae")
MAX_ITER = 300


def fixture(cfg):
    d = np.load(GOLDEN / f"cfg{cfg}_oracle.npz")
    return {k: d[k] for k in d.files}


def config_mpc(cfg, setup):
    name, make, sizes = CONFIGS[cfg]
    mpc = make(setup)
    if not setup:
        mpc.prepare_nlp()
    r = mpc.recipe
    assert sizes is None or (r.n_w, r.n_g, r.n_p) == sizes, (name, (r.n_w, r.n_g, r.n_p))
    return mpc


def compare_with_oracle(cfg, fx, idx, status, iters, u0, u_scaling, tol_u=1e-5):
    """Success flags equal; on instances where both succeed: |u0 - u0_oracle| <= tol_u * u_scaling and iterations within 1.
    Basin mismatches (a different local solution of a non-convex NLP) are returned, not folded into
    the tolerance.  Returns dict(n, n_ok, basin, iter_diff_max)."""
    ok_o = (fx["status"][idx] == 0) | (fx["status"][idx] == 1)
    ok_d = (status == 0) | (status == 1)
    assert np.array_equal(ok_o, ok_d), (cfg, "success flags differ", np.flatnonzero(ok_o != ok_d), fx["status"][idx][ok_o != ok_d], status[ok_o != ok_d])
    both = ok_o & ok_d
    du = np.abs(u0[both] - fx["u0"][idx][both]) / u_scaling
    same_basin = du.max(axis=1) <= tol_u if both.any() else np.zeros(0, bool)
    di = np.abs(iters[both].astype(int) - fx["iter_count"][idx][both].astype(int))
    return {"n": len(idx), "n_ok": int(both.sum()), "basin": int((~same_basin).sum()),
            "iter_diff_max": int(di[same_basin].max()) if same_basin.any() else 0,
            "iter_diff": di, "same_basin": same_basin}
