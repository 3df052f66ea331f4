"""The latency build (do_mpc_b200/solver/build.py::latency_threads, BatchedNlpSolver._for_batch): a larger CTA for small
batches — the reference's real-time `make_step(x0)` is a batch of one."""
import numpy as np
import pytest


def _recipe(make):
    from do_mpc_b200.solver.codegen import generate
    mpc = make(); mpc.prepare_nlp()
    return mpc.recipe, generate(mpc.recipe)


def test_latency_rule_on_the_example_problems():
    import batch_reactor, cstr, oscillating_masses
    from do_mpc_b200.solver.build import MAX_LATENCY_THREADS, latency_threads, launch_config
    for make, expect in ((lambda: cstr.cstr_robust_mpc(setup=False), 256),            # 180 edges: 23 rounds of 8 groups -> 12 of 16
                         (lambda: batch_reactor.reactor_mpc(setup=False), 224),          # 20 edges, 32-lane groups: 4 rounds -> 3
                         (lambda: cstr.cstr_nominal_mpc(setup=False), None),             # 20 edges: 2 rounds either way
                         (lambda: oscillating_masses.osc_mpc(setup=False), None)):       # 7 edges: one round already
        r, src = _recipe(make)
        lat = latency_threads(r, src)
        assert lat == expect
        if lat is not None:
            base = launch_config(r, src)[0]
            gsz = 16 if r.n_y <= 16 else 32
            rounds = lambda t: -(-len(r.edge_k) // (t // gsz))
            assert base < lat <= MAX_LATENCY_THREADS and lat % 32 == 0 and 4 * rounds(lat) <= 3 * rounds(base)
            assert launch_config(r, src, threads=lat)[0] == lat


@pytest.mark.gpu
def test_small_batches_use_the_latency_build_and_agree_with_the_throughput_build():
    """Robust CSTR (scenario tree, 180 edges): the golden instance alone (latency build, 256 threads) and inside a batch of 64
    (throughput build, 128 threads) — same status, iteration count and solution; the reference's IPOPT result for u0."""
    import cstr
    from examples_common import golden
    mpc = cstr.cstr_robust_mpc()
    g = golden("CSTR")
    x0 = g["x"][0]; up = np.zeros(2)
    assert mpc.S._lat_threads == 256 and mpc.S.latency_max_batch >= 1
    one = mpc.make_step_batch(x0[None, :], up[None, :])
    info1 = mpc.S.launch_info()
    X0 = np.repeat(x0[None, :], 64, axis=0) * (1 + 1e-3 * np.arange(64)[:, None]); Up = np.zeros((64, 2))
    many = mpc.make_step_batch(X0, Up)
    info64 = mpc.S.launch_info()
    assert info1["block"] == 256 and info1["grid"] == 1 and info64["block"] == 128
    assert one["status"][0] == many["status"][0] == 0
    assert abs(int(one["iter_count"][0]) - int(many["iter_count"][0])) <= 1
    assert np.abs((one["u0"][0] - many["u0"][0]) / mpc.recipe.u_scaling).max() < 1e-7
    assert np.abs((one["u0"][0] - g["u"][0]) / mpc.recipe.u_scaling).max() < 1e-5
    # the switch is the user's: latency_max_batch = 0 keeps every batch on the throughput build
    mpc.S.latency_max_batch = 0
    again = mpc.make_step_batch(x0[None, :], up[None, :])
    assert mpc.S.launch_info()["block"] == 128 and int(again["iter_count"][0]) == int(many["iter_count"][0])
    print("kernel ms, B = 1:", info1["last_kernel_ms"], "(latency build)", mpc.S.launch_info()["last_kernel_ms"], "(throughput build)")
