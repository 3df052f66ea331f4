"""World-size-2 test of the batch sharding + result gather (gloo, CPU).  The per-rank compute is the CPU emulation
of the kernel source (the product uses the CUDA solver; the host-side sharding / collective logic is identical)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "examples")); sys.path.insert(0, str(ROOT / "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cstr
    from conftest import HostSim, cstr_batch_inputs
    from do_mpc_b200.parallel import shard_bounds, gather_results
    d = np.load(ROOT / "tests" / "golden" / "cstr_dataset.npz")
    mpc = cstr.cstr_nominal_mpc(setup=False); mpc.prepare_nlp()
    sim = HostSim(mpc, "cstr_nominal")
    lo, hi = shard_bounds(n_total, world, rank)
    P, x0s, u0s = cstr_batch_inputs(mpc.recipe, d["x0"][lo:hi], d["u_prev"][lo:hi])
    out = sim.solve(P, x0s, u0s)
    local = {"u0_scaled": torch.tensor(out["u0_scaled"]), "status": torch.tensor(out["status"]),
             "iter_count": torch.tensor(out["iter_count"])}
    g = gather_results(local, n_total)
    if rank == 0:
        q.put({k: v.numpy() for k, v in g.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_solve_equals_single_process(dataset):
    from do_mpc_b200.parallel import shard_bounds
    n_total = 37                                  # ragged: 19 + 18
    assert shard_bounds(n_total, 2, 0) == (0, 19) and shard_bounds(n_total, 2, 1) == (19, 37)
    assert shard_bounds(7, 3, 2) == (5, 7) and sum(b - a for a, b in (shard_bounds(10, 4, r) for r in range(4))) == 10
    # build the CPU emulation once here so that the two ranks only load it
    sys.path.insert(0, str(ROOT / "examples"))
    import cstr
    from conftest import HostSim
    m = cstr.cstr_nominal_mpc(setup=False); m.prepare_nlp()
    HostSim(m, "cstr_nominal")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs: p.start()
    got = q.get(timeout=240)
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    ok = dataset["success"][:n_total].astype(bool)
    assert got["u0_scaled"].shape == (n_total, 2) and got["status"].shape == (n_total,)
    assert np.array_equal((got["status"] == 0) | (got["status"] == 1), ok)
    assert np.abs(got["u0_scaled"][ok] - dataset["u0"][:n_total][ok] / np.array([100.0, 2000.0])).max() < 1e-5
