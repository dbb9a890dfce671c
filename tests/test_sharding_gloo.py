"""N > 1 path on the CPU: two gloo ranks each solve their contiguous shard of one batch (with
the SIMT-emulated solver standing in for the GPU) and gather; the result equals the
single-process solve of the whole batch, in global problem order."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from mpc_racing_b200.sharding import shard_range

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_ranges_cover_the_batch():
    for B in (1, 7, 16, 16384, 16385):
        for world in (1, 2, 3, 8):
            edges = [shard_range(B, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == B
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, B, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    from emu_util import emu_solve
    from mpc_racing_b200.sharding import gather_results, shard_range
    from mpc_racing_b200.workload import load_fixture
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    d = load_fixture("prod_t4_256.npz")
    lo, hi = shard_range(B, rank, world)
    out = emu_solve(d["x0"][lo:hi], d["u_prev"][lo:hi], d["s0"][lo:hi], d["cx"][lo:hi], d["cy"][lo:hi], N=15)
    local = dict(status=out["status"], iters=out["iters"], obj=out["obj"], u0=out["U"][:, :, 0])
    full = gather_results(local, B, rank, world, dist=dist, device="cpu")
    dist.barrier()
    if rank == 0:
        q.put({k: v for k, v in full.items()})
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    from emu_util import emu_solve, build_emu
    from mpc_racing_b200.workload import load_fixture
    build_emu()
    B, world = 7, 2
    d = load_fixture("prod_t4_256.npz")
    ref = emu_solve(d["x0"][:B], d["u_prev"][:B], d["s0"][:B], d["cx"][:B], d["cy"][:B], N=15)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    full = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_array_equal(full["status"], ref["status"])
    np.testing.assert_array_equal(full["iters"], ref["iters"])
    np.testing.assert_array_equal(full["obj"], ref["obj"])
    np.testing.assert_array_equal(full["u0"], ref["U"][:, :, 0])
