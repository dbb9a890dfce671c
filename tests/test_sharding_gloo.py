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


# ---- config 5: the GA's (track, individual) rollouts sharded over ranks, fitness table gathered -----
class _FakeRollouts:
    """Stands in for ClosedLoopRollouts on the CPU: deterministic segment times from the parameters."""

    def __init__(self, track_id):
        self.track_id = track_id

    def run(self, x0, s0, params=None, max_steps=0, seg_bounds=None):
        n_cp = len(seg_bounds) - 1
        base = 5.0 + 0.1 * self.track_id + 1e-4 * params[:, 0] + params[:, 1]
        seg = base[:, None] + 0.05 * np.arange(n_cp)[None, :]
        seg[params[:, 3] == 4.0, -1] = np.nan          # some individuals do not finish the last segment
        return dict(seg_times=seg, fail_steps=np.zeros(len(params), dtype=np.int32))


def _ga(rank, world, dist):
    from mpc_racing_b200.GA import mpcGA
    courses = [dict(ro=_FakeRollouts(t), seg_bounds=[0, 10, 20, 30], lap_time_guess=18.0, x0=np.zeros(6), s0=0.0)
               for t in range(5)]
    ga = mpcGA.MultiTrackGA(courses, pop=13, seed=5, rank=rank, world=world, dist=dist)
    out = [ga.run_step(max_steps=10) for _ in range(3)]
    return ga.population.copy(), np.stack([o["rewards"] for o in out]), ga.avg_times.copy()


def _ga_worker(rank, world, port, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    pop, rewards, avg = _ga(rank, world, dist)
    q.put((rank, pop, rewards, avg))
    dist.destroy_process_group()


def test_multitrack_ga_sharded_over_two_ranks_equals_one_rank():
    pop1, rew1, avg1 = _ga(0, 1, None)
    assert rew1.shape == (3, 13) and (rew1 > 0).all()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_ga_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, pop, rew, avg in got:          # every rank ends with the single-process population
        np.testing.assert_array_equal(pop, pop1)
        np.testing.assert_array_equal(rew, rew1)
        np.testing.assert_array_equal(avg, avg1)
