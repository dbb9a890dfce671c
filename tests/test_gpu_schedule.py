"""mpcr_config.schedule: the dispatch order of a batch (expected-longest-first by reference-path
curvature) changes kernel time only - every output of every problem is bit-identical to the
index-order run, i.e. the order is a permutation and the solve does not depend on it."""
import numpy as np
import pytest

from mpc_racing_b200.workload import load_fixture, make_batch

pytestmark = pytest.mark.gpu


def test_schedule_changes_nothing_but_the_order():
    from mpc_racing_b200.solver import BatchedMPC
    b = make_batch(3000)                                     # > one wave of resident warps (1036 on B200)
    args = (b["x0"], b["u_prev"], b["s0"], b["cx"], b["cy"])
    ref = BatchedMPC(N=30, Ts=0.05, schedule="index").solve(*args)
    for sched in ("auto", "curvature"):
        s = BatchedMPC(N=30, Ts=0.05, schedule=sched)
        r = s.solve(*args)
        assert s.launch_info()["kernels_per_solve"] == 3
        for k in ("U", "X", "S", "eC", "eL", "lam_g", "obj", "status", "iters"):
            assert np.array_equal(getattr(r, k), getattr(ref, k), equal_nan=True), (sched, k)
    # a batch of one wave or less is not reordered under "auto"
    s = BatchedMPC(N=30, Ts=0.05)
    r = s.solve(*[a[:64] for a in args])
    assert s.launch_info()["kernels_per_solve"] == 1
    assert np.array_equal(r.U, ref.U[:64])


def test_schedule_with_degenerate_inputs():
    """NaN / zero polynomials (no curvature defined) must still give a permutation: every problem
    gets a status, the good ones their usual result."""
    from mpc_racing_b200.solver import BatchedMPC
    d = load_fixture("cfg2_shanghai_4096.npz")
    n = 1500
    x0, up, s0, cx, cy = (d[k][:n].copy() for k in ("x0", "u_prev", "s0", "cx", "cy"))
    cx[5] = np.nan
    cy[6] = 0.0
    cx[6] = 0.0
    x0[7, 3] = np.inf
    ref = BatchedMPC(N=30, Ts=0.05, schedule="index").solve(x0, up, s0, cx, cy)
    r = BatchedMPC(N=30, Ts=0.05, schedule="curvature").solve(x0, up, s0, cx, cy)
    assert np.array_equal(r.status, ref.status)
    ok = ref.status >= 0
    assert ok.sum() >= n - 10
    assert np.array_equal(r.U[ok], ref.U[ok])


def test_one_handle_from_two_streams_is_serialised_on_the_device():
    """A handle owns one dispatch counter / order buffer / event pair (include/mpcr.h): launches from
    different streams on the SAME handle are ordered on the device by the library (the second waits
    for the first), so interleaving them cannot skip or double-solve problems."""
    import torch
    from mpc_racing_b200.solver import BatchedMPC
    d = load_fixture("cfg2_shanghai_4096.npz")
    dev = torch.device("cuda", 0)
    s = BatchedMPC(N=30, Ts=0.05, on_device=True)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    A = [t(d[k][:1500]) for k in ("x0", "u_prev", "s0", "cx", "cy")]
    B = [t(d[k][1500:3000]) for k in ("x0", "u_prev", "s0", "cx", "cy")]
    ref_a, ref_b = s.solve(*A), s.solve(*B)
    torch.cuda.synchronize()
    ra, rb = ref_a.U.clone(), ref_b.U.clone()
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    outs = []
    for rep in range(3):
        oa = s.solve(*A, stream=s1.cuda_stream)
        ob = s.solve(*B, stream=s2.cuda_stream)
        outs.append((oa, ob))
    torch.cuda.synchronize()
    for oa, ob in outs:
        assert torch.equal(oa.U, ra) and torch.equal(ob.U, rb)
        assert torch.equal(oa.iters, ref_a.iters) and torch.equal(ob.status, ref_b.status)


def test_dispatch_keys_change_the_order_not_the_results():
    """mpcr_set_dispatch_keys (closed loop: the previous step's iteration counts): any key array gives
    the results of the index-order run, bit for bit."""
    import torch
    from mpc_racing_b200.solver import BatchedMPC
    d = make_batch(3000)
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    A = [t(d[k]) for k in ("x0", "u_prev", "s0", "cx", "cy")]
    ref = BatchedMPC(N=30, Ts=0.05, on_device=True, schedule="index").solve(*A)
    s = BatchedMPC(N=30, Ts=0.05, on_device=True, schedule="curvature")
    rng = np.random.default_rng(0)
    for keys in (ref.iters.clone(), torch.from_numpy(rng.integers(-5, 400, 3000).astype(np.int32)).to(dev), None):
        s.set_dispatch_keys(keys)
        r = s.solve(*A)
        torch.cuda.synchronize()
        for k in ("U", "X", "S", "lam_g", "obj", "status", "iters"):
            assert torch.equal(getattr(r, k), getattr(ref, k)), k


def test_pipelined_batches_are_bit_identical_to_serial_solves():
    """mpc_racing_b200/pipeline.py: two handles on two streams, batches enqueued back to back
    (mpcr_solve_batch_async with host arrays) - overlapping launches change nothing in the results."""
    import torch
    from mpc_racing_b200.pipeline import SolvePipeline
    from mpc_racing_b200.solver import BatchedMPC
    b = make_batch(4 * 1500)
    keys = ("x0", "u_prev", "s0", "cx", "cy")
    parts = [{k: np.ascontiguousarray(b[k][i * 1500:(i + 1) * 1500]) for k in keys} for i in range(4)]
    serial = BatchedMPC(N=30, Ts=0.05, restoration=True)
    ref = [serial.solve(*[p[k] for k in keys]) for p in parts]
    names = ("U", "X", "S", "eC", "eL", "lam_g", "obj", "status", "iters")
    # host arrays (page-locked), results copied out before the slot is reused
    pipe = SolvePipeline(depth=2, host=True, N=30, Ts=0.05, restoration=True)
    pinned = [{k: torch.from_numpy(p[k]).pin_memory().numpy() for k in keys} for p in parts]
    got = {}
    tickets = []
    for i, p in enumerate(pinned):
        tickets.append(pipe.submit(*[p[k] for k in keys]))
        if i >= 1:                                     # consume the previous batch while this one runs
            r = pipe.result(tickets[i - 1])
            got[i - 1] = {k: np.array(getattr(r, k)) for k in names}
    r = pipe.result(tickets[-1])
    got[3] = {k: np.array(getattr(r, k)) for k in names}
    with pytest.raises(ValueError):
        pipe.result(tickets[0])                        # its slot has been reused
    pipe.close()
    for i in range(4):
        for k in names:
            assert np.array_equal(got[i][k], getattr(ref[i], k), equal_nan=True), (i, k)
    # device arrays
    dpipe = SolvePipeline(depth=2, host=False, N=30, Ts=0.05, restoration=True)
    dev = [{k: torch.from_numpy(p[k]).cuda() for k in keys} for p in parts]
    for i in (0, 1):
        dpipe.submit(*[dev[i][k] for k in keys])
    for i in (0, 1):
        r = dpipe.result(i)
        for k in names:
            assert np.array_equal(getattr(r, k).cpu().numpy(), getattr(ref[i], k), equal_nan=True), (i, k)
    dpipe.close()
