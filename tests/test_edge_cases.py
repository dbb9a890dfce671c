"""Edge cases of the solve path: horizon limits, iteration cap, empty / ragged batches, invalid
arguments (no exceptions cross the C ABI: return codes + mpcr_last_error)."""
import ctypes as C

import numpy as np
import pytest

from mpc_racing_b200 import _capi
from mpc_racing_b200.workload import load_fixture
from parity_util import assert_parity, oracle_solve


@pytest.fixture(scope="module")
def cfg2():
    return load_fixture("cfg2_shanghai_4096.npz")


# ---- CPU: the kernel sources compiled for the host (tests/emu) and the config checks ----------
@pytest.mark.parametrize("N", [2, 32])
def test_horizon_limits(cfg2, N):
    """N = 2 (smallest) and N = 32 (one warp lane per stage: the largest) against the oracle."""
    from emu_util import emu_solve
    i = [7]
    out = emu_solve(cfg2['x0'][i], cfg2['u_prev'][i], cfg2['s0'][i], cfg2['cx'][i], cfg2['cy'][i], N=N)
    ref = oracle_solve(cfg2['x0'][7], cfg2['u_prev'][7], cfg2['s0'][7], cfg2['cx'][7], cfg2['cy'][7], N=N)
    assert_parity(out, 0, ref)


def test_iteration_cap_reports_maxiter(cfg2):
    from emu_util import emu_solve
    i = [0, 1]
    out = emu_solve(cfg2['x0'][i], cfg2['u_prev'][i], cfg2['s0'][i], cfg2['cx'][i], cfg2['cy'][i], max_iter=5)
    assert (out['status'] == -1).all() and (out['iters'] == 5).all()      # Maximum_Iterations_Exceeded
    assert np.isfinite(out['U']).all() and np.isfinite(out['X']).all()
    ref = oracle_solve(cfg2['x0'][0], cfg2['u_prev'][0], cfg2['s0'][0], cfg2['cx'][0], cfg2['cy'][0],
                       options=dict(max_iter=5))
    assert ref['status'] == -1 and ref['iters'] == 5
    np.testing.assert_allclose(out['U'][0], ref['U'], rtol=0, atol=1e-6)   # same 5 iterations


@pytest.mark.parametrize("field,value", [("N", 33), ("N", 1), ("Ts", 0.0), ("model", 3), ("max_iter", -1),
                                         ("solve_kernel", 7), ("integrator", 2), ("polish_tol", -1.0), ("tol", 0.0)])
def test_invalid_config_is_rejected(field, value):
    lib = _capi.load_library()
    cfg = _capi.default_config(lib)
    setattr(cfg, field, value)
    h = C.c_void_p()
    assert lib.mpcr_create(C.byref(cfg), C.byref(h)) == -1        # MPCR_ERR_INVALID_ARG, before any CUDA call
    assert not h


@pytest.mark.parametrize("N", [30, 15])
def test_pathological_inputs_terminate_with_failure_codes(cfg2, N):
    """NaN / Inf / singular inputs never hang or poison neighbours: each problem ends with a failure
    status (control/MPC.py:172-181 convention: last iterate, sol = None)."""
    from emu_util import emu_solve
    idx = [0, 1, 2, 3, 4, 5]
    x0, s0, cx = cfg2['x0'][idx].copy(), cfg2['s0'][idx].copy(), cfg2['cx'][idx].copy()
    x0[0, 3] = np.nan                  # NaN speed
    x0[1, 0] = np.inf                  # infinite position
    cx[2, :] = np.nan                  # NaN centerline coefficients
    x0[4, 3] = 0.0
    x0[4, 4] = 0.0                     # |v| = 0: the steering-gain derivative divides by it
    s0[5] = 1e9                        # absurd progress
    r = emu_solve(x0, cfg2['u_prev'][idx], s0, cx, cfg2['cy'][idx], N=N, max_iter=60)
    assert list(r['status'][:3]) == [-4, -4, -4]            # Invalid_Number_Detected
    assert r['status'][3] == 0                               # the untouched neighbour converges
    assert r['status'][4] < 0 and r['status'][5] < 0
    assert (r['iters'] <= 60).all()


# ---- GPU: through the C ABI ------------------------------------------------------------------
@pytest.mark.gpu
def test_empty_batch_and_null_arguments(cfg2):
    from mpc_racing_b200.solver import BatchedMPC
    s = BatchedMPC(N=30, Ts=0.05)
    lib, h = s.lib, s._h
    z = np.zeros(8)
    p = z.ctypes.data
    st = np.zeros(2, np.int32)
    # B = 0: success, nothing touched
    assert lib.mpcr_solve_batch(h, 0, p, p, p, p, p, None, None, None, p, p, p, p, p, None, p, st.ctypes.data,
                                st.ctypes.data, None) == 0
    # missing required output
    assert lib.mpcr_solve_batch(h, 1, p, p, p, p, p, None, None, None, None, p, p, p, p, None, p, st.ctypes.data,
                                st.ctypes.data, None) == -1
    assert b"null" in lib.mpcr_last_error(h)
    # the lane row needs max_err
    sl = BatchedMPC(N=30, Ts=0.05, lane_constraint=True)
    with pytest.raises(_capi.MpcrError, match="max_err"):
        sl.solve(cfg2['x0'][:2], cfg2['u_prev'][:2], cfg2['s0'][:2], cfg2['cx'][:2], cfg2['cy'][:2])


@pytest.mark.gpu
def test_ragged_batches_and_trace(cfg2):
    """Batch sizes around the launch granularity give per-problem identical answers; the debug
    trace has exactly iters + 1 rows per problem."""
    from mpc_racing_b200.solver import BatchedMPC
    s = BatchedMPC(N=30, Ts=0.05)
    big = s.solve(cfg2['x0'][:1100], cfg2['u_prev'][:1100], cfg2['s0'][:1100], cfg2['cx'][:1100], cfg2['cy'][:1100])
    for B in (1, 31, 33, 1037):        # 1036 = resident warps of the grid on a 148-SM part
        r = s.solve(cfg2['x0'][:B], cfg2['u_prev'][:B], cfg2['s0'][:B], cfg2['cx'][:B], cfg2['cy'][:B])
        assert np.array_equal(r.U, big.U[:B]) and np.array_equal(r.iters, big.iters[:B])
    t = s.solve(cfg2['x0'][:4], cfg2['u_prev'][:4], cfg2['s0'][:4], cfg2['cx'][:4], cfg2['cy'][:4], trace=True)
    for j in range(4):
        n = int(t.iters[j])
        assert (t.trace[j, :n + 1, 5] > 0).all() and (t.trace[j, n + 1:] == 0).all()      # mu column
        assert abs(t.trace[j, n, 0] - t.obj[j]) <= 1e-9 * abs(t.obj[j])


@pytest.mark.gpu
def test_iteration_cap_and_horizons_on_gpu(cfg2):
    from mpc_racing_b200.solver import BatchedMPC
    r = BatchedMPC(N=30, Ts=0.05, max_iter=5).solve(cfg2['x0'][:64], cfg2['u_prev'][:64], cfg2['s0'][:64],
                                                   cfg2['cx'][:64], cfg2['cy'][:64])
    assert (r.status == -1).all() and (r.iters == 5).all()
    for N in (2, 32):
        s = BatchedMPC(N=N, Ts=0.05)
        res = s.solve(cfg2['x0'][7:8], cfg2['u_prev'][7:8], cfg2['s0'][7:8], cfg2['cx'][7:8], cfg2['cy'][7:8])
        ref = oracle_solve(cfg2['x0'][7], cfg2['u_prev'][7], cfg2['s0'][7], cfg2['cx'][7], cfg2['cy'][7], N=N)
        assert_parity(dict(U=res.U, X=res.X, S=res.S, lam_g=res.lam_g, obj=res.obj, status=res.status,
                           iters=res.iters), 0, ref)


def test_invalid_parameter_rows_poison_only_their_problem(cfg2):
    """A per-problem parameter row outside the accepted range (include/mpcr.h: not finite, n < 1,
    negative weights) ends THAT problem with Invalid_Number_Detected; its neighbours are untouched."""
    from emu_util import emu_solve
    idx = [0, 1, 2, 3]
    good = np.array([1000.0, 0.85, 50.0, 2.0, 5000.0])
    prm = np.tile(good, (4, 1))
    prm[1, 3] = 0.0            # exponent 0
    prm[2, 0] = np.nan         # NaN weight
    prm[3, 4] = -1.0           # negative steer-rate weight
    r = emu_solve(cfg2['x0'][idx], cfg2['u_prev'][idx], cfg2['s0'][idx], cfg2['cx'][idx], cfg2['cy'][idx], params=prm)
    ref = emu_solve(cfg2['x0'][:1], cfg2['u_prev'][:1], cfg2['s0'][:1], cfg2['cx'][:1], cfg2['cy'][:1])
    assert r['status'][0] == 0 and np.array_equal(r['U'][0], ref['U'][0])
    assert list(r['status'][1:]) == [-4, -4, -4]
