"""The optimum itself, checked against a solver that shares nothing with IPOPT's algorithm:
tests/golden/independent_optimum.npz holds scipy trust-constr optima of the restated NLP
(generator: tests/golden/make_independent_optimum.py).  The oracle IPM (CPU) and the CUDA kernels
(GPU) must land on the same points within the north-star tolerances."""
import os

import numpy as np
import pytest

from mpc_racing_b200.workload import load_fixture
from parity_util import OBJ_RTOL, U_TOL, oracle_solve

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "independent_optimum.npz"))


def _check(obj, U, j):
    assert abs(obj - GOLD['obj'][j]) <= OBJ_RTOL * abs(GOLD['obj'][j])
    dU = np.abs(U - GOLD['U'][j])
    dU[0, -1] = 0.0          # degenerate last throttle
    assert dU.max() <= U_TOL, dU.max()


def test_oracle_ipm_reaches_the_independent_optimum():
    d = load_fixture("cfg2_shanghai_4096.npz")
    for j, b in enumerate(GOLD['index']):
        r = oracle_solve(d['x0'][b], d['u_prev'][b], d['s0'][b], d['cx'][b], d['cy'][b])
        assert r['status'] == 0
        _check(r['obj'], r['U'], j)


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["warp"])
def test_cuda_kernels_reach_the_independent_optimum(kernel):
    from mpc_racing_b200.solver import BatchedMPC
    d = load_fixture("cfg2_shanghai_4096.npz")
    idx = [int(b) for b in GOLD['index']]
    r = BatchedMPC(N=30, Ts=0.05, solve_kernel=kernel).solve(d['x0'][idx], d['u_prev'][idx], d['s0'][idx],
                                                            d['cx'][idx], d['cy'][idx])
    assert (r.status == 0).all()
    for j in range(len(idx)):
        _check(r.obj[j], r.U[j], j)
