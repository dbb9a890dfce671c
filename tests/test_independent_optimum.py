"""The optimum itself, checked against a solver that shares nothing with IPOPT's algorithm:
tests/golden/independent_optimum.npz holds scipy trust-constr results for 256 problems of config 2
(the first 192 + the 64 hardest for the interior-point path; generator
tests/golden/make_independent_optimum.py).  Of these

  * 221: trust-constr's objective equals the oracle IPM's to 1e-6 relative (``lower == 0``),
  * 28:  trust-constr ended infeasible (constraint violation > 1e-6): no optimum to compare,
  * 7:   trust-constr is feasible but HIGHER than the IPM (4 not converged, 3 converged to a worse
         local minimum of the non-convex NLP: problems 60, 1186, 1954),
  * 0:   trust-constr found a lower feasible point than the IPM.

The compiled oracle (CPU) and the CUDA kernel (GPU) must reproduce the objective on all 221 and
never exceed a feasible trust-constr objective.  Inputs: the objective is flat along some input
directions at the 1e-6 level, and trust-constr stops with a projected-gradient norm between 1e-7 and
1e-3; where it is below 1e-5 the (polished) IPM inputs agree to 1e-4, everywhere to 1e-2."""
import os

import numpy as np
import pytest

from mpc_racing_b200.workload import load_fixture

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "independent_optimum.npz"))
OBJ_RTOL = 1e-6


def _args():
    d = load_fixture("cfg2_shanghai_4096.npz")
    idx = GOLD["index"]
    return tuple(d[k][idx] for k in ("x0", "u_prev", "s0", "cx", "cy"))


def _check(obj, U, status):
    assert len(GOLD["index"]) == 256 and int((GOLD["lower"] == 0).sum()) >= 220
    assert (status >= 0).all()
    eq = GOLD["lower"] == 0
    rel = np.abs(obj - GOLD["obj"]) / np.abs(GOLD["obj"])
    assert (rel[eq] <= OBJ_RTOL).all(), (np.where(eq & (rel > OBJ_RTOL))[0], rel[eq].max())
    # never above a feasible point trust-constr found
    feas = GOLD["feasible"]
    assert ((obj - GOLD["obj"])[feas] <= OBJ_RTOL * np.abs(GOLD["obj"][feas])).all()
    dU = np.abs(U - GOLD["U"])
    dU[:, 0, -1] = 0.0          # degenerate last throttle
    du = dU.reshape(len(dU), -1).max(axis=1)
    sharp = eq & (GOLD["info"][:, 1] <= 1e-5)          # trust-constr itself converged tightly
    assert sharp.sum() >= 20 and (du[sharp] <= 1e-4).all(), du[sharp].max()
    assert (du[eq] <= 1e-2).all(), du[eq].max()


def test_compiled_oracle_reaches_the_independent_optima():
    from oracle import cxx
    r = cxx.solve_batch(*_args(), polish_tol=1e-9)
    _check(r["obj"], r["U"], r["status"])
    plain = cxx.solve_batch(*_args())
    eq = GOLD["lower"] == 0
    rel = np.abs(plain["obj"] - GOLD["obj"]) / np.abs(GOLD["obj"])
    assert (rel[eq] <= OBJ_RTOL).all()                 # already at tol = 1e-4 the objective is there


def test_numpy_oracle_reaches_the_independent_optimum():
    from parity_util import oracle_solve
    d = load_fixture("cfg2_shanghai_4096.npz")
    for j in (0, 7, 11):
        b = int(GOLD["index"][j])
        assert GOLD["lower"][j] == 0
        r = oracle_solve(d['x0'][b], d['u_prev'][b], d['s0'][b], d['cx'][b], d['cy'][b])
        assert r['status'] == 0 and abs(r['obj'] - GOLD['obj'][j]) <= OBJ_RTOL * abs(GOLD['obj'][j])


@pytest.mark.gpu
def test_cuda_kernel_reaches_the_independent_optima():
    from mpc_racing_b200.solver import BatchedMPC
    r = BatchedMPC(N=30, Ts=0.05, polish_tol=1e-9).solve(*_args())
    _check(r.obj, r.U, r.status)
