"""Lane-per-problem solver (mpc_racing_b200/csrc/mpcr_ipm_lanes.cuh), compiled for the host by
tests/emu (test infrastructure), against the numpy oracle and against the emulated
warp-per-problem solver on seeded inputs.  The CUDA build of the same source is
checked on the GPU in test_gpu_parity.py."""
import numpy as np
import pytest

from emu_util import emu_solve
from mpc_racing_b200.workload import load_fixture
from oracle import vehicle as V
from parity_util import assert_parity, oracle_solve


@pytest.fixture(scope="module")
def cfg2():
    return load_fixture("cfg2_shanghai_4096.npz")


def test_lanes_vs_oracle_default_horizon(cfg2):
    idx = list(range(8, 14))
    out = emu_solve(cfg2['x0'][idx], cfg2['u_prev'][idx], cfg2['s0'][idx], cfg2['cx'][idx], cfg2['cy'][idx],
                    lanes=True)
    for j, b in enumerate(idx):
        ref = oracle_solve(cfg2['x0'][b], cfg2['u_prev'][b], cfg2['s0'][b], cfg2['cx'][b], cfg2['cy'][b])
        assert_parity(out, j, ref)


def test_lanes_follow_the_warp_solver_iteration_by_iteration(cfg2):
    """Same algorithm, different mapping: traces agree to round-off while the paths coincide."""
    idx = list(range(64, 96))
    args = (cfg2['x0'][idx], cfg2['u_prev'][idx], cfg2['s0'][idx], cfg2['cx'][idx], cfg2['cy'][idx])
    a = emu_solve(*args, trace=True)
    b = emu_solve(*args, trace=True, lanes=True)
    assert (a['status'] == b['status']).all()
    same = a['iters'] == b['iters']
    assert same.mean() >= 0.9
    assert (np.abs(a['obj'] - b['obj']) / np.abs(a['obj']))[same].max() <= 1e-8
    for j in np.where(same)[0][:8]:
        n = min(10, int(a['iters'][j]))
        ta, tb = a['trace'][j, :n], b['trace'][j, :n]
        assert np.allclose(ta, tb, rtol=1e-6, atol=1e-9), (j, np.abs(ta - tb).max())


def test_lanes_variants(cfg2):
    d3 = load_fixture("cfg3_t1_triple_512.npz")
    out = emu_solve(d3['x0'][[0]], d3['u_prev'][[0]], d3['s0'][[0]], d3['cx'][[0]], d3['cy'][[0]],
                    max_err=d3['max_err'][[0]], lane_constraint=1, lanes=True)
    ref = oracle_solve(d3['x0'][0], d3['u_prev'][0], d3['s0'][0], d3['cx'][0], d3['cy'][0], lane=True,
                       max_error=float(d3['max_err'][0]))
    assert_parity(out, 0, ref)
    # production horizon, no wrap rows, kinematic model, warm start without a previous input
    dp = load_fixture("prod_t4_256.npz")
    out = emu_solve(dp['x0'][[1]], dp['u_prev'][[1]], dp['s0'][[1]], dp['cx'][[1]], dp['cy'][[1]], N=15, lanes=True)
    assert_parity(out, 0, oracle_solve(dp['x0'][1], dp['u_prev'][1], dp['s0'][1], dp['cx'][1], dp['cy'][1], N=15))
    i = [3]
    out = emu_solve(cfg2['x0'][i], cfg2['u_prev'][i], cfg2['s0'][i], cfg2['cx'][i], cfg2['cy'][i],
                    wrap_constraint=0, lanes=True)
    assert_parity(out, 0, oracle_solve(cfg2['x0'][3], cfg2['u_prev'][3], cfg2['s0'][3], cfg2['cx'][3],
                                       cfg2['cy'][3], wrap=False))
    out = emu_solve(cfg2['x0'][i], cfg2['u_prev'][i], cfg2['s0'][i], cfg2['cx'][i], cfg2['cy'][i], model=1,
                    lanes=True)
    assert_parity(out, 0, oracle_solve(cfg2['x0'][3], cfg2['u_prev'][3], cfg2['s0'][3], cfg2['cx'][3],
                                       cfg2['cy'][3], model=V.MODEL_KINEMATIC))
    cold = emu_solve(cfg2['x0'][i], cfg2['u_prev'][i], cfg2['s0'][i], cfg2['cx'][i], cfg2['cy'][i], lanes=True)
    nan = np.full((1, 2), np.nan)
    out = emu_solve(cfg2['x0'][i], nan, cfg2['s0'][i], cfg2['cx'][i], cfg2['cy'][i], u_warm=cold['U'].copy(),
                    lanes=True)
    assert_parity(out, 0, oracle_solve(cfg2['x0'][3], None, cfg2['s0'][3], cfg2['cx'][3], cfg2['cy'][3],
                                       u_warm=cold['U'][0]))


def test_rk4_integrator_nlp(cfg2):
    """RK4 extension wired into the NLP (lane solver): same IPM, dynamics rows x+ = RK4(x, u);
    against the oracle with the same integrator (oracle/vehicle.py:rk4_derivs)."""
    idx = [0, 7]
    out = emu_solve(cfg2['x0'][idx], cfg2['u_prev'][idx], cfg2['s0'][idx], cfg2['cx'][idx], cfg2['cy'][idx],
                    lanes=True, integrator=1)
    for j, b in enumerate(idx):
        ref = oracle_solve(cfg2['x0'][b], cfg2['u_prev'][b], cfg2['s0'][b], cfg2['cx'][b], cfg2['cy'][b],
                           integrator='rk4')
        assert ref['status'] == 0
        assert_parity(out, j, ref)
    # the returned trajectory satisfies the RK4 dynamics, not the Euler ones
    X, U = out['X'][0], out['U'][0]
    f4 = np.stack([V.rk4_step(V.MODEL_DYNAMIC, X[:, i], U[:, i], 0.05) for i in range(30)], axis=1)
    fe = V.f_dynamic(X[:, :-1].T, U.T, 0.05).T
    assert np.abs(X[:, 1:] - f4).max() <= 1e-4
    assert np.abs(X[:, 1:] - fe).max() > 1e-3


def test_restoration_phase_rescues_lane_row_failures():
    """mpcr_config.restoration: when the filter line search fails the lane solver minimises
    wq/2 |d(x) - s|^2 + zeta/2 |D_R (x - x_R)|^2 over the same (x, s) until the original filter
    accepts a point with theta reduced by kappa_resto, then resumes.  The oracle's
    restoration study (oracle/restoration.py, RestoQuadNLP) reaches the same optima."""
    d = load_fixture("cfg3_t1_triple_512.npz")
    idx = [478, 499]                         # status -2 without restoration (tests/golden/*_oracle.npz)
    a = lambda k: d[k][idx]
    plain = emu_solve(a('x0'), a('u_prev'), a('s0'), a('cx'), a('cy'), max_err=a('max_err'), lane_constraint=1,
                      lanes=True)
    assert (plain['status'] == -2).all()
    out = emu_solve(a('x0'), a('u_prev'), a('s0'), a('cx'), a('cy'), max_err=a('max_err'), lane_constraint=1,
                    lanes=True, restoration=1)
    assert (out['status'] == 0).all()
    for j, b in enumerate(idx):
        ref = oracle_solve(d['x0'][b], d['u_prev'][b], d['s0'][b], d['cx'][b], d['cy'][b], lane=True,
                           max_error=float(d['max_err'][b]), options=dict(restoration=True))
        assert ref['status'] == 0
        assert abs(out['obj'][j] - ref['obj']) <= 1e-6 * abs(ref['obj']), (b, out['obj'][j], ref['obj'])
        dU = np.abs(out['U'][j] - ref['U'])
        dU[0, -1] = 0.0
        assert dU.max() <= 2e-3, (b, dU.max())      # different restoration paths, same optimum
    assert abs(out['obj'][0] - (-112828.3055199883)) <= 1e-6 * 112828.3     # scipy trust-constr optimum of #478
    # problems that do not need it are untouched by the option
    cfg2 = load_fixture("cfg2_shanghai_4096.npz")
    i = [0, 1, 2]
    r0 = emu_solve(cfg2['x0'][i], cfg2['u_prev'][i], cfg2['s0'][i], cfg2['cx'][i], cfg2['cy'][i], lanes=True)
    r1 = emu_solve(cfg2['x0'][i], cfg2['u_prev'][i], cfg2['s0'][i], cfg2['cx'][i], cfg2['cy'][i], lanes=True,
                   restoration=1)
    assert np.array_equal(r0['U'], r1['U']) and np.array_equal(r0['iters'], r1['iters'])
