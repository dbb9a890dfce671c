"""Kernel LOGIC parity on the CPU: the warp-level solver source
(mpc_racing_b200/csrc/mpcr_ipm.cuh), run through the fiber SIMT emulator of
tests/emu (test infrastructure), against the numpy oracle on seeded inputs.
The same comparisons run on the real GPU in test_gpu_parity.py."""
import numpy as np
import pytest

from emu_util import emu_solve
from mpc_racing_b200.workload import load_fixture
from oracle import vehicle as V
from parity_util import assert_parity, oracle_solve


@pytest.fixture(scope="module")
def cfg2():
    return load_fixture("cfg2_shanghai_4096.npz")


def _run(d, idx, N=30, Ts=0.05, emu_kw=None, orc_kw=None, u_warm=None, u_prev=None, max_err=None,
         params=None):
    emu_kw = emu_kw or {}
    orc_kw = orc_kw or {}
    up = d['u_prev'][idx] if u_prev is None else u_prev
    out = emu_solve(d['x0'][idx], up, d['s0'][idx], d['cx'][idx], d['cy'][idx], N=N, Ts=Ts,
                    u_warm=u_warm, max_err=max_err, params=params, **emu_kw)
    for j, b in enumerate(idx):
        ref = oracle_solve(d['x0'][b], None if np.isnan(up[j]).all() else up[j], d['s0'][b],
                           d['cx'][b], d['cy'][b], N=N, Ts=Ts,
                           u_warm=None if u_warm is None else u_warm[j], **orc_kw)
        assert_parity(out, j, ref)
    return out


def test_default_horizon(cfg2):
    out = _run(cfg2, [0, 3, 5])
    assert (out['status'] == 0).all()


def test_production_horizon_N15():
    d = load_fixture("prod_t4_256.npz")
    _run(d, [0, 1, 2, 3], N=15)


def test_no_wrap_rows(cfg2):
    _run(cfg2, [0, 3], emu_kw=dict(wrap_constraint=0), orc_kw=dict(wrap=False))


def test_kinematic_and_blended_models(cfg2):
    _run(cfg2, [1, 2], emu_kw=dict(model=1), orc_kw=dict(model=V.MODEL_KINEMATIC))
    slow = [int(i) for i in np.where(cfg2['x0'][:400, 3] < 12.0)[0][:2]]
    _run(cfg2, slow, emu_kw=dict(model=2), orc_kw=dict(model=V.MODEL_BLENDED))


def test_lane_rows():
    d = load_fixture("cfg3_t1_triple_512.npz")
    idx = [1]
    for j, b in enumerate(idx):
        out = emu_solve(d['x0'][[b]], d['u_prev'][[b]], d['s0'][[b]], d['cx'][[b]], d['cy'][[b]],
                        max_err=d['max_err'][[b]], lane_constraint=1)
        ref = oracle_solve(d['x0'][b], d['u_prev'][b], d['s0'][b], d['cx'][b], d['cy'][b], lane=True,
                           max_error=float(d['max_err'][b]))
        assert ref['m'] == 13 * 30 + 9 + 30
        assert_parity(out, 0, ref)


def test_warm_start_and_no_previous_input(cfg2):
    idx = [4, 5]
    cold = emu_solve(cfg2['x0'][idx], cfg2['u_prev'][idx], cfg2['s0'][idx], cfg2['cx'][idx], cfg2['cy'][idx])
    uw = cold['U'].copy()
    _run(cfg2, idx, u_warm=uw)
    nan = np.full((2, 2), np.nan)
    _run(cfg2, idx, u_warm=uw, u_prev=nan)


def test_runtime_parameter_overrides(cfg2):
    idx = [6]
    prm = np.array([[800.0, 0.7, 30.0, 2, 2000.0]])
    out = emu_solve(cfg2['x0'][idx], cfg2['u_prev'][idx], cfg2['s0'][idx], cfg2['cx'][idx], cfg2['cy'][idx],
                    params=prm)
    ref = oracle_solve(cfg2['x0'][6], cfg2['u_prev'][6], cfg2['s0'][6], cfg2['cx'][6], cfg2['cy'][6],
                       params=dict(alpha_c=800.0, d_max=0.7, q_v_y=30.0, n=2, beta_delta=2000.0))
    assert_parity(out, 0, ref)
    assert out['U'][0, 0].max() <= 0.7 * (1 + 1e-6)


def test_failure_status_is_reported_not_raised():
    """config 1 (script/test_mpc.py input, Ts = 0.1, N = 25) is a stress case: the plain filter
    line search needs IPOPT's restoration phase, which neither the oracle nor the kernel has."""
    d = load_fixture("cfg1_test_mpc.npz")
    out = emu_solve(d['x0'], d['u_prev'], d['s0'], d['cx'], d['cy'], N=25, Ts=0.1)
    from oracle import cxx
    ref = cxx.solve_batch(d['x0'], d['u_prev'], d['s0'], d['cx'], d['cy'], N=25, Ts=0.1)
    # both report a failure status (the exact code on a diverging path is not compared)
    assert int(out['status'][0]) < 0 and int(ref['status'][0]) < 0
    assert np.isfinite(out['U']).all() and np.isfinite(out['X']).all()


def test_restoration_phase_in_the_warp_solver():
    """mpcr_config.restoration: the warp solver's restoration build against the oracle's
    restoration study (oracle/restoration.py, RestoQuadNLP) and the independent optimum."""
    d = load_fixture("cfg3_t1_triple_512.npz")
    idx = [478]
    a = lambda k: d[k][idx]
    plain = emu_solve(a('x0'), a('u_prev'), a('s0'), a('cx'), a('cy'), max_err=a('max_err'), lane_constraint=1)
    assert (plain['status'] == -2).all()
    out = emu_solve(a('x0'), a('u_prev'), a('s0'), a('cx'), a('cy'), max_err=a('max_err'), lane_constraint=1,
                    restoration=1)
    assert (out['status'] == 0).all()
    for j, b in enumerate(idx):
        ref = oracle_solve(d['x0'][b], d['u_prev'][b], d['s0'][b], d['cx'][b], d['cy'][b], lane=True,
                           max_error=float(d['max_err'][b]), options=dict(restoration=True))
        assert ref['status'] == 0
        assert abs(out['obj'][j] - ref['obj']) <= 1e-6 * abs(ref['obj'])
    assert abs(out['obj'][0] - (-112828.3055199883)) <= 1e-6 * 112828.3
