"""Parity tests proper: the CUDA path, through the C ABI, against the oracle."""
import numpy as np
import pytest

from mpc_racing_b200.workload import load_fixture, make_batch
from oracle import vehicle as V
from parity_util import assert_parity, oracle_solve

pytestmark = pytest.mark.gpu


def _asdict(res):
    return dict(U=res.U, X=res.X, S=res.S, eC=res.eC, eL=res.eL, lam_g=res.lam_g, obj=res.obj,
                status=res.status, iters=res.iters)


@pytest.fixture(scope="module")
def cfg2():
    return load_fixture("cfg2_shanghai_4096.npz")


@pytest.fixture(scope="module")
def solver():
    from mpc_racing_b200.solver import BatchedMPC
    return BatchedMPC(N=30, Ts=0.05, model="dynamic")


def test_parity_default_horizon(cfg2, solver):
    idx = list(range(24))
    res = _asdict(solver.solve(cfg2['x0'][idx], cfg2['u_prev'][idx], cfg2['s0'][idx], cfg2['cx'][idx],
                               cfg2['cy'][idx]))
    for j, b in enumerate(idx):
        ref = oracle_solve(cfg2['x0'][b], cfg2['u_prev'][b], cfg2['s0'][b], cfg2['cx'][b], cfg2['cy'][b])
        assert_parity(res, j, ref)


def test_matches_emulated_kernel_logic(cfg2, solver):
    """CUDA build vs the CPU emulation of the same source: identical iteration counts."""
    from emu_util import emu_solve
    idx = list(range(32, 64))
    res = solver.solve(cfg2['x0'][idx], cfg2['u_prev'][idx], cfg2['s0'][idx], cfg2['cx'][idx], cfg2['cy'][idx])
    emu = emu_solve(cfg2['x0'][idx], cfg2['u_prev'][idx], cfg2['s0'][idx], cfg2['cx'][idx], cfg2['cy'][idx])
    assert (res.status == emu['status']).all()
    assert np.mean(res.iters == emu['iters']) >= 0.9
    same = res.iters == emu['iters']
    assert np.abs(res.obj - emu['obj'])[same].max() <= 1e-9 * np.abs(emu['obj']).max()


@pytest.mark.parametrize("kw,okw", [
    (dict(model="kinematic"), dict(model=V.MODEL_KINEMATIC)),
    (dict(wrap_constraint=False), dict(wrap=False)),
    (dict(model="blended"), dict(model=V.MODEL_BLENDED)),
    (dict(restoration=True), dict(options=dict(restoration=True))),
    (dict(model="blended", integrator="rk4"), dict(model=V.MODEL_BLENDED, integrator="rk4")),
])
def test_parity_variants(cfg2, kw, okw):
    from mpc_racing_b200.solver import BatchedMPC
    s = BatchedMPC(N=30, Ts=0.05, **kw)
    idx = [0, 1, 2, 3]
    res = _asdict(s.solve(cfg2['x0'][idx], cfg2['u_prev'][idx], cfg2['s0'][idx], cfg2['cx'][idx], cfg2['cy'][idx]))
    for j, b in enumerate(idx):
        ref = oracle_solve(cfg2['x0'][b], cfg2['u_prev'][b], cfg2['s0'][b], cfg2['cx'][b], cfg2['cy'][b], **okw)
        assert_parity(res, j, ref)


def test_parity_lane_rows_and_production_horizon():
    from mpc_racing_b200.solver import BatchedMPC
    d = load_fixture("cfg3_t1_triple_512.npz")
    s = BatchedMPC(N=30, Ts=0.05, lane_constraint=True)
    idx = [0, 1, 2]
    res = _asdict(s.solve(d['x0'][idx], d['u_prev'][idx], d['s0'][idx], d['cx'][idx], d['cy'][idx],
                          max_err=d['max_err'][idx]))
    for j, b in enumerate(idx):
        ref = oracle_solve(d['x0'][b], d['u_prev'][b], d['s0'][b], d['cx'][b], d['cy'][b], lane=True,
                           max_error=float(d['max_err'][b]))
        assert_parity(res, j, ref)
    d = load_fixture("prod_t4_256.npz")
    s = BatchedMPC(N=15, Ts=0.05)
    idx = [0, 1, 2, 3]
    res = _asdict(s.solve(d['x0'][idx], d['u_prev'][idx], d['s0'][idx], d['cx'][idx], d['cy'][idx]))
    for j, b in enumerate(idx):
        ref = oracle_solve(d['x0'][b], d['u_prev'][b], d['s0'][b], d['cx'][b], d['cy'][b], N=15)
        assert_parity(res, j, ref)


def test_device_pointer_path_is_bit_identical(cfg2, solver):
    import torch
    from mpc_racing_b200.solver import BatchedMPC
    B = 256
    ds = BatchedMPC(N=30, Ts=0.05, on_device=True)
    host = solver.solve(cfg2['x0'][:B], cfg2['u_prev'][:B], cfg2['s0'][:B], cfg2['cx'][:B], cfg2['cy'][:B])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a[:B])).cuda()
    dev = ds.solve(t(cfg2['x0']), t(cfg2['u_prev']), t(cfg2['s0']), t(cfg2['cx']), t(cfg2['cy']))
    torch.cuda.synchronize()
    for k in ("U", "X", "S", "eC", "eL", "lam_g", "obj", "status", "iters"):
        assert np.array_equal(getattr(dev, k).cpu().numpy(), getattr(host, k)), k


def test_full_size_properties(cfg2, solver):
    """BASELINE config 2 at full size (4096): size-independent properties of a solution."""
    B = 4096
    r = solver.solve(cfg2['x0'], cfg2['u_prev'], cfg2['s0'], cfg2['cx'], cfg2['cy'])
    conv = r.status >= 0
    assert conv.mean() >= 0.97, conv.mean()
    U, X, S = r.U[conv], r.X[conv], r.S[conv]
    # dynamics residual of the returned trajectory (explicit Euler, control/MPC.py:186-229)
    f = V.f_dynamic(np.moveaxis(X[:, :, :-1], 1, 2), np.moveaxis(U, 1, 2), 0.05)
    resid = np.abs(np.moveaxis(X[:, :, 1:], 1, 2) - f).reshape(len(X), -1).max(axis=1)
    st = r.status[conv]
    assert resid[st == 0].max() <= 1e-4 * 1.01        # constr_viol_tol
    assert resid.max() <= 1e-2 * 1.01                 # acceptable_constr_viol_tol (status 1)
    # bounds (relaxed by IPOPT's 1e-8) and rate rows, incl. the wrap rows
    assert U[:, 0].max() <= 0.85 + 1e-6 and U[:, 0].min() >= -1 - 1e-6
    assert np.abs(U[:, 1]).max() <= 0.9 + 1e-6
    dthr = U[:, 0] - np.roll(U[:, 0], 1, axis=1)
    dst = U[:, 1] - np.roll(U[:, 1], 1, axis=1)
    assert dthr.min() >= -0.4 - 1e-6 and dthr.max() <= 2 + 1e-6
    assert np.abs(dst).max() <= 0.2 + 1e-6
    ds = np.diff(S, axis=1)
    assert ds.min() >= 0.1 - 1e-6 and ds.max() <= 2.5 + 1e-6
    assert np.allclose(X[:, :, 0], cfg2['x0'][conv]) and np.allclose(S[:, 0], cfg2['s0'][conv])
    assert np.abs(U[:, 0, 0] - cfg2['u_prev'][conv, 0]).max() <= 2.0 + 1e-6
    # idempotence / determinism: a second launch returns bit-identical results
    r2 = solver.solve(cfg2['x0'], cfg2['u_prev'], cfg2['s0'], cfg2['cx'], cfg2['cy'])
    assert np.array_equal(r.U, r2.U) and np.array_equal(r.iters, r2.iters)
    # batch-position independence: a permuted batch gives the permuted answers
    perm = np.random.default_rng(0).permutation(B)[:512]
    r3 = solver.solve(cfg2['x0'][perm], cfg2['u_prev'][perm], cfg2['s0'][perm], cfg2['cx'][perm], cfg2['cy'][perm])
    assert np.array_equal(r3.U, r.U[perm]) and np.array_equal(r3.status, r.status[perm])


def test_dropin_mpc_class(cfg2):
    """The reference's constructor surface (control/MPC.py:10-22, :183-184)."""
    from mpc_racing_b200.control.MPC import MPC
    from mpc_racing_b200.control.ControllerParameters import RuntimeControllerParameters
    from mpc_racing_b200.models import State
    b = 3
    x = cfg2['x0'][b]
    st = State(x=x[0], y=x[1], yaw=x[2], v_x=x[3], v_y=x[4], yaw_dot=x[5],
               throttle=cfg2['u_prev'][b, 0], steer=cfg2['u_prev'][b, 1])
    mpc = MPC(state0=st, s0=cfg2['s0'][b], centerline_x_poly_coeffs=list(cfg2['cx'][b]),
              centerline_y_poly_coeffs=list(cfg2['cy'][b]), max_error=float(cfg2['max_err'][b]),
              runtime_params=RuntimeControllerParameters())
    sol, ret, dual = mpc.solution()
    States, U, S_hat, e_c, e_l = ret
    assert States.shape == (6, 31) and U.shape == (2, 30) and S_hat.shape == (31,)
    assert len(e_c) == 30 and len(e_l) == 30 and dual.shape == (13 * 30 + 9,)
    assert sol is not None and sol.stats()['success']
    ref = oracle_solve(cfg2['x0'][b], cfg2['u_prev'][b], cfg2['s0'][b], cfg2['cx'][b], cfg2['cy'][b])
    assert abs(sol.obj - ref['obj']) <= 1e-6 * abs(ref['obj'])
    # warm-started receding-horizon call (agent.py:171-205): last_controls as list of tuples
    last = [(t, s) for t, s in zip(U[0], U[1])]
    mpc2 = MPC(state0=st, s0=cfg2['s0'][b], centerline_x_poly_coeffs=list(cfg2['cx'][b]),
               centerline_y_poly_coeffs=list(cfg2['cy'][b]), max_error=1.0,
               runtime_params=RuntimeControllerParameters(), last_controls=last, Ts=0.05, N=30)
    assert mpc2.solution()[0] is not None
    # failure convention: config 1 returns the last iterate with sol = None, dual = None
    d = load_fixture("cfg1_test_mpc.npz")
    x = d['x0'][0]
    st1 = State(x=x[0], y=x[1], yaw=x[2], v_x=x[3], v_y=x[4], yaw_dot=x[5], throttle=.19, steer=0.63)
    m1 = MPC(state0=st1, s0=69.6, centerline_x_poly_coeffs=list(d['cx'][0]),
             centerline_y_poly_coeffs=list(d['cy'][0]), max_error=float(d['max_err'][0]),
             runtime_params=RuntimeControllerParameters(), Ts=0.1, N=25)
    s1, r1, d1 = m1.solution()
    ref1 = oracle_solve(d['x0'][0], d['u_prev'][0], 69.6, d['cx'][0], d['cy'][0], N=25, Ts=0.1)
    assert (s1 is None) == (ref1['status'] < 0)
    assert r1[0].shape == (6, 26) and np.isfinite(r1[0]).all()


@pytest.mark.parametrize("name,kw,N", [("cfg2_shanghai_4096", {}, 30),
                                       ("cfg3_t1_triple_512", dict(lane_constraint=True), 30),
                                       ("cfg3_t1_triple_512_resto", dict(lane_constraint=True, restoration=True), 30),
                                       ("prod_t4_256", {}, 15)])
def test_whole_batches_against_oracle_golden(name, kw, N):
    """EVERY problem of BASELINE config 2 (4096), config 3 (512, lane rows, with and without the
    restoration phase) and the production batch (256, N = 15) against the committed oracle solutions
    (tests/golden/*_oracle.npz, generator make_oracle_solutions.py).  Kernel (Riccati + Woodbury) and
    oracle (dense LDL') follow the same algorithm but not the same round-off.  Measured on B200, round 2
    (tests/parity_report.py, DESIGN.md "Parity status"): converged / failed class equal 100 % on every
    batch, objective within 1e-6 on 100 %, inputs within 1e-4 on 99.90 % / 100 % / 100 % / 100 % - the
    config-2 rest are 4 problems (245, 910, 1622, 3231) with the objective equal to 1e-6 and one
    throttle component off by up to 5.7e-3: a direction the tol = 1e-4 termination leaves undetermined."""
    import os
    from mpc_racing_b200.solver import BatchedMPC
    d = load_fixture(name.replace("_resto", "") + ".npz")
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + "_oracle.npz"))
    if "resto" in name:
        assert (g["status"] >= 0).all()      # with the restoration phase the oracle solves all 512
    r = BatchedMPC(N=N, Ts=0.05, **kw).solve(d["x0"], d["u_prev"], d["s0"], d["cx"], d["cy"],
                                            max_err=d["max_err"] if kw else None)
    n = len(g["status"])
    assert ((r.status >= 0) != (g["status"] >= 0)).sum() <= 1          # measured: 0
    assert (r.status != g["status"]).sum() <= 2                        # measured: 1 (-3 vs -2), 0, 0, 0
    both = (r.status >= 0) & (g["status"] >= 0)
    rel = (np.abs(r.obj - g["obj"]) / np.abs(g["obj"]))[both]
    assert (rel > 1e-6).sum() <= 1, np.where(rel > 1e-6)[0]            # measured: 0 - no different minima
    dU = np.abs(r.U - g["U"]) / np.maximum(1.0, np.abs(g["U"]))
    dU[:, 0, -1] = 0.0
    du = dU.reshape(len(dU), -1).max(axis=1)[both]
    assert (du > 1e-4).sum() <= max(2, n // 500), np.where(du > 1e-4)[0]     # measured: 4 of 4096, else 0
    assert du.max() <= 1e-2                                            # measured: 5.7e-3
    assert (r.iters == g["iters"]).mean() >= 0.95                      # measured: 97.3 - 99.6 %


def test_polished_solutions_agree_with_the_compiled_oracle(cfg2):
    """The tol-polish argument made testable: kernel and compiled oracle both continue
    to 1e-9 after their tol = 1e-4 termination; every config-2 problem on which both converge to the
    same minimum (measured: all 4095 converged ones) then agrees to 2e-4 in the inputs and 1e-6 in the
    objective, 99.9 % to 1e-4 (measured: one problem, 3981, at 1.15e-4), with status and iteration
    count of the FIRST termination unchanged."""
    from mpc_racing_b200.solver import BatchedMPC
    from oracle import cxx
    args = (cfg2['x0'], cfg2['u_prev'], cfg2['s0'], cfg2['cx'], cfg2['cy'])
    plain = BatchedMPC(N=30, Ts=0.05).solve(*args)
    r = BatchedMPC(N=30, Ts=0.05, polish_tol=1e-9).solve(*args)
    assert np.array_equal(r.status, plain.status) and np.array_equal(r.iters, plain.iters)
    c = cxx.solve_batch(*args, polish_tol=1e-9)
    assert ((r.status >= 0) == (c["status"] >= 0)).all()
    both = (r.status >= 0) & (c["status"] >= 0)
    rel = np.abs(r.obj - c["obj"]) / np.abs(c["obj"])
    same_min = both & (rel <= 1e-6)
    assert same_min.sum() >= both.sum() - 1                            # measured: no different minimum
    dU = np.abs(r.U - c["U"]) / np.maximum(1.0, np.abs(c["U"]))
    dU[:, 0, -1] = 0.0
    du = dU.reshape(len(dU), -1).max(axis=1)
    assert du[same_min].max() <= 2e-4, (du[same_min].max(), np.where(same_min & (du > 1e-4))[0])
    assert (du[same_min] <= 1e-4).mean() >= 0.999
    # multipliers in lam_g order (control/MPC.py:171): measured max 2.4e-6 relative to the largest one
    m = min(r.lam_g.shape[1], c["lam_g"].shape[1])
    den = np.maximum(1.0, np.abs(c["lam_g"][:, :m]).max(axis=1))
    dl = np.abs(r.lam_g[:, :m] - c["lam_g"][:, :m]).max(axis=1) / den
    assert dl[same_min].max() <= 1e-4, dl[same_min].max()


def test_rk4_integrator_on_gpu(cfg2):
    """RK4 extension (32-entry stage Jacobian in the same kernel) against the oracle with the same integrator."""
    from mpc_racing_b200.solver import BatchedMPC
    s = BatchedMPC(N=30, Ts=0.05, integrator="rk4")
    idx = [0, 3, 7, 100]
    res = _asdict(s.solve(cfg2['x0'][idx], cfg2['u_prev'][idx], cfg2['s0'][idx], cfg2['cx'][idx], cfg2['cy'][idx]))
    for j, b in enumerate(idx):
        ref = oracle_solve(cfg2['x0'][b], cfg2['u_prev'][b], cfg2['s0'][b], cfg2['cx'][b], cfg2['cy'][b],
                           integrator='rk4')
        assert_parity(res, j, ref)


def test_restoration_phase_on_gpu():
    """mpcr_config.restoration on the GPU (the kernel build with the feasibility-restoration phase):
    config 3 (lane rows) has 24 of 512 problems on which the plain filter line search gives up."""
    import os
    from mpc_racing_b200.solver import BatchedMPC
    d = load_fixture("cfg3_t1_triple_512.npz")
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cfg3_t1_triple_512_oracle.npz"))
    args = (d['x0'], d['u_prev'], d['s0'], d['cx'], d['cy'])
    plain = BatchedMPC(N=30, Ts=0.05, lane_constraint=True).solve(*args, max_err=d['max_err'])
    resc = BatchedMPC(N=30, Ts=0.05, lane_constraint=True, restoration=True).solve(*args, max_err=d['max_err'])
    assert (plain.status == -2).sum() >= 20
    assert (resc.status >= 0).mean() >= 0.995
    # problems that never needed it: same answers (the restoration build is a separate compilation)
    untouched = plain.status >= 0
    assert (resc.iters[untouched] == plain.iters[untouched]).mean() >= 0.97
    same = untouched & (resc.iters == plain.iters)
    assert np.abs(resc.U[same] - plain.U[same]).max() <= 1e-7
    # a rescued problem against the independent optimum (scipy trust-constr) of the oracle study
    b = 478
    assert plain.status[b] == -2 and g['status'][b] == -2 and resc.status[b] == 0
    # the two launch modes of the phase - every problem in the build that carries it ("always", what
    # restoration=True picks with lane rows) against the two-stage mode (plain solver + re-solve of
    # the problems that need the phase) - return the same solutions
    two = BatchedMPC(N=30, Ts=0.05, lane_constraint=True, restoration="two-stage").solve(*args, max_err=d['max_err'])
    assert np.array_equal(two.status >= 0, resc.status >= 0)
    same2 = two.iters == resc.iters
    assert same2.mean() >= 0.97
    assert np.abs(two.obj[same2] - resc.obj[same2]).max() <= 1e-9 * np.abs(resc.obj).max()
    assert abs(resc.obj[b] - (-112828.3055199883)) <= 1e-6 * 112828.3
