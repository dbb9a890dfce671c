"""Shared helpers for the parity tests (oracle vs emulated / CUDA solver)."""
import numpy as np

from oracle import ipm
from oracle.nlp import RacingNLP

# Tolerances (BASELINE.json north_star): |du| <= 1e-4 relative, objective 1e-6 relative, same
# status.  The last throttle U[0, N-1] is a degenerate component of the reference NLP (it only
# enters the cost through exp(q_vmax (v_x,N - v_max)) ~ e^-60): it is compared
# with the loose bound below instead.
U_TOL = 1e-4
OBJ_RTOL = 1e-6
LAST_THROTTLE_TOL = 0.1
STATE_TOL = 1e-4
LAM_RTOL = 1e-4


def oracle_solve(x0, u_prev, s0, cx, cy, N=30, Ts=0.05, options=None, **kw):
    nlp = RacingNLP(x0, u_prev, s0, cx, cy, N=N, Ts=Ts, **kw)
    r = ipm.solve(nlp, options)
    U, S, X = nlp.split(r.x)
    return dict(U=U.T.copy(), X=X.T.copy(), S=S.copy(), obj=r.obj, status=r.status, iters=r.iters,
                lam_g=r.lam_g, m=nlp.m, nlp=nlp)


def assert_parity(res, b, ref, check_lam=True):
    """res: batched result (arrays indexed [b]); ref: oracle_solve() output."""
    assert int(res['status'][b]) == ref['status'], (int(res['status'][b]), ref['status'])
    if ref['status'] < 0:
        return
    assert abs(res['obj'][b] - ref['obj']) <= OBJ_RTOL * abs(ref['obj']), (res['obj'][b], ref['obj'])
    dU = np.abs(res['U'][b] - ref['U'])
    scale = np.maximum(1.0, np.abs(ref['U']))
    last = dU[0, -1]
    dU[0, -1] = 0.0
    assert (dU / scale).max() <= U_TOL, (dU / scale).max()
    assert last <= LAST_THROTTLE_TOL, last
    dX = np.abs(res['X'][b] - ref['X']) / np.maximum(1.0, np.abs(ref['X']))
    dX[3:, -1] = 0.0            # (v_x, v_y, r) at stage N depend on the degenerate input
    assert dX.max() <= STATE_TOL, dX.max()
    assert np.abs(res['S'][b] - ref['S']).max() <= STATE_TOL * max(1.0, np.abs(ref['S']).max())
    if check_lam:
        lam = res['lam_g'][b][:ref['m']]
        den = max(1.0, np.abs(ref['lam_g']).max())
        assert np.abs(lam - ref['lam_g']).max() / den <= LAM_RTOL, np.abs(lam - ref['lam_g']).max() / den
