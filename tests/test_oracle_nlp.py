"""The oracle NLP (oracle/nlp.py) against an independent torch restatement of
control/MPC.py:62-149 differentiated with autograd, and structural checks
(sizes, row order, initial guess) from SURVEY 8a."""
import numpy as np
import pytest
import torch

from oracle import vehicle as V
from oracle.nlp import RacingNLP, DEFAULT_PARAMS
from test_oracle_derivatives import _torch_model

CX = [1.2339473453246648e-06, -0.0005915402107889742, 0.08892755002816716, -4.445749591460687,
      219.69532953595657]
CY = [1.5622873116791219e-06, -0.0004883413317220176, 0.04308071072198656, -1.0575964331053849,
      86.02026634631152]
X0 = [171, 91.8, -0.219, 20, 0.48, -0.059]


def _poly(c, s):
    f = 0
    for i, ck in enumerate(reversed(c)):       # control/util.py:4-8
        f = f + ck * s ** i
    return f


def _dpoly(c, s):
    f = 0
    for i, ck in enumerate(reversed(c)):
        if i >= 1:
            f = f + i * ck * s ** (i - 1)
    return f


def torch_nlp(w, N, Ts, x0, u_prev, s0, p, model=V.MODEL_DYNAMIC, wrap=True, lane=None):
    """(J, g) exactly in the order control/MPC.py builds them."""
    U = w[:2 * N].reshape(N, 2).T            # (2,N)
    S = w[2 * N:3 * N + 1]
    X = w[3 * N + 1:].reshape(N + 1, 6).T    # (6,N+1)

    def eC(s, x): return _dpoly(CY, s) * (x[0] - _poly(CX, s)) - _dpoly(CX, s) * (x[1] - _poly(CY, s))
    def eL(s, x): return -_dpoly(CX, s) * (x[0] - _poly(CX, s)) - _dpoly(CY, s) * (x[1] - _poly(CY, s))

    J = -p['lambda_s'] * S[N]
    J = J + p['q_v_y'] * X[4, N] ** 2
    J = J + p['alpha_c'] * eC(S[N], X[:, N]) ** p['n']
    J = J + p['alpha_L'] * eL(S[N], X[:, N]) ** 2
    J = J + torch.exp(p['q_v_max'] * (X[3, N] - p['v_max']))
    for i in range(1, N):
        J = J + p['q_v_y'] * X[4, i] ** 2
        J = J + p['alpha_c'] * eC(S[i], X[:, i]) ** p['n']
        J = J + p['alpha_L'] * eL(S[i], X[:, i]) ** 2
        J = J + p['beta_delta'] * (U[1, i] - U[1, i - 1]) ** 2
        J = J + torch.exp(p['q_v_max'] * (X[3, i] - p['v_max']))
    g = [S[0]] + [X[k, 0] for k in range(6)]
    for i in range(1, N + 1):
        f = _torch_model(model, torch.cat([X[:, i - 1], U[:, i - 1]]), Ts)
        g += list(X[:, i] - f)
        g.append(S[i] - S[i - 1])
        if lane is not None:
            g.append(eC(S[i], X[:, i]))
    for i in range(N):
        g += [U[0, i], U[0, i], U[1, i], U[1, i]]
        if i > 0 or wrap:
            g += [U[0, i] - U[0, i - 1], U[1, i] - U[1, i - 1]]     # i-1 = -1 wraps (MPC.py:142-143)
    if u_prev is not None:
        g += [U[0, 0] - u_prev[0], U[1, 0] - u_prev[1]]
    return J, torch.stack(g)


@pytest.mark.parametrize("N,Ts,model,wrap,lane", [
    (6, 0.05, V.MODEL_DYNAMIC, True, None),
    (5, 0.1, V.MODEL_DYNAMIC, False, 3.0),
    (5, 0.05, V.MODEL_KINEMATIC, True, None),
    (4, 0.05, V.MODEL_BLENDED, True, 2.0),
])
def test_nlp_against_autograd(N, Ts, model, wrap, lane):
    p = dict(DEFAULT_PARAMS)
    x0 = list(X0)
    if model == V.MODEL_BLENDED:
        x0[3] = 9.0
    nlp = RacingNLP(x0, [0.19, 0.3], 69.6, CX, CY, N=N, Ts=Ts, model=model, wrap=wrap,
                    lane=lane is not None, max_error=lane)
    rng = np.random.default_rng(3)
    w = nlp.w0 + rng.normal(0, 0.05, nlp.n)
    lam = rng.normal(0, 1, nlp.m)
    wt = torch.tensor(w, dtype=torch.float64, requires_grad=True)
    fn = lambda ww: torch_nlp(ww, N, Ts, x0, [0.19, 0.3], 69.6, p, model, wrap, lane)
    Jt, gt = fn(wt)
    assert gt.shape[0] == nlp.m
    np.testing.assert_allclose(nlp.f(w), Jt.item(), rtol=1e-12)
    np.testing.assert_allclose(nlp.g(w), gt.detach().numpy(), rtol=1e-12, atol=1e-12)
    gJ = torch.autograd.grad(Jt, wt, retain_graph=True)[0].numpy()
    np.testing.assert_allclose(nlp.grad_f(w), gJ, rtol=1e-10, atol=1e-7)
    jac_t = torch.autograd.functional.jacobian(lambda ww: fn(ww)[1], wt).numpy()
    np.testing.assert_allclose(nlp.jac(w), jac_t, rtol=1e-10, atol=1e-11)
    lam_t = torch.tensor(lam)
    sig = 0.37
    H_t = torch.autograd.functional.hessian(
        lambda ww: sig * fn(ww)[0] + (lam_t * fn(ww)[1]).sum(), wt).numpy()
    H_o = nlp.hess(w, sig, lam)
    scale = np.abs(H_t).max()
    np.testing.assert_allclose(H_o, H_t, rtol=1e-9, atol=1e-12 * scale)


def test_sizes_and_bounds():
    for N in (30, 25, 15):
        nlp = RacingNLP(X0, [0.19, 0.63], 69.6, CX, CY, N=N, Ts=0.05)
        assert nlp.n == 9 * N + 7
        assert nlp.m == 13 * N + 9                      # lam_g length
        eq = nlp.gl == nlp.gu
        assert eq.sum() == 6 * N + 7
        assert (~eq).sum() == 7 * N + 2
    nlp = RacingNLP(X0, None, 69.6, CX, CY, N=15, Ts=0.05, u_warm=np.zeros((2, 15)))
    assert nlp.m == 13 * 15 + 7
    # row order: s0, x0(6), then per stage 6 dyn + 1 progress row
    kinds = [r[0] for r in nlp.rows]
    assert kinds[:7] == ['s0'] + ['x0'] * 6
    assert kinds[7:14] == ['dyn'] * 6 + ['ds']
    assert kinds[7 + 7 * 15:7 + 7 * 15 + 6] == ['thr_ub', 'thr_lb', 'st_ub', 'st_lb', 'dthr', 'dst']


def test_initial_guess():
    """control/MPC.py:109-131."""
    N, Ts = 25, 0.1
    nlp = RacingNLP(X0, [0.19, 0.63], 69.6, CX, CY, N=N, Ts=Ts)
    U, S, X = nlp.split(nlp.w0)
    assert np.all(U[:, 0] == 0.19) and np.all(U[:, 1] == 0.63)
    np.testing.assert_allclose(S, 69.6 + np.arange(N + 1) * Ts * 50.0, rtol=0, atol=1e-12)
    x = np.array(X0, dtype=float)
    for i in range(1, N + 1):
        x = V.f_dynamic(x, [0.19, 0.63], Ts)
        np.testing.assert_array_equal(X[i], x)
    # warm start = shifted last controls, last one repeated
    uw = np.arange(2 * N, dtype=float).reshape(2, N) / 100
    nlp = RacingNLP(X0, [0.19, 0.63], 69.6, CX, CY, N=N, Ts=Ts, u_warm=uw)
    U, _, _ = nlp.split(nlp.w0)
    np.testing.assert_array_equal(U[:-1, 0], uw[0, 1:])
    assert U[-1, 0] == uw[0, -1] and U[-1, 1] == uw[1, -1]


def test_rk4_nlp_derivatives_by_finite_differences():
    """NLP with the RK4 integrator (extension): Jacobian and Lagrangian Hessian against central
    differences of g and of grad L."""
    from mpc_racing_b200.workload import load_fixture
    d = load_fixture("prod_t4_256.npz")
    nlp = RacingNLP(d['x0'][2], d['u_prev'][2], d['s0'][2], d['cx'][2], d['cy'][2], N=6, Ts=0.05,
                    integrator='rk4')
    rng = np.random.default_rng(0)
    w = nlp.w0 + 1e-2 * rng.normal(size=nlp.n)
    lam = rng.normal(size=nlp.m)
    J = nlp.jac(w)
    H = nlp.hess(w, 0.0, lam)     # constraint part only: the cost's degree-4 polynomials in absolute
    #                               progress (terms ~1e7) drown a finite-difference check in round-off
    h = 1e-6
    Jfd = np.zeros_like(J)
    Hfd = np.zeros_like(H)
    gl = lambda ww: nlp.jac(ww).T @ lam
    for k in range(nlp.n):
        e = np.zeros(nlp.n)
        e[k] = h
        Jfd[:, k] = (nlp.g(w + e) - nlp.g(w - e)) / (2 * h)
        Hfd[:, k] = (gl(w + e) - gl(w - e)) / (2 * h)
    assert np.abs(J - Jfd).max() <= 1e-6 * max(1.0, np.abs(J).max())
    assert np.abs(H - Hfd).max() <= 1e-5 * max(1.0, np.abs(H).max())
