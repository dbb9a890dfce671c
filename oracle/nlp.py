"""Oracle (test infrastructure): the NLP of control/MPC.py restated in numpy.

Follows control/MPC.py line by line:
  :57-59   Ts, max/min progress increment
  :62-64   decision variables U(2,N), S_hat(1,N+1), States(6,N+1); Opti's flat
           vector is creation order, each matrix column-major
  :71-81   Gx, Gy, dGx, dGy (control/util.py:4-8 make_poly), e_hat_C, e_hat_L
  :86-98   cost
  :101-107 initial-condition equalities
  :109-131 initial guess
  :133-149 constraints, in ``subject_to`` order (= order of ``lam_g``)
with the bound / canonical-expression convention of CasADi's Opti: a constraint
against a constant becomes ``lb <= expr <= ub`` on the non-constant side, an
equality between two expressions becomes ``lhs - rhs == 0``.

Every function is dense; sizes are n = 9N+7 variables, m = 13N+7 (+2) rows.
"""
import numpy as np

from . import vehicle as V

DEFAULT_PARAMS = dict(
    # control/ControllerParameters.py:4-23 (fixed)
    lambda_s=300.0, alpha_L=500.0, min_steer=-0.9, max_steer=0.9, min_throttle=-1.0,
    max_steer_delta=0.2, min_steer_delta=-0.2, max_throttle_delta=2.0,
    min_throttle_delta=-0.4, q_v_max=2.0, v_max=50.0, Ts=0.05, N=30, max_iter=500,
    # control/ControllerParameters.py:26-32 (runtime)
    alpha_c=1000.0, d_max=0.85, q_v_y=50.0, n=2, beta_delta=5000.0,
    # control/MPC.py:59
    min_s_delta=0.1,
)


def poly_all(c, s):
    """Value and first three derivatives of make_poly (control/util.py:4-8):
    sum_k c[deg-k] * s**k, coefficients highest order first."""
    c = np.asarray(c, dtype=np.float64)
    deg = len(c) - 1
    lo = c[::-1]
    s = np.asarray(s, dtype=np.float64)
    p = np.zeros_like(s)
    d1 = np.zeros_like(s)
    d2 = np.zeros_like(s)
    d3 = np.zeros_like(s)
    for k in range(deg + 1):
        p = p + lo[k] * s ** k
        if k >= 1:
            d1 = d1 + k * lo[k] * s ** (k - 1)
        if k >= 2:
            d2 = d2 + k * (k - 1) * lo[k] * s ** (k - 2)
        if k >= 3:
            d3 = d3 + k * (k - 1) * (k - 2) * lo[k] * s ** (k - 3)
    return p, d1, d2, d3


class RacingNLP:
    """min f(w) s.t. gl <= g(w) <= gu, as CasADi's Opti hands it to IPOPT."""

    def __init__(self, x0, u_prev, s0, cx, cy, params=None, N=None, Ts=None,
                 u_warm=None, model=V.MODEL_DYNAMIC, wrap=True, lane=False,
                 max_error=None, integrator='euler'):
        p = dict(DEFAULT_PARAMS)
        if params:
            p.update(params)
        self.p = p
        self.N = N = int(p['N'] if N is None else N)
        self.Ts = Ts = float(p['Ts'] if Ts is None else Ts)
        self.model = model
        # 'euler' = the maps control/MPC.py uses (parity); 'rk4' = extension (oracle/vehicle.py:rk4_step)
        self.integrator = integrator
        self.wrap = bool(wrap)
        self.lane = bool(lane)
        self.max_error = max_error
        self.x0 = np.asarray(x0, dtype=np.float64)
        u_prev = np.array([np.nan, np.nan]) if u_prev is None else np.asarray(u_prev, dtype=np.float64)
        self.u_prev = u_prev
        self.s0 = float(s0)
        self.cx = np.asarray(cx, dtype=np.float64)
        self.cy = np.asarray(cy, dtype=np.float64)
        self.n = 9 * N + 7
        # index maps
        self.iU = lambda k, i: 2 * i + k
        self.iS = lambda i: 2 * N + i
        self.iX = lambda k, i: 3 * N + 1 + 6 * i + k
        self._build_rows()
        self._build_w0(u_warm)

    # -- constraint bookkeeping ------------------------------------------------
    def _build_rows(self):
        N, p, Ts = self.N, self.p, self.Ts
        rows = []  # (kind, stage, lb, ub)
        rows.append(('s0', 0, self.s0, self.s0))
        for k in range(6):
            rows.append(('x0', k, self.x0[k], self.x0[k]))
        for i in range(1, N + 1):
            for k in range(6):
                rows.append(('dyn', (i, k), 0.0, 0.0))
            rows.append(('ds', i, p['min_s_delta'], Ts * p['v_max']))
            if self.lane:
                rows.append(('lane', i, -self.max_error, self.max_error))
        for i in range(N):
            rows.append(('thr_ub', i, -np.inf, p['d_max']))
            rows.append(('thr_lb', i, p['min_throttle'], np.inf))
            rows.append(('st_ub', i, -np.inf, p['max_steer']))
            rows.append(('st_lb', i, p['min_steer'], np.inf))
            if i > 0 or self.wrap:
                rows.append(('dthr', i, p['min_throttle_delta'], p['max_throttle_delta']))
                rows.append(('dst', i, p['min_steer_delta'], p['max_steer_delta']))
        if not np.isnan(self.u_prev[0]):
            rows.append(('dthr0', 0, p['min_throttle_delta'], p['max_throttle_delta']))
        if not np.isnan(self.u_prev[1]):
            rows.append(('dst0', 0, p['min_steer_delta'], p['max_steer_delta']))
        self.rows = rows
        self.m = len(rows)
        self.gl = np.array([r[2] for r in rows], dtype=np.float64)
        self.gu = np.array([r[3] for r in rows], dtype=np.float64)
        # constant (linear) part of the Jacobian and row index of the dyn/lane blocks
        n = self.n
        Jlin = np.zeros((self.m, n))
        self.dyn_row = np.zeros(N + 1, dtype=int)
        self.lane_row = np.zeros(N + 1, dtype=int)
        for j, (kind, st, _, _) in enumerate(rows):
            if kind == 's0':
                Jlin[j, self.iS(0)] = 1
            elif kind == 'x0':
                Jlin[j, self.iX(st, 0)] = 1
            elif kind == 'dyn':
                i, k = st
                if k == 0:
                    self.dyn_row[i] = j
                Jlin[j, self.iX(k, i)] = 1
            elif kind == 'ds':
                Jlin[j, self.iS(st)] = 1
                Jlin[j, self.iS(st - 1)] = -1
            elif kind == 'lane':
                self.lane_row[st] = j
            elif kind in ('thr_ub', 'thr_lb'):
                Jlin[j, self.iU(0, st)] = 1
            elif kind in ('st_ub', 'st_lb'):
                Jlin[j, self.iU(1, st)] = 1
            elif kind == 'dthr':
                Jlin[j, self.iU(0, st)] += 1
                Jlin[j, self.iU(0, (st - 1) % N)] += -1   # control/MPC.py:142 (i=0 wraps)
            elif kind == 'dst':
                Jlin[j, self.iU(1, st)] += 1
                Jlin[j, self.iU(1, (st - 1) % N)] += -1   # control/MPC.py:143
            elif kind == 'dthr0':
                Jlin[j, self.iU(0, 0)] = 1
            elif kind == 'dst0':
                Jlin[j, self.iU(1, 0)] = 1
        self.Jlin = Jlin
        self.goff = np.zeros(self.m)
        for j, (kind, st, _, _) in enumerate(rows):
            if kind == 'dthr0':
                self.goff[j] = -self.u_prev[0]
            elif kind == 'dst0':
                self.goff[j] = -self.u_prev[1]

    # -- initial guess (control/MPC.py:109-131) -----------------------------------
    def _build_w0(self, u_warm):
        N, Ts, p = self.N, self.Ts, self.p
        if u_warm is not None:
            uw = np.asarray(u_warm, dtype=np.float64).reshape(2, N)
            u_pred = np.concatenate([uw[:, 1:], uw[:, -1:]], axis=1)   # :120-121
        else:
            u_pred = np.repeat(self.u_prev.reshape(2, 1), N, axis=1)   # :123
        w = np.zeros(self.n)
        for i in range(N):
            w[self.iU(0, i)] = u_pred[0, i]
            w[self.iU(1, i)] = u_pred[1, i]
        x = self.x0.copy()
        w[self.iS(0)] = self.s0
        for k in range(6):
            w[self.iX(k, 0)] = x[k]
        for i in range(1, N + 1):
            w[self.iS(i)] = self.s0 + i * Ts * p['v_max']              # :127
            if self.integrator == 'rk4':
                x = V.rk4_step(V.MODEL_DYNAMIC, x, u_pred[:, i - 1], Ts)
            else:
                x = V.f_dynamic(x, u_pred[:, i - 1], Ts)               # :129 (always f_vehicle)
            for k in range(6):
                w[self.iX(k, i)] = x[k]
        self.w0 = w

    # -- unpack ---------------------------------------------------------------------
    def split(self, w):
        N = self.N
        U = w[:2 * N].reshape(N, 2)             # (N,2): row i = (thr_i, st_i)
        S = w[2 * N:3 * N + 1]                  # (N+1,)
        X = w[3 * N + 1:].reshape(N + 1, 6)     # (N+1,6)
        return U, S, X

    def errors(self, S, X):
        """e_hat_C, e_hat_L (control/MPC.py:79-80) at (S_hat[i], States[:, i])."""
        px, dpx, _, _ = poly_all(self.cx, S)
        py, dpy, _, _ = poly_all(self.cy, S)
        eC = dpy * (X[:, 0] - px) - dpx * (X[:, 1] - py)
        eL = -dpx * (X[:, 0] - px) - dpy * (X[:, 1] - py)
        return eC, eL

    def _stage_cost(self, S, X, order):
        """Stage cost l(x_i, s_i) for i = 1..N, gradient over (X,Y,s,vx,vy) and
        the dense 3x3 Hessian block over (X,Y,s) + diagonal (vx, vy) terms."""
        p = self.p
        n_exp = int(p['n'])
        px, dpx, d2px, d3px = poly_all(self.cx, S)
        py, dpy, d2py, d3py = poly_all(self.cy, S)
        dX = X[:, 0] - px
        dY = X[:, 1] - py
        eC = dpy * dX - dpx * dY
        eL = -dpx * dX - dpy * dY
        vx, vy = X[:, 3], X[:, 4]
        ex = np.exp(p['q_v_max'] * (vx - p['v_max']))
        val = p['q_v_y'] * vy ** 2 + p['alpha_c'] * eC ** n_exp + p['alpha_L'] * eL ** 2 + ex
        if order == 0:
            return val
        # gradients of eC, eL over (X, Y, s)
        gC = np.stack([dpy, -dpx, d2py * dX - d2px * dY], axis=-1)
        gL = np.stack([-dpx, -dpy, -d2px * dX + dpx ** 2 - d2py * dY + dpy ** 2], axis=-1)
        dC = p['alpha_c'] * n_exp * eC ** (n_exp - 1)
        dL = p['alpha_L'] * 2 * eL
        g3 = dC[:, None] * gC + dL[:, None] * gL
        gvx = p['q_v_max'] * ex
        gvy = 2 * p['q_v_y'] * vy
        if order == 1:
            return val, g3, gvx, gvy
        z = np.zeros_like(S)
        HC = np.empty(S.shape + (3, 3))
        HC[:, 0, 0] = z; HC[:, 0, 1] = z; HC[:, 1, 0] = z; HC[:, 1, 1] = z
        HC[:, 0, 2] = HC[:, 2, 0] = d2py
        HC[:, 1, 2] = HC[:, 2, 1] = -d2px
        HC[:, 2, 2] = d3py * dX - d2py * dpx - d3px * dY + d2px * dpy
        HL = np.empty(S.shape + (3, 3))
        HL[:, 0, 0] = z; HL[:, 0, 1] = z; HL[:, 1, 0] = z; HL[:, 1, 1] = z
        HL[:, 0, 2] = HL[:, 2, 0] = -d2px
        HL[:, 1, 2] = HL[:, 2, 1] = -d2py
        HL[:, 2, 2] = -d3px * dX - d3py * dY + 3 * (dpx * d2px + dpy * d2py)
        ddC = p['alpha_c'] * n_exp * (n_exp - 1) * eC ** (n_exp - 2) if n_exp >= 2 else z
        H3 = (ddC[:, None, None] * gC[:, :, None] * gC[:, None, :] + dC[:, None, None] * HC
              + 2 * p['alpha_L'] * gL[:, :, None] * gL[:, None, :] + dL[:, None, None] * HL)
        hvx = p['q_v_max'] ** 2 * ex
        hvy = 2 * p['q_v_y'] * np.ones_like(vy)
        return val, g3, gvx, gvy, H3, hvx, hvy

    # -- objective -------------------------------------------------------------------
    def f(self, w):
        U, S, X = self.split(w)
        p, N = self.p, self.N
        val = self._stage_cost(S[1:], X[1:], 0).sum()
        dst = U[1:N, 1] - U[0:N - 1, 1]
        return -p['lambda_s'] * S[N] + val + p['beta_delta'] * np.sum(dst ** 2)

    def grad_f(self, w):
        U, S, X = self.split(w)
        p, N = self.p, self.N
        g = np.zeros(self.n)
        _, g3, gvx, gvy = self._stage_cost(S[1:], X[1:], 1)
        for i in range(1, N + 1):
            g[self.iX(0, i)] += g3[i - 1, 0]
            g[self.iX(1, i)] += g3[i - 1, 1]
            g[self.iS(i)] += g3[i - 1, 2]
            g[self.iX(3, i)] += gvx[i - 1]
            g[self.iX(4, i)] += gvy[i - 1]
        g[self.iS(N)] += -p['lambda_s']
        dst = U[1:N, 1] - U[0:N - 1, 1]
        for i in range(1, N):
            g[self.iU(1, i)] += 2 * p['beta_delta'] * dst[i - 1]
            g[self.iU(1, i - 1)] -= 2 * p['beta_delta'] * dst[i - 1]
        return g

    # -- discretised model (Euler = reference; RK4 = extension) -----------------------------
    def _step(self, X, U):
        if self.integrator == 'rk4':
            return V.rk4_step(self.model, X, U, self.Ts)
        return V.f_model(self.model, X, U, self.Ts)

    def _derivs(self, X, U, lam):
        if self.integrator != 'rk4':
            return V.model_derivs(self.model, X, U, self.Ts, lam)
        J = np.zeros((X.shape[0], 6, 8))
        H = np.zeros((X.shape[0], 8, 8)) if lam is not None else None
        for i in range(X.shape[0]):
            Ji, Hi = V.rk4_derivs(self.model, X[i], U[i], self.Ts, None if lam is None else lam[i])
            J[i] = Ji
            if lam is not None:
                H[i] = Hi
        return J, H

    # -- constraints ------------------------------------------------------------------
    def g(self, w):
        U, S, X = self.split(w)
        N = self.N
        out = self.Jlin @ w + self.goff
        fx = self._step(X[:N], U)                           # f(x_{i-1}, u_{i-1}), i = 1..N
        for i in range(1, N + 1):
            r = self.dyn_row[i]
            out[r:r + 6] -= fx[i - 1]
        if self.lane:
            eC, _ = self.errors(S, X)
            for i in range(1, N + 1):
                out[self.lane_row[i]] = eC[i]
        return out

    def _lane_grad(self, S, X):
        px, dpx, d2px, d3px = poly_all(self.cx, S)
        py, dpy, d2py, d3py = poly_all(self.cy, S)
        dX = X[:, 0] - px
        dY = X[:, 1] - py
        gC = np.stack([dpy, -dpx, d2py * dX - d2px * dY], axis=-1)
        HC22 = d3py * dX - d2py * dpx - d3px * dY + d2px * dpy
        return gC, d2py, -d2px, HC22

    def jac(self, w):
        U, S, X = self.split(w)
        N = self.N
        J = self.Jlin.copy()
        jm, _ = self._derivs(X[:N], U, None)
        for i in range(1, N + 1):
            r = self.dyn_row[i]
            cols = [self.iX(k, i - 1) for k in range(6)] + [self.iU(0, i - 1), self.iU(1, i - 1)]
            J[r:r + 6, cols] -= jm[i - 1]
        if self.lane:
            gC, _, _, _ = self._lane_grad(S, X)
            for i in range(1, N + 1):
                r = self.lane_row[i]
                J[r, self.iX(0, i)] = gC[i, 0]
                J[r, self.iX(1, i)] = gC[i, 1]
                J[r, self.iS(i)] = gC[i, 2]
        return J

    # -- Hessian of sigma_f * f + lam^T g ---------------------------------------------
    def hess(self, w, sigma_f, lam):
        U, S, X = self.split(w)
        p, N = self.p, self.N
        H = np.zeros((self.n, self.n))
        _, _, _, _, H3, hvx, hvy = self._stage_cost(S[1:], X[1:], 2)
        for i in range(1, N + 1):
            idx = [self.iX(0, i), self.iX(1, i), self.iS(i)]
            H[np.ix_(idx, idx)] += sigma_f * H3[i - 1]
            H[self.iX(3, i), self.iX(3, i)] += sigma_f * hvx[i - 1]
            H[self.iX(4, i), self.iX(4, i)] += sigma_f * hvy[i - 1]
        b = 2 * p['beta_delta'] * sigma_f
        for i in range(1, N):
            a, c = self.iU(1, i), self.iU(1, i - 1)
            H[a, a] += b
            H[c, c] += b
            H[a, c] -= b
            H[c, a] -= b
        lam_dyn = np.stack([lam[self.dyn_row[i]:self.dyn_row[i] + 6] for i in range(1, N + 1)])
        _, hm = self._derivs(X[:N], U, lam_dyn)
        for i in range(1, N + 1):
            cols = [self.iX(k, i - 1) for k in range(6)] + [self.iU(0, i - 1), self.iU(1, i - 1)]
            H[np.ix_(cols, cols)] -= hm[i - 1]
        if self.lane:
            _, h02, h12, h22 = self._lane_grad(S, X)
            for i in range(1, N + 1):
                l = lam[self.lane_row[i]]
                a, b_, c = self.iX(0, i), self.iX(1, i), self.iS(i)
                H[a, c] += l * h02[i]; H[c, a] += l * h02[i]
                H[b_, c] += l * h12[i]; H[c, b_] += l * h12[i]
                H[c, c] += l * h22[i]
        return H
