"""Oracle (test infrastructure): STUDIES of the restoration phase - the kernels' phase itself is restated
inline in oracle/ipm.py (options['restoration'] = True).  Here:
IPOPT's feasibility-restoration phase (RestoMinC_1Nrm) restated for the dense oracle.

When the filter line search fails, IPOPT minimises the l1 constraint violation

    min  rho * sum(n_c + p_c + n_d + p_d) + zeta/2 * || D_R (x - x_R) ||^2
    s.t. c(x) + n_c - p_c = 0,   dL <= d(x) + n_d - p_d <= dU,   n, p >= 0

(rho = 1000, zeta = sqrt(mu), D_R = diag(1 / max(1, |x_R|))) with the same interior-point
algorithm, until a point is acceptable to the ORIGINAL filter and has reduced the original
infeasibility theta by the factor kappa_resto = 0.9; the regular algorithm then resumes from that
point with least-squares multipliers.  This file builds that NLP around the scaled constraint
functions of the outer problem; oracle/ipm.py calls it with options['restoration'] = 'l1' (this class)
or 'quad' (RestoQuadNLP below: the penalty variant the kernels use, as a nested NLP).
"""
import numpy as np


class RestoNLP:
    def __init__(self, n, cfun, dfun, jacs, hess_con, dL, dU, x_R, s_R, mu, rho=1000.0, relax_eq=True):
        c0 = cfun(x_R)
        d0 = dfun(x_R)
        self.nx, self.mc, self.md = n, c0.size, d0.size
        self.cfun, self.dfun, self.jacs, self.hess_con = cfun, dfun, jacs, hess_con
        self.rho = rho
        self.zeta = np.sqrt(mu)
        self.x_R = x_R.copy()
        self.DR2 = (1.0 / np.maximum(1.0, np.abs(x_R))) ** 2
        mc, md = self.mc, self.md
        self.n = n + 2 * mc + 2 * md
        self.m = mc + md + 2 * mc + 2 * md
        self.gl = np.concatenate([np.zeros(mc), dL, np.zeros(2 * mc + 2 * md)])
        self.gu = np.concatenate([np.zeros(mc), dU, np.full(2 * mc + 2 * md, np.inf)])
        if not relax_eq:
            self.gu[mc + md:mc + md + 2 * mc] = 2e-8

        # IpRestoIterateInitializer: n, p that minimise the barrier problem for fixed x
        def np_init(v):
            a = (mu - rho * v) / (2.0 * rho)
            nn = a + np.sqrt(a * a + mu * v / (2.0 * rho))
            return nn, v + nn
        nc, pc = np_init(c0)
        self.relax_eq = relax_eq
        if not relax_eq:
            # variant: equality rows stay hard (n_c = p_c pinned to ~0 by their bounds rows below)
            nc, pc = np.full(mc, 1e-8), np.full(mc, 1e-8)
        nd, pd = np_init(d0 - s_R)
        self.w0 = np.concatenate([x_R, nc, pc, nd, pd])

    def split(self, w):
        n, mc, md = self.nx, self.mc, self.md
        o = n
        return (w[:n], w[o:o + mc], w[o + mc:o + 2 * mc], w[o + 2 * mc:o + 2 * mc + md],
                w[o + 2 * mc + md:])

    def f(self, w):
        x, nc, pc, nd, pd = self.split(w)
        dx = x - self.x_R
        return self.rho * (nc.sum() + pc.sum() + nd.sum() + pd.sum()) + 0.5 * self.zeta * np.sum(self.DR2 * dx * dx)

    def grad_f(self, w):
        x = w[:self.nx]
        g = np.full(self.n, self.rho)
        g[:self.nx] = self.zeta * self.DR2 * (x - self.x_R)
        return g

    def g(self, w):
        x, nc, pc, nd, pd = self.split(w)
        return np.concatenate([self.cfun(x) + nc - pc, self.dfun(x) + nd - pd, nc, pc, nd, pd])

    def jac(self, w):
        n, mc, md = self.nx, self.mc, self.md
        Jc, Jd = self.jacs(w[:n])
        J = np.zeros((self.m, self.n))
        J[:mc, :n] = Jc
        J[mc:mc + md, :n] = Jd
        o = n
        J[:mc, o:o + mc] = np.eye(mc)
        J[:mc, o + mc:o + 2 * mc] = -np.eye(mc)
        J[mc:mc + md, o + 2 * mc:o + 2 * mc + md] = np.eye(md)
        J[mc:mc + md, o + 2 * mc + md:] = -np.eye(md)
        J[mc + md:, n:] = np.eye(2 * mc + 2 * md)
        return J

    def hess(self, w, sigma_f, lam):
        n, mc, md = self.nx, self.mc, self.md
        H = np.zeros((self.n, self.n))
        H[:n, :n] = self.hess_con(w[:n], lam[:mc], lam[mc:mc + md])
        H[np.arange(n), np.arange(n)] += sigma_f * self.zeta * self.DR2
        return H


class RestoQuadNLP:
    """Variant studied for the kernels: no elastic variables.  The inequality rows' defining
    equations d(x) - s = 0 are replaced by a quadratic penalty,

        min  w/2 * || d(x) - s ||^2 + zeta/2 * || D_R (x - x_R) ||^2
        s.t. c(x) = 0,   dL <= s <= dU

    over the SAME variables (x, s) as the original problem, so every row keeps its single slack and
    the stage-wise elimination only changes its (D_r, e_r) formulas."""

    def __init__(self, n, cfun, dfun, jacs, hess_con, dL, dU, x_R, s_R, mu, weight=1.0):
        c0 = cfun(x_R)
        self.nx, self.mc, self.md = n, c0.size, s_R.size
        self.cfun, self.dfun, self.jacs, self.hess_con = cfun, dfun, jacs, hess_con
        self.wq = weight
        self.zeta = np.sqrt(mu)
        self.x_R = x_R.copy()
        self.DR2 = (1.0 / np.maximum(1.0, np.abs(x_R))) ** 2
        self.n = n + self.md
        self.m = self.mc + self.md
        self.gl = np.concatenate([np.zeros(self.mc), dL])
        self.gu = np.concatenate([np.zeros(self.mc), dU])
        self.w0 = np.concatenate([x_R, s_R])

    def f(self, w):
        x, s = w[:self.nx], w[self.nx:]
        r = self.dfun(x) - s
        dx = x - self.x_R
        return 0.5 * self.wq * r @ r + 0.5 * self.zeta * np.sum(self.DR2 * dx * dx)

    def grad_f(self, w):
        x, s = w[:self.nx], w[self.nx:]
        r = self.dfun(x) - s
        _, Jd = self.jacs(x)
        return np.concatenate([self.wq * Jd.T @ r + self.zeta * self.DR2 * (x - self.x_R), -self.wq * r])

    def g(self, w):
        return np.concatenate([self.cfun(w[:self.nx]), w[self.nx:]])

    def jac(self, w):
        n, mc, md = self.nx, self.mc, self.md
        Jc, _ = self.jacs(w[:n])
        J = np.zeros((self.m, self.n))
        J[:mc, :n] = Jc
        J[mc:, n:] = np.eye(md)
        return J

    def hess(self, w, sigma_f, lam):
        n, mc, md = self.nx, self.mc, self.md
        x, s = w[:n], w[n:]
        r = self.dfun(x) - s
        _, Jd = self.jacs(x)
        H = np.zeros((self.n, self.n))
        H[:n, :n] = self.hess_con(x, lam[:mc], sigma_f * self.wq * r) + sigma_f * self.wq * Jd.T @ Jd
        H[np.arange(n), np.arange(n)] += sigma_f * self.zeta * self.DR2
        H[:n, n:] = -sigma_f * self.wq * Jd.T
        H[n:, :n] = -sigma_f * self.wq * Jd
        H[n + np.arange(md), n + np.arange(md)] += sigma_f * self.wq
        return H
