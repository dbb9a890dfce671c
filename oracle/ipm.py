"""Oracle (test infrastructure): IPOPT's algorithm restated for dense NLPs.

PARITY UNPINNED: CasADi/IPOPT (casadi==3.6.5, reference requirements.txt:8) is
not vendored in the reference and cannot be installed here, and the reference
stores no solver output.  This file restates the published algorithm
(A. Waechter, L. Biegler, "On the implementation of an interior-point filter
line-search algorithm for large-scale nonlinear programming", Math. Prog. 2006)
with IPOPT 3.14's documented default options and the options the reference sets
at control/MPC.py:152-161 (max_iter 500, tol 1e-4, acceptable_tol 1e-2,
warm_start_init_point no).  The problem reaches IPOPT as
    min f(x)  s.t.  gL <= g(x) <= gU,    x free
so every inequality row gets a slack (d(x) - s = 0, dL <= s <= dU) and x itself
has no bounds.

Deliberate simplifications (documented in DESIGN.md, shared by the CUDA path):
  * restoration phase: IPOPT's l1 feasibility NLP is replaced by a quadratic-penalty variant over the
    same variables (options['restoration'] = True, off by default - the solve then stops with
    status RESTORATION_NEEDED, a failure status - restated inline below exactly as the kernels do
    it; oracle/restoration.py holds the studies of IPOPT's own formulation);
  * no watchdog procedure, no "tiny step" heuristics, no slack re-basing
    (IPOPT's slack_move);
  * the linear solver is a dense LDL^T (scipy) whose inertia is read from the
    block diagonal; IPOPT uses MUMPS.
"""
import numpy as np
import scipy.linalg as sla

# return statuses (0/1 = success for CasADi: Solve_Succeeded / Solved_To_Acceptable_Level)
SOLVE_SUCCEEDED = 0
SOLVED_TO_ACCEPTABLE_LEVEL = 1
MAXITER_EXCEEDED = -1
RESTORATION_NEEDED = -2
ERROR_IN_STEP_COMPUTATION = -3
INVALID_NUMBER = -4

STATUS_NAMES = {0: 'Solve_Succeeded', 1: 'Solved_To_Acceptable_Level',
                -1: 'Maximum_Iterations_Exceeded', -2: 'Restoration_Failed',
                -3: 'Error_In_Step_Computation', -4: 'Invalid_Number_Detected'}

DEFAULT_OPTIONS = dict(
    max_iter=500, tol=1e-4, acceptable_tol=1e-2, acceptable_iter=15,
    dual_inf_tol=1.0, constr_viol_tol=1e-4, compl_inf_tol=1e-4,
    acceptable_dual_inf_tol=1e10, acceptable_constr_viol_tol=1e-2,
    acceptable_compl_inf_tol=1e-2,
    nlp_scaling_max_gradient=100.0, nlp_scaling_min_value=1e-8,
    bound_relax_factor=1e-8, bound_push=0.01, bound_frac=0.01,
    mu_init=0.1, mu_linear_decrease_factor=0.2, mu_superlinear_decrease_power=1.5,
    barrier_tol_factor=10.0, mu_min=1e-11, tau_min=0.99, kappa_d=1e-5, kappa_sigma=1e10,
    constr_mult_init_max=1000.0, bound_mult_init_val=1.0, s_max=100.0,
    # filter line search
    theta_max_fact=1e4, theta_min_fact=1e-4, gamma_theta=1e-5, gamma_phi=1e-8,
    delta=1.0, s_theta=1.1, s_phi=2.3, eta_phi=1e-8, alpha_min_frac=0.05,
    alpha_red_factor=0.5, max_soc=4, kappa_soc=0.99,
    # inertia correction
    delta_w_init=1e-4, delta_w_min=1e-20, delta_w_max=1e40,
    kappa_w_minus=1.0 / 3.0, kappa_w_plus=8.0, kappa_w_plus_first=100.0,
    # restoration: True = the kernels' feasibility-restoration phase (quadratic penalty over the same (x, s),
    # DESIGN.md section 3), restated below; 'l1' / 'quad' = the studies of oracle/restoration.py (IPOPT's l1
    # formulation / the penalty variant as a nested NLP), which the kernels do not mirror
    restoration=False, resto_weight=1.0, kappa_resto=0.9, resto_max_iter=100, resto_relax_eq=True, resto_quad_weight=1.0,
)


class Result:
    pass


def _inertia(lu, d):
    """(#pos, #neg, #zero) eigenvalues from scipy.linalg.ldl's block diagonal."""
    n = d.shape[0]
    pos = neg = zero = 0
    i = 0
    while i < n:
        if i + 1 < n and d[i + 1, i] != 0.0:
            ev = np.linalg.eigvalsh(d[i:i + 2, i:i + 2])
            for e in ev:
                if e > 0: pos += 1
                elif e < 0: neg += 1
                else: zero += 1
            i += 2
        else:
            e = d[i, i]
            if e > 0: pos += 1
            elif e < 0: neg += 1
            else: zero += 1
            i += 1
    return pos, neg, zero


class _KKT:
    """Factorisation of the condensed augmented system

        [ H + dw I + Jd' D Jd   Jc' ] [dx ]   [r1]
        [ Jc                    0   ] [dyc] = [r2]

    where D = Sigma_s + dw (slack block eliminated exactly)."""

    def __init__(self, H, Jc, Jd, sig_s, dw):
        n, mc = H.shape[0], Jc.shape[0]
        self.n, self.mc = n, mc
        self.Jd, self.D = Jd, sig_s + dw
        K = np.zeros((n + mc, n + mc))
        K[:n, :n] = H + dw * np.eye(n) + Jd.T @ (self.D[:, None] * Jd)
        K[:n, n:] = Jc.T
        K[n:, :n] = Jc
        self.K = K
        lu, d, perm = sla.ldl(K, lower=True)
        self.lu, self.d, self.perm = lu, d, perm
        self.inertia = _inertia(lu, d)

    def ok(self):
        return self.inertia == (self.n, self.mc, 0)

    def _ldl_solve(self, b):
        if not hasattr(self, '_Lp'):
            self._Lp = np.ascontiguousarray(self.lu[self.perm])      # unit lower triangular
            self._dinv = np.linalg.inv(self.d)                       # block diagonal (1x1 / 2x2)
        y = sla.solve_triangular(self._Lp, b[self.perm], lower=True, unit_diagonal=True,
                                 check_finite=False)
        z = self._dinv @ y
        xp = sla.solve_triangular(self._Lp.T, z, lower=False, unit_diagonal=True,
                                  check_finite=False)
        out = np.empty_like(b)
        out[self.perm] = xp
        return out

    def solve(self, rx, rs, rc, rd):
        """Solve the full 4-block system
            (H+dw) dx + Jc' dyc + Jd' dyd = rx
            D ds - dyd                    = rs
            Jc dx                         = rc
            Jd dx - ds                    = rd."""
        # ds = Jd dx - rd ; dyd = D ds - rs = D (Jd dx - rd) - rs
        # => (H+dw + Jd' D Jd) dx + Jc' dyc = rx + Jd'(D rd + rs)
        r1 = rx + self.Jd.T @ (self.D * rd + rs)
        rhs = np.concatenate([r1, rc])
        sol = self._ldl_solve(rhs)
        # one step of iterative refinement (IPOPT refines too)
        sol += self._ldl_solve(rhs - self.K @ sol)
        dx, dyc = sol[:self.n], sol[self.n:]
        ds = self.Jd @ dx - rd
        dyd = self.D * ds - rs
        return dx, ds, dyc, dyd


def solve(nlp, options=None, trace=None, stop=None):
    o = dict(DEFAULT_OPTIONS)
    if options:
        o.update(options)
    n, m = nlp.n, nlp.m
    gl, gu = nlp.gl.copy(), nlp.gu.copy()
    is_eq = gl == gu
    ic = np.where(is_eq)[0]
    idn = np.where(~is_eq)[0]
    mc, md = len(ic), len(idn)

    x = nlp.w0.copy()

    # ---- gradient-based scaling at the starting point ---------------------------
    g0 = nlp.grad_f(x)
    J0 = nlp.jac(x)
    gmax = np.max(np.abs(g0)) if n else 0.0
    df = 1.0
    if gmax > o['nlp_scaling_max_gradient']:
        df = max(o['nlp_scaling_max_gradient'] / gmax, o['nlp_scaling_min_value'])
    rowmax = np.max(np.abs(J0), axis=1)
    dg = np.ones(m)
    big = rowmax > o['nlp_scaling_max_gradient']
    dg[big] = np.maximum(o['nlp_scaling_max_gradient'] / rowmax[big], o['nlp_scaling_min_value'])
    dc_s, dd_s = dg[ic], dg[idn]

    cval = gl[ic]
    dL = gl[idn] * dd_s
    dU = gu[idn] * dd_s
    hasL = np.isfinite(dL)
    hasU = np.isfinite(dU)
    # bound relaxation
    relax = lambda b: np.minimum(o['constr_viol_tol'], o['bound_relax_factor'] * np.maximum(1.0, np.abs(b)))
    dL = np.where(hasL, dL - relax(np.where(hasL, dL, 0.0)), -np.inf)
    dU = np.where(hasU, dU + relax(np.where(hasU, dU, 0.0)), np.inf)

    def evals(xx):
        gg = nlp.g(xx)
        return df * nlp.f(xx), dc_s * (gg[ic] - cval), dd_s * gg[idn]

    def jacs(xx):
        J = nlp.jac(xx)
        return dc_s[:, None] * J[ic], dd_s[:, None] * J[idn]

    def _lam_of(yc, yd):
        lam = np.zeros(m)
        lam[ic] = yc * dc_s
        lam[idn] = yd * dd_s
        return lam

    def hess(xx, yc, yd):
        return nlp.hess(xx, df, _lam_of(yc, yd))

    # ---- starting point: slacks pushed into their bounds -----------------------
    fval, c, d = evals(x)
    s = d.copy()
    kp, kf = o['bound_push'], o['bound_frac']
    both = hasL & hasU
    pL = np.where(both, np.minimum(kp * np.maximum(1, np.abs(np.where(hasL, dL, 0))), kf * (np.where(both, dU - dL, 0))),
                  kp * np.maximum(1, np.abs(np.where(hasL, dL, 0))))
    pU = np.where(both, np.minimum(kp * np.maximum(1, np.abs(np.where(hasU, dU, 0))), kf * (np.where(both, dU - dL, 0))),
                  kp * np.maximum(1, np.abs(np.where(hasU, dU, 0))))
    s = np.where(hasL, np.maximum(s, dL + pL), s)
    s = np.where(hasU, np.minimum(s, dU - pU), s)

    vL = np.where(hasL, o['bound_mult_init_val'], 0.0)
    vU = np.where(hasU, o['bound_mult_init_val'], 0.0)

    # ---- least-squares equality multipliers ---------------------------------------
    gf = df * nlp.grad_f(x)
    Jc, Jd = jacs(x)
    kk = _KKT(np.zeros((n, n)), Jc, Jd, np.zeros(md), 1.0)
    _, _, yc, yd = kk.solve(gf, vU - vL, np.zeros(mc), np.zeros(md))
    yc, yd = -yc, -yd
    if max(np.max(np.abs(yc), initial=0.0), np.max(np.abs(yd), initial=0.0)) > o['constr_mult_init_max']:
        yc[:] = 0.0
        yd[:] = 0.0

    mu = o['mu_init']
    tau = max(o['tau_min'], 1.0 - mu)
    kd = o['kappa_d']
    onlyL = hasL & ~hasU
    onlyU = hasU & ~hasL

    def slacks(ss):
        return np.where(hasL, ss - dL, 1.0), np.where(hasU, dU - ss, 1.0)

    def barrier_obj(fv, ss, mu_):
        sl, su = slacks(ss)
        phi = fv - mu_ * (np.sum(np.log(sl[hasL])) + np.sum(np.log(su[hasU])))
        phi += kd * mu_ * (np.sum(sl[onlyL]) + np.sum(su[onlyU]))
        return phi

    def theta_of(cc, dd, ss, inner=False):
        return np.sum(np.abs(cc)) + (0.0 if inner else np.sum(np.abs(dd - ss)))

    def errors(mu_, gf_, Jc_, Jd_, c_, d_, s_, yc_, yd_, vL_, vU_):
        sl, su = slacks(s_)
        rx = gf_ + Jc_.T @ yc_ + Jd_.T @ yd_
        rs = -yd_ - vL_ + vU_
        dual_inf = max(np.max(np.abs(rx), initial=0.0), np.max(np.abs(rs), initial=0.0))
        pr_inf = np.max(np.abs(c_), initial=0.0)
        if not resto:
            pr_inf = max(pr_inf, np.max(np.abs(d_ - s_), initial=0.0))
        compl = max(np.max(np.abs(sl * vL_ - mu_)[hasL], initial=0.0),
                    np.max(np.abs(su * vU_ - mu_)[hasU], initial=0.0))
        nb = int(hasL.sum() + hasU.sum())
        sum_y = np.sum(np.abs(yc_)) + np.sum(np.abs(yd_))
        sum_v = np.sum(np.abs(vL_)) + np.sum(np.abs(vU_))
        smax = o['s_max']
        s_d = max(smax, (sum_y + sum_v) / max(1, mc + md + nb)) / smax
        s_c = max(smax, sum_v / max(1, nb)) / smax
        return max(dual_inf / s_d, pr_inf, compl / s_c), dual_inf, pr_inf, compl

    filt = []
    theta0 = theta_of(c, d, s)
    theta_max = o['theta_max_fact'] * max(1.0, theta0)
    theta_min = o['theta_min_fact'] * max(1.0, theta0)
    dw_last = 0.0
    acceptable_count = 0
    status = MAXITER_EXCEEDED
    it = 0
    res = Result()
    n_fact = 0
    ls_total = 0
    n_resto = 0
    resto = False
    resto_it = 0
    wq = o['resto_weight']
    zeta = 0.0
    x_R = None
    DR2 = None
    saved = None
    theta_R = 0.0

    def f_resto(xx, dd, ss):
        return 0.5 * wq * np.sum((dd - ss) ** 2) + 0.5 * zeta * np.sum(DR2 * (xx - x_R) ** 2)

    while True:
        gf = zeta * DR2 * (x - x_R) if resto else df * nlp.grad_f(x)
        Jc, Jd = jacs(x)
        if resto and resto_it >= 1:
            # ---- leave the restoration phase?  acceptable to the ORIGINAL filter, theta reduced ------
            fo = df * nlp.f(x)
            th_o = theta_of(c, d, s)
            mu_o, tau_o, thmax_o, thmin_o, filt_o, acc_o = saved
            leave = np.isfinite(fo) and th_o <= o['kappa_resto'] * theta_R and th_o <= thmax_o
            if leave:
                ph_o = barrier_obj(fo, s, mu_o)
                leave = np.isfinite(ph_o) and all((th_o <= tf or ph_o <= pf) for (tf, pf) in filt_o)
            if leave:
                resto = False
                mu, tau, theta_max, theta_min, filt, acceptable_count = mu_o, tau_o, thmax_o, thmin_o, filt_o, acc_o
                vL = np.where(hasL, o['bound_mult_init_val'], 0.0)
                vU = np.where(hasU, o['bound_mult_init_val'], 0.0)
                gf = df * nlp.grad_f(x)
                kk = _KKT(np.zeros((n, n)), Jc, Jd, np.zeros(md), 1.0)
                _, _, yc, yd = kk.solve(gf, vU - vL, np.zeros(mc), np.zeros(md))
                yc, yd = -yc, -yd
                if max(np.max(np.abs(yc), initial=0.0), np.max(np.abs(yd), initial=0.0)) > o['constr_mult_init_max']:
                    yc[:] = 0.0
                    yd[:] = 0.0
                fval = fo
            elif resto_it >= o['resto_max_iter']:
                status = RESTORATION_NEEDED
                break
        if not (np.all(np.isfinite(gf)) and np.all(np.isfinite(Jc)) and np.all(np.isfinite(Jd))
                and np.isfinite(fval)):
            status = INVALID_NUMBER
            break
        if stop is not None and it > 0 and stop(x, s):
            status = SOLVE_SUCCEEDED
            break
        # ---- convergence tests --------------------------------------------------
        E0, dual_inf, pr_inf, compl0 = errors(0.0, gf, Jc, Jd, c, d, s, yc, yd, vL, vU)
        gg = nlp.g(x)
        viol = max(np.max(np.maximum(nlp.gl - gg, 0.0), initial=0.0),
                   np.max(np.maximum(gg - nlp.gu, 0.0), initial=0.0))
        dfe = 1.0 if resto else df          # the restoration objective is not rescaled
        if trace is not None:
            trace.append(dict(it=it, obj=fval / dfe, E0=E0, dual_inf=dual_inf, pr_inf=pr_inf,
                              compl=compl0, mu=mu, x=x.copy()))
        if (E0 <= o['tol'] and dual_inf / dfe <= o['dual_inf_tol'] and (resto or viol <= o['constr_viol_tol'])
                and compl0 / dfe <= o['compl_inf_tol']):
            # restoration: a stationary point of the infeasibility measure that the filter rejects
            status = RESTORATION_NEEDED if resto else SOLVE_SUCCEEDED
            break
        if (not resto and E0 <= o['acceptable_tol'] and dual_inf / df <= o['acceptable_dual_inf_tol']
                and viol <= o['acceptable_constr_viol_tol']
                and compl0 / df <= o['acceptable_compl_inf_tol']):
            acceptable_count += 1
            if acceptable_count >= o['acceptable_iter']:
                status = SOLVED_TO_ACCEPTABLE_LEVEL
                break
        else:
            acceptable_count = 0
        # What a failed line search does at THIS iterate (IpBacktrackingLineSearch.cpp, the block
        # before PerformRestoration): at an almost feasible point (scaled violation <= 1e-2 tol) the
        # restoration phase is not started; the stored acceptable point is restored ("Solved To
        # Acceptable Level") or, with none, the run ends "Restoration Failed".  The stored point is
        # the latest iterate that passed the acceptable test; only the case "that is the current
        # iterate" (acceptable_count > 0) is followed here, an older stored point is not kept.
        ls_fail_mode = 0
        if not resto and pr_inf <= 1e-2 * o['tol']:
            ls_fail_mode = 1 if acceptable_count > 0 else 2
        if it >= o['max_iter']:
            status = MAXITER_EXCEEDED
            break

        # ---- barrier parameter update (monotone) --------------------------------
        mu_floor = max(o['mu_min'], min(o['tol'], o['compl_inf_tol'] * dfe) / (o['barrier_tol_factor'] + 1))
        while True:
            Emu = errors(mu, gf, Jc, Jd, c, d, s, yc, yd, vL, vU)[0]
            if Emu <= o['barrier_tol_factor'] * mu and mu > mu_floor:
                mu = max(mu_floor, min(o['mu_linear_decrease_factor'] * mu,
                                       mu ** o['mu_superlinear_decrease_power']))
                tau = max(o['tau_min'], 1.0 - mu)
                filt = []
            else:
                break

        # ---- search direction ---------------------------------------------------
        sl, su = slacks(s)
        sig_s = np.where(hasL, vL / sl, 0.0) + np.where(hasU, vU / su, 0.0)
        H = hess(x, yc, yd)
        gs_bar = (-np.where(hasL, mu / sl, 0.0) + np.where(hasU, mu / su, 0.0)
                  + kd * mu * (onlyL.astype(float) - onlyU.astype(float)))
        if resto:
            # rows in the objective (yd = wq (d - s) carries their curvature weights in H): eliminating
            # ds from  wq/2 (Jd dx - ds + rd)^2 + (Sigma + dw)/2 ds^2 + gs_bar ds  scales the row's
            # condensed terms by kappa = wq / (wq + Sigma + dw)
            H = nlp.hess(x, 0.0, _lam_of(yc, yd)) + np.diag(zeta * DR2)
            rx = gf + Jc.T @ yc
            rs = None

            def resto_kkt(dw_):
                Dt = sig_s + dw_
                kap = wq / (wq + Dt)
                return _KKT(H, Jc, Jd, kap * Dt - dw_, dw_), kap, Dt
        else:
            rx = gf + Jc.T @ yc + Jd.T @ yd
            rs = gs_bar - yd
        dw = 0.0
        kkt = resto_kkt(dw)[0] if resto else _KKT(H, Jc, Jd, sig_s, dw)
        n_fact += 1
        first_try = True
        failed = False
        while not kkt.ok():
            if first_try:
                dw = o['delta_w_init'] if dw_last == 0.0 else max(o['delta_w_min'], dw_last * o['kappa_w_minus'])
                first_try = False
            else:
                if dw_last == 0.0 or 1e5 * dw_last < dw:
                    dw = dw * o['kappa_w_plus_first']
                else:
                    dw = dw * o['kappa_w_plus']
            if dw > o['delta_w_max']:
                failed = True
                break
            kkt = resto_kkt(dw)[0] if resto else _KKT(H, Jc, Jd, sig_s, dw)
            n_fact += 1
        if failed:
            status = ERROR_IN_STEP_COMPUTATION
            break
        if dw > 0.0:
            dw_last = dw
        if resto:
            _, kap, Dt = resto_kkt(dw)

            def resto_solve(c_rhs, d_rhs):
                e_eff = kap * (Dt * d_rhs + gs_bar)
                dx_, _, dyc_, _ = kkt.solve(-rx, -e_eff, -c_rhs, np.zeros(md))
                ds_ = kap * (Jd @ dx_ + d_rhs) - gs_bar / (wq + Dt)
                return dx_, ds_, dyc_, np.zeros(md)
            dx, ds, dyc, dyd = resto_solve(c, d - s)
        else:
            dx, ds, dyc, dyd = kkt.solve(-rx, -rs, -c, -(d - s))
        dvL = np.where(hasL, mu / sl - vL - vL / sl * ds, 0.0)
        dvU = np.where(hasU, mu / su - vU + vU / su * ds, 0.0)

        # ---- fraction to the boundary -------------------------------------------
        def max_step(val, dval, mask):
            sel = mask & (dval < 0)
            if not np.any(sel):
                return 1.0
            return min(1.0, np.min(-tau * val[sel] / dval[sel]))
        alpha_max = min(max_step(sl, ds, hasL), max_step(su, -ds, hasU))
        alpha_z = min(max_step(vL, dvL, hasL), max_step(vU, dvU, hasU))

        # ---- filter line search -------------------------------------------------
        theta = theta_of(c, d, s, resto)
        phi = barrier_obj(fval, s, mu)
        gphi_d = gf @ dx + gs_bar @ ds
        if resto:
            gphi_d += wq * (d - s) @ (Jd @ dx - ds)
        if gphi_d < 0:
            amin = min(o['gamma_theta'], o['gamma_phi'] * theta / (-gphi_d))
            if theta <= theta_min:
                amin = min(amin, o['delta'] * theta ** o['s_theta'] / (-gphi_d) ** o['s_phi'])
        else:
            amin = o['gamma_theta']
        amin *= o['alpha_min_frac']

        def acceptable_to_filter(th, ph):
            for (tf, pf) in filt:
                if not (th <= tf or ph <= pf):   # dominated by a filter entry
                    return False
            return True

        def le(a, b, ref):
            return a - b <= 10.0 * np.finfo(float).eps * abs(ref)

        alpha = alpha_max
        accepted = False
        n_ls = 0
        soc_done = False
        dx_c, ds_c, dyc_c, dyd_c = dx, ds, dyc, dyd
        while alpha >= amin or n_ls == 0:
            xt = x + alpha * dx_c
            st = s + alpha * ds_c
            ft, ct, dt = evals(xt)
            if resto:
                ft = f_resto(xt, dt, st)
            n_ls += 1
            ok = np.isfinite(ft) and np.all(np.isfinite(ct)) and np.all(np.isfinite(dt))
            if ok:
                tht = theta_of(ct, dt, st, resto)
                pht = barrier_obj(ft, st, mu)
                ok = np.isfinite(pht) and tht <= theta_max
            if ok:
                switching = (gphi_d < 0 and
                             alpha * (-gphi_d) ** o['s_phi'] > o['delta'] * theta ** o['s_theta'])
                if theta <= theta_min and switching:
                    armijo = le(pht - phi, o['eta_phi'] * alpha * gphi_d, phi)
                    ok = armijo
                    f_type = True
                else:
                    ok = (le(tht, (1 - o['gamma_theta']) * theta, theta)
                          or le(pht - phi, -o['gamma_phi'] * theta, phi))
                    f_type = False
                if ok:
                    ok = acceptable_to_filter(tht, pht)
            if ok:
                accepted = True
                break
            # second-order correction on the first trial step
            if (not soc_done and n_ls == 1 and np.isfinite(ft) and np.all(np.isfinite(ct))
                    and theta_of(ct, dt, st, resto) >= theta):
                soc_done = True
                c_soc = alpha * c + ct
                d_soc = alpha * (d - s) + (dt - st)
                th_old = theta
                for _ in range(o['max_soc']):
                    dxs, dss, dycs, dyds = (resto_solve(c_soc, d_soc) if resto
                                            else kkt.solve(-rx, -rs, -c_soc, -d_soc))
                    a_soc = min(max_step(sl, dss, hasL), max_step(su, -dss, hasU))
                    xt = x + a_soc * dxs
                    st = s + a_soc * dss
                    ft, ct, dt = evals(xt)
                    if resto:
                        ft = f_resto(xt, dt, st)
                    if not (np.isfinite(ft) and np.all(np.isfinite(ct))):
                        break
                    tht = theta_of(ct, dt, st, resto)
                    pht = barrier_obj(ft, st, mu)
                    if not np.isfinite(pht) or tht > theta_max:
                        break
                    switching = (gphi_d < 0 and
                                 alpha * (-gphi_d) ** o['s_phi'] > o['delta'] * theta ** o['s_theta'])
                    if theta <= theta_min and switching:
                        ok = le(pht - phi, o['eta_phi'] * alpha * gphi_d, phi)
                        f_type = True
                    else:
                        ok = (le(tht, (1 - o['gamma_theta']) * theta, theta)
                              or le(pht - phi, -o['gamma_phi'] * theta, phi))
                        f_type = False
                    if ok and acceptable_to_filter(tht, pht):
                        accepted = True
                        dx_c, ds_c, dyc_c, dyd_c = dxs, dss, dycs, dyds
                        alpha = a_soc
                        dvL = np.where(hasL, mu / sl - vL - vL / sl * dss, 0.0)
                        dvU = np.where(hasU, mu / su - vU + vU / su * dss, 0.0)
                        alpha_z = min(max_step(vL, dvL, hasL), max_step(vU, dvU, hasU))
                        break
                    if tht > o['kappa_soc'] * th_old:
                        break
                    th_old = tht
                    c_soc = a_soc * c_soc + ct
                    d_soc = a_soc * d_soc + (dt - st)
                if accepted:
                    break
            alpha *= o['alpha_red_factor']
        ls_total += n_ls
        if not accepted and ls_fail_mode == 1:
            status = SOLVED_TO_ACCEPTABLE_LEVEL
            break
        if not accepted and ls_fail_mode == 2:
            status = RESTORATION_NEEDED
            break
        if not accepted and o['restoration'] is True and not resto:
            # ---- enter the restoration phase at the current iterate (x_R, s_R): the kernels' text ------
            theta_R = theta
            filt.append(((1 - o['gamma_theta']) * theta, phi - o['gamma_phi'] * theta))
            saved = (mu, tau, theta_max, theta_min, filt, acceptable_count)
            x_R = x.copy()
            DR2 = (1.0 / np.maximum(1.0, np.abs(x_R))) ** 2
            vmax = max(np.max(np.abs(c), initial=0.0), np.max(np.abs(d - s), initial=0.0))
            s = np.where(hasL, np.maximum(s, dL + pL), s)
            s = np.where(hasU, np.minimum(s, dU - pU), s)
            vL = np.where(hasL, o['bound_mult_init_val'], 0.0)
            vU = np.where(hasU, o['bound_mult_init_val'], 0.0)
            yc = np.zeros(mc)
            yd = wq * (d - s)
            mu = max(mu, vmax)
            tau = max(o['tau_min'], 1.0 - mu)
            zeta = np.sqrt(mu)
            filt = []
            acceptable_count = 0
            resto = True
            resto_it = 0
            fval = f_resto(x, d, s)
            th0 = theta_of(c, d, s, True)
            theta_max = o['theta_max_fact'] * max(1.0, th0)
            theta_min = o['theta_min_fact'] * max(1.0, th0)
            continue
        if not accepted and o['restoration'] in ('l1', 'quad'):
            from .restoration import RestoNLP
            theta_R, phi_R = theta, phi
            filt.append(((1 - o['gamma_theta']) * theta_R, phi_R - o['gamma_phi'] * theta_R))
            cfun = lambda xx: dc_s * (nlp.g(xx)[ic] - cval)
            dfun = lambda xx: dd_s * nlp.g(xx)[idn]

            def hess_con(xx, yc_, yd_):
                lam_ = np.zeros(m)
                lam_[ic] = yc_ * dc_s
                lam_[idn] = yd_ * dd_s
                return nlp.hess(xx, 0.0, lam_)
            mu_R = max(mu, float(np.max(np.abs(np.concatenate([c, d - s])), initial=0.0)))
            if o['restoration'] == 'quad':
                from .restoration import RestoQuadNLP
                rn = RestoQuadNLP(n, cfun, dfun, jacs, hess_con, dL, dU, x, s, mu_R, weight=o['resto_quad_weight'])
            else:
                rn = RestoNLP(n, cfun, dfun, jacs, hess_con, dL, dU, x, s, mu_R, relax_eq=o['resto_relax_eq'])
            found = {}

            def resto_stop(w, s_in):
                xx = w[:n]
                ss = w[n:n + md] if o['restoration'] == 'quad' else s_in[:md]
                fv, cc, dd = evals(xx)
                th = theta_of(cc, dd, ss)
                if not (np.isfinite(fv) and th <= o['kappa_resto'] * theta_R and th <= theta_max):
                    return False
                ph = barrier_obj(fv, ss, mu)
                if not (np.isfinite(ph) and acceptable_to_filter(th, ph)):
                    return False
                found.update(x=xx.copy(), s=ss.copy())
                return True
            ro = dict(o)
            ro.update(restoration=False, nlp_scaling_max_gradient=1e300, mu_init=mu_R, max_iter=o['resto_max_iter'],
                      tol=1e-12, acceptable_tol=1e-12)
            rr = solve(rn, ro, stop=resto_stop)
            n_resto += rr.iters
            if found:
                x, s = found['x'], found['s']
                fval, c, d = evals(x)
                vL = np.where(hasL, o['bound_mult_init_val'], 0.0)
                vU = np.where(hasU, o['bound_mult_init_val'], 0.0)
                gf = df * nlp.grad_f(x)
                Jc, Jd = jacs(x)
                kk = _KKT(np.zeros((n, n)), Jc, Jd, np.zeros(md), 1.0)
                _, _, yc, yd = kk.solve(gf, vU - vL, np.zeros(mc), np.zeros(md))
                yc, yd = -yc, -yd
                if max(np.max(np.abs(yc), initial=0.0), np.max(np.abs(yd), initial=0.0)) > o['constr_mult_init_max']:
                    yc[:] = 0.0
                    yd[:] = 0.0
                it += rr.iters
                continue
        if not accepted:
            status = RESTORATION_NEEDED
            break
        if not f_type:
            filt.append(((1 - o['gamma_theta']) * theta, phi - o['gamma_phi'] * theta))

        # ---- accept the trial point ---------------------------------------------
        x, s = xt, st
        fval, c, d = ft, ct, dt
        yc = yc + alpha * dyc_c
        yd = wq * (d - s) if resto else yd + alpha * dyd_c
        if resto:
            resto_it += 1
            n_resto += 1
        vL = vL + alpha_z * dvL
        vU = vU + alpha_z * dvU
        sl, su = slacks(s)
        ks = o['kappa_sigma']
        vL = np.where(hasL, np.maximum(np.minimum(vL, ks * mu / sl), mu / (ks * sl)), 0.0)
        vU = np.where(hasU, np.maximum(np.minimum(vU, ks * mu / su), mu / (ks * su)), 0.0)
        it += 1

    lam = np.zeros(m)
    lam[ic] = yc * dc_s / df
    lam[idn] = yd * dd_s / df
    res.x = x
    res.lam_g = lam
    res.obj = nlp.f(x)
    res.status = status
    res.iters = it
    res.mu = mu
    res.df = df
    res.n_fact = n_fact
    res.ls_total = ls_total
    res.n_resto = n_resto
    res.success = status in (SOLVE_SUCCEEDED, SOLVED_TO_ACCEPTABLE_LEVEL)
    return res
