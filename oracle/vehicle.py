"""Oracle (test infrastructure): the vehicle model used INSIDE the NLP.

Restates control/MPC.py:186-283 (``f_vehicle``, ``f_vehicle_kinematic``,
``steer_cmd_to_angle``, ``Fx``), control/util.py:10-11 (``deg2rad`` with
pi ~ 3.14) and the constants of models/VehicleParameters.py:8-41, in numpy,
vectorised over leading axes, with hand-derived first and second derivatives
(checked against torch.autograd in tests/test_oracle_derivatives.py).

State x = (X, Y, yaw, v_x, v_y, r); input u = (throttle, steer).
"""
import numpy as np

# models/VehicleParameters.py:8-41
M = 1845.0
MAX_STEER_DEG = 70.0
T_MAX = 743.0
R_WHEEL = 0.37
C_WHEEL = 2 * 3.14 * R_WHEEL          # VehicleParameters.py:16 (pi ~ 3.14, as written)
R_GEAR = 9.0
RHO = 1.225
C_D = 0.23
A_F = 2.2
C_ROLL = 0.012
IZ = 3960.0
LF = 0.8
LR = 2.0
CF = 65000.0
CR = 65000.0
G = 9.81
VBLENDMIN = 2.0
VBLENDMAX = 15.0
CAR_WIDTH = 1.85

# control/MPC.py:273-283
ETA1 = -0.00004428225806
ETA0 = 1.282413306
# control/MPC.py:262-271
GAIN1 = -0.001971664699
GAIN0 = 0.986547

MODEL_DYNAMIC = 0
MODEL_KINEMATIC = 1
MODEL_BLENDED = 2


def deg2rad(z):
    """control/util.py:10-11."""
    return (z / 360) * 2 * 3.14


def Fx(throttle, v_x):
    """control/MPC.py:273-283."""
    wheel_rpm = (v_x / C_WHEEL) * 60
    rpm = wheel_rpm * R_GEAR * 4.5
    eta = ETA1 * rpm + ETA0
    wheel_force = throttle * eta * T_MAX * R_GEAR / R_WHEEL
    drag_force = 0.5 * RHO * C_D * A_F * (v_x ** 2)
    rolling_resistance = C_ROLL * M * G
    return wheel_force - drag_force - rolling_resistance


def steer_cmd_to_angle(steer_cmd, v_x, v_y):
    """control/MPC.py:262-271."""
    vel = np.sqrt(v_x ** 2 + v_y ** 2) * 3.6
    gain = GAIN1 * vel + GAIN0
    return deg2rad(steer_cmd * gain * MAX_STEER_DEG)


def f_dynamic(x, u, Ts):
    """control/MPC.py:186-229: explicit-Euler dynamic single-track model."""
    x = np.asarray(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    X, Y, yaw, v_x, v_y, r = (x[..., i] for i in range(6))
    fx = Fx(u[..., 0], v_x)
    delta = steer_cmd_to_angle(u[..., 1], v_x, v_y)
    theta_f = np.arctan2(v_y + LF * r, v_x + 0.1)
    theta_r = np.arctan2(v_y - LR * r, v_x + 0.1)
    Fyf = CF * (delta - theta_f)
    Fyr = CR * (-theta_r)
    v_x_dot = ((fx - Fyf * np.sin(delta)) / M) + (v_y * r)
    v_y_dot = ((Fyf * np.cos(delta) + Fyr) / M) - (v_x * r)
    r_dot = ((Fyf * np.cos(delta) * LF) - (Fyr * LR)) / IZ
    out = np.empty(np.broadcast(X, fx).shape + (6,))
    out[..., 0] = X + (v_x * np.cos(yaw) - v_y * np.sin(yaw)) * Ts
    out[..., 1] = Y + (v_x * np.sin(yaw) + v_y * np.cos(yaw)) * Ts
    out[..., 2] = yaw + r * Ts
    out[..., 3] = v_x + v_x_dot * Ts
    out[..., 4] = v_y + v_y_dot * Ts
    out[..., 5] = r + r_dot * Ts
    return out


def f_kinematic(x, u, Ts):
    """control/MPC.py:231-260 (defined by the reference, never called by it)."""
    x = np.asarray(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    X, Y, yaw, v_x, v_y, r = (x[..., i] for i in range(6))
    fx = Fx(u[..., 0], v_x)
    delta = steer_cmd_to_angle(u[..., 1], v_x, v_y)
    out = np.empty(np.broadcast(X, fx).shape + (6,))
    out[..., 0] = X + (v_x * np.cos(yaw) - v_y * np.sin(yaw)) * Ts
    out[..., 1] = Y + (v_x * np.sin(yaw) + v_y * np.cos(yaw)) * Ts
    out[..., 2] = yaw + r * Ts
    out[..., 3] = v_x + (fx / M) * Ts
    out[..., 4] = r * LR
    out[..., 5] = (v_x / (LR + LF)) * np.tan(delta)
    return out


def blend_lambda(v_x, v_y):
    """models/BlendedBicycleModel.py:23-26."""
    vel = np.hypot(v_x, v_y)
    return np.clip((vel - VBLENDMIN) / (VBLENDMAX - VBLENDMIN), 0, 1)


def f_blended(x, u, Ts):
    """lambda-blend of the two NLP models (models/BlendedBicycleModel.py:46);
    extension - the reference NLP never uses it."""
    x = np.asarray(x, dtype=np.float64)
    lam = blend_lambda(x[..., 3], x[..., 4])[..., None]
    return lam * f_dynamic(x, u, Ts) + (1 - lam) * f_kinematic(x, u, Ts)


def f_model(model, x, u, Ts):
    return (f_dynamic, f_kinematic, f_blended)[model](x, u, Ts)


# --------------------------------------------------------------------------
# Derivatives.  Variables are ordered z = (X, Y, yaw, v_x, v_y, r, thr, steer).
# --------------------------------------------------------------------------
IVX, IVY, IR, ITH, IST = 3, 4, 5, 6, 7


def _common(x, u):
    """Shared sub-expressions with gradients/Hessians over z (8 vars)."""
    x = np.asarray(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    shp = x.shape[:-1]
    v_x, v_y, r = x[..., 3], x[..., 4], x[..., 5]
    thr, st = u[..., 0], u[..., 1]

    def zeros_g():
        return np.zeros(shp + (8,))

    def zeros_h():
        return np.zeros(shp + (8, 8))

    # Fx
    kr = 60.0 * R_GEAR * 4.5 / C_WHEEL
    Gf = T_MAX * R_GEAR / R_WHEEL
    kd = 0.5 * RHO * C_D * A_F
    eta = ETA1 * kr * v_x + ETA0
    fx = thr * eta * Gf - kd * v_x ** 2 - C_ROLL * M * G
    gfx = zeros_g()
    gfx[..., ITH] = eta * Gf
    gfx[..., IVX] = thr * Gf * ETA1 * kr - 2 * kd * v_x
    hfx = zeros_h()
    hfx[..., IVX, IVX] = -2 * kd
    hfx[..., IVX, ITH] = hfx[..., ITH, IVX] = Gf * ETA1 * kr

    # delta
    cd = MAX_STEER_DEG * 2 * 3.14 / 360.0
    g1 = 3.6 * GAIN1
    V = np.sqrt(v_x ** 2 + v_y ** 2)
    gain = GAIN0 + g1 * V
    delta = cd * st * gain
    gd = zeros_g()
    gd[..., IST] = cd * gain
    gd[..., IVX] = cd * st * g1 * v_x / V
    gd[..., IVY] = cd * st * g1 * v_y / V
    hd = zeros_h()
    hd[..., IST, IVX] = hd[..., IVX, IST] = cd * g1 * v_x / V
    hd[..., IST, IVY] = hd[..., IVY, IST] = cd * g1 * v_y / V
    V3 = V ** 3
    hd[..., IVX, IVX] = cd * st * g1 * v_y ** 2 / V3
    hd[..., IVY, IVY] = cd * st * g1 * v_x ** 2 / V3
    hd[..., IVX, IVY] = hd[..., IVY, IVX] = -cd * st * g1 * v_x * v_y / V3

    # slip angles atan2(a, b), b = v_x + 0.1
    def atan2_d(a, ga, b, gb):
        D = a * a + b * b
        th = np.arctan2(a, b)
        da, db = b / D, -a / D
        g = da[..., None] * ga + db[..., None] * gb
        haa = -2 * a * b / D ** 2
        hbb = 2 * a * b / D ** 2
        hab = (a * a - b * b) / D ** 2
        h = (haa[..., None, None] * ga[..., :, None] * ga[..., None, :]
             + hbb[..., None, None] * gb[..., :, None] * gb[..., None, :]
             + hab[..., None, None] * (ga[..., :, None] * gb[..., None, :]
                                       + gb[..., :, None] * ga[..., None, :]))
        return th, g, h

    gb = zeros_g()
    gb[..., IVX] = 1.0
    gaf = zeros_g()
    gaf[..., IVY] = 1.0
    gaf[..., IR] = LF
    gar = zeros_g()
    gar[..., IVY] = 1.0
    gar[..., IR] = -LR
    b = v_x + 0.1
    thf, gthf, hthf = atan2_d(v_y + LF * r, gaf, b, gb)
    thr_, gthr, hthr = atan2_d(v_y - LR * r, gar, b, gb)
    return dict(fx=fx, gfx=gfx, hfx=hfx, delta=delta, gd=gd, hd=hd,
                thf=thf, gthf=gthf, hthf=hthf, thr=thr_, gthr=gthr, hthr=hthr)


def _outer(a, b):
    return a[..., :, None] * b[..., None, :]


def _sym(a, b):
    return _outer(a, b) + _outer(b, a)


def _position_rows(x, Ts, jac, lam, hess):
    yaw, v_x, v_y = x[..., 2], x[..., 3], x[..., 4]
    c, s = np.cos(yaw), np.sin(yaw)
    if jac is not None:
        jac[..., 0, 0] = 1.0
        jac[..., 0, 2] = (-v_x * s - v_y * c) * Ts
        jac[..., 0, 3] = c * Ts
        jac[..., 0, 4] = -s * Ts
        jac[..., 1, 1] = 1.0
        jac[..., 1, 2] = (v_x * c - v_y * s) * Ts
        jac[..., 1, 3] = s * Ts
        jac[..., 1, 4] = c * Ts
        jac[..., 2, 2] = 1.0
        jac[..., 2, 5] = Ts
    if hess is not None:
        lX, lY = lam[..., 0], lam[..., 1]
        hess[..., 2, 2] += Ts * (lX * (-v_x * c + v_y * s) + lY * (-v_x * s - v_y * c))
        t = Ts * (-lX * s + lY * c)
        hess[..., 2, 3] += t
        hess[..., 3, 2] += t
        t = Ts * (-lX * c - lY * s)
        hess[..., 2, 4] += t
        hess[..., 4, 2] += t


def dynamic_derivs(x, u, Ts, lam=None):
    """Jacobian (...,6,8) of f_dynamic and, if ``lam`` (...,6) is given, the
    lam-weighted Hessian sum_k lam_k * d2 f_k / dz2 (...,8,8)."""
    x = np.asarray(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    shp = x.shape[:-1]
    c = _common(x, u)
    v_x, v_y, r = x[..., 3], x[..., 4], x[..., 5]
    delta, gd, hd = c['delta'], c['gd'], c['hd']
    sd, cdl = np.sin(delta), np.cos(delta)
    Fyf = CF * (delta - c['thf'])
    gFyf = CF * (gd - c['gthf'])
    Fyr = -CR * c['thr']
    gFyr = -CR * c['gthr']

    jac = np.zeros(shp + (6, 8))
    _position_rows(x, Ts, jac, None, None)
    # v_x row
    gA = gFyf * sd[..., None] + (Fyf * cdl)[..., None] * gd          # grad(Fyf sin d)
    gB = gFyf * cdl[..., None] - (Fyf * sd)[..., None] * gd          # grad(Fyf cos d)
    gvx = (c['gfx'] - gA) / M
    gvx[..., IVY] += r
    gvx[..., IR] += v_y
    gvy = (gB + gFyr) / M
    gvy[..., IVX] -= r
    gvy[..., IR] -= v_x
    gr = (gB * LF - gFyr * LR) / IZ
    jac[..., 3, :] = gvx * Ts
    jac[..., 3, 3] += 1.0
    jac[..., 4, :] = gvy * Ts
    jac[..., 4, 4] += 1.0
    jac[..., 5, :] = gr * Ts
    jac[..., 5, 5] += 1.0
    if lam is None:
        return jac, None

    lam = np.asarray(lam, dtype=np.float64)
    hess = np.zeros(shp + (8, 8))
    _position_rows(x, Ts, None, lam, hess)
    c1 = lam[..., 3] / M
    c2 = lam[..., 4] / M + lam[..., 5] * LF / IZ
    c3 = lam[..., 4] / M - lam[..., 5] * LR / IZ
    w = -c1 * sd + c2 * cdl
    wp = -c1 * cdl - c2 * sd
    e = lambda a: a[..., None, None]
    H = (e(CF * w + Fyf * wp) * hd - e(CF * w) * c['hthf']
         + e(wp) * _sym(gFyf, gd) - e(Fyf * w) * _outer(gd, gd)
         - e(c3 * CR) * c['hthr'] + e(c1) * c['hfx'])
    # bilinear terms lam_vx * v_y * r - lam_vy * v_x * r
    H[..., IVY, IR] += lam[..., 3]
    H[..., IR, IVY] += lam[..., 3]
    H[..., IVX, IR] -= lam[..., 4]
    H[..., IR, IVX] -= lam[..., 4]
    hess += Ts * H
    return jac, hess


def kinematic_derivs(x, u, Ts, lam=None):
    x = np.asarray(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    shp = x.shape[:-1]
    c = _common(x, u)
    v_x = x[..., 3]
    delta, gd, hd = c['delta'], c['gd'], c['hd']
    L = LR + LF
    td = np.tan(delta)
    sec2 = 1.0 + td * td
    jac = np.zeros(shp + (6, 8))
    _position_rows(x, Ts, jac, None, None)
    jac[..., 3, :] = c['gfx'] / M * Ts
    jac[..., 3, 3] += 1.0
    jac[..., 4, IR] = LR
    gr = (v_x / L * sec2)[..., None] * gd
    gr[..., IVX] += td / L
    jac[..., 5, :] = gr
    if lam is None:
        return jac, None
    lam = np.asarray(lam, dtype=np.float64)
    hess = np.zeros(shp + (8, 8))
    _position_rows(x, Ts, None, lam, hess)
    e = lambda a: a[..., None, None]
    hess += e(lam[..., 3] * Ts / M) * c['hfx']
    evx = np.zeros(shp + (8,))
    evx[..., IVX] = 1.0
    Hr = (e(sec2 / L) * _sym(evx, gd)
          + e(v_x / L * sec2) * hd
          + e(v_x / L * 2 * sec2 * td) * _outer(gd, gd))
    hess += e(lam[..., 5]) * Hr
    return jac, hess


def blended_derivs(x, u, Ts, lam=None):
    x = np.asarray(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    shp = x.shape[:-1]
    v_x, v_y = x[..., 3], x[..., 4]
    V = np.hypot(v_x, v_y)
    span = VBLENDMAX - VBLENDMIN
    raw = (V - VBLENDMIN) / span
    inside = (raw > 0) & (raw < 1)
    lm = np.clip(raw, 0, 1)
    gl = np.zeros(shp + (8,))
    gl[..., IVX] = np.where(inside, v_x / V / span, 0.0)
    gl[..., IVY] = np.where(inside, v_y / V / span, 0.0)
    hl = np.zeros(shp + (8, 8))
    V3 = V ** 3
    hl[..., IVX, IVX] = np.where(inside, v_y ** 2 / V3 / span, 0.0)
    hl[..., IVY, IVY] = np.where(inside, v_x ** 2 / V3 / span, 0.0)
    hl[..., IVX, IVY] = hl[..., IVY, IVX] = np.where(inside, -v_x * v_y / V3 / span, 0.0)
    fd = f_dynamic(x, u, Ts)
    fk = f_kinematic(x, u, Ts)
    jd, hdn = dynamic_derivs(x, u, Ts, lam)
    jk, hk = kinematic_derivs(x, u, Ts, lam)
    diff = fd - fk
    jac = lm[..., None, None] * jd + (1 - lm)[..., None, None] * jk \
        + diff[..., :, None] * gl[..., None, :]
    if lam is None:
        return jac, None
    lam = np.asarray(lam, dtype=np.float64)
    ldiff = np.einsum('...k,...k->...', lam, diff)
    gdiff = np.einsum('...k,...kj->...j', lam, jd - jk)
    hess = lm[..., None, None] * hdn + (1 - lm)[..., None, None] * hk \
        + _sym(gl, gdiff) + ldiff[..., None, None] * hl
    return jac, hess


def model_derivs(model, x, u, Ts, lam=None):
    return (dynamic_derivs, kinematic_derivs, blended_derivs)[model](x, u, Ts, lam)


# ---------------------------------------------------------------------------
# RK4 discretisation (extension; the reference's NLP uses the explicit-Euler maps above,
# control/MPC.py:186-260).  The vector field is the model's Euler increment per unit time,
# g(y, u) = (F(y, u; Ts) - y) / Ts: for the dynamic single-track model this is the ODE
# right-hand side itself, for the kinematic / blended maps (which assign v_y, r algebraically)
# it is the relaxation field those maps define.  With D_i = F(y_i, u) - y_i:
#     y_1 = x, y_2 = x + D_1/2, y_3 = x + D_2/2, y_4 = x + D_3,
#     x+  = x + (D_1 + 2 D_2 + 2 D_3 + D_4) / 6.
# ---------------------------------------------------------------------------
_RK4_W = (1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0)
_RK4_C = (0.0, 0.5, 0.5, 1.0)


def rk4_step(model, x, u, Ts):
    x = np.asarray(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    acc = np.zeros_like(x)
    d = np.zeros_like(x)
    for i in range(4):
        y = x + _RK4_C[i] * d
        d = f_model(model, y, u, Ts) - y
        acc = acc + _RK4_W[i] * d
    return x + acc


def rk4_derivs(model, x, u, Ts, lam=None):
    """Jacobian (6,8) of rk4_step over z = (x, u) and, if ``lam`` is given, the lam-weighted
    Hessian (8,8): forward sensitivities of the four stage points, adjoint stage weights
    mu_i, H = sum_i W_i' Hess(mu_i' D)(y_i, u) W_i."""
    x = np.asarray(x, dtype=np.float64)
    u = np.asarray(u, dtype=np.float64)
    I0 = np.hstack([np.eye(6), np.zeros((6, 2))])
    E = np.hstack([np.zeros((2, 6)), np.eye(2)])
    ys, Ds, Ws = [], [], []
    d = np.zeros(6)
    S = np.zeros((6, 8))
    J = I0.copy()
    for i in range(4):
        y = x + _RK4_C[i] * d
        Y = I0 + _RK4_C[i] * S
        jac, _ = model_derivs(model, y, u, Ts, None)
        D = jac - I0
        W = np.vstack([Y, E])
        d = f_model(model, y, u, Ts) - y
        S = D @ W
        J = J + _RK4_W[i] * S
        ys.append(y); Ds.append(D); Ws.append(W)
    if lam is None:
        return J, None
    lam = np.asarray(lam, dtype=np.float64)
    H = np.zeros((8, 8))
    mu_next = None
    for i in (3, 2, 1, 0):
        mu = _RK4_W[i] * lam
        if i < 3:
            mu = mu + _RK4_C[i + 1] * (Ds[i + 1][:, :6].T @ mu_next)
        _, Hi = model_derivs(model, ys[i], u, Ts, mu)
        H += Ws[i].T @ Hi @ Ws[i]
        mu_next = mu
    return J, H
