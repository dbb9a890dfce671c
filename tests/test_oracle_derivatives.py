"""The oracle's hand-derived model derivatives vs torch.autograd (fp64)."""
import numpy as np
import pytest
import torch

from oracle import vehicle as V


def _torch_model(model, z, Ts):
    """Independent torch restatement of control/MPC.py:186-283 for autograd."""
    X, Y, yaw, v_x, v_y, r, thr, st = z
    rpm = (v_x / V.C_WHEEL) * 60 * V.R_GEAR * 4.5
    eta = V.ETA1 * rpm + V.ETA0
    fx = thr * eta * V.T_MAX * V.R_GEAR / V.R_WHEEL - 0.5 * V.RHO * V.C_D * V.A_F * v_x ** 2 \
        - V.C_ROLL * V.M * V.G
    vel = torch.sqrt(v_x ** 2 + v_y ** 2) * 3.6
    delta = (st * (V.GAIN1 * vel + V.GAIN0) * 70.0 / 360) * 2 * 3.14
    Xn = X + (v_x * torch.cos(yaw) - v_y * torch.sin(yaw)) * Ts
    Yn = Y + (v_x * torch.sin(yaw) + v_y * torch.cos(yaw)) * Ts
    yawn = yaw + r * Ts

    def dyn():
        thf = torch.atan2(v_y + V.LF * r, v_x + 0.1)
        thr_ = torch.atan2(v_y - V.LR * r, v_x + 0.1)
        Fyf = V.CF * (delta - thf)
        Fyr = V.CR * (-thr_)
        vxd = (fx - Fyf * torch.sin(delta)) / V.M + v_y * r
        vyd = (Fyf * torch.cos(delta) + Fyr) / V.M - v_x * r
        rd = (Fyf * torch.cos(delta) * V.LF - Fyr * V.LR) / V.IZ
        return torch.stack([Xn, Yn, yawn, v_x + vxd * Ts, v_y + vyd * Ts, r + rd * Ts])

    def kin():
        return torch.stack([Xn, Yn, yawn, v_x + fx / V.M * Ts, r * V.LR,
                            v_x / (V.LR + V.LF) * torch.tan(delta)])

    if model == V.MODEL_DYNAMIC:
        return dyn()
    if model == V.MODEL_KINEMATIC:
        return kin()
    lam = torch.clamp((torch.sqrt(v_x ** 2 + v_y ** 2) - V.VBLENDMIN) / (V.VBLENDMAX - V.VBLENDMIN), 0, 1)
    return lam * dyn() + (1 - lam) * kin()


@pytest.mark.parametrize("model", [V.MODEL_DYNAMIC, V.MODEL_KINEMATIC, V.MODEL_BLENDED])
def test_model_value_jac_hess(model):
    rng = np.random.default_rng(0)
    for trial in range(12):
        x = np.array([rng.uniform(-100, 100), rng.uniform(-100, 100), rng.uniform(-3, 3),
                      rng.uniform(1, 40) if trial % 3 else rng.uniform(3, 14),
                      rng.normal(0, 1.0), rng.normal(0, 0.3)])
        u = np.array([rng.uniform(-1, 0.85), rng.uniform(-0.9, 0.9)])
        lam = rng.normal(0, 1, 6)
        Ts = 0.05 if trial % 2 else 0.1
        z = torch.tensor(np.concatenate([x, u]), dtype=torch.float64)
        f_t = _torch_model(model, z, Ts).numpy()
        J_t = torch.autograd.functional.jacobian(lambda zz: _torch_model(model, zz, Ts), z).numpy()
        lam_t = torch.tensor(lam)
        H_t = torch.autograd.functional.hessian(
            lambda zz: (lam_t * _torch_model(model, zz, Ts)).sum(), z).numpy()
        f_o = V.f_model(model, x, u, Ts)
        J_o, H_o = V.model_derivs(model, x, u, Ts, lam)
        np.testing.assert_allclose(f_o, f_t, rtol=1e-13, atol=1e-12)
        np.testing.assert_allclose(J_o, J_t, rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(H_o, H_t, rtol=1e-9, atol=1e-10)


def test_vectorised_matches_scalar():
    rng = np.random.default_rng(1)
    x = rng.normal(0, 1, (5, 6)) + np.array([0, 0, 0, 15, 0, 0])
    u = rng.uniform(-0.5, 0.5, (5, 2))
    lam = rng.normal(0, 1, (5, 6))
    J, H = V.dynamic_derivs(x, u, 0.05, lam)
    for k in range(5):
        Jk, Hk = V.dynamic_derivs(x[k], u[k], 0.05, lam[k])
        np.testing.assert_array_equal(J[k], Jk)
        np.testing.assert_array_equal(H[k], Hk)


def _torch_rk4(model, z, Ts):
    x, u = z[:6], z[6:]
    acc = torch.zeros(6, dtype=torch.float64)
    d = torch.zeros(6, dtype=torch.float64)
    for c, w in ((0.0, 1 / 6), (0.5, 1 / 3), (0.5, 1 / 3), (1.0, 1 / 6)):
        y = x + c * d
        d = _torch_model(model, torch.cat([y, u]), Ts) - y
        acc = acc + w * d
    return x + acc


@pytest.mark.parametrize("model", [V.MODEL_DYNAMIC, V.MODEL_KINEMATIC, V.MODEL_BLENDED])
def test_rk4_value_jac_hess(model):
    """RK4 extension (north star item 2): value, Jacobian and lam-weighted Hessian of the RK4 map
    built from the Euler-map derivatives, against torch.autograd through the composed map."""
    rng = np.random.default_rng(3)
    for trial in range(8):
        x = np.array([rng.uniform(-100, 100), rng.uniform(-100, 100), rng.uniform(-3, 3),
                      rng.uniform(16, 40) if trial % 2 else rng.uniform(3, 13),
                      rng.normal(0, 1.0), rng.normal(0, 0.3)])
        u = np.array([rng.uniform(-1, 0.85), rng.uniform(-0.9, 0.9)])
        lam = rng.normal(0, 1, 6)
        Ts = 0.05 if trial % 2 else 0.1
        z = torch.tensor(np.concatenate([x, u]), dtype=torch.float64)
        f_t = _torch_rk4(model, z, Ts).numpy()
        J_t = torch.autograd.functional.jacobian(lambda zz: _torch_rk4(model, zz, Ts), z).numpy()
        lam_t = torch.tensor(lam)
        H_t = torch.autograd.functional.hessian(lambda zz: (lam_t * _torch_rk4(model, zz, Ts)).sum(), z).numpy()
        f_o = V.rk4_step(model, x, u, Ts)
        J_o, H_o = V.rk4_derivs(model, x, u, Ts, lam)
        np.testing.assert_allclose(f_o, f_t, rtol=1e-13, atol=1e-12)
        np.testing.assert_allclose(J_o, J_t, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(H_o, H_t, rtol=1e-8, atol=1e-9)
