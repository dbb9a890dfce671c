"""RK4 / Euler model step with analytic derivatives (mpc_racing_b200/csrc/mpcr_rk4.cuh): the device
source compiled for the host (tests/emu) and the CUDA kernel through the C ABI, against the oracle
restatement (oracle/vehicle.py, itself checked against torch.autograd in test_oracle_derivatives)."""
import ctypes as C

import numpy as np
import pytest

from oracle import vehicle as V


def _inputs(B=24, seed=5):
    rng = np.random.default_rng(seed)
    x = np.stack([np.array([rng.uniform(-100, 100), rng.uniform(-100, 100), rng.uniform(-3, 3),
                            rng.uniform(3, 40), rng.normal(0, 1), rng.normal(0, .3)]) for _ in range(B)])
    u = np.stack([np.array([rng.uniform(-1, .85), rng.uniform(-.9, .9)]) for _ in range(B)])
    return x, u, rng.normal(0, 1, (B, 6))


def _check(model, integ, Ts, x, u, lam, f, J, H):
    for b in range(x.shape[0]):
        if integ:
            fo = V.rk4_step(model, x[b], u[b], Ts)
            Jo, Ho = V.rk4_derivs(model, x[b], u[b], Ts, lam[b])
        else:
            fo = V.f_model(model, x[b], u[b], Ts)
            Jo, Ho = V.model_derivs(model, x[b], u[b], Ts, lam[b])
        np.testing.assert_allclose(f[b], fo, rtol=1e-13, atol=1e-12)
        np.testing.assert_allclose(J[b], Jo, rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(H[b], Ho, rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("model", [0, 1, 2])
@pytest.mark.parametrize("integ", [0, 1])
def test_device_source_on_host(model, integ):
    from emu_util import emu_lib
    lib = emu_lib()
    lib.emu_model_eval.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double] + [C.c_void_p] * 4
    x, u, lam = _inputs()
    B = x.shape[0]
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    f, J, H = np.zeros((B, 6)), np.zeros((B, 6, 8)), np.zeros((B, 8, 8))
    lib.emu_model_eval(model, integ, B, p(x), p(u), 0.05, p(lam), p(f), p(J), p(H))
    _check(model, integ, 0.05, x, u, lam, f, J, H)


def test_rk4_is_fourth_order_for_the_dynamic_model():
    """Halving the step of the dynamic single-track ODE cuts the RK4 error ~16x (Euler: ~2x)."""
    x = np.array([0.0, 0.0, 0.3, 18.0, 0.4, 0.1])
    u = np.array([0.4, 0.15])

    def roll(step, n, h):
        y = x.copy()
        for _ in range(n):
            y = step(V.MODEL_DYNAMIC, y, u, h)
        return y
    ref = roll(V.rk4_step, 1024, 0.2 / 1024)
    e1 = np.abs(roll(V.rk4_step, 4, 0.05) - ref).max()
    e2 = np.abs(roll(V.rk4_step, 8, 0.025) - ref).max()
    assert 10.0 < e1 / e2 < 24.0, e1 / e2
    ee1 = np.abs(roll(V.f_model, 4, 0.05) - ref).max()
    assert e1 < 1e-3 * ee1


@pytest.mark.gpu
@pytest.mark.parametrize("integ", ["euler", "rk4"])
def test_cuda_model_eval(integ):
    from mpc_racing_b200.solver import BatchedMPC
    s = BatchedMPC(N=30, Ts=0.05)
    x, u, lam = _inputs(B=200, seed=9)
    for model in (0, 1, 2):
        f, J, H = s.model_eval(x, u, lam, model=model, integrator=integ)
        _check(model, 1 if integ == "rk4" else 0, 0.05, x[:40], u[:40], lam[:40], f, J, H)
