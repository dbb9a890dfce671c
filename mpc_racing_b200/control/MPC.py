"""Drop-in replacement for the reference's ``control.MPC`` (control/MPC.py:9-184).

Same constructor signature (the solve happens in the constructor), same
``solution() -> (sol, ret, dual)`` surface and the same ``ret`` layout:
``(States[6,N+1], U[2,N], S_hat[N+1], [e_hat_C_i], [e_hat_L_i])``.  The NLP is
solved on the GPU by libmpcr.so (one warp per problem); there is no CasADi, no
IPOPT and no CPU fallback.  ``sol is None`` and ``dual is None`` when the solver
reports a failure status, exactly the reference's error convention
(control/MPC.py:172-181) minus its ``print``/``breakpoint()``.
"""
import math
from typing import List, Type

import numpy as np

from ..models.State import State
from ..models.VehicleParameters import VehicleParameters
from ..solver import BatchedMPC, STATUS_NAMES
from .ControllerParameters import FixedControllerParameters, RuntimeControllerParameters

_SOLVERS = {}


def _solver_for(N, Ts, fixed, model, wrap, device, restoration):
    key = (N, Ts, model, wrap, device, restoration) + tuple(
        getattr(fixed, k) for k in ("lambda_s", "alpha_L", "min_steer", "max_steer", "min_throttle",
                                    "max_steer_delta", "min_steer_delta", "max_throttle_delta",
                                    "min_throttle_delta", "q_v_max", "v_max"))
    s = _SOLVERS.get(key)
    if s is None:
        s = BatchedMPC(N=N, Ts=Ts, model=model, fixed_params=fixed, wrap_constraint=wrap,
                       device=device, on_device=False, restoration=restoration)
        _SOLVERS[key] = s
    return s


class MPCSolution:
    """Stand-in for CasADi's OptiSol: the handful of accessors callers use."""

    def __init__(self, States, U, S_hat, obj, iters, status):
        self.States, self.U, self.S_hat = States, U, S_hat
        self.obj, self.iters, self.status = obj, iters, status

    def stats(self):
        return {"return_status": STATUS_NAMES[self.status], "success": self.status >= 0,
                "iter_count": self.iters}


class MPC:
    model = "dynamic"        # control/MPC.py:83 always uses f_vehicle (the dynamic model)
    wrap_constraint = True   # control/MPC.py:142-143 at i = 0
    restoration = True       # IPOPT always has its restoration phase; here: mpcr_config.restoration
    device = 0

    def __init__(self,
                 state0: Type[State],
                 s0: float,
                 centerline_x_poly_coeffs: List[float],
                 centerline_y_poly_coeffs: List[float],
                 max_error: float,
                 runtime_params: Type[RuntimeControllerParameters],
                 sol0=None,       # accepted and ignored, as in the reference
                 duals=None,      # accepted and ignored, as in the reference
                 last_controls=None,
                 Ts=None,
                 N=None) -> None:
        self.fixed_params = FixedControllerParameters()
        self.runtime_params = runtime_params
        self.state0 = state0
        N = self.fixed_params.N if N is None else int(N)
        Ts = FixedControllerParameters.Ts if Ts is None else Ts
        solver = _solver_for(N, float(Ts), self.fixed_params, self.model, self.wrap_constraint,
                             self.device, self.restoration)
        nan = float("nan")
        x0 = np.array([[state0.x, state0.y, state0.yaw, state0.v_x, state0.v_y, state0.yaw_dot]],
                      dtype=np.float64)
        u_prev = np.array([[nan if state0.throttle is None else state0.throttle,
                            nan if state0.steer is None else state0.steer]], dtype=np.float64)
        u_warm = None
        if last_controls is not None:
            lc = np.asarray(last_controls, dtype=np.float64)      # [N, 2] rows (throttle, steer)
            if lc.shape != (N, 2):
                raise ValueError(f"last_controls must hold N={N} (throttle, steer) pairs")
            u_warm = np.ascontiguousarray(lc.T)[None]
        elif state0.throttle is None or state0.steer is None:
            raise ValueError("state0.throttle/steer are None and no last_controls were given "
                             "(the reference cannot build an initial guess either)")
        # control/MPC.py:50: the throttle bound comes from the class attribute
        prm = np.array([[runtime_params.alpha_c, RuntimeControllerParameters.d_max,
                         runtime_params.q_v_y, runtime_params.n, runtime_params.beta_delta]],
                       dtype=np.float64)
        res = solver.solve(x0, u_prev, np.array([s0], dtype=np.float64),
                           np.asarray(centerline_x_poly_coeffs, dtype=np.float64)[None],
                           np.asarray(centerline_y_poly_coeffs, dtype=np.float64)[None],
                           max_err=None, params=prm, u_warm=u_warm)
        status = int(res.status[0])
        self.status = status
        self.iters = int(res.iters[0])
        self.ret = (res.X[0].copy(), res.U[0].copy(), res.S[0].copy(),
                    [float(v) for v in res.eC[0]], [float(v) for v in res.eL[0]])
        if status >= 0:
            m = 13 * N + 7 + int(state0.throttle is not None) + int(state0.steer is not None)
            self.dual = res.lam_g[0, :m].copy()
            self.sol = MPCSolution(self.ret[0], self.ret[1], self.ret[2], float(res.obj[0]),
                                   self.iters, status)
        else:
            self.sol = None
            self.dual = None

    def solution(self):
        return self.sol, self.ret, self.dual

    # ---- numeric versions of the public model helpers (control/MPC.py:186-283) ----------
    def f_vehicle(self, x_k, u_k, Ts):
        m, Iz, lf, lr = VehicleParameters.m, VehicleParameters.Iz, VehicleParameters.lf, VehicleParameters.lr
        Cf, Cr = VehicleParameters.Cf, VehicleParameters.Cr
        x, y, yaw, v_x, v_y, r = (float(x_k[i]) for i in range(6))
        Fx = self.Fx(float(u_k[0]), v_x)
        delta = self.steer_cmd_to_angle(float(u_k[1]), v_x, v_y)
        th_f = math.atan2(v_y + lf * r, v_x + 0.1)
        th_r = math.atan2(v_y - lr * r, v_x + 0.1)
        Fyf, Fyr = Cf * (delta - th_f), Cr * (-th_r)
        vxd = (Fx - Fyf * math.sin(delta)) / m + v_y * r
        vyd = (Fyf * math.cos(delta) + Fyr) / m - v_x * r
        rd = (Fyf * math.cos(delta) * lf - Fyr * lr) / Iz
        return np.array([x + (v_x * math.cos(yaw) - v_y * math.sin(yaw)) * Ts,
                         y + (v_x * math.sin(yaw) + v_y * math.cos(yaw)) * Ts,
                         yaw + r * Ts, v_x + vxd * Ts, v_y + vyd * Ts, r + rd * Ts])

    def f_vehicle_kinematic(self, x_k, u_k, Ts):
        m, lf, lr = VehicleParameters.m, VehicleParameters.lf, VehicleParameters.lr
        x, y, yaw, v_x, v_y, r = (float(x_k[i]) for i in range(6))
        Fx = self.Fx(float(u_k[0]), v_x)
        delta = self.steer_cmd_to_angle(float(u_k[1]), v_x, v_y)
        return np.array([x + (v_x * math.cos(yaw) - v_y * math.sin(yaw)) * Ts,
                         y + (v_x * math.sin(yaw) + v_y * math.cos(yaw)) * Ts,
                         yaw + r * Ts, v_x + (Fx / m) * Ts, r * lr,
                         (v_x / (lr + lf)) * math.tan(delta)])

    def steer_cmd_to_angle(self, steer_cmd, v_x, v_y):
        vel = math.sqrt(v_x ** 2 + v_y ** 2) * 3.6
        gain = -0.001971664699 * vel + 0.986547
        z = steer_cmd * gain * VehicleParameters.max_steer
        return (z / 360) * 2 * 3.14                      # control/util.py:10-11

    def Fx(self, throttle, v_x):
        P = VehicleParameters
        rpm = (v_x / P.C_wheel) * 60 * P.R * 4.5
        eta = -0.00004428225806 * rpm + 1.282413306
        return (throttle * eta * P.T_max * P.R / P.r_wheel
                - 0.5 * P.rho * P.C_d * P.A_f * (v_x ** 2) - P.C_roll * P.m * P.g)
