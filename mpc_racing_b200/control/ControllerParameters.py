"""Controller parameter surface of the reference, kept field-for-field so that
callers can switch imports (reference: control/ControllerParameters.py:4-32).

Both are dataclasses whose attributes are also readable on the class itself -
the reference reads e.g. ``FixedControllerParameters.Ts`` and
``RuntimeControllerParameters.d_max`` without instantiating
(control/MPC.py:50,57,154).
"""
from dataclasses import dataclass, fields


@dataclass
class FixedControllerParameters:
    lambda_s: float = 300          # reward on progress at the end of the horizon
    alpha_L: float = 500           # lag-error penalty
    min_steer: float = -0.9
    max_steer: float = 0.9
    min_throttle: float = -1.0
    max_steer_delta: float = 0.2
    min_steer_delta: float = -0.2
    max_throttle_delta: float = 2.0
    min_throttle_delta: float = -0.4
    q_v_max: float = 2             # soft speed limit: exp(q_v_max (v_x - v_max))
    v_max: float = 50
    Ts: float = 0.05
    N: int = 30
    lookahead_distance: float = N * Ts * v_max
    max_iter: int = 500


@dataclass
class RuntimeControllerParameters:
    alpha_c: float = 1000          # contouring-error penalty
    d_max: float = 0.85            # throttle upper bound
    q_v_y: float = 50              # lateral-velocity penalty
    n: int = 2                     # exponent on the contouring error
    beta_delta: float = 5000       # steering-rate penalty

    def as_row(self):
        """(alpha_c, d_max, q_v_y, n, beta_delta) - the per-problem override row of
        ``mpcr_solve_batch(params=...)`` (include/mpcr.h)."""
        return [float(getattr(self, f.name)) for f in fields(self)]
