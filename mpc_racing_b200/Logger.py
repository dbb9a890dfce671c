"""Run logs in the reference's on-disk format (reference: Logger.py:1-54, agent.py:298-309) so that
script/replay_mpc.py, script/test_model_error.py and script/replay.py work unchanged on
GPU-generated closed-loop runs: ``<runs>/<name>/steps.csv`` with the reference's 18 columns and one
pickle per step under ``<runs>/<name>/mpc/<step>`` holding the MPC prediction.

The pickles reference ``models.State.State`` - the module path the reference's scripts import -
not this package's module path."""
import os
import pickle
import sys
import types
from contextlib import contextmanager
from datetime import datetime

import numpy as np

from .models.State import State

# reference: Logger.py:5-8
member_names = ['steps', 'X', 'Y', 'yaw', 'vx', 'vy', 'yawdot',
                'progress', 'error', 'cmd_throttle', 'cmd_steer', 'cmd_brake',
                'next_left_lane_point_x', 'next_left_lane_point_y', 'next_right_lane_point_x',
                'next_right_lane_point_y', 'last_ts', 'mpc_time']


@contextmanager
def _reference_state_class():
    """Yield the class to pickle State objects with, importable as ``models.State.State``."""
    ref = sys.modules.get('models.State')
    if ref is not None and hasattr(ref, 'State'):
        yield ref.State                   # running inside the reference tree: use its class
        return
    created = []
    if 'models' not in sys.modules:
        sys.modules['models'] = types.ModuleType('models')
        created.append('models')
    mod = types.ModuleType('models.State')
    mod.State = State
    sys.modules['models.State'] = mod
    created.append('models.State')
    old = State.__module__
    State.__module__ = 'models.State'
    try:
        yield State
    finally:
        State.__module__ = old
        for k in created:
            del sys.modules[k]


class Logger:
    def __init__(self, runs_fp, name=None) -> None:
        self.log_dir = os.path.join(runs_fp, name or datetime.now().strftime("%Y-%m-%d-%H-%M-%S"))
        self.fp = os.path.join(self.log_dir, 'steps.csv')
        self.mpc_fp = os.path.join(self.log_dir, 'mpc')
        os.makedirs(self.mpc_fp, exist_ok=False)
        self.write_header()

    def write_header(self):
        with open(self.fp, 'w') as f:
            f.write(','.join(member_names))
            f.write('\n')

    def log(self, agent):
        """``agent``: any object (or dict) carrying the 18 members (reference: Logger.py:26-38)."""
        agent_vars = agent if isinstance(agent, dict) else vars(agent)
        with open(self.fp, 'a') as f:
            f.write(",".join(str(agent_vars[prop]) for prop in member_names) + "\n")

    def pickle_mpc_res(self, agent):
        """reference: Logger.py:40-53 (same keys, same types)."""
        a = agent if isinstance(agent, dict) else vars(agent)
        with _reference_state_class() as cls:
            states = [s if type(s) is cls else
                      cls(x=s.x, y=s.y, yaw=s.yaw, v_x=s.v_x, v_y=s.v_y, yaw_dot=s.yaw_dot,
                          steer=s.steer, throttle=s.throttle) for s in a['predicted_states']]
            data = {
                'controlled': (a['steps'] >= a.get('start_control_at', 0)),
                'step': a['steps'],
                'predicted_states': states,
                'controls': a['last_controls'],
                'mean_ts': a['mean_ts'],
                'time': a['mpc_time'],
                's_hat': a['s_hat'],
                'e_hat_c': a['e_hat_c'],
                'e_hat_l': a['e_hat_l'],
            }
            with open(os.path.join(self.mpc_fp, str(a['steps'])), 'wb') as f:
                pickle.dump(data, f, protocol=pickle.HIGHEST_PROTOCOL)


def write_rollout(history, vehicle, runs_fp, Ts, name=None, mpc_time=-1):
    """Write one vehicle of a recorded ``ClosedLoop.run(record=True)`` history as a reference run.

    Per step the reference logs the state BEFORE the control is applied and the command chosen
    in that step (agent.py:246-309); lane-boundary points come from CARLA and are logged as nan.
    Returns the run directory."""
    lg = Logger(runs_fp, name)
    v = int(vehicle)
    for k, h in enumerate(history):
        x = h['x'][v]
        thr, st = float(h['u0'][v, 0]), float(h['u0'][v, 1])
        X, U = h['X'][v], h['U'][v]
        agent = dict(
            steps=k, X=float(x[0]), Y=float(x[1]), yaw=float(x[2]), vx=float(x[3]), vy=float(x[4]),
            yawdot=float(x[5]), progress=float(h['progress'][v]),
            error=float(h['dist'][v] * h['sign'][v]),
            cmd_throttle=max(thr, 0.0), cmd_steer=st, cmd_brake=max(-thr, 0.0),     # agent.py:289-296
            next_left_lane_point_x=float('nan'), next_left_lane_point_y=float('nan'),
            next_right_lane_point_x=float('nan'), next_right_lane_point_y=float('nan'),
            last_ts=Ts, mpc_time=mpc_time,
            predicted_states=[State(x=float(c[0]), y=float(c[1]), yaw=float(c[2]), v_x=float(c[3]),
                                    v_y=float(c[4]), yaw_dot=float(c[5])) for c in X.T],   # agent.py:190-200
            last_controls=[(float(a), float(b)) for a, b in zip(U[0], U[1])],             # agent.py:201-205
            mean_ts=Ts, s_hat=np.array(h['S'][v]), e_hat_c=[float(e) for e in h['eC'][v]],
            e_hat_l=[float(e) for e in h['eL'][v]], start_control_at=0)
        lg.pickle_mpc_res(agent)
        lg.log(agent)
    return lg.log_dir
