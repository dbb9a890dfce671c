"""Run logs in the reference's format (reference: Logger.py:1-54): steps.csv columns and the
per-step MPC pickles must load in an environment where `models.State` is the REFERENCE's class."""
import dataclasses
import os
import pickle
import sys
import types

import numpy as np
import pandas as pd

from mpc_racing_b200 import Logger as L


def _history(T=4, R=3, N=15, seed=0):
    rng = np.random.default_rng(seed)
    return [dict(x=rng.normal(size=(R, 6)), progress=rng.uniform(0, 100, R), u0=rng.uniform(-1, 0.85, (R, 2)),
                 U=rng.normal(size=(R, 2, N)), X=rng.normal(size=(R, 6, N + 1)), S=rng.normal(size=(R, N + 1)),
                 eC=rng.normal(size=(R, N)), eL=rng.normal(size=(R, N)), dist=rng.uniform(0, 1, R),
                 sign=rng.choice([-1, 1], R), status=np.zeros(R, np.int32)) for _ in range(T)]


def test_steps_csv_and_pickles_round_trip(tmp_path):
    hist = _history()
    d = L.write_rollout(hist, vehicle=1, runs_fp=str(tmp_path), Ts=0.05, name="run0")
    df = pd.read_csv(os.path.join(d, "steps.csv"))
    assert list(df.columns) == L.member_names          # reference: Logger.py:5-8
    assert len(df) == len(hist) and list(df["steps"]) == [0, 1, 2, 3]
    assert np.allclose(df["X"], [h["x"][1, 0] for h in hist])
    thr = np.array([h["u0"][1, 0] for h in hist])
    assert np.allclose(df["cmd_throttle"], np.maximum(thr, 0)) and np.allclose(df["cmd_brake"], np.maximum(-thr, 0))
    assert sorted(int(f) for f in os.listdir(os.path.join(d, "mpc"))) == [0, 1, 2, 3]

    # load the pickles the way script/replay_mpc.py does, with a stand-in for the reference's
    # models/State.py (same dataclass fields) registered as `models.State`
    @dataclasses.dataclass
    class State:
        x: float
        y: float
        yaw: float
        v_x: float
        v_y: float
        yaw_dot: float
        steer: float = None
        throttle: float = None
    State.__module__ = "models.State"
    State.__qualname__ = "State"
    pkg, mod = types.ModuleType("models"), types.ModuleType("models.State")
    mod.State = State
    sys.modules["models"], sys.modules["models.State"] = pkg, mod
    try:
        with open(os.path.join(d, "mpc", "2"), "rb") as f:
            data = pickle.load(f)
    finally:
        del sys.modules["models.State"], sys.modules["models"]
    assert set(data) == {"controlled", "step", "predicted_states", "controls", "mean_ts", "time",
                         "s_hat", "e_hat_c", "e_hat_l"}                        # reference: Logger.py:41-50
    assert data["step"] == 2 and data["controlled"] is True
    ps = data["predicted_states"]
    assert len(ps) == 16 and type(ps[0]) is State
    assert np.isclose(ps[3].v_x, hist[2]["X"][1, 3, 3])
    assert len(data["controls"]) == 15 and np.isclose(data["controls"][4][1], hist[2]["U"][1, 1, 4])
    assert len(data["e_hat_c"]) == 15 and data["s_hat"].shape == (16,)
