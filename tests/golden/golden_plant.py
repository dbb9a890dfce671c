"""Golden vectors for the plant models (models/*.py), produced by the reference's own classes."""
import os

import numpy as np


def make_plant(outdir):
    from models.State import State
    from models.KinematicBicycleModel import KinematicBicycleModel
    from models.DynamicBicycleModel import DynamicBicycleModel
    from models.BlendedBicycleModel import BlendedBicycleModel
    rng = np.random.default_rng(1234)
    n = 96
    x0 = np.stack([rng.uniform(-100, 100, n), rng.uniform(-100, 100, n), rng.uniform(-3.2, 3.2, n),
                   rng.uniform(0.5, 45, n), rng.normal(0, 0.5, n), rng.normal(0, 0.2, n)], axis=1)
    x0[:16, 3] = rng.uniform(1.0, 16.0, 16)          # blending range
    u = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)], axis=1)
    u[::7, 0] = 0.0                                  # regen-brake branch (throttle == 0)
    u[1::9, 0] = 0.5                                 # carla_penalty peak
    dt = np.where(np.arange(n) % 2 == 0, 0.05, 0.1)
    out = {"x0": x0, "u": u, "dt": dt}
    for name, cls in (("kinematic", KinematicBicycleModel), ("dynamic", DynamicBicycleModel),
                      ("blended", BlendedBicycleModel)):
        res = []
        for i in range(n):
            m = cls(State(*x0[i]))
            st, _ = m.step(u[i, 0], u[i, 1], dt=dt[i])
            res.append([st.x, st.y, st.yaw, st.v_x, st.v_y, st.yaw_dot])
        out[name] = np.array(res, dtype=np.float64)
    # a 40-step rollout of the blended model (closed-loop plant of SURVEY 8(d) config 4)
    m = BlendedBicycleModel(State(0.0, 0.0, 0.3, 5.0, 0.0, 0.0))
    traj = []
    for k in range(40):
        st, _ = m.step(0.6 if k < 25 else -0.4, 0.2 * np.sin(0.3 * k), dt=0.05)
        traj.append([st.x, st.y, st.yaw, st.v_x, st.v_y, st.yaw_dot])
    out["blended_rollout"] = np.array(traj)
    np.savez_compressed(os.path.join(outdir, "plant_golden.npz"), **out)
