"""Synthetic problem batches for tests, smoke() and bench.

The start states of config 2 were sampled with the reference's own spline code
(tests/golden/make_golden.py, seed 1234) and are committed as fixtures inside the package
(mpc_racing_b200/data/fixtures: an installed package needs no tests/ tree); larger
batches are built from them by seeded perturbations of the vehicle state so that
every problem in a batch is distinct.
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "fixtures")


def load_fixture(name="cfg2_shanghai_4096.npz"):
    d = np.load(os.path.join(GOLDEN, name))
    return {k: np.ascontiguousarray(d[k], dtype=np.float64) for k in d.files}


def make_batch(B, seed=1234, fixture="cfg2_shanghai_4096.npz"):
    """B problems: the fixture's problems, cycled, copies perturbed (seeded)."""
    d = load_fixture(fixture)
    n = d["x0"].shape[0]
    idx = np.arange(B) % n
    out = {k: np.ascontiguousarray(v[idx]) for k, v in d.items() if v.shape[0] == n}
    rep = np.arange(B) // n
    rng = np.random.default_rng(seed)
    pert = rep > 0
    m = int(pert.sum())
    if m:
        x0 = out["x0"]
        yaw = x0[pert, 2]
        lat = rng.uniform(-0.3, 0.3, m)                       # lateral offset (m)
        x0[pert, 0] += -np.sin(yaw) * lat
        x0[pert, 1] += np.cos(yaw) * lat
        x0[pert, 2] += rng.normal(0, 0.02, m)
        x0[pert, 3] = np.clip(x0[pert, 3] + rng.uniform(-2, 2, m), 5.0, 35.0)
        x0[pert, 4] += rng.normal(0, 0.1, m)
        x0[pert, 5] += rng.normal(0, 0.02, m)
        out["u_prev"][pert, 0] = np.clip(out["u_prev"][pert, 0] + rng.uniform(-0.1, 0.1, m), 0, 0.85)
    return out
