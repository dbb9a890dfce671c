"""Generate the golden fixtures under tests/golden/ by IMPORTING THE REFERENCE'S OWN
PYTHON MODULES from /root/reference (read-only).  Run in the build container only:

    python tests/golden/make_golden.py [--what problems|splines|plant|all]

The GPU box has no /root/reference; tests only read the committed .npz files.
matplotlib is absent here, so an empty stub package (tests/golden/_stub) is put on
sys.path - splines/ParameterizedLine.py:3 imports pyplot at module top.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# problem-input fixtures live inside the package (bench.py and smoke() use them too)
FIX = os.path.join(HERE, "..", "..", "mpc_racing_b200", "data", "fixtures")
REF = "/root/reference"
SEED = 1234      # the harness's default seed (automatic_control_GRAIC.py:928-930)


def ref_imports():
    sys.path.insert(0, os.path.join(HERE, "_stub"))
    sys.path.insert(0, REF)
    os.chdir(REF)            # the reference opens "waypoints/..." and "lanes/..." cwd-relative


def sample_problems(cl, B, rng, lookback=5.0, lookahead=75.0, car_half_width=0.925):
    """start states sampled along a track, perturbation recipe of
    script/test_mpc.py:134-139; coefficients by the reference's x/y_as_coeffs."""
    out = dict(x0=[], u_prev=[], s0=[], cx=[], cy=[], max_err=[])
    L = cl.length
    for _ in range(B):
        s0 = rng.uniform(0, L)
        max_err = cl.lookup_error(s0, lookahead) - car_half_width
        e = rng.uniform(-max_err / 2, max_err / 2)
        upn = np.array(cl.unit_principal_normal(s0))
        X = cl.Gx(s0) + e * upn[0]
        Y = cl.Gy(s0) + e * upn[1]
        yaw = cl.unit_tangent_yaw(s0) + rng.normal(0, 0.05)
        vx = rng.uniform(5, 35)
        vy = rng.normal(0, 0.3)
        r = rng.normal(0, 0.05)
        thr = rng.uniform(0, 0.85)
        st = rng.uniform(-0.2, 0.2)
        out['x0'].append([X, Y, yaw, vx, vy, r])
        out['u_prev'].append([thr, st])
        out['s0'].append(s0)
        out['cx'].append(cl.x_as_coeffs(s0 - lookback, lookahead, 4))
        out['cy'].append(cl.y_as_coeffs(s0 - lookback, lookahead, 4))
        out['max_err'].append(max_err)
    return {k: np.array(v, dtype=np.float64) for k, v in out.items()}


def make_problems():
    from splines.ParameterizedCenterline import ParameterizedCenterline
    cl = ParameterizedCenterline("shanghai_intl_circuit")
    rng = np.random.default_rng(SEED)
    d = sample_problems(cl, 4096, rng)
    np.savez_compressed(os.path.join(FIX, "cfg2_shanghai_4096.npz"), **d)
    # config 1: script/test_mpc.py:18-51 verbatim
    s0 = 69.6
    cfg1 = dict(x0=np.array([[171, 91.8, -0.219, 20, 0.48, -0.059]], dtype=np.float64),
                u_prev=np.array([[0.19, 0.63]]), s0=np.array([s0]),
                cx=np.array([cl.x_as_coeffs(s0 - 5, 75, deg=4)]),
                cy=np.array([cl.y_as_coeffs(s0 - 5, 75, deg=4)]),
                max_err=np.array([cl.lookup_error(s0, 75) - 1.85 / 2]),
                Ts=np.array([0.1]), N=np.array([25]))
    np.savez_compressed(os.path.join(FIX, "cfg1_test_mpc.npz"), **cfg1)
    # config 3 inputs: t1_triple with lane bounds
    cl = ParameterizedCenterline("t1_triple")
    d = sample_problems(cl, 512, np.random.default_rng(SEED + 3))
    np.savez_compressed(os.path.join(FIX, "cfg3_t1_triple_512.npz"), **d)
    # production settings (agent.py:139-181): N = 15, lookahead 45, on t4
    cl = ParameterizedCenterline("t4")
    d = sample_problems(cl, 256, np.random.default_rng(SEED + 4), lookahead=45.0)
    np.savez_compressed(os.path.join(FIX, "prod_t4_256.npz"), **d)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="all")
    a = ap.parse_args()
    ref_imports()
    if a.what in ("problems", "all"):
        make_problems()
    if a.what in ("splines", "all"):
        from golden_splines import make_splines      # noqa
        make_splines(HERE)
    if a.what in ("plant", "all"):
        from golden_plant import make_plant          # noqa
        make_plant(HERE)
    print("golden fixtures written to", HERE)
