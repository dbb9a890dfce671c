"""Closed-loop rollouts and the GA fitness loop (config 5) on the GPU.

The reference never closes the loop outside CARLA; parity here means: same plant step + same
per-step preparation + same MPC solve, compared step by step against a CPU loop built from the
oracle pieces (oracle/splines.py, oracle/ipm.py, oracle/plant.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def t4():
    from mpc_racing_b200.splines import ParameterizedCenterline
    return ParameterizedCenterline("t4")


def test_plant_kernel_matches_reference_models():
    from mpc_racing_b200.solver import BatchedMPC
    import ctypes as C
    g = np.load(os.path.join(GOLD, "plant_golden.npz"))
    s = BatchedMPC()
    for name, mid in (("dynamic", 0), ("kinematic", 1), ("blended", 2)):
        for dt in (0.05, 0.1):
            sel = g["dt"] == dt
            x = np.ascontiguousarray(g["x0"][sel])
            u = np.ascontiguousarray(g["u"][sel])
            out = np.empty_like(x)
            rc = s.lib.mpcr_plant_step(s._h, mid, len(x), x.ctypes.data, u.ctypes.data, dt, out.ctypes.data, None)
            assert rc == 0
            np.testing.assert_allclose(out, g[name][sel], rtol=1e-12, atol=1e-11)


def test_closed_loop_steps_match_cpu_loop(t4):
    """32 vehicles x 50 control steps (graph-free recording mode) against the CPU pipeline built from
    the oracle pieces, one step ahead from every recorded state: projection + refit
    (oracle/splines.py = the reference's SciPy path), the solve (oracle/cxx, warm-started with the
    previous plan), the plant step (oracle/plant.py)."""
    from mpc_racing_b200.rollout import ClosedLoopRollouts, start_states
    from oracle import cxx, plant, splines as osp
    ro = ClosedLoopRollouts(t4, N=15, Ts=0.05, plant="blended")
    R, K = 32, 50
    x0, s0 = start_states(t4, R, seed=7, v0=8.0)
    res = ro.run(x0, s0, max_steps=K, record=True)
    hist = res["history"]
    assert len(hist) == K
    ocl = osp.Centerline("t4", lanes=False)
    prog_prev = s0.copy()
    u_prev = np.zeros((R, 2))
    checked = 0
    for k in range(K):
        x = hist[k]["x"]                                  # the state the MPC saw = plant state before the step
        prog = np.empty(R)
        cx, cy = np.empty((R, 5)), np.empty((R, 5))
        skip = np.zeros(R, dtype=bool)
        for r in range(R):
            lo, hi = (prog_prev[r] - 2) % ocl.length, (prog_prev[r] + 2) % ocl.length
            b = (lo, hi) if lo < hi else (hi, lo)
            if 5 < abs(b[1] - b[0]):
                skip[r] = True                            # at the wrap the reference calls dual_annealing
                prog[r] = hist[k]["progress"][r]
            else:
                prog[r], _ = ocl.projection_local(x[r, 0], x[r, 1], bounds=b)
            cx[r] = ocl.x_as_coeffs(prog[r] - 5, 45, 4)
            cy[r] = ocl.y_as_coeffs(prog[r] - 5, 45, 4)
        ok = ~skip
        np.testing.assert_allclose(hist[k]["progress"][ok], prog[ok], rtol=0, atol=1e-6)
        np.testing.assert_allclose(hist[k]["cx"][ok], cx[ok], rtol=1e-5, atol=1e-7)      # np.polyfit vs normal equations
        ref = cxx.solve_batch(x, u_prev, hist[k]["progress"], hist[k]["cx"], hist[k]["cy"], N=15,
                              u_warm=hist[k - 1]["U"] if k > 0 else None)
        assert (hist[k]["status"] == ref["status"]).all()
        conv = ref["status"] >= 0
        np.testing.assert_allclose(hist[k]["u0"][conv], ref["U"][conv, :, 0], rtol=0, atol=2e-4)
        for r in range(R):
            np.testing.assert_allclose(hist[k]["x_plant_next"][r], plant.blended_step(x[r], hist[k]["u0"][r], 0.05),
                                       rtol=1e-12, atol=1e-11)
        checked += int(conv.sum())
        prog_prev = hist[k]["progress"].copy()
        u_prev = hist[k]["u0"].copy()
    assert checked >= 0.99 * R * K
    # the graphed engine (no recording, no host round trips) ends in exactly the same states
    fast = ro.run(x0, s0, max_steps=K)
    np.testing.assert_array_equal(fast["final_state"], hist[-1]["x_plant_next"])
    assert fast["converged_solves"] == sum(int((h["status"] >= 0).sum()) for h in hist)
    # the recorded history writes out as a reference-format run (reference: Logger.py)
    import pandas as pd
    import tempfile
    from mpc_racing_b200.Logger import write_rollout, member_names
    with tempfile.TemporaryDirectory() as tmp:
        d = write_rollout(hist, vehicle=0, runs_fp=tmp, Ts=0.05, name="gpu_run")
        df = pd.read_csv(d + "/steps.csv")
        assert list(df.columns) == member_names and len(df) == K
        np.testing.assert_allclose(df["progress"], [h["progress"][0] for h in hist])
    ro.close()


def test_agent_state_estimate_and_warmup(t4):
    """estimate_state / start_control_at reproduce what the reference's agent feeds the MPC
    (agent.py:240-254: yaw rate by finite difference of the wrapped yaw, velocities world -> body)
    and its open-loop warm-up (agent.py:278-283: throttle 0.5, steer 0 before start_control_at)."""
    from mpc_racing_b200.rollout import ClosedLoopRollouts, start_states
    R, K, W = 16, 16, 5
    # the dynamic plant: the reference's BLENDED plant mixes the two models' wrapped yaw angles linearly
    # (models/BlendedBicycleModel.py:46), which yields a meaningless heading at the +-pi crossing
    ro = ClosedLoopRollouts(t4, N=15, Ts=0.05, plant="dynamic", estimate_state=True, start_control_at=W)
    # start the fleet 0.3 ... 5 m before the point where the track heading crosses +-pi, at 8 m/s:
    # every vehicle's (wrapped) plant yaw jumps by 2 pi within the K ticks
    sg = np.linspace(0.0, t4.length, 4000, endpoint=False)
    eg = t4.eval_all(sg)
    yg = np.arctan2(eg[:, 3], eg[:, 2])
    s_star = sg[np.where(np.abs(np.diff(yg)) > 3)[0][0]]
    s0 = (s_star - np.linspace(0.3, 5.0, R)) % t4.length
    e = t4.eval_all(s0)
    x0 = np.stack([e[:, 0], e[:, 1], np.arctan2(e[:, 3], e[:, 2]), np.full(R, 8.0), np.zeros(R), np.zeros(R)], axis=1)
    hist = ro.run(x0, s0, max_steps=K, record=True)["history"]
    xp = x0.copy()
    wraps = 0
    for k in range(K):
        seen = hist[k]["x"]
        yaw = xp[:, 2]
        wx = xp[:, 3] * np.cos(yaw) - xp[:, 4] * np.sin(yaw)
        wy = xp[:, 3] * np.sin(yaw) + xp[:, 4] * np.cos(yaw)
        vx = wx * np.cos(-yaw) - wy * np.sin(-yaw)
        vy = wx * np.sin(-yaw) + wy * np.cos(-yaw)
        np.testing.assert_allclose(seen[:, :3], xp[:, :3], rtol=0, atol=0)
        np.testing.assert_allclose(seen[:, 3], vx, rtol=1e-14, atol=1e-14)
        np.testing.assert_allclose(seen[:, 4], vy, rtol=1e-14, atol=1e-13)
        if k == 0:
            np.testing.assert_array_equal(seen[:, 5], xp[:, 5])       # no previous yaw yet
        else:
            d = yaw - yaw_old
            d = np.where(d > 3, yaw - (yaw_old + np.pi * 2), np.where(d < -3, yaw - (yaw_old - np.pi * 2), d))
            np.testing.assert_allclose(seen[:, 5], d / 0.05, rtol=1e-12, atol=1e-12)
            crossed = np.abs(yaw - yaw_old) > 3                        # the plant wraps its yaw into (-pi, pi]
            wraps += int(crossed.sum())
            # at a crossing the estimate is the small physical rate, not -+2 pi / Ts = 126 rad/s
            assert (np.abs(seen[crossed, 5] - xp[crossed, 5]) < 1.0).all(), (seen[crossed, 5], xp[crossed, 5])
        if k < W:
            np.testing.assert_array_equal(hist[k]["u0"], np.tile([0.5, 0.0], (R, 1)))
        else:
            ok = hist[k]["status"] >= 0                                    # the MPC has taken over
            assert ok.mean() >= 0.75, (k, hist[k]["status"])
            assert (np.abs(hist[k]["u0"][ok] - np.array([0.5, 0.0])).max(axis=1) > 0).any(), k
        yaw_old = yaw.copy()
        xp = hist[k]["x_plant_next"]
    assert wraps >= 1, "no vehicle crossed the heading wrap"
    ro.close()


def test_lap_rollouts_and_segments(t4):
    """A batch of full laps on t4: every vehicle finishes, stays on the track, segment times add
    up to the lap time."""
    from mpc_racing_b200.rollout import ClosedLoopRollouts, start_states
    from mpc_racing_b200.splines.TrackSegments import TrackSegments
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        seg = TrackSegments(t4, 5, 30, 1500, 1500, 250)
    gold = np.load(os.path.join(GOLD, "splines_golden.npz"))["t4/segments"]
    assert abs(seg.lap_time - gold[0]) <= 1e-9 * gold[0]            # the reference's quad + fsolve values
    np.testing.assert_allclose(seg.bounds, gold[1:], rtol=1e-9, atol=1e-9)
    fast = TrackSegments(t4, 5, 30, 1500, 1500, 250, method="table")     # device-evaluated Simpson table
    assert abs(fast.lap_time - gold[0]) <= 2e-4 * gold[0]
    np.testing.assert_allclose(fast.bounds, gold[1:], rtol=2e-4, atol=0.2)
    ro = ClosedLoopRollouts(t4, N=15, Ts=0.05, plant="blended")
    R = 64
    x0, s0 = start_states(t4, R, seed=11, v0=5.0)
    bounds = [b for b in seg.bounds]
    bounds[-1] = t4.length
    res = ro.run(x0, s0, max_steps=2500, seg_bounds=bounds)
    fin = ~np.isnan(res["lap_time"])
    assert fin.mean() >= 0.9, fin.mean()
    assert np.nanmedian(res["lap_time"]) < 125.0
    assert (res["max_abs_error"][fin] < 9.2).all()            # inside the lane half-width of t4
    st = res["seg_times"][fin]
    np.testing.assert_allclose(st.sum(axis=1), res["lap_time"][fin], rtol=0, atol=1e-9)
    ro.close()


def test_ga_generation(t4):
    from mpc_racing_b200.GA import mpcGA
    from mpc_racing_b200.rollout import ClosedLoopRollouts, start_states
    from mpc_racing_b200.splines.TrackSegments import TrackSegments
    seg = TrackSegments(t4, mpcGA.num_checkpoints, 30, 1500, 1500, 250)
    ro = ClosedLoopRollouts(t4, N=15, Ts=0.05, plant="blended")
    # a short course: the first checkpoint segment only, population 16
    bounds = [0.0] + [seg.bounds[1] * f for f in (0.2, 0.4, 0.6, 0.8, 1.0)]
    ga = mpcGA.GeneticAlgorithm(ro, bounds, seg.lap_time / mpcGA.num_checkpoints, pop=16, seed=3)
    x0, s0 = start_states(t4, 1, seed=5, v0=10.0)
    before = ga.population.copy()
    out = ga.run_step(x0[0], s0, max_steps=260)
    assert out["rewards"].shape == (16,) and np.isfinite(out["rewards"]).all() and out["best_reward"] > 0
    assert (ga.population != before).any()
    prm = mpcGA.decode(ga.population)
    assert ((prm[:, 3] == 2) | (prm[:, 3] == 4)).all() and (prm[:, 1] >= 0.5).all()
    ro.close()


def test_ga_generation_over_two_tracks():
    """BASELINE config 5 in small: one generation of the multi-track GA (population 8, tracks t3 + t4,
    a short course), fitness table gathered (single rank here; the 2-rank path is the gloo test)."""
    import warnings
    from mpc_racing_b200.GA import mpcGA
    from mpc_racing_b200.rollout import ClosedLoopRollouts, start_states
    from mpc_racing_b200.splines import ParameterizedCenterline
    from mpc_racing_b200.splines.TrackSegments import TrackSegments
    courses = []
    for track in ("t3", "t4"):
        cl = ParameterizedCenterline(track, lanes=False, error=False)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            seg = TrackSegments(cl, mpcGA.num_checkpoints, 30, 1500, 1500, 250)
        bounds = [0.0] + [seg.bounds[1] * f for f in (0.2, 0.4, 0.6, 0.8, 1.0)]
        x0, s0 = start_states(cl, 1, seed=5, v0=10.0)
        courses.append(dict(ro=ClosedLoopRollouts(cl, N=15, Ts=0.05, plant="blended", restoration=True),
                            seg_bounds=bounds, lap_time_guess=seg.lap_time / mpcGA.num_checkpoints,
                            x0=x0[0], s0=float(s0[0])))
    ga = mpcGA.MultiTrackGA(courses, pop=8, seed=3)
    before = ga.population.copy()
    out = ga.run_step(max_steps=260)
    assert out["rewards"].shape == (8,) and np.isfinite(out["rewards"]).all() and out["best_reward"] > 0
    assert out["rollouts"]["seg_times"].shape == (2, 8, 5)
    assert np.isfinite(out["rollouts"]["seg_times"]).mean() >= 0.8
    assert (ga.population != before).any()
    for c in courses:
        c["ro"].close()
