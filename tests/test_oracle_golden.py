"""Pin the oracle against golden vectors produced by the REFERENCE'S OWN classes
(tests/golden/make_golden.py imports /root/reference in the build container)."""
import os

import numpy as np
import pytest

from oracle import plant, splines

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def sg():
    return np.load(os.path.join(GOLD, "splines_golden.npz"))


@pytest.fixture(scope="module")
def lines():
    return {t: splines.Centerline(t) for t in splines.TRACKS}


@pytest.mark.parametrize("track", splines.TRACKS)
def test_spline_lengths_and_evaluation(sg, lines, track):
    cl = lines[track]
    L = sg[f"{track}/length"]
    assert cl.length == L[0] and cl.right_lane.length == L[1] and cl.left_lane.length == L[2]
    np.testing.assert_array_equal(cl.spline_x.t, sg[f"{track}/knots"])
    s = sg[f"{track}/s"]
    ev = np.array([[cl.Gx(v), cl.Gy(v), cl.dGx(v), cl.dGy(v), cl.ddGx(v), cl.ddGy(v)] for v in s])
    np.testing.assert_array_equal(ev, sg[f"{track}/eval"])        # same SciPy code path: bit-exact
    geo = np.array([[cl.unit_tangent_yaw(v), cl.curvature(v), *cl.unit_principal_normal(v)] for v in s])
    np.testing.assert_array_equal(geo, sg[f"{track}/geo"])
    sl = sg[f"{track}/lane_s"]
    lane = np.array([[cl.right_lane.Gx(v), cl.right_lane.Gy(v), cl.left_lane.Gx(v % cl.left_lane.length),
                      cl.left_lane.Gy(v % cl.left_lane.length)] for v in sl])
    np.testing.assert_array_equal(lane, sg[f"{track}/lane_eval"])


@pytest.mark.parametrize("track", splines.TRACKS)
def test_coeffs_lookup_projection(sg, lines, track):
    cl = lines[track]
    sq = sg[f"{track}/sq"]
    cxy = np.array([cl.x_as_coeffs(v - 5, la, 4) + cl.y_as_coeffs(v - 5, la, 4)
                    for v in sq for la in (45.0, 75.0)])
    np.testing.assert_array_equal(cxy, sg[f"{track}/coeffs"])
    lerr = np.array([[cl.lookup_error(v, 45), cl.lookup_error(v, 75)] for v in sq])
    np.testing.assert_array_equal(lerr, sg[f"{track}/lookup_error"])
    pq, pr = sg[f"{track}/proj_q"], sg[f"{track}/proj_r"]
    for q, r, sgn in zip(pq, pr, sg[f"{track}/error_sign"]):
        sp, dist = cl.projection_local(q[0], q[1], bounds=(q[2], q[3]))
        assert sp == r[0] and dist == r[1]
        assert cl.error_sign(q[0], q[1], sp) == sgn


def test_known_values_from_survey(lines):
    """SURVEY section 4 probes of the reference."""
    cl = lines["shanghai_intl_circuit"]
    assert cl.length == 1899.9319231385196
    np.testing.assert_allclose(cl.x_as_coeffs(64.6, 75, 4),
                               [1.2339473453246648e-06, -0.0005915402107889742, 0.08892755002816716,
                                -4.445749591460687, 219.69532953595657], rtol=1e-9)
    assert abs(cl.lookup_error(69.6, 75) - 5.805978107108841) < 1e-12
    sp, dist = cl.projection_local(171, 91.8, (60, 80))
    assert abs(sp - 70.07249095647715) < 1e-6 and abs(dist - 1.2688563634347347) < 1e-9
    assert [lines[t].length for t in splines.TRACKS[1:]] == [777.0884426619975, 1343.5442277778864,
                                                             1279.1813095110865, 924.6139125082169]


def test_track_segments(sg, lines):
    for t in ("t1_triple", "t4"):
        with pytest.warns(None) if False else np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                seg = splines.Segments(lines[t], 5, 30, 1500, 1500, 250)
        np.testing.assert_allclose([seg.lap_time] + list(seg.bounds), sg[f"{t}/segments"], rtol=1e-12)


def test_plant_models():
    g = np.load(os.path.join(GOLD, "plant_golden.npz"))
    for name in ("kinematic", "dynamic", "blended"):
        out = np.array([plant.PLANT_STEP[name](g["x0"][i], g["u"][i], g["dt"][i]) for i in range(len(g["x0"]))])
        np.testing.assert_allclose(out, g[name], rtol=1e-13, atol=1e-13)
    x = np.array([0.0, 0.0, 0.3, 5.0, 0.0, 0.0])
    for k in range(40):
        x = plant.blended_step(x, [0.6 if k < 25 else -0.4, 0.2 * np.sin(0.3 * k)], 0.05)
        np.testing.assert_allclose(x, g["blended_rollout"][k], rtol=1e-11, atol=1e-12)
