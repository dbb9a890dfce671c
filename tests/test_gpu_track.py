"""Track kernels on the GPU (mpcr_track_* through the product's splines classes) against the
golden vectors produced by the reference's own splines code."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRACKS = ["shanghai_intl_circuit", "t1_triple", "t2_triple", "t3", "t4"]


@pytest.fixture(scope="module")
def sg():
    return np.load(os.path.join(GOLD, "splines_golden.npz"))


@pytest.fixture(scope="module")
def lines():
    from mpc_racing_b200.splines import ParameterizedCenterline
    return {t: ParameterizedCenterline(t) for t in TRACKS}


@pytest.mark.parametrize("track", TRACKS)
def test_eval_and_geometry(sg, lines, track):
    cl = lines[track]
    L = sg[f"{track}/length"]
    assert cl.length == L[0] and cl.right_lane.length == L[1] and cl.left_lane.length == L[2]
    s = sg[f"{track}/s"]
    ev = cl.eval_all(s)
    ref = sg[f"{track}/eval"]
    np.testing.assert_allclose(ev[:, :2], ref[:, :2], rtol=0, atol=2e-10)
    np.testing.assert_allclose(ev[:, 2:4], ref[:, 2:4], rtol=0, atol=1e-10)
    np.testing.assert_allclose(ev[:, 4:], ref[:, 4:], rtol=0, atol=1e-9)
    geo = sg[f"{track}/geo"]
    np.testing.assert_allclose(cl.unit_tangent_yaw(s), geo[:, 0], rtol=0, atol=1e-9)
    np.testing.assert_allclose(cl.curvature(s), geo[:, 1], rtol=0, atol=1e-9)
    np.testing.assert_allclose(cl.unit_principal_normal(s), geo[:, 2:], rtol=0, atol=1e-9)
    # scalar call surface of the reference
    assert abs(cl.Gx(float(s[0])) - ref[0, 0]) < 2e-10 and isinstance(cl.Gx(float(s[0])), float)
    # lanes: thousands of pieces -> global-memory table path
    sl = sg[f"{track}/lane_s"]
    le = sg[f"{track}/lane_eval"]
    np.testing.assert_allclose(cl.right_lane.eval_all(sl)[:, :2], le[:, :2], rtol=0, atol=1e-9)
    np.testing.assert_allclose(cl.left_lane.eval_all(sl % cl.left_lane.length)[:, :2], le[:, 2:], rtol=0, atol=1e-9)


@pytest.mark.parametrize("track", TRACKS)
def test_coeffs_projection_lookup(sg, lines, track):
    cl = lines[track]
    sq = sg[f"{track}/sq"]
    gold = sg[f"{track}/coeffs"].reshape(len(sq), 2, 10)
    for li, la in enumerate((45.0, 75.0)):
        cx, cy = cl.coeffs(sq - 5, la)
        for j in range(len(sq)):
            tt = np.linspace(sq[j] - 5, sq[j] - 5 + la, 25)
            np.testing.assert_allclose(np.polyval(cx[j], tt), np.polyval(gold[j, li, :5], tt), rtol=0, atol=1e-6)
            np.testing.assert_allclose(np.polyval(cy[j], tt), np.polyval(gold[j, li, 5:], tt), rtol=0, atol=1e-6)
        np.testing.assert_array_equal(cl.lookup_error(sq, la), sg[f"{track}/lookup_error"][:, li])
    pq, pr = sg[f"{track}/proj_q"], sg[f"{track}/proj_r"]
    s, d, sgn = cl.project(pq[:, 0], pq[:, 1], pq[:, 2], pq[:, 3])
    np.testing.assert_allclose(s, pr[:, 0], rtol=0, atol=1e-6)
    np.testing.assert_allclose(d, pr[:, 1], rtol=0, atol=1e-9)
    np.testing.assert_array_equal(sgn, sg[f"{track}/error_sign"])
    s2, d2, _ = cl.project(pq[:, 0], pq[:, 1])              # global search
    assert (d2 <= pr[:, 1] + 1e-6).all()
    sp, dist = cl.projection(float(pq[0, 0]), float(pq[0, 1]), bounds=(float(pq[0, 2]), float(pq[0, 3])))
    assert abs(sp - pr[0, 0]) < 1e-6 and abs(dist - pr[0, 1]) < 1e-9
    assert cl.error_sign(float(pq[0, 0]), float(pq[0, 1]), float(s[0])) == int(sg[f"{track}/error_sign"][0])


def test_survey_known_values(lines):
    cl = lines["shanghai_intl_circuit"]
    assert cl.length == 1899.9319231385196
    assert abs(cl.lookup_error(69.6, 75) - 5.805978107108841) < 1e-12
    sp, dist = cl.projection_local(171, 91.8, (60, 80))     # a 20 m window -> global path here
    assert abs(sp - 70.07249095647715) < 1e-4 and abs(dist - 1.2688563634347347) < 1e-8
    ref = [1.2339473453246648e-06, -0.0005915402107889742, 0.08892755002816716, -4.445749591460687,
           219.69532953595657]
    tt = np.linspace(64.6, 64.6 + 75, 30)
    np.testing.assert_allclose(np.polyval(cl.x_as_coeffs(64.6, 75, 4), tt), np.polyval(ref, tt), rtol=0, atol=1e-6)


def test_workload_regenerated_on_device(lines):
    """The config-2 fixture (sampled with the reference's code) can be regenerated with the GPU
    track kernels: same max_err, same centerline polynomials."""
    from mpc_racing_b200.workload import load_fixture
    d = load_fixture()
    cl = lines["shanghai_intl_circuit"]
    idx = np.arange(0, 4096, 17)
    np.testing.assert_allclose(cl.lookup_error(d["s0"][idx], 75) - 0.925, d["max_err"][idx], rtol=0, atol=1e-12)
    cx, cy = cl.coeffs(d["s0"][idx] - 5, 75)
    for j, b in enumerate(idx):
        tt = np.linspace(d["s0"][b] - 5, d["s0"][b] + 70, 20)
        np.testing.assert_allclose(np.polyval(cx[j], tt), np.polyval(d["cx"][b], tt), rtol=0, atol=1e-6)
        np.testing.assert_allclose(np.polyval(cy[j], tt), np.polyval(d["cy"][b], tt), rtol=0, atol=1e-6)
