"""Track geometry on the CPU: (1) the product's host-built PPoly tables against the oracle
(SciPy B-splines, pinned to the reference by test_oracle_golden.py), (2) the device functions of
csrc/mpcr_track.cuh compiled for the host (tests/emu) against the oracle and the golden vectors."""
import ctypes as C
import os

import numpy as np
import pytest

from emu_util import emu_lib
from mpc_racing_b200.splines import tables
from oracle import splines as osp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
V = C.c_void_p


def _tables_for(track):
    d = osp.load_track_arrays(track)
    wp = [tuple(p) for p in d["waypoints"]]
    if osp.euclidean(wp[-1], wp[0]) > 0.1:
        wp.append(osp.midpoint(wp[-1], wp[0], alpha=0.9))
    return tables.line_tables(wp), d


@pytest.fixture(scope="module")
def lines():
    return {t: osp.Centerline(t) for t in osp.TRACKS}


@pytest.fixture(scope="module")
def sg():
    return np.load(os.path.join(GOLD, "splines_golden.npz"))


@pytest.fixture(scope="module")
def lib():
    L = emu_lib()
    L.emu_track_eval.argtypes = [C.c_int, V, V, V, C.c_int, V, V]
    L.emu_track_coeffs.argtypes = [C.c_int, V, V, V, C.c_int, V, C.c_double, V, V]
    L.emu_track_project.argtypes = [C.c_int, V, V, V, C.c_int, V, V, V, V, V, V, V]
    L.emu_track_lookup_error.argtypes = [C.c_int, V, V, V, C.c_int, V, V, C.c_int, V, C.c_double, V]
    return L


def _p(a):
    assert a.flags["C_CONTIGUOUS"] and a.dtype in (np.float64, np.int32)
    return a.ctypes.data


@pytest.mark.parametrize("track", osp.TRACKS)
def test_tables_and_eval(lines, sg, lib, track):
    t, _ = _tables_for(track)
    cl = lines[track]
    assert t["length"] == cl.length                       # bit-identical chord length
    # break points == distinct knots of SciPy's not-a-knot spline (bit-exact segment indexing)
    np.testing.assert_array_equal(t["brk"], np.unique(cl.spline_x.t))
    s = np.ascontiguousarray(sg[f"{track}/s"])
    out = np.empty((len(s), 6))
    lib.emu_track_eval(len(t["brk"]) - 1, _p(t["brk"]), _p(t["cx"]), _p(t["cy"]), len(s), _p(s), _p(out))
    ref = sg[f"{track}/eval"]
    np.testing.assert_allclose(out[:, :2], ref[:, :2], rtol=0, atol=2e-10)
    np.testing.assert_allclose(out[:, 2:4], ref[:, 2:4], rtol=0, atol=1e-10)
    np.testing.assert_allclose(out[:, 4:], ref[:, 4:], rtol=0, atol=1e-9)


def test_lane_tables_large(lines, lib):
    """Lane-boundary splines have thousands of points: banded not-a-knot solve."""
    d = osp.load_track_arrays("t1_triple")
    t = tables.line_tables([tuple(p) for p in d["left"]])
    ln = lines["t1_triple"].right_lane                   # built from *_left.csv in the reference
    assert t["length"] == ln.length
    s = np.linspace(0, ln.length * 0.999, 400)
    out = np.empty((len(s), 6))
    lib.emu_track_eval(len(t["brk"]) - 1, _p(t["brk"]), _p(t["cx"]), _p(t["cy"]), len(s), _p(s), _p(out))
    np.testing.assert_allclose(out[:, 0], [ln.Gx(v) for v in s], rtol=0, atol=1e-9)
    np.testing.assert_allclose(out[:, 1], [ln.Gy(v) for v in s], rtol=0, atol=1e-9)


@pytest.mark.parametrize("track", osp.TRACKS)
def test_coeffs_projection_lookup(lines, sg, lib, track):
    t, d = _tables_for(track)
    n = len(t["brk"]) - 1
    sq = np.ascontiguousarray(sg[f"{track}/sq"])
    gold = sg[f"{track}/coeffs"].reshape(len(sq), 2, 10)
    for li, la in enumerate((45.0, 75.0)):
        s0 = np.ascontiguousarray(sq - 5)
        cx = np.empty((len(sq), 5))
        cy = np.empty((len(sq), 5))
        lib.emu_track_coeffs(n, _p(t["brk"]), _p(t["cx"]), _p(t["cy"]), len(sq), _p(s0), la, _p(cx), _p(cy))
        for j in range(len(sq)):
            # compare the POLYNOMIALS over the window: raw monomial coefficients in absolute
            # progress (s up to 1900) are ill-conditioned - np.polyfit's own result carries
            # ~1e-7 m of round-off there; tolerance 1e-6 m on values up to ~1e3
            tt = np.linspace(s0[j], s0[j] + la, 25)
            np.testing.assert_allclose(np.polyval(cx[j], tt), np.polyval(gold[j, li, :5], tt), rtol=0, atol=1e-6)
            np.testing.assert_allclose(np.polyval(cy[j], tt), np.polyval(gold[j, li, 5:], tt), rtol=0, atol=1e-6)
    # projection: same bounded-Brent sequence as SciPy's
    pq, pr = sg[f"{track}/proj_q"], sg[f"{track}/proj_r"]
    B = len(pq)
    so = np.empty(B)
    do = np.empty(B)
    sgn = np.empty(B, dtype=np.int32)
    cols = [np.ascontiguousarray(pq[:, k]) for k in range(4)]
    lib.emu_track_project(n, _p(t["brk"]), _p(t["cx"]), _p(t["cy"]), B, *[_p(c) for c in cols], _p(so), _p(do), _p(sgn))
    np.testing.assert_allclose(so, pr[:, 0], rtol=0, atol=1e-6)
    np.testing.assert_allclose(do, pr[:, 1], rtol=0, atol=1e-9)
    np.testing.assert_array_equal(sgn, sg[f"{track}/error_sign"])
    # global path (NaN bounds): at least as close as the local search around the truth
    nan = np.full(B, np.nan)
    lib.emu_track_project(n, _p(t["brk"]), _p(t["cx"]), _p(t["cy"]), B, _p(cols[0]), _p(cols[1]), _p(nan), _p(nan),
                          _p(so), _p(do), _p(sgn))
    assert (do <= pr[:, 1] + 1e-6).all()
    # lookup_error: bit-exact table minimum
    er = np.ascontiguousarray(d["max_error"][:, 1])
    el = np.ascontiguousarray(d["max_error"][:, 2])
    for li, la in enumerate((45.0, 75.0)):
        out = np.empty(len(sq))
        lib.emu_track_lookup_error(n, _p(t["brk"]), _p(t["cx"]), _p(t["cy"]), len(er), _p(er), _p(el), len(sq),
                                   _p(sq), la, _p(out))
        np.testing.assert_array_equal(out, sg[f"{track}/lookup_error"][:, li])


def test_plant_device_functions_match_reference(lib):
    """csrc/mpcr_plant.cuh (host build) vs golden vectors of the reference's models/*.py."""
    g = np.load(os.path.join(GOLD, "plant_golden.npz"))
    lib.emu_plant_step.argtypes = [C.c_int, C.c_int, V, V, V, V]
    x0 = np.ascontiguousarray(g["x0"])
    u = np.ascontiguousarray(g["u"])
    dt = np.ascontiguousarray(g["dt"])
    for name, mid in (("dynamic", 0), ("kinematic", 1), ("blended", 2)):
        out = np.empty_like(x0)
        lib.emu_plant_step(mid, len(x0), _p(x0), _p(u), _p(dt), _p(out))
        np.testing.assert_allclose(out, g[name], rtol=1e-13, atol=1e-12)


def test_track_segments_reproduce_the_reference_quad_and_fsolve_values():
    """splines/TrackSegments.py:8-35 with the reference's own recipe (scipy quad + fsolve) over this
    package's spline table: lap time and checkpoint bounds of the reference golden to 1e-9 (the
    Simpson-table variant converges to the exact integral instead, ~1e-4 away from QUADPACK's value)."""
    import warnings
    from mpc_racing_b200.splines import tables
    from mpc_racing_b200.splines.TrackSegments import TrackSegments
    from mpc_racing_b200.splines.util import euclidean, midpoint
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "splines_golden.npz"))

    class HostLine:          # the host part of ParameterizedCenterline (no device needed for the precompute)
        pass
    for t in ("t1_triple", "t4"):
        d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mpc_racing_b200", "data",
                                 "tracks", t + ".npz"))
        wp = [tuple(p) for p in d["waypoints"]]
        if euclidean(wp[-1], wp[0]) > 0.1:
            wp = wp + [midpoint(wp[-1], wp[0], alpha=0.9)]
        line = HostLine()
        line._tables = tables.line_tables(wp)
        line.length = line._tables["length"]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")          # QUADPACK's round-off warning, as in the reference
            seg = TrackSegments(line, 5, 30, 1500, 1500, 250)
        gold = g[f"{t}/segments"]
        assert abs(seg.lap_time - gold[0]) <= 1e-9 * gold[0]
        np.testing.assert_allclose(seg.bounds, gold[1:], rtol=1e-9, atol=1e-9)
