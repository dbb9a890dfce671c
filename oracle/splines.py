"""Oracle (test infrastructure): track geometry restated from the reference.

Follows, line by line:
  splines/util.py:3-10                      euclidean, midpoint
  splines/ParameterizedLine.py:19-64        Gx..ddGy (s % length, BSpline), x/y_as_coeffs
  splines/ParameterizedLine.py:66-105       projection / projection_local / projection_global
  splines/ParameterizedLine.py:107-149      unit_tangent(_yaw), curvature, mean_curvature,
                                            unit_principal_normal
  splines/ParameterizedLine.py:162-178      from_waypoints (cumulative chord length, k=3
                                            not-a-knot make_interp_spline)
  splines/ParameterizedCenterline.py:61-105 lookup_error, error_sign, from_file
  splines/ParameterizedLane.py:11-25        progress_bounds, projection, from_file
  splines/TrackSegments.py:8-35             curve_radius/velocity, segment_time, bounds
Like the reference it delegates to SciPy (make_interp_spline, bounded Brent
``minimize_scalar``, ``dual_annealing``, ``quad``, ``fsolve``) and numpy.polyfit -
third-party dependencies that exist on the GPU box as well.  Track data is read
from the converted .npz files under mpc_racing_b200/data/tracks (generated from
the reference's pickles/CSVs by tools/convert_tracks.py).

PINNED: tests/test_oracle_golden.py checks this module against golden vectors
produced by the reference's own classes (tests/golden/make_golden.py --what splines).
"""
import os
from math import sqrt

import numpy as np
from scipy.interpolate import make_interp_spline
from scipy.optimize import dual_annealing, minimize_scalar

TRACK_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mpc_racing_b200",
                         "data", "tracks")
TRACKS = ["shanghai_intl_circuit", "t1_triple", "t2_triple", "t3", "t4"]


def euclidean(p1, p2):
    return sqrt((p1[0] - p2[0]) ** 2 + (p1[1] - p2[1]) ** 2)


def midpoint(p1, p2, alpha=0.5):
    dx = (p2[0] - p1[0]) * alpha
    dy = (p2[1] - p1[1]) * alpha
    return dx + p1[0], dy + p1[1]


class Line:
    """splines/ParameterizedLine.py"""

    def __init__(self, waypoints):
        ss = [0]
        cum = 0
        for i in range(len(waypoints) - 1):
            cum += euclidean(waypoints[i], waypoints[i + 1])
            ss.append(cum)
        s = np.array(ss)
        x = np.array([p[0] for p in waypoints])
        y = np.array([p[1] for p in waypoints])
        self.spline_x = make_interp_spline(s, x)
        self.spline_y = make_interp_spline(s, y)
        self.length = ss[-1]
        self.s_data = s
        self.waypoints = list(zip(x, y))
        self.last_progress = None

    def Gx(self, s): return self.spline_x(s % self.length)
    def Gy(self, s): return self.spline_y(s % self.length)
    def dGx(self, s): return self.spline_x.derivative()(s % self.length)
    def dGy(self, s): return self.spline_y.derivative()(s % self.length)
    def ddGx(self, s): return self.spline_x.derivative().derivative()(s % self.length)
    def ddGy(self, s): return self.spline_y.derivative().derivative()(s % self.length)

    def x_as_coeffs(self, s, lookahead, deg=4):
        interps = np.linspace(0, lookahead, 50) + s
        return list(np.polyfit(interps, np.array([self.Gx(t) for t in interps]), deg=deg))

    def y_as_coeffs(self, s, lookahead, deg=4):
        interps = np.linspace(0, lookahead, 50) + s
        return list(np.polyfit(interps, np.array([self.Gy(t) for t in interps]), deg=deg))

    def projection(self, X, Y, bounds=None, seed=None):
        if bounds is None or 5 < abs(bounds[1] - bounds[0]):
            return self.projection_global(X, Y, seed=seed)
        return self.projection_local(X, Y, bounds=bounds)

    def projection_local(self, X, Y, bounds=None):
        if bounds is None:
            bounds = (0, self.length)
        dist = lambda s: sqrt((self.Gx(s) - X) ** 2 + (self.Gy(s) - Y) ** 2)
        ret = minimize_scalar(dist, bounds=bounds)
        return ret.x, dist(ret.x)

    def projection_global(self, X, Y, seed=None):
        """The reference's dual_annealing is unseeded; ``seed`` only makes the oracle repeatable."""
        dist = lambda s: sqrt((self.Gx(s) - X) ** 2 + (self.Gy(s) - Y) ** 2)
        ret = dual_annealing(lambda v: dist(v[0]), bounds=[(0, self.length)], seed=seed)
        return ret.x[0], dist(ret.x[0])

    def unit_tangent(self, s):
        dr = np.array([self.dGx(s), self.dGy(s)])
        return dr / np.linalg.norm(dr)

    def unit_tangent_yaw(self, s):
        ut = self.unit_tangent(s)
        return np.arctan2(ut[1], ut[0])

    def curvature(self, s):
        return np.abs(self.dGx(s) * self.ddGy(s) - self.dGy(s) * self.ddGx(s))

    def mean_curvature(self, s, lookahead, N=10):
        ss = np.linspace(s, s + lookahead, N)
        return (1 / N) * sum(self.curvature(t) for t in ss)

    def unit_principal_normal(self, s):
        ut_x, ut_y = self.unit_tangent(s)
        return ut_y, -ut_x

    # splines/ParameterizedLane.py:11-20
    def progress_bounds(self, step=0.5):
        if self.last_progress is None:
            return None
        return (self.last_progress - (2 * step), self.last_progress + (step * 6))

    def lane_projection(self, X, Y, bounds=None, seed=None):
        ret = self.projection(X, Y, bounds, seed=seed)
        self.last_progress = ret[0]
        return ret


def load_track_arrays(track):
    d = np.load(os.path.join(TRACK_DIR, track + ".npz"))
    return {k: d[k] for k in d.files}


class Centerline(Line):
    """splines/ParameterizedCenterline.py"""

    def __init__(self, track="shanghai_intl_circuit", lanes=True):
        d = load_track_arrays(track)
        waypoints = [tuple(p) for p in d["waypoints"]]           # track id already dropped
        if euclidean(waypoints[-1], waypoints[0]) > 0.1:         # :99-103 closing point
            waypoints.append(midpoint(waypoints[-1], waypoints[0], alpha=0.9))
        super().__init__(waypoints)
        self.track = track
        # NOTE the reference loads right_lane from *_left.csv and vice versa (:17-21)
        if lanes:
            self.right_lane = Line([tuple(p) for p in d["left"]])
            self.left_lane = Line([tuple(p) for p in d["right"]])
        self.err_ss = d["max_error"][:, 0]
        self.err_right = d["max_error"][:, 1]
        self.err_left = d["max_error"][:, 2]
        self._err_index = {float(k): i for i, k in enumerate(self.err_ss)}

    def lookup_error(self, s, lookahead):
        def round_to_half(x):
            return round(x * 2) / 2
        s_round = round_to_half(s)
        lookahead_round = round_to_half(lookahead)
        ss = np.arange(s_round, s + lookahead_round, 0.5)
        left_min = 10000
        right_min = 10000
        for t in ss:
            key = round_to_half(t % self.length)
            i = self._err_index[float(key)]
            left_min = min(left_min, self.err_left[i])
            right_min = min(right_min, self.err_right[i])
        return min(right_min, left_min)

    def error_sign(self, X, Y, s):
        upn = np.array(self.unit_principal_normal(s))
        disp = np.array([X - self.Gx(s), Y - self.Gy(s)])
        return 1 if np.linalg.norm(disp - upn) < np.linalg.norm(disp + upn) else -1


class Segments:
    """splines/TrackSegments.py:8-35 (GA checkpoints; host precompute)."""

    def __init__(self, centerline, n_cp, v_max, d_f, d_r, m):
        import scipy.integrate as integrate
        import scipy.optimize as optimize
        import math
        self.line = centerline
        self.n_cp = n_cp
        self.v_max = v_max
        self.curve_c = (2 * d_f + 2 * d_r) / m
        radius = lambda s: 1 / abs((self.line.dGx(s) - self.line.dGy(s)) * self.line.ddGy(s))   # sic
        self.curve_radius = radius
        self.curve_velocity = lambda s: min(self.v_max, math.sqrt(radius(s) * self.curve_c))
        self.segment_time = lambda a, b: integrate.quad(lambda x: 1 / self.curve_velocity(x), a, b,
                                                        limit=500)[0]
        self.lap_time = self.segment_time(0, centerline.length)
        bounds = [0]
        for _ in range(n_cp):
            b = optimize.fsolve(lambda x: self.segment_time(bounds[-1], x) - (self.lap_time / n_cp),
                                bounds[-1])[0]
            bounds.append(b)
        self.bounds = bounds
