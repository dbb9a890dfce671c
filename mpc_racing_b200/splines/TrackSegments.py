"""Mirror of splines/TrackSegments.py:8-35: equal-TIME checkpoints from a curvature-limited
speed profile.  Host precompute, once per track (like the spline tables themselves).

``method="quad"`` (default) is the reference's own numerical recipe - ``scipy.integrate.quad``
(limit 500) for the segment time and ``scipy.optimize.fsolve`` for every checkpoint - fed with the
speed profile evaluated from this package's spline table by the device kernels' segment rule, so
``lap_time`` and ``bounds`` reproduce the reference's numbers to the precision of the spline
table (1e-9; tests/test_gpu_closed_loop.py, tests/test_track_cpu.py).
``method="table"`` tabulates the speed profile with one batched DEVICE evaluation on a 1 cm grid
and integrates with Simpson's rule: milliseconds instead of seconds, and it converges to the exact
integral - which QUADPACK itself misses by ~1e-4 on this integrand (it has kinks wherever the
speed limit switches on) - so it differs from the reference's numbers at that level.

``curve_radius`` reproduces the reference's expression ``1 / |(x' - y') y''|`` as written (it
is not the curvature formula).
"""
import math
from bisect import bisect_right

import numpy as np


class TrackSegments:
    def __init__(self, centerline, n_cp, v_max, d_f, d_r, m, grid=0.01, method="quad"):
        self.line = centerline
        self.n_cp = n_cp
        self.v_max = v_max
        self.curve_c = (2 * d_f + 2 * d_r) / m
        self.method = method
        if method == "quad":
            t = centerline._tables
            self._brk = [float(v) for v in t["brk"]]
            self._cx = np.asarray(t["cx"], dtype=np.float64).tolist()
            self._cy = np.asarray(t["cy"], dtype=np.float64).tolist()
            self._L = float(centerline.length)
            self.lap_time = self.segment_time(0, centerline.length)
            self.bounds = self.calculate_segment_bounds()
            return
        L = centerline.length
        n = int(np.ceil(L / grid)) // 2 * 2 + 2            # even number of Simpson intervals
        self._s = np.linspace(0.0, L, n + 1)
        inv_v = 1.0 / self.curve_velocity(self._s)
        h = self._s[1] - self._s[0]
        # cumulative Simpson on pairs of intervals, linear in between
        pair = h / 3.0 * (inv_v[0:-2:2] + 4 * inv_v[1:-1:2] + inv_v[2::2])
        cum_even = np.concatenate([[0.0], np.cumsum(pair)])
        self._T = np.interp(self._s, self._s[::2], cum_even)
        self.lap_time = float(cum_even[-1])
        self.bounds = self.calculate_segment_bounds()

    # ---- the reference's recipe (host precompute over the spline table) --------------------------
    def _inv_velocity_scalar(self, s):
        """1 / curve_velocity(s) from the spline table: piece rule and Horner forms of
        csrc/mpcr_track.cuh:track_eval (``s % length``, right-continuous pieces)."""
        sm = s % self._L
        i = min(max(bisect_right(self._brk, sm) - 1, 0), len(self._cx) - 1)
        d = sm - self._brk[i]
        a, b = self._cx[i], self._cy[i]
        dgx = a[1] + d * (2.0 * a[2] + d * 3.0 * a[3])
        dgy = b[1] + d * (2.0 * b[2] + d * 3.0 * b[3])
        ddgy = 2.0 * b[2] + 6.0 * b[3] * d
        radius = 1 / abs((dgx - dgy) * ddgy)
        return 1 / min(self.v_max, math.sqrt(radius * self.curve_c))

    def curve_radius(self, s):
        e = self.line.eval_all(np.minimum(s, self.line.length * (1 - 1e-16)))
        return 1 / np.abs((e[:, 2] - e[:, 3]) * e[:, 5])

    def curve_velocity(self, s):
        s = np.atleast_1d(np.asarray(s, dtype=np.float64))
        return np.minimum(self.v_max, np.sqrt(self.curve_radius(s) * self.curve_c))

    def segment_time(self, s_0, s_1):
        if self.method == "quad":
            import scipy.integrate as integrate
            return integrate.quad(self._inv_velocity_scalar, s_0, s_1, limit=500)[0]
        return float(np.interp(s_1, self._s, self._T) - np.interp(s_0, self._s, self._T))

    def calculate_segment_bounds(self):
        bounds = [0]
        if self.method == "quad":
            import scipy.optimize as optimize
            for i in range(self.n_cp):
                seg_bound = optimize.fsolve(
                    lambda x: self.segment_time(bounds[-1], x) - (self.lap_time / self.n_cp), bounds[-1])[0]
                bounds.append(seg_bound)
            return bounds
        for i in range(self.n_cp):
            target = min((i + 1) * self.lap_time / self.n_cp, self._T[-1])
            bounds.append(float(np.interp(target, self._T, self._s)))
        return bounds

    def get_bound(self, i):
        assert (i >= 0)
        assert (i < self.n_cp)
        return self.bounds[i]
