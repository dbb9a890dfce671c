"""Mirror of splines/TrackSegments.py:8-35: equal-TIME checkpoints from a curvature-limited
speed profile.  Host precompute, once per track: the speed profile is tabulated with one batched
device evaluation on a fine grid and integrated with Simpson's rule (the reference calls
QUADPACK + MINPACK point by point, 9.5 s on Shanghai; this takes milliseconds and agrees to
~1e-5 relative - QUADPACK itself warns about round-off on this integrand).

``curve_radius`` reproduces the reference's expression ``1 / |(x' - y') y''|`` as written (it
is not the curvature formula).
"""
import numpy as np


class TrackSegments:
    def __init__(self, centerline, n_cp, v_max, d_f, d_r, m, grid=0.01):
        self.line = centerline
        self.n_cp = n_cp
        self.v_max = v_max
        self.curve_c = (2 * d_f + 2 * d_r) / m
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

    def curve_radius(self, s):
        e = self.line.eval_all(np.minimum(s, self.line.length * (1 - 1e-16)))
        return 1 / np.abs((e[:, 2] - e[:, 3]) * e[:, 5])

    def curve_velocity(self, s):
        s = np.atleast_1d(np.asarray(s, dtype=np.float64))
        return np.minimum(self.v_max, np.sqrt(self.curve_radius(s) * self.curve_c))

    def segment_time(self, s_0, s_1):
        return float(np.interp(s_1, self._s, self._T) - np.interp(s_0, self._s, self._T))

    def calculate_segment_bounds(self):
        bounds = [0]
        for i in range(self.n_cp):
            target = min((i + 1) * self.lap_time / self.n_cp, self._T[-1])
            bounds.append(float(np.interp(target, self._T, self._s)))
        return bounds

    def get_bound(self, i):
        assert (i >= 0)
        assert (i < self.n_cp)
        return self.bounds[i]
