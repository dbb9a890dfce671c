"""Host precompute (once per track): the not-a-knot cubic interpolating spline of
``splines/ParameterizedLine.py:162-178`` (``make_interp_spline(s, x)``, k = 3) as a
piecewise power-basis table for the device kernels.

Break points are the data abscissae WITHOUT the 2nd and the 2nd-last one - exactly the
distinct knots of SciPy's not-a-knot B-spline - so the piece index chosen for a progress
value is the interval SciPy's evaluator would choose.
"""
from math import sqrt

import numpy as np
from scipy.linalg import solve_banded


def chord_lengths(waypoints):
    """Cumulative chord length (ParameterizedLine.py:163-169), summed sequentially like the
    reference so that ``length`` is bit-identical."""
    ss = [0]
    cum = 0
    for i in range(len(waypoints) - 1):
        p1, p2 = waypoints[i], waypoints[i + 1]
        cum += sqrt((p1[0] - p2[0]) ** 2 + (p1[1] - p2[1]) ** 2)
        ss.append(cum)
    return np.array(ss, dtype=np.float64)


def notaknot_second_derivs(s, y):
    n = len(s)
    h = np.diff(s)
    ab = np.zeros((5, n))            # banded storage, l = u = 2: ab[2 + i - j, j] = A[i, j]
    rhs = np.zeros(n)

    def put(i, j, v):
        ab[2 + i - j, j] = v
    put(0, 0, h[1]); put(0, 1, -(h[0] + h[1])); put(0, 2, h[0])
    for i in range(1, n - 1):
        put(i, i - 1, h[i - 1]); put(i, i, 2 * (h[i - 1] + h[i])); put(i, i + 1, h[i])
        rhs[i] = 6 * ((y[i + 1] - y[i]) / h[i] - (y[i] - y[i - 1]) / h[i - 1])
    put(n - 1, n - 3, h[n - 2]); put(n - 1, n - 2, -(h[n - 3] + h[n - 2])); put(n - 1, n - 1, h[n - 3])
    return solve_banded((2, 2), ab, rhs)


def ppoly_table(s, y):
    """Pieces on the reduced break set: returns (brk[m+1], coef[m,4] lowest order first)."""
    s = np.asarray(s, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n = len(s)
    assert n >= 5, "need at least 5 data points"
    M = notaknot_second_derivs(s, y)
    h = np.diff(s)
    a = y[:-1]
    b = (y[1:] - y[:-1]) / h - h * (2 * M[:-1] + M[1:]) / 6
    c = M[:-1] / 2
    d = (M[1:] - M[:-1]) / (6 * h)
    coef = np.stack([a, b, c, d], axis=1)               # piece i lives on [s_i, s_{i+1}]
    keep = np.ones(n, dtype=bool)
    keep[1] = False
    keep[n - 2] = False
    brk = s[keep]
    idx = np.nonzero(keep)[0][:-1]                      # data piece used from each kept break
    return brk, np.ascontiguousarray(coef[idx])


def line_tables(waypoints):
    s = chord_lengths(waypoints)
    x = np.array([p[0] for p in waypoints], dtype=np.float64)
    y = np.array([p[1] for p in waypoints], dtype=np.float64)
    brk, cx = ppoly_table(s, x)
    _, cy = ppoly_table(s, y)
    return dict(brk=np.ascontiguousarray(brk), cx=cx, cy=cy, length=float(s[-1]), s_data=s)
