// Track geometry on the device: piecewise-cubic (PPoly) coefficient tables of the
// reference's not-a-knot interpolating splines and the per-step preparation that
// the reference does on the host around every MPC solve:
//   splines/ParameterizedLine.py:19-41   Gx, Gy, dGx, dGy, ddGx, ddGy   -> track_eval
//   splines/ParameterizedLine.py:43-64   x/y_as_coeffs (50-point deg-4 fit) -> track_fit4
//   splines/ParameterizedLine.py:66-105  projection (bounded Brent / global)  -> track_project
//   splines/ParameterizedCenterline.py:61-91 lookup_error, error_sign
// Plain host/device functions (the CPU tests call them through tests/emu); the
// kernels stage the table into shared memory with one TMA bulk copy when it fits.
#pragma once
#include <math.h>
#include "mpcr_types.h"

namespace mpcr {

struct TrackView {
  int n;                  // number of polynomial pieces
  double length;
  const double* brk;      // [n+1] break points (brk[0] = 0, brk[n] = length)
  const double* cx;       // [n][4] local power basis, lowest order first: sum c_k (s - brk_i)^k
  const double* cy;       // [n][4]
  int n_err;              // rows of the 0.5 m max-error table
  const double* err_right;
  const double* err_left;
};

// Python's float modulo (result takes the sign of the divisor)
MPCR_HD double py_mod(double a, double b) {
  double r = fmod(a, b);
  if (r != 0.0 && ((r < 0.0) != (b < 0.0))) r += b;
  return r;
}

// piece i with brk[i] <= s < brk[i+1] (right-continuous; last piece closed)
MPCR_HD int track_piece(const TrackView& t, double s) {
  int lo = 0, hi = t.n;          // invariant: brk[lo] <= s < brk[hi] (with brk[n] treated as +inf)
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (s >= t.brk[mid]) lo = mid; else hi = mid;
  }
  return lo;
}

// out = (Gx, Gy, dGx, dGy, ddGx, ddGy) at progress s (any real; wrapped by s % length)
MPCR_HD void track_eval(const TrackView& t, double s, int order, double* out) {
  const double sm = py_mod(s, t.length);
  const int i = track_piece(t, sm);
  const double d = sm - t.brk[i];
  const double* a = t.cx + 4 * i;
  const double* b = t.cy + 4 * i;
  out[0] = a[0] + d * (a[1] + d * (a[2] + d * a[3]));
  out[1] = b[0] + d * (b[1] + d * (b[2] + d * b[3]));
  if (order >= 1) {
    out[2] = a[1] + d * (2.0 * a[2] + d * 3.0 * a[3]);
    out[3] = b[1] + d * (2.0 * b[2] + d * 3.0 * b[3]);
  }
  if (order >= 2) {
    out[4] = 2.0 * a[2] + 6.0 * a[3] * d;
    out[5] = 2.0 * b[2] + 6.0 * b[3] * d;
  }
}

MPCR_HD double track_dist(const TrackView& t, double s, double X, double Y) {
  double g[2];
  track_eval(t, s, 0, g);
  return sqrt((g[0] - X) * (g[0] - X) + (g[1] - Y) * (g[1] - Y));
}

// SciPy's bounded Brent (scipy.optimize.minimize_scalar(bounds=...), xatol 1e-5, maxiter 500),
// the routine splines/ParameterizedLine.py:96 calls; same sequence of evaluations.
MPCR_HD void track_fminbound(const TrackView& t, double X, double Y, double x1, double x2,
                             double* xs, double* fs) {
  const double xatol = 1e-5;
  const int maxfun = 500;
  const double sqrt_eps = sqrt(2.2e-16);
  const double golden_mean = 0.5 * (3.0 - sqrt(5.0));
  double a = x1, b = x2;
  double fulc = a + golden_mean * (b - a);
  double nfc = fulc, xf = fulc;
  double rat = 0.0, e = 0.0;
  double x = xf;
  double fx = track_dist(t, x, X, Y);
  int num = 1;
  double fu;
  double ffulc = fx, fnfc = fx;
  double xm = 0.5 * (a + b);
  double tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
  double tol2 = 2.0 * tol1;
  while (fabs(xf - xm) > (tol2 - 0.5 * (b - a))) {
    bool golden = true;
    if (fabs(e) > tol1) {
      golden = false;
      double r = (xf - nfc) * (fx - ffulc);
      double q = (xf - fulc) * (fx - fnfc);
      double p = (xf - fulc) * q - (xf - nfc) * r;
      q = 2.0 * (q - r);
      if (q > 0.0) p = -p;
      q = fabs(q);
      r = e;
      e = rat;
      if ((fabs(p) < fabs(0.5 * q * r)) && (p > q * (a - xf)) && (p < q * (b - xf))) {
        rat = (p + 0.0) / q;
        x = xf + rat;
        if (((x - a) < tol2) || ((b - x) < tol2)) {
          double d = xm - xf;
          double si = (d > 0.0 ? 1.0 : (d < 0.0 ? -1.0 : 0.0)) + (d == 0.0 ? 1.0 : 0.0);
          rat = tol1 * si;
        }
      } else {
        golden = true;
      }
    }
    if (golden) {
      e = (xf >= xm) ? a - xf : b - xf;
      rat = golden_mean * e;
    }
    double si = (rat > 0.0 ? 1.0 : (rat < 0.0 ? -1.0 : 0.0)) + (rat == 0.0 ? 1.0 : 0.0);
    x = xf + si * fmax(fabs(rat), tol1);
    fu = track_dist(t, x, X, Y);
    num += 1;
    if (fu <= fx) {
      if (x >= xf) a = xf; else b = xf;
      fulc = nfc; ffulc = fnfc;
      nfc = xf; fnfc = fx;
      xf = x; fx = fu;
    } else {
      if (x < xf) a = x; else b = x;
      if ((fu <= fnfc) || (nfc == xf)) {
        fulc = nfc; ffulc = fnfc;
        nfc = x; fnfc = fu;
      } else if ((fu <= ffulc) || (fulc == xf) || (fulc == nfc)) {
        fulc = x; ffulc = fu;
      }
    }
    xm = 0.5 * (a + b);
    tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
    tol2 = 2.0 * tol1;
    if (num >= maxfun) break;
  }
  *xs = xf;
  *fs = fx;
}

// projection (splines/ParameterizedLine.py:66-105).  lo/hi NaN or a window wider than 5 m
// selects the global path; the reference's unseeded dual_annealing is replaced by a
// deterministic coarse scan (step <= 1 m) followed by the same bounded Brent on the best cell.
// Returns (s*, dist(s*)) like the reference (dist re-evaluated at s*).
MPCR_HD void track_project(const TrackView& t, double X, double Y, double lo, double hi,
                           double* s_out, double* d_out) {
  const bool local = (lo == lo) && (hi == hi) && !(5.0 < fabs(hi - lo));
  double xs, fs;
  if (local) {
    track_fminbound(t, X, Y, lo, hi, &xs, &fs);
  } else {
    const int cells = (int)ceil(t.length);
    const double h = t.length / cells;
    double best = 1.0e300;
    int bi = 0;
    for (int i = 0; i < cells; ++i) {
      const double d = track_dist(t, (i + 0.5) * h, X, Y);
      if (d < best) { best = d; bi = i; }
    }
    double a = (bi - 0.5) * h, b = (bi + 1.5) * h;
    if (a < 0.0) a = 0.0;
    if (b > t.length) b = t.length;
    track_fminbound(t, X, Y, a, b, &xs, &fs);
  }
  *s_out = xs;
  *d_out = track_dist(t, xs, X, Y);
}

// error_sign (splines/ParameterizedCenterline.py:82-91): +1 to the right of the centerline
MPCR_HD int track_error_sign(const TrackView& t, double X, double Y, double s) {
  double g[4];
  track_eval(t, s, 1, g);
  const double nrm = sqrt(g[2] * g[2] + g[3] * g[3]);
  const double nx = g[3] / nrm, ny = -(g[2] / nrm);       // unit_principal_normal = (t_y, -t_x)
  const double dx = X - g[0], dy = Y - g[1];
  const double a = sqrt((dx - nx) * (dx - nx) + (dy - ny) * (dy - ny));
  const double b = sqrt((dx + nx) * (dx + nx) + (dy + ny) * (dy + ny));
  return a < b ? 1 : -1;
}

// round_to_half of the reference: round(x * 2) / 2 with Python's round-half-to-even
MPCR_HD double round_to_half(double x) { return rint(x * 2.0) / 2.0; }

// lookup_error (splines/ParameterizedCenterline.py:61-80): min over the window of both lanes
MPCR_HD double track_lookup_error(const TrackView& t, double s, double lookahead) {
  const double s_round = round_to_half(s);
  const double la_round = round_to_half(lookahead);
  const double stop = s + la_round;
  // np.arange(start, stop, 0.5): ceil((stop - start) / step) values start + i * step
  int cnt = (int)ceil((stop - s_round) / 0.5);
  if (cnt < 0) cnt = 0;
  double left_min = 10000.0, right_min = 10000.0;
  for (int i = 0; i < cnt; ++i) {
    const double v = s_round + i * 0.5;
    const double key = round_to_half(py_mod(v, t.length));
    int idx = (int)(key * 2.0 + 0.5);
    if (idx >= t.n_err) idx = t.n_err - 1;      // the reference raises KeyError here (key == table end)
    if (idx < 0) idx = 0;
    left_min = fmin(left_min, t.err_left[idx]);
    right_min = fmin(right_min, t.err_right[idx]);
  }
  return fmin(right_min, left_min);
}

// x/y_as_coeffs (splines/ParameterizedLine.py:43-64): least-squares degree-4 polynomial through
// 50 samples G(s + linspace(0, lookahead, 50)), returned highest order first in ABSOLUTE
// progress like np.polyfit.  Solved in the centred/scaled variable tau = (t - mid) / half by
// Cholesky of the 5x5 normal equations (well conditioned there), then shifted back.
MPCR_HD void track_fit4(const TrackView& t, double s, double lookahead, double* cx, double* cy) {
  const int M = 50;
  const double mid = s + 0.5 * lookahead;
  const double half = (lookahead != 0.0) ? 0.5 * lookahead : 1.0;
  double A[5][5], bx[5], by[5];
  for (int i = 0; i < 5; ++i) {
    bx[i] = by[i] = 0.0;
    for (int j = 0; j < 5; ++j) A[i][j] = 0.0;
  }
  for (int k = 0; k < M; ++k) {
    // np.linspace(0, lookahead, 50)[k] + s
    const double step = lookahead / (M - 1);
    const double tk = (k == M - 1 ? lookahead : k * step) + s;
    double g[2];
    track_eval(t, tk, 0, g);
    const double tau = (tk - mid) / half;
    double pw[9];
    pw[0] = 1.0;
    for (int i = 1; i < 9; ++i) pw[i] = pw[i - 1] * tau;
    for (int i = 0; i < 5; ++i) {
      bx[i] += pw[i] * g[0];
      by[i] += pw[i] * g[1];
      for (int j = 0; j <= i; ++j) A[i][j] += pw[i + j];
    }
  }
  // Cholesky A = L L'
  double L[5][5];
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j <= i; ++j) {
      double v = A[i][j];
      for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k];
      L[i][j] = (i == j) ? sqrt(v) : v / L[j][j];
    }
  for (int pass = 0; pass < 2; ++pass) {
    double* b = pass == 0 ? bx : by;
    double y[5], c[5];
    for (int i = 0; i < 5; ++i) {
      double v = b[i];
      for (int k = 0; k < i; ++k) v -= L[i][k] * y[k];
      y[i] = v / L[i][i];
    }
    for (int i = 4; i >= 0; --i) {
      double v = y[i];
      for (int k = i + 1; k < 5; ++k) v -= L[k][i] * c[k];
      c[i] = v / L[i][i];
    }
    // p(t) = sum_i c_i ((t - mid)/half)^i  ->  monomials in t (binomial expansion)
    double mono[5] = {0, 0, 0, 0, 0};
    const double binom[5][5] = {{1, 0, 0, 0, 0}, {1, 1, 0, 0, 0}, {1, 2, 1, 0, 0}, {1, 3, 3, 1, 0},
                                {1, 4, 6, 4, 1}};
    double hpow = 1.0;
    for (int i = 0; i < 5; ++i) {
      const double ci = c[i] / hpow;        // c_i / half^i
      double mp = 1.0;                      // (-mid)^(i-j)
      for (int j = i; j >= 0; --j) {
        mono[j] += ci * binom[i][j] * mp;
        mp *= -mid;
      }
      hpow *= half;
    }
    double* dst = pass == 0 ? cx : cy;
    for (int i = 0; i < 5; ++i) dst[i] = mono[4 - i];     // highest order first
  }
}

}  // namespace mpcr
