// Oracle (TEST INFRASTRUCTURE, never the product path): compiled CPU restatement of the
// reference's MPC solve, independent of the CUDA kernels' source.
//
//   * the NLP is control/MPC.py:57-149 of AlexGisi/mpc-racing restated expression by expression
//     (variables :62-64, track-error functions :71-81 + control/util.py:4-8, cost :86-98,
//     constraints in subject_to order :101-149, initial guess :109-131); the vehicle model is
//     control/MPC.py:186-283 with the constants of models/VehicleParameters.py:8-41 (pi ~ 3.14);
//   * derivatives come from forward-mode second-order automatic differentiation of those
//     expressions (what CasADi does for the reference) - NOT from the hand-derived formulas of
//     oracle/vehicle.py or of the kernels, so it cross-checks both;
//   * the solver is a transcription of oracle/ipm.py (IPOPT's algorithm, Waechter & Biegler 2006,
//     IPOPT 3.14 default options + control/MPC.py:152-161): same text, same order of operations;
//   * the linear algebra is a symmetric-indefinite block-tridiagonal LDL^T with Bunch-Kaufman
//     pivoting inside the (<= 16 x 16) stage blocks and the two wrap-around rows
//     (control/MPC.py:142-143 at i = 0) as a 2-row border; the inertia is read from the pivots
//     (Haynsworth additivity), as IPOPT reads it from MUMPS.  A dense Bunch-Kaufman LDL^T of the
//     whole KKT matrix (linsol = 1, what oracle/ipm.py does with scipy) is kept as a cross-check.
//     Neither is the kernels' Riccati recursion.
//
// PARITY UNPINNED against IPOPT itself (CasADi is not installable offline, the reference stores
// no solver output); pinned against oracle/ipm.py (tests/test_oracle_cxx.py) and through it against
// the independent trust-constr optima (tests/golden/independent_optimum.npz).
//
// Build: oracle/cxx/Makefile  ->  oracle/cxx/liboracle_cxx.so  (C ABI, OpenMP over problems).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr double kInf = std::numeric_limits<double>::infinity();

// =====================================================================================
// Second-order forward-mode AD ("hyper-dual" numbers over NV independent variables)
// =====================================================================================
template <int NV>
struct HD {
  static constexpr int NH = NV * (NV + 1) / 2;
  double v;
  double g[NV];
  double h[NH];   // packed lower triangle: (i, j), i >= j -> i (i + 1) / 2 + j
  HD() : v(0.0) { std::fill(g, g + NV, 0.0); std::fill(h, h + NH, 0.0); }
  HD(double c) : v(c) { std::fill(g, g + NV, 0.0); std::fill(h, h + NH, 0.0); }   // constant
  static HD var(double val, int i) { HD r(val); r.g[i] = 1.0; return r; }
  double hess(int i, int j) const { return i >= j ? h[i * (i + 1) / 2 + j] : h[j * (j + 1) / 2 + i]; }
};

template <int NV>
HD<NV> chain(const HD<NV>& a, double f, double f1, double f2) {   // f(a) with f' = f1, f'' = f2
  HD<NV> r;
  r.v = f;
  for (int i = 0; i < NV; ++i) r.g[i] = f1 * a.g[i];
  int q = 0;
  for (int i = 0; i < NV; ++i)
    for (int j = 0; j <= i; ++j, ++q) r.h[q] = f1 * a.h[q] + f2 * a.g[i] * a.g[j];
  return r;
}
template <int NV> HD<NV> operator+(const HD<NV>& a, const HD<NV>& b) {
  HD<NV> r; r.v = a.v + b.v;
  for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] + b.g[i];
  for (int q = 0; q < HD<NV>::NH; ++q) r.h[q] = a.h[q] + b.h[q];
  return r;
}
template <int NV> HD<NV> operator-(const HD<NV>& a, const HD<NV>& b) {
  HD<NV> r; r.v = a.v - b.v;
  for (int i = 0; i < NV; ++i) r.g[i] = a.g[i] - b.g[i];
  for (int q = 0; q < HD<NV>::NH; ++q) r.h[q] = a.h[q] - b.h[q];
  return r;
}
template <int NV> HD<NV> operator-(const HD<NV>& a) { return chain(a, -a.v, -1.0, 0.0); }
template <int NV> HD<NV> operator*(const HD<NV>& a, const HD<NV>& b) {
  HD<NV> r; r.v = a.v * b.v;
  for (int i = 0; i < NV; ++i) r.g[i] = a.v * b.g[i] + b.v * a.g[i];
  int q = 0;
  for (int i = 0; i < NV; ++i)
    for (int j = 0; j <= i; ++j, ++q)
      r.h[q] = a.v * b.h[q] + b.v * a.h[q] + a.g[i] * b.g[j] + a.g[j] * b.g[i];
  return r;
}
template <int NV> HD<NV> inv(const HD<NV>& a) {
  const double r = 1.0 / a.v;
  return chain(a, r, -r * r, 2.0 * r * r * r);
}
template <int NV> HD<NV> operator/(const HD<NV>& a, const HD<NV>& b) { return a * inv(b); }
template <int NV> HD<NV> operator+(const HD<NV>& a, double c) { HD<NV> r = a; r.v += c; return r; }
template <int NV> HD<NV> operator+(double c, const HD<NV>& a) { return a + c; }
template <int NV> HD<NV> operator-(const HD<NV>& a, double c) { HD<NV> r = a; r.v -= c; return r; }
template <int NV> HD<NV> operator-(double c, const HD<NV>& a) { return (-a) + c; }
template <int NV> HD<NV> operator*(const HD<NV>& a, double c) { return chain(a, a.v * c, c, 0.0); }
template <int NV> HD<NV> operator*(double c, const HD<NV>& a) { return a * c; }
template <int NV> HD<NV> operator/(const HD<NV>& a, double c) { return chain(a, a.v / c, 1.0 / c, 0.0); }
template <int NV> HD<NV> sin(const HD<NV>& a) { return chain(a, std::sin(a.v), std::cos(a.v), -std::sin(a.v)); }
template <int NV> HD<NV> cos(const HD<NV>& a) { return chain(a, std::cos(a.v), -std::sin(a.v), -std::cos(a.v)); }
template <int NV> HD<NV> tan(const HD<NV>& a) {
  const double t = std::tan(a.v), s2 = 1.0 + t * t;
  return chain(a, t, s2, 2.0 * t * s2);
}
template <int NV> HD<NV> exp(const HD<NV>& a) { const double e = std::exp(a.v); return chain(a, e, e, e); }
template <int NV> HD<NV> sqrt(const HD<NV>& a) {
  const double s = std::sqrt(a.v);
  return chain(a, s, 0.5 / s, -0.25 / (a.v * s));
}
template <int NV> HD<NV> atan2(const HD<NV>& a, const HD<NV>& b) {
  const double D = a.v * a.v + b.v * b.v;
  const double da = b.v / D, db = -a.v / D;
  const double haa = -2.0 * a.v * b.v / (D * D), hbb = -haa, hab = (a.v * a.v - b.v * b.v) / (D * D);
  HD<NV> r; r.v = std::atan2(a.v, b.v);
  for (int i = 0; i < NV; ++i) r.g[i] = da * a.g[i] + db * b.g[i];
  int q = 0;
  for (int i = 0; i < NV; ++i)
    for (int j = 0; j <= i; ++j, ++q)
      r.h[q] = da * a.h[q] + db * b.h[q] + haa * a.g[i] * a.g[j] + hbb * b.g[i] * b.g[j] +
               hab * (a.g[i] * b.g[j] + a.g[j] * b.g[i]);
  return r;
}
template <int NV> HD<NV> powi(const HD<NV>& a, int e) {      // integer power (cost exponent n in {2, 4, ...})
  HD<NV> r(1.0);
  for (int k = 0; k < e; ++k) r = r * a;
  return r;
}
inline double sin(double a) { return std::sin(a); }
inline double cos(double a) { return std::cos(a); }
inline double tan(double a) { return std::tan(a); }
inline double exp(double a) { return std::exp(a); }
inline double sqrt(double a) { return std::sqrt(a); }
inline double atan2(double a, double b) { return std::atan2(a, b); }
inline double powi(double a, int e) { double r = 1.0; for (int k = 0; k < e; ++k) r *= a; return r; }
inline double value_of(double a) { return a; }
template <int NV> double value_of(const HD<NV>& a) { return a.v; }

// =====================================================================================
// The vehicle model inside the NLP: control/MPC.py:186-283, models/VehicleParameters.py:8-41
// =====================================================================================
namespace vp {
constexpr double m = 1845.0, max_steer = 70.0, T_max = 743.0, r_wheel = 0.37;
constexpr double C_wheel = 2 * 3.14 * r_wheel;     // VehicleParameters.py:16 (pi ~ 3.14 as written)
constexpr double R = 9.0, rho = 1.225, C_d = 0.23, A_f = 2.2, C_roll = 0.012, Iz = 3960.0;
constexpr double lf = 0.8, lr = 2.0, Cf = 65000.0, Cr = 65000.0, g = 9.81;
constexpr double Vblendmin = 2.0, Vblendmax = 15.0;
}  // namespace vp

template <class T> T deg2rad(const T& z) { return (z / 360.0) * 2.0 * 3.14; }      // control/util.py:10-11

template <class T> T Fx(const T& throttle, const T& v_x) {                           // control/MPC.py:273-283
  T wheel_rpm = (v_x / vp::C_wheel) * 60.0;
  T rpm = wheel_rpm * vp::R * 4.5;
  T eta = -0.00004428225806 * rpm + 1.282413306;
  T wheel_force = throttle * eta * vp::T_max * vp::R / vp::r_wheel;
  T drag_force = 0.5 * vp::rho * vp::C_d * vp::A_f * (v_x * v_x);
  const double rolling_resistance = vp::C_roll * vp::m * vp::g;
  return wheel_force - drag_force - rolling_resistance;
}

template <class T> T steer_cmd_to_angle(const T& steer_cmd, const T& v_x, const T& v_y) {   // :262-271
  T vel = sqrt(v_x * v_x + v_y * v_y) * 3.6;
  T gain = -0.001971664699 * vel + 0.986547;
  return deg2rad(steer_cmd * gain * vp::max_steer);
}

template <class T> void f_dynamic(const T* x, const T* u, double Ts, T* out) {       // control/MPC.py:186-229
  const T &X = x[0], &Y = x[1], &yaw = x[2], &v_x = x[3], &v_y = x[4], &r = x[5];
  T fx = Fx(u[0], v_x);
  T delta = steer_cmd_to_angle(u[1], v_x, v_y);
  T theta_f = atan2(v_y + vp::lf * r, v_x + 0.1);
  T theta_r = atan2(v_y - vp::lr * r, v_x + 0.1);
  T Fyf = vp::Cf * (delta - theta_f);
  T Fyr = vp::Cr * (-theta_r);
  T v_x_dot = ((fx - Fyf * sin(delta)) / vp::m) + (v_y * r);
  T v_y_dot = ((Fyf * cos(delta) + Fyr) / vp::m) - (v_x * r);
  T r_dot = ((Fyf * cos(delta) * vp::lf) - (Fyr * vp::lr)) / vp::Iz;
  out[0] = X + (v_x * cos(yaw) - v_y * sin(yaw)) * Ts;
  out[1] = Y + (v_x * sin(yaw) + v_y * cos(yaw)) * Ts;
  out[2] = yaw + r * Ts;
  out[3] = v_x + v_x_dot * Ts;
  out[4] = v_y + v_y_dot * Ts;
  out[5] = r + r_dot * Ts;
}

template <class T> void f_kinematic(const T* x, const T* u, double Ts, T* out) {     // control/MPC.py:231-260
  const T &X = x[0], &Y = x[1], &yaw = x[2], &v_x = x[3], &v_y = x[4], &r = x[5];
  T fx = Fx(u[0], v_x);
  T delta = steer_cmd_to_angle(u[1], v_x, v_y);
  out[0] = X + (v_x * cos(yaw) - v_y * sin(yaw)) * Ts;
  out[1] = Y + (v_x * sin(yaw) + v_y * cos(yaw)) * Ts;
  out[2] = yaw + r * Ts;
  out[3] = v_x + (fx / vp::m) * Ts;
  out[4] = r * vp::lr;
  out[5] = (v_x / (vp::lr + vp::lf)) * tan(delta);
}

template <class T> void f_blended(const T* x, const T* u, double Ts, T* out) {       // models/BlendedBicycleModel.py:23-46
  T fd[6], fk[6];
  f_dynamic(x, u, Ts, fd);
  f_kinematic(x, u, Ts, fk);
  const double vel = std::sqrt(value_of(x[3]) * value_of(x[3]) + value_of(x[4]) * value_of(x[4]));
  const double raw = (vel - vp::Vblendmin) / (vp::Vblendmax - vp::Vblendmin);
  if (raw > 0.0 && raw < 1.0) {
    T lam = (sqrt(x[3] * x[3] + x[4] * x[4]) - vp::Vblendmin) / (vp::Vblendmax - vp::Vblendmin);
    for (int k = 0; k < 6; ++k) out[k] = lam * fd[k] + (1.0 - lam) * fk[k];
  } else {
    const double lam = raw <= 0.0 ? 0.0 : 1.0;     // clip: constant outside the blending band
    for (int k = 0; k < 6; ++k) out[k] = lam * fd[k] + (1.0 - lam) * fk[k];
  }
}

template <class T> void f_model(int model, const T* x, const T* u, double Ts, T* out) {
  if (model == 0) f_dynamic(x, u, Ts, out);
  else if (model == 1) f_kinematic(x, u, Ts, out);
  else f_blended(x, u, Ts, out);
}

// classical RK4 of the model's Euler increment per unit time (extension; oracle/vehicle.py:rk4_step)
template <class T> void f_step(int model, int integrator, const T* x, const T* u, double Ts, T* out) {
  if (integrator == 0) { f_model(model, x, u, Ts, out); return; }
  const double W[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0}, C[4] = {0.0, 0.5, 0.5, 1.0};
  T acc[6], d[6], y[6], f[6];
  for (int k = 0; k < 6; ++k) { acc[k] = T(0.0); d[k] = T(0.0); }
  for (int i = 0; i < 4; ++i) {
    for (int k = 0; k < 6; ++k) y[k] = x[k] + C[i] * d[k];
    f_model(model, y, u, Ts, f);
    for (int k = 0; k < 6; ++k) { d[k] = f[k] - y[k]; acc[k] = acc[k] + W[i] * d[k]; }
  }
  for (int k = 0; k < 6; ++k) out[k] = x[k] + acc[k];
}

// =====================================================================================
// Parameters: control/ControllerParameters.py:4-32, control/MPC.py:59
// =====================================================================================
struct Params {
  double lambda_s = 300.0, alpha_L = 500.0, min_steer = -0.9, max_steer = 0.9, min_throttle = -1.0;
  double max_steer_delta = 0.2, min_steer_delta = -0.2, max_throttle_delta = 2.0, min_throttle_delta = -0.4;
  double q_v_max = 2.0, v_max = 50.0, min_s_delta = 0.1;
  double alpha_c = 1000.0, d_max = 0.85, q_v_y = 50.0, beta_delta = 5000.0;
  int n_exp = 2;
};

// make_poly (control/util.py:4-8): sum_k c[deg - k] s^k, coefficients highest order first, monomials
template <class T> T make_poly(const double* c, const T& s) {
  T acc(0.0);
  T pw(1.0);
  for (int k = 0; k <= 4; ++k) {
    acc = acc + c[4 - k] * pw;
    pw = pw * s;
  }
  return acc;
}
template <class T> T make_dpoly(const double* c, const T& s) {     // ca.gradient of make_poly
  T acc(0.0);
  T pw(1.0);
  for (int k = 1; k <= 4; ++k) {
    acc = acc + (k * c[4 - k]) * pw;
    pw = pw * s;
  }
  return acc;
}
// e_hat_C, e_hat_L: control/MPC.py:79-80
template <class T> void track_errors(const double* cx, const double* cy, const T& X, const T& Y, const T& s,
                                     T& eC, T& eL) {
  T gx = make_poly(cx, s), gy = make_poly(cy, s), dgx = make_dpoly(cx, s), dgy = make_dpoly(cy, s);
  eC = dgy * (X - gx) - dgx * (Y - gy);
  eL = -dgx * (X - gx) - dgy * (Y - gy);
}
// stage cost of control/MPC.py:88-94 over (X, Y, s, v_x, v_y)
template <class T> T stage_cost(const Params& p, const double* cx, const double* cy, const T& X, const T& Y,
                                const T& s, const T& vx, const T& vy) {
  T eC, eL;
  track_errors(cx, cy, X, Y, s, eC, eL);
  return p.q_v_y * (vy * vy) + p.alpha_c * powi(eC, p.n_exp) + p.alpha_L * (eL * eL) +
         exp(p.q_v_max * (vx - p.v_max));
}

// =====================================================================================
// Dense symmetric-indefinite LDL^T with Bunch-Kaufman (partial) pivoting.  Full storage,
// lower triangle + L in place.  P A P' = L D L'.
// =====================================================================================
struct BK {
  int n = 0;
  std::vector<double> a;       // n x n row-major; after factor(): L (strict lower) and D (diag / 2x2 blocks)
  std::vector<int> perm;       // position -> original index
  std::vector<int> piv;        // 1 = 1x1 pivot, 2 = first row of a 2x2 pivot, 0 = second row of a 2x2 pivot
  int npos = 0, nneg = 0, nzero = 0;
  double& A(int i, int j) { return a[(size_t)i * n + j]; }
  double A(int i, int j) const { return a[(size_t)i * n + j]; }

  void swap_sym(int p, int q) {   // symmetric interchange of rows/columns p and q (full storage)
    if (p == q) return;
    for (int j = 0; j < n; ++j) std::swap(A(p, j), A(q, j));
    for (int i = 0; i < n; ++i) std::swap(A(i, p), A(i, q));
    std::swap(perm[p], perm[q]);
  }

  void factor(const double* src, int n_) {
    n = n_;
    a.assign(src, src + (size_t)n * n);
    perm.resize(n);
    piv.assign(n, 1);
    for (int i = 0; i < n; ++i) perm[i] = i;
    npos = nneg = nzero = 0;
    const double alpha = (1.0 + std::sqrt(17.0)) / 8.0;
    std::vector<double> col1(n), col2(n);
    int k = 0;
    while (k < n) {
      const double absakk = std::fabs(A(k, k));
      int r = -1;
      double colmax = 0.0;
      for (int i = k + 1; i < n; ++i)
        if (std::fabs(A(i, k)) > colmax) { colmax = std::fabs(A(i, k)); r = i; }
      int kstep = 1;
      if (std::max(absakk, colmax) == 0.0) {          // singular column: zero pivot
        nzero++;
        piv[k] = 1;
        k += 1;
        continue;
      }
      if (absakk < alpha * colmax) {
        double rowmax = 0.0;
        for (int j = k; j < n; ++j)
          if (j != r) rowmax = std::max(rowmax, std::fabs(A(r, j)));
        if (absakk >= alpha * colmax * (colmax / rowmax)) {
          // 1x1, no interchange
        } else if (std::fabs(A(r, r)) >= alpha * rowmax) {
          swap_sym(k, r);
        } else {
          kstep = 2;
          swap_sym(k + 1, r);
        }
      }
      if (kstep == 1) {
        const double d = A(k, k);
        if (d > 0) npos++; else if (d < 0) nneg++; else nzero++;
        for (int i = k + 1; i < n; ++i) col1[i] = A(i, k);
        for (int i = k + 1; i < n; ++i) {
          const double l = col1[i] / d;
          if (l != 0.0) {
            double* ai = &a[(size_t)i * n];
            for (int j = k + 1; j <= i; ++j) ai[j] -= l * col1[j];
          }
          A(i, k) = l;
        }
        piv[k] = 1;
      } else {
        const double d11 = A(k, k), d21 = A(k + 1, k), d22 = A(k + 1, k + 1);
        const double det = d11 * d22 - d21 * d21;
        if (det < 0) { npos++; nneg++; }
        else if (det > 0) { if (d11 > 0) npos += 2; else nneg += 2; }
        else nzero += 2;
        for (int i = k + 2; i < n; ++i) { col1[i] = A(i, k); col2[i] = A(i, k + 1); }
        for (int i = k + 2; i < n; ++i) {
          const double l1 = (col1[i] * d22 - col2[i] * d21) / det, l2 = (col2[i] * d11 - col1[i] * d21) / det;
          double* ai = &a[(size_t)i * n];
          for (int j = k + 2; j <= i; ++j) ai[j] -= l1 * col1[j] + l2 * col2[j];
          A(i, k) = l1;
          A(i, k + 1) = l2;
        }
        piv[k] = 2;
        piv[k + 1] = 0;
      }
      // keep the trailing block fully stored (row searches and interchanges read both triangles)
      for (int i = k + kstep; i < n; ++i)
        for (int j = k + kstep; j < i; ++j) A(j, i) = A(i, j);
      k += kstep;
    }
  }

  void solve(double* b) const {   // in place, b of length n (original ordering)
    std::vector<double> y(n);
    for (int i = 0; i < n; ++i) y[i] = b[perm[i]];
    for (int i = 0; i < n; ++i) {          // L y = b (unit lower; 2x2 blocks have no sub-entry between them)
      double acc = y[i];
      const double* ai = &a[(size_t)i * n];
      int jmax = i;
      if (piv[i] == 0) jmax = i - 1;       // second row of a 2x2 block: A(i, i-1) is D, not L
      for (int j = 0; j < jmax; ++j) acc -= ai[j] * y[j];
      y[i] = acc;
    }
    for (int i = 0; i < n;) {              // D z = y
      if (piv[i] == 2) {
        const double d11 = A(i, i), d21 = A(i + 1, i), d22 = A(i + 1, i + 1);
        const double det = d11 * d22 - d21 * d21;
        const double y1 = y[i], y2 = y[i + 1];
        y[i] = (y1 * d22 - y2 * d21) / det;
        y[i + 1] = (y2 * d11 - y1 * d21) / det;
        i += 2;
      } else {
        y[i] = (A(i, i) != 0.0) ? y[i] / A(i, i) : 0.0;
        i += 1;
      }
    }
    for (int i = n - 1; i >= 0; --i) {     // L' x = z
      double acc = y[i];
      for (int j = i + 1; j < n; ++j) {
        if (piv[j] == 0 && j - 1 == i) continue;     // A(j, i) is the off-diagonal of D
        acc -= A(j, i) * y[j];
      }
      y[i] = acc;
    }
    for (int i = 0; i < n; ++i) b[perm[i]] = y[i];
  }
};


// =====================================================================================
// The NLP of control/MPC.py as CasADi's Opti hands it to IPOPT:  min f(w)  s.t. gl <= g(w) <= gu.
// Flat vector in Opti's creation order, each matrix column-major (control/MPC.py:62-64):
//   w = [U (2 x N) | S_hat (N + 1) | States (6 x (N + 1))]
// =====================================================================================
enum RowKind { kS0, kX0, kDyn, kDs, kLane, kThrUb, kThrLb, kStUb, kStLb, kDthr, kDst, kDthr0, kDst0 };

struct Row { int kind, stage, comp; double lb, ub; };

struct SparseRow { int n = 0; int col[12]; double val[12]; void add(int c, double v) { col[n] = c; val[n] = v; ++n; } };

struct Triplet { int i, j; double v; };      // symmetric entry of the lower triangle (i >= j)

struct Problem {
  int N, model, wrap, lane, integrator;
  double Ts;
  Params p;
  double x0[6], u_prev[2], s0, cx[5], cy[5], max_err;
  bool has_thr, has_st;
  std::vector<double> u_warm;     // [2][N] or empty
  int n, m;
  std::vector<Row> rows;
  std::vector<double> gl, gu, w0;

  int iU(int k, int i) const { return 2 * i + k; }
  int iS(int i) const { return 2 * N + i; }
  int iX(int k, int i) const { return 3 * N + 1 + 6 * i + k; }

  void build() {
    n = 9 * N + 7;
    rows.clear();
    rows.push_back({kS0, 0, 0, s0, s0});                                        // control/MPC.py:101
    for (int k = 0; k < 6; ++k) rows.push_back({kX0, 0, k, x0[k], x0[k]});       // :102-107
    for (int i = 1; i <= N; ++i) {
      for (int k = 0; k < 6; ++k) rows.push_back({kDyn, i, k, 0.0, 0.0});        // :133
      rows.push_back({kDs, i, 0, p.min_s_delta, Ts * p.v_max});                 // :134
      if (lane) rows.push_back({kLane, i, 0, -max_err, max_err});               // :135 (commented out upstream)
    }
    for (int i = 0; i < N; ++i) {                                               // :137-143
      rows.push_back({kThrUb, i, 0, -kInf, p.d_max});
      rows.push_back({kThrLb, i, 0, p.min_throttle, kInf});
      rows.push_back({kStUb, i, 0, -kInf, p.max_steer});
      rows.push_back({kStLb, i, 0, p.min_steer, kInf});
      if (i > 0 || wrap) {
        rows.push_back({kDthr, i, 0, p.min_throttle_delta, p.max_throttle_delta});
        rows.push_back({kDst, i, 0, p.min_steer_delta, p.max_steer_delta});
      }
    }
    if (has_thr) rows.push_back({kDthr0, 0, 0, p.min_throttle_delta, p.max_throttle_delta});   // :145-149
    if (has_st) rows.push_back({kDst0, 0, 0, p.min_steer_delta, p.max_steer_delta});
    m = (int)rows.size();
    gl.resize(m);
    gu.resize(m);
    for (int r = 0; r < m; ++r) { gl[r] = rows[r].lb; gu[r] = rows[r].ub; }
    // ---- initial guess (control/MPC.py:109-131) ----
    w0.assign(n, 0.0);
    std::vector<double> up(2 * N);
    for (int i = 0; i < N; ++i) {
      if (!u_warm.empty()) {                       // :120-121: shift by one, repeat the last
        const int src = (i + 1 < N) ? i + 1 : N - 1;
        up[2 * i] = u_warm[src];
        up[2 * i + 1] = u_warm[N + src];
      } else {                                     // :123
        up[2 * i] = u_prev[0];
        up[2 * i + 1] = u_prev[1];
      }
      w0[iU(0, i)] = up[2 * i];
      w0[iU(1, i)] = up[2 * i + 1];
    }
    double x[6], xn[6];
    for (int k = 0; k < 6; ++k) { x[k] = x0[k]; w0[iX(k, 0)] = x[k]; }
    w0[iS(0)] = s0;
    for (int i = 1; i <= N; ++i) {
      w0[iS(i)] = s0 + i * Ts * p.v_max;           // :127
      f_step<double>(0, integrator, x, &up[2 * (i - 1)], Ts, xn);      // :129 always f_vehicle
      for (int k = 0; k < 6; ++k) { x[k] = xn[k]; w0[iX(k, i)] = x[k]; }
    }
  }

  int prevU(int i) const { return (i - 1 + N) % N; }     // control/MPC.py:142-143: U[:, i-1] at i = 0 wraps

  // ---- objective (control/MPC.py:86-98) ----
  double f(const double* w) const {
    double acc = -p.lambda_s * w[iS(N)];
    for (int i = 1; i <= N; ++i)
      acc += stage_cost<double>(p, cx, cy, w[iX(0, i)], w[iX(1, i)], w[iS(i)], w[iX(3, i)], w[iX(4, i)]);
    for (int i = 1; i < N; ++i) {
      const double d = w[iU(1, i)] - w[iU(1, i - 1)];
      acc += p.beta_delta * d * d;
    }
    return acc;
  }

  void grad_f(const double* w, double* g) const {
    std::fill(g, g + n, 0.0);
    for (int i = 1; i <= N; ++i) {
      typedef HD<5> T;
      T c = stage_cost<T>(p, cx, cy, T::var(w[iX(0, i)], 0), T::var(w[iX(1, i)], 1), T::var(w[iS(i)], 2),
                          T::var(w[iX(3, i)], 3), T::var(w[iX(4, i)], 4));
      g[iX(0, i)] += c.g[0]; g[iX(1, i)] += c.g[1]; g[iS(i)] += c.g[2]; g[iX(3, i)] += c.g[3]; g[iX(4, i)] += c.g[4];
    }
    g[iS(N)] += -p.lambda_s;
    for (int i = 1; i < N; ++i) {
      const double d = w[iU(1, i)] - w[iU(1, i - 1)];
      g[iU(1, i)] += 2 * p.beta_delta * d;
      g[iU(1, i - 1)] -= 2 * p.beta_delta * d;
    }
  }

  // ---- constraints g(w), in subject_to order ----
  void g(const double* w, double* out) const {
    double fx[6];
    int cur = -1;
    for (int r = 0; r < m; ++r) {
      const Row& R = rows[r];
      const int i = R.stage;
      switch (R.kind) {
        case kS0: out[r] = w[iS(0)]; break;
        case kX0: out[r] = w[iX(R.comp, 0)]; break;
        case kDyn:
          if (cur != i) { f_step<double>(model, integrator, &w[iX(0, i - 1)], &w[iU(0, i - 1)], Ts, fx); cur = i; }
          out[r] = w[iX(R.comp, i)] - fx[R.comp];
          break;
        case kDs: out[r] = w[iS(i)] - w[iS(i - 1)]; break;
        case kLane: {
          double eC, eL;
          track_errors<double>(cx, cy, w[iX(0, i)], w[iX(1, i)], w[iS(i)], eC, eL);
          out[r] = eC;
        } break;
        case kThrUb: case kThrLb: out[r] = w[iU(0, i)]; break;
        case kStUb: case kStLb: out[r] = w[iU(1, i)]; break;
        case kDthr: out[r] = w[iU(0, i)] - w[iU(0, prevU(i))]; break;
        case kDst: out[r] = w[iU(1, i)] - w[iU(1, prevU(i))]; break;
        case kDthr0: out[r] = w[iU(0, 0)] - u_prev[0]; break;
        case kDst0: out[r] = w[iU(1, 0)] - u_prev[1]; break;
      }
    }
  }

  // model step of stage i (rows x_i - F(x_{i-1}, u_{i-1})) with first and second derivatives over
  // z = (x_{i-1} (6), u_{i-1} (2))
  void model_ad(const double* w, int i, HD<8>* F) const {
    typedef HD<8> T;
    T x[6], u[2];
    for (int k = 0; k < 6; ++k) x[k] = T::var(w[iX(k, i - 1)], k);
    u[0] = T::var(w[iU(0, i - 1)], 6);
    u[1] = T::var(w[iU(1, i - 1)], 7);
    f_step<T>(model, integrator, x, u, Ts, F);
  }
  int zcol(int i, int a) const { return a < 6 ? iX(a, i - 1) : iU(a - 6, i - 1); }

  void jac(const double* w, std::vector<SparseRow>& J) const {
    J.assign(m, SparseRow());
    HD<8> F[6];
    int cur = -1;
    for (int r = 0; r < m; ++r) {
      const Row& R = rows[r];
      const int i = R.stage;
      SparseRow& row = J[r];
      switch (R.kind) {
        case kS0: row.add(iS(0), 1.0); break;
        case kX0: row.add(iX(R.comp, 0), 1.0); break;
        case kDyn:
          if (cur != i) { model_ad(w, i, F); cur = i; }
          row.add(iX(R.comp, i), 1.0);
          for (int a = 0; a < 8; ++a)
            if (F[R.comp].g[a] != 0.0) row.add(zcol(i, a), -F[R.comp].g[a]);
          break;
        case kDs: row.add(iS(i), 1.0); row.add(iS(i - 1), -1.0); break;
        case kLane: {
          typedef HD<3> T;
          T eC, eL;
          track_errors<T>(cx, cy, T::var(w[iX(0, i)], 0), T::var(w[iX(1, i)], 1), T::var(w[iS(i)], 2), eC, eL);
          row.add(iX(0, i), eC.g[0]); row.add(iX(1, i), eC.g[1]); row.add(iS(i), eC.g[2]);
        } break;
        case kThrUb: case kThrLb: case kDthr0: row.add(iU(0, i), 1.0); break;
        case kStUb: case kStLb: case kDst0: row.add(iU(1, i), 1.0); break;
        case kDthr: row.add(iU(0, i), 1.0); row.add(iU(0, prevU(i)), -1.0); break;
        case kDst: row.add(iU(1, i), 1.0); row.add(iU(1, prevU(i)), -1.0); break;
      }
    }
  }

  // Hessian of sigma_f f + lam' g, lower triangle as triplets
  void hess(const double* w, double sigma_f, const double* lam, std::vector<Triplet>& H) const {
    H.clear();
    auto put = [&](int a, int b, double v) {
      if (v == 0.0) return;
      if (a < b) std::swap(a, b);
      H.push_back({a, b, v});
    };
    if (sigma_f != 0.0) {
      for (int i = 1; i <= N; ++i) {
        typedef HD<5> T;
        T c = stage_cost<T>(p, cx, cy, T::var(w[iX(0, i)], 0), T::var(w[iX(1, i)], 1), T::var(w[iS(i)], 2),
                            T::var(w[iX(3, i)], 3), T::var(w[iX(4, i)], 4));
        const int idx[5] = {iX(0, i), iX(1, i), iS(i), iX(3, i), iX(4, i)};
        for (int a = 0; a < 5; ++a)
          for (int b = 0; b <= a; ++b) put(idx[a], idx[b], sigma_f * c.hess(a, b));
      }
      const double b2 = 2 * p.beta_delta * sigma_f;
      for (int i = 1; i < N; ++i) {
        put(iU(1, i), iU(1, i), b2);
        put(iU(1, i - 1), iU(1, i - 1), b2);
        put(iU(1, i), iU(1, i - 1), -b2);
      }
    }
    HD<8> F[6];
    int r = 7;
    for (int i = 1; i <= N; ++i) {
      bool any = false;
      for (int k = 0; k < 6; ++k) any = any || lam[r + k] != 0.0;
      if (any) {
        model_ad(w, i, F);
        for (int a = 0; a < 8; ++a)
          for (int b = 0; b <= a; ++b) {
            double acc = 0.0;
            for (int k = 0; k < 6; ++k) acc += lam[r + k] * F[k].hess(a, b);
            put(zcol(i, a), zcol(i, b), -acc);
          }
      }
      r += 6 + 1;
      if (lane) {
        if (lam[r] != 0.0) {
          typedef HD<3> T;
          T eC, eL;
          track_errors<T>(cx, cy, T::var(w[iX(0, i)], 0), T::var(w[iX(1, i)], 1), T::var(w[iS(i)], 2), eC, eL);
          const int idx[3] = {iX(0, i), iX(1, i), iS(i)};
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b <= a; ++b) put(idx[a], idx[b], lam[r] * eC.hess(a, b));
        }
        r += 1;
      }
    }
  }
};

// =====================================================================================
// KKT solver:  [ W + Jd' D Jd   Jc' ] [dx ]   [r1]
//              [ Jc             0   ] [dyc] = [r2]
// stage ordering -> block tridiagonal; "border" inequality rows (the wrap rows) stay explicit:
//              [ K    A       ] [ z   ]   [ r  ]
//              [ A'  -D_b^-1  ] [ dyb ] = [ rb ]
// =====================================================================================
struct KKTSolver {
  int refine = 1;     // steps of iterative refinement per solve (oracle/ipm.py: 1)
  virtual ~KKTSolver() {}
  // entries of the symmetric matrix K (lower triangle, indices in [0, n + mc)), border columns
  virtual void begin(int ntot, int nborder) = 0;
  virtual void add(int i, int j, double v) = 0;
  virtual void add_border(int b, int i, double v) = 0;
  virtual void set_border_diag(int b, double v) = 0;
  virtual void factor(int& npos, int& nneg, int& nzero) = 0;
  virtual void solve(const double* r, const double* rb, double* z, double* zb) const = 0;   // with one refinement step
};

// ---- dense cross-check: one Bunch-Kaufman LDL^T of the whole bordered matrix -----------------
struct DenseKKT : KKTSolver {
  int nt = 0, nb = 0, dim = 0;
  std::vector<double> M;
  BK bk;
  void begin(int ntot, int nborder) override { nt = ntot; nb = nborder; dim = nt + nb; M.assign((size_t)dim * dim, 0.0); }
  void add(int i, int j, double v) override { M[(size_t)i * dim + j] += v; if (i != j) M[(size_t)j * dim + i] += v; }
  void add_border(int b, int i, double v) override { M[(size_t)(nt + b) * dim + i] += v; M[(size_t)i * dim + nt + b] += v; }
  void set_border_diag(int b, double v) override { M[(size_t)(nt + b) * dim + nt + b] = v; }
  void factor(int& npos, int& nneg, int& nzero) override {
    bk.factor(M.data(), dim);
    npos = bk.npos; nneg = bk.nneg; nzero = bk.nzero;
  }
  void solve(const double* r, const double* rb, double* z, double* zb) const override {
    std::vector<double> rhs(dim), x(dim), res(dim);
    for (int i = 0; i < nt; ++i) rhs[i] = r[i];
    for (int b = 0; b < nb; ++b) rhs[nt + b] = rb[b];
    x = rhs;
    bk.solve(x.data());
    std::fill(res.begin(), res.end(), 0.0);
    if (refine) for (int i = 0; i < dim; ++i) {
      double acc = rhs[i];
      for (int j = 0; j < dim; ++j) acc -= M[(size_t)i * dim + j] * x[j];
      res[i] = acc;
    }
    if (refine) bk.solve(res.data());
    for (int i = 0; i < nt; ++i) z[i] = x[i] + res[i];
    for (int b = 0; b < nb; ++b) zb[b] = x[nt + b] + res[nt + b];
  }
};

// ---- block tridiagonal + border -----------------------------------------------------------------
struct BlockTriKKT : KKTSolver {
  static constexpr int BS = 16;
  int nblk = 0, nt = 0, nb = 0;
  std::vector<int> blk_of, off_of;      // global index -> (block, offset)
  std::vector<int> bsz, bstart;         // block sizes / start in the permuted ordering
  std::vector<double> Kd, Ko;           // [nblk][BS][BS]: diagonal blocks, sub-diagonal blocks K_{k,k-1}
  std::vector<double> Ab;               // border columns in permuted ordering [nb][nt]
  double bdiag[4];
  std::vector<BK> S;                    // factorised Schur complements of the diagonal blocks
  std::vector<double> Yb;               // K^-1 A  [nb][nt]
  double Sb[4][4], Sbinv[4][4];

  void layout(const std::vector<int>& block_of_index, int nblocks) {
    nblk = nblocks;
    nt = (int)block_of_index.size();
    blk_of = block_of_index;
    off_of.assign(nt, 0);
    bsz.assign(nblk, 0);
    for (int i = 0; i < nt; ++i) off_of[i] = bsz[blk_of[i]]++;
    bstart.assign(nblk + 1, 0);
    for (int k = 0; k < nblk; ++k) bstart[k + 1] = bstart[k] + bsz[k];
    S.resize(nblk);
  }
  int pos(int i) const { return bstart[blk_of[i]] + off_of[i]; }

  void begin(int ntot, int nborder) override {
    (void)ntot;
    nb = nborder;
    Kd.assign((size_t)nblk * BS * BS, 0.0);
    Ko.assign((size_t)nblk * BS * BS, 0.0);
    Ab.assign((size_t)nb * nt, 0.0);
  }
  void add(int i, int j, double v) override {
    int bi = blk_of[i], bj = blk_of[j], oi = off_of[i], oj = off_of[j];
    if (bi == bj) {
      Kd[((size_t)bi * BS + oi) * BS + oj] += v;
      if (oi != oj) Kd[((size_t)bi * BS + oj) * BS + oi] += v;
    } else {
      if (bi < bj) { std::swap(bi, bj); std::swap(oi, oj); }
      // bi == bj + 1 by construction of the stage ordering
      Ko[((size_t)bi * BS + oi) * BS + oj] += v;
    }
  }
  void add_border(int b, int i, double v) override { Ab[(size_t)b * nt + pos(i)] += v; }
  void set_border_diag(int b, double v) override { bdiag[b] = v; }

  // y <- K y for the block-tridiagonal part (permuted ordering)
  void matvec(const double* x, double* y) const {
    for (int k = 0; k < nblk; ++k) {
      const int n = bsz[k];
      for (int a = 0; a < n; ++a) {
        double acc = 0.0;
        const double* row = &Kd[((size_t)k * BS + a) * BS];
        for (int c = 0; c < n; ++c) acc += row[c] * x[bstart[k] + c];
        if (k > 0) {
          const double* ro = &Ko[((size_t)k * BS + a) * BS];
          for (int c = 0; c < bsz[k - 1]; ++c) acc += ro[c] * x[bstart[k - 1] + c];
        }
        y[bstart[k] + a] = acc;
      }
      if (k + 1 < nblk) {
        const int n1 = bsz[k + 1];
        for (int a = 0; a < n; ++a) {
          double acc = 0.0;
          for (int c = 0; c < n1; ++c) acc += Ko[((size_t)(k + 1) * BS + c) * BS + a] * x[bstart[k + 1] + c];
          y[bstart[k] + a] += acc;
        }
      }
    }
  }

  void tri_solve(double* y) const {     // in place, permuted ordering
    double tmp[BS];
    for (int k = 1; k < nblk; ++k) {    // forward: y_k -= M_k S_{k-1}^-1 y_{k-1}
      const int np = bsz[k - 1], n = bsz[k];
      for (int c = 0; c < np; ++c) tmp[c] = y[bstart[k - 1] + c];
      S[k - 1].solve(tmp);
      for (int a = 0; a < n; ++a) {
        double acc = 0.0;
        const double* ro = &Ko[((size_t)k * BS + a) * BS];
        for (int c = 0; c < np; ++c) acc += ro[c] * tmp[c];
        y[bstart[k] + a] -= acc;
      }
    }
    for (int k = nblk - 1; k >= 0; --k) {   // backward: x_k = S_k^-1 (y_k - M_{k+1}' x_{k+1})
      const int n = bsz[k];
      if (k + 1 < nblk) {
        const int n1 = bsz[k + 1];
        for (int a = 0; a < n; ++a) {
          double acc = 0.0;
          for (int c = 0; c < n1; ++c) acc += Ko[((size_t)(k + 1) * BS + c) * BS + a] * y[bstart[k + 1] + c];
          y[bstart[k] + a] -= acc;
        }
      }
      S[k].solve(&y[bstart[k]]);
    }
  }

  void factor(int& npos, int& nneg, int& nzero) override {
    npos = nneg = nzero = 0;
    double Sk[BS * BS], col[BS], Z[BS * BS];
    for (int k = 0; k < nblk; ++k) {
      const int n = bsz[k];
      for (int a = 0; a < n; ++a)
        for (int c = 0; c < n; ++c) Sk[a * n + c] = Kd[((size_t)k * BS + a) * BS + c];
      if (k > 0) {
        const int np = bsz[k - 1];
        // Z = S_{k-1}^-1 M_k'  (np x n), column by column
        for (int c = 0; c < n; ++c) {
          for (int r = 0; r < np; ++r) col[r] = Ko[((size_t)k * BS + c) * BS + r];
          S[k - 1].solve(col);
          for (int r = 0; r < np; ++r) Z[r * n + c] = col[r];
        }
        for (int a = 0; a < n; ++a)
          for (int c = 0; c <= a; ++c) {
            double acc = 0.0;
            const double* ro = &Ko[((size_t)k * BS + a) * BS];
            for (int r = 0; r < np; ++r) acc += ro[r] * Z[r * n + c];
            Sk[a * n + c] -= acc;
            if (c != a) Sk[c * n + a] -= acc;
          }
      }
      S[k].factor(Sk, n);
      npos += S[k].npos; nneg += S[k].nneg; nzero += S[k].nzero;
    }
    if (nb > 0) {
      Yb = Ab;
      for (int b = 0; b < nb; ++b) tri_solve(&Yb[(size_t)b * nt]);
      for (int a = 0; a < nb; ++a)
        for (int b = 0; b < nb; ++b) {
          double acc = (a == b) ? bdiag[a] : 0.0;
          for (int i = 0; i < nt; ++i) acc -= Ab[(size_t)a * nt + i] * Yb[(size_t)b * nt + i];
          Sb[a][b] = acc;
        }
      double flat[16];
      for (int a = 0; a < nb; ++a)
        for (int b = 0; b < nb; ++b) flat[a * nb + b] = 0.5 * (Sb[a][b] + Sb[b][a]);
      BK sb;
      sb.factor(flat, nb);
      npos += sb.npos; nneg += sb.nneg; nzero += sb.nzero;
      for (int b = 0; b < nb; ++b) {
        double e[4] = {0, 0, 0, 0};
        e[b] = 1.0;
        sb.solve(e);
        for (int a = 0; a < nb; ++a) Sbinv[a][b] = e[a];
      }
    }
  }

  void solve_once(const double* rp, const double* rb, double* zp, double* zb) const {   // permuted ordering
    for (int i = 0; i < nt; ++i) zp[i] = rp[i];
    tri_solve(zp);
    if (nb > 0) {
      double t[4];
      for (int a = 0; a < nb; ++a) {
        double acc = rb[a];
        for (int i = 0; i < nt; ++i) acc -= Ab[(size_t)a * nt + i] * zp[i];
        t[a] = acc;
      }
      for (int a = 0; a < nb; ++a) {
        double acc = 0.0;
        for (int b = 0; b < nb; ++b) acc += Sbinv[a][b] * t[b];
        zb[a] = acc;
      }
      for (int b = 0; b < nb; ++b)
        for (int i = 0; i < nt; ++i) zp[i] -= Yb[(size_t)b * nt + i] * zb[b];
    }
  }

  void solve(const double* r, const double* rb, double* z, double* zb) const override {
    std::vector<double> rp(nt), zp(nt), res(nt), cp(nt);
    double resb[4] = {0, 0, 0, 0}, cb[4] = {0, 0, 0, 0}, zb0[4] = {0, 0, 0, 0};
    for (int i = 0; i < nt; ++i) rp[pos(i)] = r[i];
    solve_once(rp.data(), rb, zp.data(), zb0);
    if (!refine) {
      for (int i = 0; i < nt; ++i) z[i] = zp[pos(i)];
      for (int b = 0; b < nb; ++b) zb[b] = zb0[b];
      return;
    }
    // one step of iterative refinement (oracle/ipm.py does the same; IPOPT refines too)
    matvec(zp.data(), res.data());
    for (int i = 0; i < nt; ++i) res[i] = rp[i] - res[i];
    for (int b = 0; b < nb; ++b) {
      for (int i = 0; i < nt; ++i) res[i] -= Ab[(size_t)b * nt + i] * zb0[b];
      double acc = rb[b] - bdiag[b] * zb0[b];
      for (int i = 0; i < nt; ++i) acc -= Ab[(size_t)b * nt + i] * zp[i];
      resb[b] = acc;
    }
    solve_once(res.data(), resb, cp.data(), cb);
    for (int i = 0; i < nt; ++i) z[i] = zp[pos(i)] + cp[pos(i)];
    for (int b = 0; b < nb; ++b) zb[b] = zb0[b] + cb[b];
  }
};

// =====================================================================================
// The interior-point solver: transcription of oracle/ipm.py::solve (IPOPT's algorithm)
// =====================================================================================
struct Options {
  int max_iter = 500;
  double tol = 1e-4, acceptable_tol = 1e-2;
  int acceptable_iter = 15;
  double dual_inf_tol = 1.0, constr_viol_tol = 1e-4, compl_inf_tol = 1e-4;
  double acceptable_dual_inf_tol = 1e10, acceptable_constr_viol_tol = 1e-2, acceptable_compl_inf_tol = 1e-2;
  double nlp_scaling_max_gradient = 100.0, nlp_scaling_min_value = 1e-8;
  double bound_relax_factor = 1e-8, bound_push = 0.01, bound_frac = 0.01;
  double mu_init = 0.1, mu_linear_decrease_factor = 0.2, mu_superlinear_decrease_power = 1.5;
  double barrier_tol_factor = 10.0, mu_min = 1e-11, tau_min = 0.99, kappa_d = 1e-5, kappa_sigma = 1e10;
  double constr_mult_init_max = 1000.0, bound_mult_init_val = 1.0, s_max = 100.0;
  double theta_max_fact = 1e4, theta_min_fact = 1e-4, gamma_theta = 1e-5, gamma_phi = 1e-8;
  double delta = 1.0, s_theta = 1.1, s_phi = 2.3, eta_phi = 1e-8, alpha_min_frac = 0.05;
  double alpha_red_factor = 0.5;
  int max_soc = 4;
  double kappa_soc = 0.99;
  double delta_w_init = 1e-4, delta_w_min = 1e-20, delta_w_max = 1e40;
  double kappa_w_minus = 1.0 / 3.0, kappa_w_plus = 8.0, kappa_w_plus_first = 100.0;
  int restoration = 0;
  double resto_weight = 1.0, kappa_resto = 0.9;
  int resto_max_iter = 100;
  double polish_tol = 0.0;      // > 0: after termination keep iterating to this tolerance (status / iters of the
                                // first termination are reported), DESIGN.md "polish"
  int linsol = 0;               // 0 = block tridiagonal + border, 1 = dense Bunch-Kaufman
  int refine = 1;               // iterative-refinement steps of every KKT solve
};

enum { SOLVE_SUCCEEDED = 0, SOLVED_TO_ACCEPTABLE_LEVEL = 1, MAXITER_EXCEEDED = -1, RESTORATION_NEEDED = -2,
       ERROR_IN_STEP_COMPUTATION = -3, INVALID_NUMBER = -4 };

struct Result {
  std::vector<double> x, lam_g;
  double obj = 0.0;
  int status = MAXITER_EXCEEDED, iters = 0, polish_iters = 0, n_fact = 0;
};

typedef std::vector<double> Vec;

static double maxabs(const Vec& v) { double m = 0.0; for (double a : v) m = std::max(m, std::fabs(a)); return m; }
static bool all_finite(const Vec& v) { for (double a : v) if (!std::isfinite(a)) return false; return true; }

struct Ipm {
  const Problem& nlp;
  Options o;
  int n, m, mc, md;
  std::vector<int> ic, idn;       // equality rows / inequality rows (indices into g)
  std::vector<int> border;        // positions in idn of the rows kept as border (wrap rows)
  std::vector<int> border_of;     // idn position -> border slot or -1
  Vec dc_s, dd_s, cval, dL, dU;
  std::vector<char> hasL, hasU;
  double df = 1.0;
  KKTSolver* lin = nullptr;
  BlockTriKKT blk;
  DenseKKT dense;
  std::vector<SparseRow> J;       // unscaled Jacobian at the current point
  std::vector<Triplet> Htrip;
  Vec gg;

  Ipm(const Problem& p, const Options& opt) : nlp(p), o(opt) {
    n = p.n;
    m = p.m;
    for (int r = 0; r < m; ++r) (p.gl[r] == p.gu[r] ? ic : idn).push_back(r);
    mc = (int)ic.size();
    md = (int)idn.size();
    border_of.assign(md, -1);
    for (int q = 0; q < md; ++q) {
      const Row& R = p.rows[idn[q]];
      if ((R.kind == kDthr || R.kind == kDst) && R.stage == 0) {
        border_of[q] = (int)border.size();
        border.push_back(q);
      }
    }
    // stage ordering of the KKT unknowns (primal n, then equality multipliers mc)
    std::vector<int> blkidx(n + mc);
    for (int i = 0; i < p.N; ++i) { blkidx[p.iU(0, i)] = i; blkidx[p.iU(1, i)] = i; }
    for (int i = 0; i <= p.N; ++i) {
      blkidx[p.iS(i)] = i;
      for (int k = 0; k < 6; ++k) blkidx[p.iX(k, i)] = i;
    }
    for (int q = 0; q < mc; ++q) blkidx[n + q] = p.rows[ic[q]].stage;     // s0 / x0 rows: stage 0; dyn rows: stage i
    blk.layout(blkidx, p.N + 1);
    lin = (o.linsol == 1) ? (KKTSolver*)&dense : (KKTSolver*)&blk;
    lin->refine = o.refine;
    gg.resize(m);
  }

  // ---- scaled evaluations -------------------------------------------------------------------
  void evals(const Vec& x, double& fv, Vec& c, Vec& d) {
    nlp.g(x.data(), gg.data());
    fv = df * nlp.f(x.data());
    for (int q = 0; q < mc; ++q) c[q] = dc_s[q] * (gg[ic[q]] - cval[q]);
    for (int q = 0; q < md; ++q) d[q] = dd_s[q] * gg[idn[q]];
  }
  void JcT(const Vec& y, Vec& out) const {       // out += Jc' y
    for (int q = 0; q < mc; ++q) {
      const SparseRow& r = J[ic[q]];
      const double s = dc_s[q] * y[q];
      for (int k = 0; k < r.n; ++k) out[r.col[k]] += s * r.val[k];
    }
  }
  void JdT(const Vec& y, Vec& out) const {
    for (int q = 0; q < md; ++q) {
      const SparseRow& r = J[idn[q]];
      const double s = dd_s[q] * y[q];
      for (int k = 0; k < r.n; ++k) out[r.col[k]] += s * r.val[k];
    }
  }
  double Jd_row(int q, const Vec& dx) const {
    const SparseRow& r = J[idn[q]];
    double acc = 0.0;
    for (int k = 0; k < r.n; ++k) acc += r.val[k] * dx[r.col[k]];
    return dd_s[q] * acc;
  }

  // ---- KKT: assemble + factor  (oracle/ipm.py::_KKT) -----------------------------------------------
  Vec Dcur;   // D = sig_s + dw of the factorised system
  bool kkt_factor(const std::vector<Triplet>* H, double hdiag_extra_is_dw, const Vec* hdiag, const Vec& D, int& n_fact) {
    const int nbord = (int)border.size();
    lin->begin(n + mc, nbord);
    if (H) for (const Triplet& t : *H) lin->add(t.i, t.j, t.v);
    for (int i = 0; i < n; ++i) lin->add(i, i, hdiag_extra_is_dw + (hdiag ? (*hdiag)[i] : 0.0));
    for (int q = 0; q < md; ++q) {
      const SparseRow& r = J[idn[q]];
      if (border_of[q] >= 0) {
        for (int k = 0; k < r.n; ++k) lin->add_border(border_of[q], r.col[k], dd_s[q] * r.val[k]);
        lin->set_border_diag(border_of[q], -1.0 / D[q]);
        continue;
      }
      const double s = D[q] * dd_s[q] * dd_s[q];
      for (int a = 0; a < r.n; ++a)
        for (int b = 0; b < r.n; ++b)
          if (r.col[a] >= r.col[b]) lin->add(r.col[a], r.col[b], s * r.val[a] * r.val[b]);
    }
    for (int q = 0; q < mc; ++q) {
      const SparseRow& r = J[ic[q]];
      for (int k = 0; k < r.n; ++k) lin->add(n + q, r.col[k], dc_s[q] * r.val[k]);
    }
    int np, nn, nz;
    lin->factor(np, nn, nz);
    n_fact++;
    Dcur = D;
    return np == n && nn == mc + nbord && nz == 0;
  }

  // oracle/ipm.py::_KKT.solve(rx, rs, rc, rd)
  void kkt_solve(const Vec& rx, const Vec& rs, const Vec& rc, const Vec& rd, Vec& dx, Vec& ds, Vec& dyc, Vec& dyd) const {
    const int nbord = (int)border.size();
    Vec rhs(n + mc, 0.0), t(md, 0.0), sol(n + mc);
    double rb[4] = {0, 0, 0, 0}, zb[4] = {0, 0, 0, 0};
    for (int q = 0; q < md; ++q) {
      if (border_of[q] >= 0) rb[border_of[q]] = rd[q] + rs[q] / Dcur[q];
      else t[q] = Dcur[q] * rd[q] + rs[q];
    }
    for (int i = 0; i < n; ++i) rhs[i] = rx[i];
    {
      Vec acc(n, 0.0);
      JdT(t, acc);
      for (int i = 0; i < n; ++i) rhs[i] += acc[i];
    }
    for (int q = 0; q < mc; ++q) rhs[n + q] = rc[q];
    lin->solve(rhs.data(), rb, sol.data(), zb);
    (void)nbord;
    dx.assign(sol.begin(), sol.begin() + n);
    dyc.assign(sol.begin() + n, sol.end());
    ds.resize(md);
    dyd.resize(md);
    for (int q = 0; q < md; ++q) {
      ds[q] = Jd_row(q, dx) - rd[q];
      dyd[q] = Dcur[q] * ds[q] - rs[q];
    }
  }

  Result solve() {
    Result res;
    Vec x = nlp.w0;
    // ---- gradient-based scaling at the starting point ---------------------------------------
    Vec g0(n);
    nlp.grad_f(x.data(), g0.data());
    nlp.jac(x.data(), J);
    const double gmax = maxabs(g0);
    df = 1.0;
    if (gmax > o.nlp_scaling_max_gradient) df = std::max(o.nlp_scaling_max_gradient / gmax, o.nlp_scaling_min_value);
    Vec dg(m, 1.0);
    for (int r = 0; r < m; ++r) {
      double rowmax = 0.0;
      for (int k = 0; k < J[r].n; ++k) rowmax = std::max(rowmax, std::fabs(J[r].val[k]));
      if (rowmax > o.nlp_scaling_max_gradient)
        dg[r] = std::max(o.nlp_scaling_max_gradient / rowmax, o.nlp_scaling_min_value);
    }
    dc_s.resize(mc); cval.resize(mc);
    dd_s.resize(md); dL.resize(md); dU.resize(md); hasL.resize(md); hasU.resize(md);
    for (int q = 0; q < mc; ++q) { dc_s[q] = dg[ic[q]]; cval[q] = nlp.gl[ic[q]]; }
    auto relax = [&](double b) { return std::min(o.constr_viol_tol, o.bound_relax_factor * std::max(1.0, std::fabs(b))); };
    for (int q = 0; q < md; ++q) {
      dd_s[q] = dg[idn[q]];
      dL[q] = nlp.gl[idn[q]] * dd_s[q];
      dU[q] = nlp.gu[idn[q]] * dd_s[q];
      hasL[q] = std::isfinite(dL[q]);
      hasU[q] = std::isfinite(dU[q]);
      dL[q] = hasL[q] ? dL[q] - relax(dL[q]) : -kInf;
      dU[q] = hasU[q] ? dU[q] + relax(dU[q]) : kInf;
    }
    auto lam_of = [&](const Vec& yc, const Vec& yd, Vec& lam) {
      lam.assign(m, 0.0);
      for (int q = 0; q < mc; ++q) lam[ic[q]] = yc[q] * dc_s[q];
      for (int q = 0; q < md; ++q) lam[idn[q]] = yd[q] * dd_s[q];
    };
    // ---- starting point: slacks pushed into their bounds --------------------------------------
    double fval;
    Vec c(mc), d(md);
    evals(x, fval, c, d);
    Vec s = d, pL(md), pU(md);
    for (int q = 0; q < md; ++q) {
      const double aL = std::max(1.0, std::fabs(hasL[q] ? dL[q] : 0.0)), aU = std::max(1.0, std::fabs(hasU[q] ? dU[q] : 0.0));
      if (hasL[q] && hasU[q]) {
        pL[q] = std::min(o.bound_push * aL, o.bound_frac * (dU[q] - dL[q]));
        pU[q] = std::min(o.bound_push * aU, o.bound_frac * (dU[q] - dL[q]));
      } else {
        pL[q] = o.bound_push * aL;
        pU[q] = o.bound_push * aU;
      }
      if (hasL[q]) s[q] = std::max(s[q], dL[q] + pL[q]);
      if (hasU[q]) s[q] = std::min(s[q], dU[q] - pU[q]);
    }
    Vec vL(md), vU(md);
    for (int q = 0; q < md; ++q) { vL[q] = hasL[q] ? o.bound_mult_init_val : 0.0; vU[q] = hasU[q] ? o.bound_mult_init_val : 0.0; }
    int n_fact = 0;
    // ---- least-squares equality multipliers ---------------------------------------------------
    Vec gf(n), yc(mc), yd(md);
    auto lsq_multipliers = [&]() {
      nlp.grad_f(x.data(), gf.data());
      for (double& v : gf) v *= df;
      Vec D1(md, 1.0);                             // _KKT(H = 0, sig_s = 0, dw = 1): D = 1, H + dw I = I
      kkt_factor(nullptr, 1.0, nullptr, D1, n_fact);
      Vec rs(md), z0c(mc, 0.0), z0d(md, 0.0), dx, ds, dyc, dyd;
      for (int q = 0; q < md; ++q) rs[q] = vU[q] - vL[q];
      kkt_solve(gf, rs, z0c, z0d, dx, ds, dyc, dyd);
      for (int q = 0; q < mc; ++q) yc[q] = -dyc[q];
      for (int q = 0; q < md; ++q) yd[q] = -dyd[q];
      if (std::max(maxabs(yc), maxabs(yd)) > o.constr_mult_init_max) {
        std::fill(yc.begin(), yc.end(), 0.0);
        std::fill(yd.begin(), yd.end(), 0.0);
      }
    };
    lsq_multipliers();

    double mu = o.mu_init, tau = std::max(o.tau_min, 1.0 - mu);
    const double kd = o.kappa_d;
    auto SL = [&](const Vec& ss, int q) { return hasL[q] ? ss[q] - dL[q] : 1.0; };
    auto SU = [&](const Vec& ss, int q) { return hasU[q] ? dU[q] - ss[q] : 1.0; };
    auto barrier_obj = [&](double fv, const Vec& ss, double mu_) {
      double lg = 0.0, lin_ = 0.0;
      for (int q = 0; q < md; ++q) {
        if (hasL[q]) lg += std::log(SL(ss, q));
        if (hasU[q]) lg += std::log(SU(ss, q));
        if (hasL[q] && !hasU[q]) lin_ += SL(ss, q);
        if (hasU[q] && !hasL[q]) lin_ += SU(ss, q);
      }
      return fv - mu_ * lg + kd * mu_ * lin_;
    };
    auto theta_of = [&](const Vec& cc, const Vec& dd, const Vec& ss, bool inner) {
      double t = 0.0;
      for (double v : cc) t += std::fabs(v);
      if (!inner) for (int q = 0; q < md; ++q) t += std::fabs(dd[q] - ss[q]);
      return t;
    };
    bool resto = false;
    struct Err { double E, dual_inf, pr_inf, compl_; };
    auto errors = [&](double mu_, const Vec& gf_, const Vec& c_, const Vec& d_, const Vec& s_, const Vec& yc_,
                      const Vec& yd_, const Vec& vL_, const Vec& vU_) {
      Vec rx = gf_;
      JcT(yc_, rx);
      JdT(yd_, rx);
      double dual_inf = maxabs(rx), pr_inf = maxabs(c_), compl_ = 0.0, sum_y = 0.0, sum_v = 0.0;
      int nb = 0;
      for (int q = 0; q < md; ++q) {
        dual_inf = std::max(dual_inf, std::fabs(-yd_[q] - vL_[q] + vU_[q]));
        if (!resto) pr_inf = std::max(pr_inf, std::fabs(d_[q] - s_[q]));
        if (hasL[q]) { compl_ = std::max(compl_, std::fabs(SL(s_, q) * vL_[q] - mu_)); nb++; }
        if (hasU[q]) { compl_ = std::max(compl_, std::fabs(SU(s_, q) * vU_[q] - mu_)); nb++; }
        sum_y += std::fabs(yd_[q]);
        sum_v += std::fabs(vL_[q]) + std::fabs(vU_[q]);
      }
      for (double v : yc_) sum_y += std::fabs(v);
      const double smax = o.s_max;
      const double s_d = std::max(smax, (sum_y + sum_v) / std::max(1, mc + md + nb)) / smax;
      const double s_c = std::max(smax, sum_v / std::max(1, nb)) / smax;
      Err e;
      e.E = std::max(dual_inf / s_d, std::max(pr_inf, compl_ / s_c));
      e.dual_inf = dual_inf; e.pr_inf = pr_inf; e.compl_ = compl_;
      return e;
    };

    typedef std::pair<double, double> FE;
    std::vector<FE> filt;
    const double theta0 = theta_of(c, d, s, false);
    double theta_max = o.theta_max_fact * std::max(1.0, theta0), theta_min = o.theta_min_fact * std::max(1.0, theta0);
    double dw_last = 0.0;
    int acceptable_count = 0, status = MAXITER_EXCEEDED, it = 0, resto_it = 0;
    const double wq = o.resto_weight;
    double zeta = 0.0, theta_R = 0.0;
    Vec x_R, DR2;
    struct Saved { double mu, tau, theta_max, theta_min; std::vector<FE> filt; int acc; } saved;
    double tol = o.tol;
    bool polishing = false;
    int first_status = 0, first_iters = 0;

    auto f_resto = [&](const Vec& xx, const Vec& dd, const Vec& ss) {
      double a = 0.0, b = 0.0;
      for (int q = 0; q < md; ++q) a += (dd[q] - ss[q]) * (dd[q] - ss[q]);
      for (int i = 0; i < n; ++i) b += DR2[i] * (xx[i] - x_R[i]) * (xx[i] - x_R[i]);
      return 0.5 * wq * a + 0.5 * zeta * b;
    };
    auto acceptable_to_filter = [&](const std::vector<FE>& F, double th, double ph) {
      for (const FE& e : F) if (!(th <= e.first || ph <= e.second)) return false;
      return true;
    };
    auto le = [&](double a, double b, double ref) { return a - b <= 10.0 * std::numeric_limits<double>::epsilon() * std::fabs(ref); };

    Vec lam(m);
    while (true) {
      if (resto) { gf.resize(n); for (int i = 0; i < n; ++i) gf[i] = zeta * DR2[i] * (x[i] - x_R[i]); }
      else { nlp.grad_f(x.data(), gf.data()); for (double& v : gf) v *= df; }
      nlp.jac(x.data(), J);
      if (resto && resto_it >= 1) {
        // ---- leave the restoration phase?  acceptable to the ORIGINAL filter, theta reduced ------
        const double fo = df * nlp.f(x.data());
        const double th_o = theta_of(c, d, s, false);
        bool leave = std::isfinite(fo) && th_o <= o.kappa_resto * theta_R && th_o <= saved.theta_max;
        if (leave) {
          const double ph_o = barrier_obj(fo, s, saved.mu);
          leave = std::isfinite(ph_o) && acceptable_to_filter(saved.filt, th_o, ph_o);
        }
        if (leave) {
          resto = false;
          mu = saved.mu; tau = saved.tau; theta_max = saved.theta_max; theta_min = saved.theta_min;
          filt = saved.filt; acceptable_count = saved.acc;
          for (int q = 0; q < md; ++q) { vL[q] = hasL[q] ? o.bound_mult_init_val : 0.0; vU[q] = hasU[q] ? o.bound_mult_init_val : 0.0; }
          lsq_multipliers();      // recomputes gf = df grad f and the multipliers at the restored point
          fval = fo;
        } else if (resto_it >= o.resto_max_iter) {
          status = RESTORATION_NEEDED;
          break;
        }
      }
      bool jfinite = true;
      for (int r = 0; r < m && jfinite; ++r)
        for (int k = 0; k < J[r].n; ++k) if (!std::isfinite(J[r].val[k])) { jfinite = false; break; }
      if (!(all_finite(gf) && jfinite && std::isfinite(fval))) { status = INVALID_NUMBER; break; }
      // ---- convergence tests ------------------------------------------------------------------
      const Err e0 = errors(0.0, gf, c, d, s, yc, yd, vL, vU);
      nlp.g(x.data(), gg.data());
      double viol = 0.0;
      for (int r = 0; r < m; ++r) viol = std::max(viol, std::max(nlp.gl[r] - gg[r], gg[r] - nlp.gu[r]));
      viol = std::max(viol, 0.0);
      const double dfe = resto ? 1.0 : df;
      if (e0.E <= tol && e0.dual_inf / dfe <= o.dual_inf_tol && (resto || viol <= o.constr_viol_tol) &&
          e0.compl_ / dfe <= o.compl_inf_tol) {
        status = resto ? RESTORATION_NEEDED : SOLVE_SUCCEEDED;
        if (!resto && !polishing && o.polish_tol > 0.0 && o.polish_tol < tol) {
          polishing = true; first_status = status; first_iters = it; tol = o.polish_tol;
          continue;
        }
        break;
      }
      if (!polishing && !resto && e0.E <= o.acceptable_tol && e0.dual_inf / df <= o.acceptable_dual_inf_tol &&
          viol <= o.acceptable_constr_viol_tol && e0.compl_ / df <= o.acceptable_compl_inf_tol) {
        acceptable_count++;
        if (acceptable_count >= o.acceptable_iter) {
          status = SOLVED_TO_ACCEPTABLE_LEVEL;
          if (o.polish_tol > 0.0 && o.polish_tol < tol) {
            polishing = true; first_status = status; first_iters = it; tol = o.polish_tol;
            continue;
          }
          break;
        }
      } else {
        acceptable_count = 0;
      }
      int ls_fail_mode = 0;
      if (!resto && e0.pr_inf <= 1e-2 * o.tol) ls_fail_mode = acceptable_count > 0 ? 1 : 2;
      if (it >= o.max_iter) { status = MAXITER_EXCEEDED; break; }

      // ---- barrier parameter update (monotone) ------------------------------------------------
      const double mu_floor = std::max(o.mu_min, std::min(tol, o.compl_inf_tol * dfe) / (o.barrier_tol_factor + 1));
      while (true) {
        const double Emu = errors(mu, gf, c, d, s, yc, yd, vL, vU).E;
        if (Emu <= o.barrier_tol_factor * mu && mu > mu_floor) {
          mu = std::max(mu_floor, std::min(o.mu_linear_decrease_factor * mu, std::pow(mu, o.mu_superlinear_decrease_power)));
          tau = std::max(o.tau_min, 1.0 - mu);
          filt.clear();
        } else {
          break;
        }
      }

      // ---- search direction -------------------------------------------------------------------
      Vec sl(md), su(md), sig_s(md), gs_bar(md);
      for (int q = 0; q < md; ++q) {
        sl[q] = SL(s, q);
        su[q] = SU(s, q);
        sig_s[q] = (hasL[q] ? vL[q] / sl[q] : 0.0) + (hasU[q] ? vU[q] / su[q] : 0.0);
        gs_bar[q] = -(hasL[q] ? mu / sl[q] : 0.0) + (hasU[q] ? mu / su[q] : 0.0) +
                    kd * mu * ((hasL[q] && !hasU[q] ? 1.0 : 0.0) - (hasU[q] && !hasL[q] ? 1.0 : 0.0));
      }
      lam_of(yc, yd, lam);
      Vec rx = gf, rs(md), hdiag;
      JcT(yc, rx);
      if (resto) {
        nlp.hess(x.data(), 0.0, lam.data(), Htrip);
        hdiag.resize(n);
        for (int i = 0; i < n; ++i) hdiag[i] = zeta * DR2[i];
      } else {
        nlp.hess(x.data(), df, lam.data(), Htrip);
        JdT(yd, rx);
        for (int q = 0; q < md; ++q) rs[q] = gs_bar[q] - yd[q];
      }
      Vec kap(md, 1.0), Dt(md);
      auto factor_with = [&](double dw_) {
        Vec D(md);
        for (int q = 0; q < md; ++q) {
          Dt[q] = sig_s[q] + dw_;
          if (resto) { kap[q] = wq / (wq + Dt[q]); D[q] = kap[q] * Dt[q]; }
          else D[q] = Dt[q];
        }
        return kkt_factor(&Htrip, dw_, resto ? &hdiag : nullptr, D, n_fact);
      };
      double dw = 0.0;
      bool okf = factor_with(dw), first_try = true, failed = false;
      while (!okf) {
        if (first_try) {
          dw = (dw_last == 0.0) ? o.delta_w_init : std::max(o.delta_w_min, dw_last * o.kappa_w_minus);
          first_try = false;
        } else {
          dw = (dw_last == 0.0 || 1e5 * dw_last < dw) ? dw * o.kappa_w_plus_first : dw * o.kappa_w_plus;
        }
        if (dw > o.delta_w_max) { failed = true; break; }
        okf = factor_with(dw);
      }
      if (failed) { status = ERROR_IN_STEP_COMPUTATION; break; }
      if (dw > 0.0) dw_last = dw;
      Vec dx, ds, dyc, dyd;
      auto resto_solve = [&](const Vec& c_rhs, const Vec& d_rhs, Vec& dx_, Vec& ds_, Vec& dyc_, Vec& dyd_) {
        Vec nrx(n), ne(md), nc(mc), z(md, 0.0), dsx, dydx;
        for (int i = 0; i < n; ++i) nrx[i] = -rx[i];
        for (int q = 0; q < md; ++q) ne[q] = -kap[q] * (Dt[q] * d_rhs[q] + gs_bar[q]);
        for (int q = 0; q < mc; ++q) nc[q] = -c_rhs[q];
        kkt_solve(nrx, ne, nc, z, dx_, dsx, dyc_, dydx);
        ds_.resize(md);
        dyd_.assign(md, 0.0);
        for (int q = 0; q < md; ++q) ds_[q] = kap[q] * (Jd_row(q, dx_) + d_rhs[q]) - gs_bar[q] / (wq + Dt[q]);
      };
      auto plain_solve = [&](const Vec& c_rhs, const Vec& d_rhs, Vec& dx_, Vec& ds_, Vec& dyc_, Vec& dyd_) {
        Vec nrx(n), nrs(md), nc(mc), nd(md);
        for (int i = 0; i < n; ++i) nrx[i] = -rx[i];
        for (int q = 0; q < md; ++q) { nrs[q] = -rs[q]; nd[q] = -d_rhs[q]; }
        for (int q = 0; q < mc; ++q) nc[q] = -c_rhs[q];
        kkt_solve(nrx, nrs, nc, nd, dx_, ds_, dyc_, dyd_);
      };
      Vec dms(md);
      for (int q = 0; q < md; ++q) dms[q] = d[q] - s[q];
      if (resto) resto_solve(c, dms, dx, ds, dyc, dyd);
      else plain_solve(c, dms, dx, ds, dyc, dyd);
      Vec dvL(md), dvU(md);
      auto dual_steps = [&](const Vec& dss) {
        for (int q = 0; q < md; ++q) {
          dvL[q] = hasL[q] ? mu / sl[q] - vL[q] - vL[q] / sl[q] * dss[q] : 0.0;
          dvU[q] = hasU[q] ? mu / su[q] - vU[q] + vU[q] / su[q] * dss[q] : 0.0;
        }
      };
      dual_steps(ds);
      // ---- fraction to the boundary -------------------------------------------------------------
      auto max_step_prim = [&](const Vec& dss) {
        double a = 1.0;
        for (int q = 0; q < md; ++q) {
          if (hasL[q] && dss[q] < 0) a = std::min(a, -tau * sl[q] / dss[q]);
          if (hasU[q] && -dss[q] < 0) a = std::min(a, -tau * su[q] / (-dss[q]));
        }
        return a;
      };
      auto max_step_dual = [&]() {
        double a = 1.0;
        for (int q = 0; q < md; ++q) {
          if (hasL[q] && dvL[q] < 0) a = std::min(a, -tau * vL[q] / dvL[q]);
          if (hasU[q] && dvU[q] < 0) a = std::min(a, -tau * vU[q] / dvU[q]);
        }
        return a;
      };
      const double alpha_max = max_step_prim(ds);
      double alpha_z = max_step_dual();

      // ---- filter line search -------------------------------------------------------------------
      const double theta = theta_of(c, d, s, resto);
      const double phi = barrier_obj(fval, s, mu);
      double gphi_d = 0.0;
      for (int i = 0; i < n; ++i) gphi_d += gf[i] * dx[i];
      for (int q = 0; q < md; ++q) gphi_d += gs_bar[q] * ds[q];
      if (resto)
        for (int q = 0; q < md; ++q) gphi_d += wq * (d[q] - s[q]) * (Jd_row(q, dx) - ds[q]);
      double amin;
      if (gphi_d < 0) {
        amin = std::min(o.gamma_theta, o.gamma_phi * theta / (-gphi_d));
        if (theta <= theta_min) amin = std::min(amin, o.delta * std::pow(theta, o.s_theta) / std::pow(-gphi_d, o.s_phi));
      } else {
        amin = o.gamma_theta;
      }
      amin *= o.alpha_min_frac;

      double alpha = alpha_max;
      bool accepted = false, f_type = false, soc_done = false;
      int n_ls = 0;
      Vec dx_c = dx, ds_c = ds, dyc_c = dyc, dyd_c = dyd;
      Vec xt(n), st(md), ct(mc), dt(md);
      double ft = 0.0;
      auto trial = [&](const Vec& dxx, const Vec& dss, double a) {
        for (int i = 0; i < n; ++i) xt[i] = x[i] + a * dxx[i];
        for (int q = 0; q < md; ++q) st[q] = s[q] + a * dss[q];
        evals(xt, ft, ct, dt);
        if (resto) ft = f_resto(xt, dt, st);
      };
      auto accept_test = [&](double tht, double pht) {
        const bool switching = gphi_d < 0 && alpha * std::pow(-gphi_d, o.s_phi) > o.delta * std::pow(theta, o.s_theta);
        bool ok;
        if (theta <= theta_min && switching) {
          ok = le(pht - phi, o.eta_phi * alpha * gphi_d, phi);
          f_type = true;
        } else {
          ok = le(tht, (1 - o.gamma_theta) * theta, theta) || le(pht - phi, -o.gamma_phi * theta, phi);
          f_type = false;
        }
        return ok;
      };
      while (alpha >= amin || n_ls == 0) {
        trial(dx_c, ds_c, alpha);
        n_ls++;
        bool ok = std::isfinite(ft) && all_finite(ct) && all_finite(dt);
        double tht = 0.0, pht = 0.0;
        if (ok) {
          tht = theta_of(ct, dt, st, resto);
          pht = barrier_obj(ft, st, mu);
          ok = std::isfinite(pht) && tht <= theta_max;
        }
        if (ok) {
          ok = accept_test(tht, pht);
          if (ok) ok = acceptable_to_filter(filt, tht, pht);
        }
        if (ok) { accepted = true; break; }
        // second-order correction on the first trial step
        if (!soc_done && n_ls == 1 && std::isfinite(ft) && all_finite(ct) && theta_of(ct, dt, st, resto) >= theta) {
          soc_done = true;
          Vec c_soc(mc), d_soc(md);
          for (int q = 0; q < mc; ++q) c_soc[q] = alpha * c[q] + ct[q];
          for (int q = 0; q < md; ++q) d_soc[q] = alpha * (d[q] - s[q]) + (dt[q] - st[q]);
          double th_old = theta;
          for (int k = 0; k < o.max_soc; ++k) {
            Vec dxs, dss, dycs, dyds;
            if (resto) resto_solve(c_soc, d_soc, dxs, dss, dycs, dyds);
            else plain_solve(c_soc, d_soc, dxs, dss, dycs, dyds);
            const double a_soc = max_step_prim(dss);
            trial(dxs, dss, a_soc);
            if (!(std::isfinite(ft) && all_finite(ct))) break;
            const double tht2 = theta_of(ct, dt, st, resto);
            const double pht2 = barrier_obj(ft, st, mu);
            if (!std::isfinite(pht2) || tht2 > theta_max) break;
            bool ok2 = accept_test(tht2, pht2);
            if (ok2 && acceptable_to_filter(filt, tht2, pht2)) {
              accepted = true;
              dx_c = dxs; ds_c = dss; dyc_c = dycs; dyd_c = dyds;
              alpha = a_soc;
              dual_steps(dss);
              alpha_z = max_step_dual();
              break;
            }
            if (tht2 > o.kappa_soc * th_old) break;
            th_old = tht2;
            for (int q = 0; q < mc; ++q) c_soc[q] = a_soc * c_soc[q] + ct[q];
            for (int q = 0; q < md; ++q) d_soc[q] = a_soc * d_soc[q] + (dt[q] - st[q]);
          }
          if (accepted) break;
        }
        alpha *= o.alpha_red_factor;
      }
      if (!accepted && ls_fail_mode == 1) { status = SOLVED_TO_ACCEPTABLE_LEVEL; break; }
      if (!accepted && ls_fail_mode == 2) { status = RESTORATION_NEEDED; break; }
      if (!accepted && o.restoration && !resto) {
        // ---- enter the restoration phase at the current iterate (x_R, s_R) ------------------------
        theta_R = theta;
        filt.push_back(FE((1 - o.gamma_theta) * theta, phi - o.gamma_phi * theta));
        saved.mu = mu; saved.tau = tau; saved.theta_max = theta_max; saved.theta_min = theta_min;
        saved.filt = filt; saved.acc = acceptable_count;
        x_R = x;
        DR2.resize(n);
        for (int i = 0; i < n; ++i) { const double t = 1.0 / std::max(1.0, std::fabs(x_R[i])); DR2[i] = t * t; }
        double vmax = maxabs(c);
        for (int q = 0; q < md; ++q) vmax = std::max(vmax, std::fabs(d[q] - s[q]));
        for (int q = 0; q < md; ++q) {
          if (hasL[q]) s[q] = std::max(s[q], dL[q] + pL[q]);
          if (hasU[q]) s[q] = std::min(s[q], dU[q] - pU[q]);
          vL[q] = hasL[q] ? o.bound_mult_init_val : 0.0;
          vU[q] = hasU[q] ? o.bound_mult_init_val : 0.0;
          yd[q] = wq * (d[q] - s[q]);
        }
        std::fill(yc.begin(), yc.end(), 0.0);
        mu = std::max(mu, vmax);
        tau = std::max(o.tau_min, 1.0 - mu);
        zeta = std::sqrt(mu);
        filt.clear();
        acceptable_count = 0;
        resto = true;
        resto_it = 0;
        fval = f_resto(x, d, s);
        const double th0 = theta_of(c, d, s, true);
        theta_max = o.theta_max_fact * std::max(1.0, th0);
        theta_min = o.theta_min_fact * std::max(1.0, th0);
        continue;
      }
      if (!accepted) {
        status = RESTORATION_NEEDED;
        break;
      }
      if (!f_type) filt.push_back(FE((1 - o.gamma_theta) * theta, phi - o.gamma_phi * theta));

      // ---- accept the trial point ---------------------------------------------------------------
      x = xt; s = st;
      fval = ft; c = ct; d = dt;
      for (int q = 0; q < mc; ++q) yc[q] += alpha * dyc_c[q];
      for (int q = 0; q < md; ++q) yd[q] = resto ? wq * (d[q] - s[q]) : yd[q] + alpha * dyd_c[q];
      if (resto) resto_it++;
      const double ks = o.kappa_sigma;
      for (int q = 0; q < md; ++q) {
        vL[q] += alpha_z * dvL[q];
        vU[q] += alpha_z * dvU[q];
        const double sl2 = SL(s, q), su2 = SU(s, q);
        vL[q] = hasL[q] ? std::max(std::min(vL[q], ks * mu / sl2), mu / (ks * sl2)) : 0.0;
        vU[q] = hasU[q] ? std::max(std::min(vU[q], ks * mu / su2), mu / (ks * su2)) : 0.0;
      }
      it++;
    }
    res.x = x;
    res.lam_g.assign(m, 0.0);
    for (int q = 0; q < mc; ++q) res.lam_g[ic[q]] = yc[q] * dc_s[q] / df;
    for (int q = 0; q < md; ++q) res.lam_g[idn[q]] = yd[q] * dd_s[q] / df;
    res.obj = nlp.f(x.data());
    if (polishing) {       // the reported outcome is the first termination's; the iterate is the polished one
      res.status = first_status;
      res.iters = first_iters;
      res.polish_iters = it - first_iters;
    } else {
      res.status = status;
      res.iters = it;
    }
    res.n_fact = n_fact;
    return res;
  }
};

}  // namespace

// =====================================================================================
// C ABI (ctypes): batch of problems, OpenMP over problems
// =====================================================================================
extern "C" {

struct orc_options {
  int32_t N, model, wrap, lane, integrator, max_iter, restoration, linsol, threads, no_refine;
  double Ts, tol, acceptable_tol, polish_tol, mu_init;
  // control/ControllerParameters.py (NaN = default)
  double lambda_s, alpha_L, min_steer, max_steer, min_throttle, max_steer_delta, min_steer_delta,
      max_throttle_delta, min_throttle_delta, q_v_max, v_max, min_s_delta, alpha_c, d_max, q_v_y, n_exp, beta_delta;
};

int orc_version(void) { return 2; }

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_options_default(orc_options* o) {
  std::memset(o, 0, sizeof(*o));
  Params p;
  o->N = 30; o->model = 0; o->wrap = 1; o->lane = 0; o->integrator = 0; o->max_iter = 500; o->restoration = 0;
  o->linsol = 0; o->threads = 0;
  o->Ts = 0.05; o->tol = 1e-4; o->acceptable_tol = 1e-2; o->polish_tol = 0.0; o->mu_init = 0.1;
  o->lambda_s = p.lambda_s; o->alpha_L = p.alpha_L; o->min_steer = p.min_steer; o->max_steer = p.max_steer;
  o->min_throttle = p.min_throttle; o->max_steer_delta = p.max_steer_delta; o->min_steer_delta = p.min_steer_delta;
  o->max_throttle_delta = p.max_throttle_delta; o->min_throttle_delta = p.min_throttle_delta;
  o->q_v_max = p.q_v_max; o->v_max = p.v_max; o->min_s_delta = p.min_s_delta; o->alpha_c = p.alpha_c;
  o->d_max = p.d_max; o->q_v_y = p.q_v_y; o->n_exp = p.n_exp; o->beta_delta = p.beta_delta;
}

// Arrays follow include/mpcr.h's mpcr_solve_batch (row-major [B, ...], NaN u_prev = None).
// lam_g has stride 13N+9 (+N with the lane row), rows of problems without u_prev zero-padded at the end.
int orc_solve_batch(const orc_options* opt, int B, const double* x0, const double* u_prev, const double* s0,
                    const double* cx, const double* cy, const double* max_err, const double* params,
                    const double* u_warm, double* U, double* X, double* S, double* lam_g, double* obj,
                    int32_t* status, int32_t* iters, int32_t* polish_iters) {
  if (!opt || B < 0 || opt->N < 2) return -1;
  const int N = opt->N;
  const int ls = 7 + N * (7 + (opt->lane ? 1 : 0)) + 6 * N + 2;
  int nthreads = opt->threads;
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
  for (int b = 0; b < B; ++b) {
    Problem pr;
    pr.N = N; pr.model = opt->model; pr.wrap = opt->wrap; pr.lane = opt->lane; pr.integrator = opt->integrator;
    pr.Ts = opt->Ts;
    Params& p = pr.p;
    p.lambda_s = opt->lambda_s; p.alpha_L = opt->alpha_L; p.min_steer = opt->min_steer; p.max_steer = opt->max_steer;
    p.min_throttle = opt->min_throttle; p.max_steer_delta = opt->max_steer_delta; p.min_steer_delta = opt->min_steer_delta;
    p.max_throttle_delta = opt->max_throttle_delta; p.min_throttle_delta = opt->min_throttle_delta;
    p.q_v_max = opt->q_v_max; p.v_max = opt->v_max; p.min_s_delta = opt->min_s_delta; p.alpha_c = opt->alpha_c;
    p.d_max = opt->d_max; p.q_v_y = opt->q_v_y; p.n_exp = (int)(opt->n_exp + 0.5); p.beta_delta = opt->beta_delta;
    if (params) {
      const double* q = params + (size_t)b * 5;
      p.alpha_c = q[0]; p.d_max = q[1]; p.q_v_y = q[2]; p.n_exp = (int)(q[3] + 0.5); p.beta_delta = q[4];
    }
    for (int k = 0; k < 6; ++k) pr.x0[k] = x0[(size_t)b * 6 + k];
    pr.u_prev[0] = u_prev[(size_t)b * 2];
    pr.u_prev[1] = u_prev[(size_t)b * 2 + 1];
    pr.has_thr = pr.u_prev[0] == pr.u_prev[0];
    pr.has_st = pr.u_prev[1] == pr.u_prev[1];
    pr.s0 = s0[b];
    for (int k = 0; k < 5; ++k) { pr.cx[k] = cx[(size_t)b * 5 + k]; pr.cy[k] = cy[(size_t)b * 5 + k]; }
    pr.max_err = max_err ? max_err[b] : 0.0;
    if (u_warm) pr.u_warm.assign(u_warm + (size_t)b * 2 * N, u_warm + (size_t)(b + 1) * 2 * N);
    // control/MPC.py:123: without last_controls the guess repeats (state0.throttle, state0.steer); a None there would
    // fail in the reference, the C ABI of the product uses 0 (mpcr_ipm.cuh solve_problem) - same here
    if (!pr.has_thr) pr.u_prev[0] = 0.0;
    if (!pr.has_st) pr.u_prev[1] = 0.0;
    pr.build();
    Options o;
    o.max_iter = opt->max_iter; o.tol = opt->tol; o.acceptable_tol = opt->acceptable_tol;
    o.restoration = opt->restoration; o.polish_tol = opt->polish_tol; o.linsol = opt->linsol; o.mu_init = opt->mu_init; o.refine = opt->no_refine ? 0 : 1;
    Ipm ipm(pr, o);
    Result r = ipm.solve();
    for (int i = 0; i < N; ++i) {
      U[(size_t)b * 2 * N + i] = r.x[pr.iU(0, i)];
      U[(size_t)b * 2 * N + N + i] = r.x[pr.iU(1, i)];
    }
    for (int i = 0; i <= N; ++i) {
      S[(size_t)b * (N + 1) + i] = r.x[pr.iS(i)];
      for (int k = 0; k < 6; ++k) X[(size_t)b * 6 * (N + 1) + (size_t)k * (N + 1) + i] = r.x[pr.iX(k, i)];
    }
    if (lam_g) {
      for (int k = 0; k < ls; ++k) lam_g[(size_t)b * ls + k] = k < pr.m ? r.lam_g[k] : 0.0;
    }
    obj[b] = r.obj;
    status[b] = r.status;
    iters[b] = r.iters;
    if (polish_iters) polish_iters[b] = r.polish_iters;
  }
  return 0;
}

// model step + AD derivatives, for cross-checking the hand-derived ones (tests): f[6], jac[6][8], hess[8][8] (lam-weighted)
int orc_model_eval(int model, int integrator, const double* x, const double* u, double Ts, const double* lam,
                   double* f, double* jac, double* hess) {
  typedef HD<8> T;
  T xx[6], uu[2], F[6];
  for (int k = 0; k < 6; ++k) xx[k] = T::var(x[k], k);
  uu[0] = T::var(u[0], 6);
  uu[1] = T::var(u[1], 7);
  f_step<T>(model, integrator, xx, uu, Ts, F);
  for (int k = 0; k < 6; ++k) {
    f[k] = F[k].v;
    if (jac) for (int a = 0; a < 8; ++a) jac[k * 8 + a] = F[k].g[a];
  }
  if (hess && lam)
    for (int a = 0; a < 8; ++a)
      for (int b = 0; b < 8; ++b) {
        double acc = 0.0;
        for (int k = 0; k < 6; ++k) acc += lam[k] * F[k].hess(a, b);
        hess[a * 8 + b] = acc;
      }
  return 0;
}

// dense Bunch-Kaufman LDL^T of a symmetric matrix (unit test of the factorisation): x = A^-1 b, inertia (pos, neg, zero)
int orc_bk_solve(int n, const double* A, const double* b, double* x, int32_t* inertia) {
  BK bk;
  bk.factor(A, n);
  for (int i = 0; i < n; ++i) x[i] = b[i];
  bk.solve(x);
  inertia[0] = bk.npos; inertia[1] = bk.nneg; inertia[2] = bk.nzero;
  return 0;
}

}  // extern "C"
