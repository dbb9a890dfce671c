// Vehicle models of the NLP (control/MPC.py:186-283) and the stage cost
// (control/MPC.py:71-98) with hand-derived first and second derivatives,
// as fp64 device functions.  Replaces CasADi's SX expression graph + AD.
//
// State x = (X, Y, yaw, v_x, v_y, r), input u = (throttle, steer).
#pragma once
#include <math.h>
#include "mpcr_types.h"

namespace mpcr {

// Double-precision libm expansions are 100-250 SASS instructions each; on the device they are
// wrapped in non-inlined functions so that the solve kernel's hot loop stays inside the SM's
// instruction cache (profiles/r1a: the first version was instruction-fetch bound).
#if defined(__CUDACC__)
#define MPCR_MATH static __host__ __device__ __noinline__
#else
#define MPCR_MATH inline
#endif
MPCR_MATH void sincos_d(double a, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  sincos(a, s, c);
#else
  *s = sin(a);
  *c = cos(a);
#endif
}
MPCR_MATH double atan2_d(double y, double x) { return atan2(y, x); }
MPCR_MATH double exp_d(double x) { return exp(x); }
MPCR_MATH double log_d(double x) { return log(x); }
MPCR_MATH double rcp_d(double x) { return 1.0 / x; }

// models/VehicleParameters.py:8-41 (pi ~ 3.14 exactly as the reference writes it)
namespace vp {
constexpr double m = 1845.0;
constexpr double max_steer = 70.0;
constexpr double T_max = 743.0;
constexpr double r_wheel = 0.37;
constexpr double C_wheel = 2 * 3.14 * r_wheel;
constexpr double R = 9.0;
constexpr double rho = 1.225;
constexpr double C_d = 0.23;
constexpr double A_f = 2.2;
constexpr double C_roll = 0.012;
constexpr double Iz = 3960.0;
constexpr double lf = 0.8;
constexpr double lr = 2.0;
constexpr double Cf = 65000.0;
constexpr double Cr = 65000.0;
constexpr double g = 9.81;
constexpr double Vblendmin = 2.0;
constexpr double Vblendmax = 15.0;
// control/MPC.py:273-283
constexpr double eta1 = -0.00004428225806;
constexpr double eta0 = 1.282413306;
// control/MPC.py:262-271
constexpr double gain1 = -0.001971664699;
constexpr double gain0 = 0.986547;
// derived
constexpr double kr = 60.0 * R * 4.5 / C_wheel;      // rpm per (m/s)
constexpr double Gf = T_max * R / r_wheel;           // wheel force per unit throttle*eta
constexpr double kd = 0.5 * rho * C_d * A_f;
constexpr double Froll = C_roll * m * g;
constexpr double cdel = max_steer * 2 * 3.14 / 360.0;  // control/util.py:10-11
// DERIVATIVE code multiplies by these instead of dividing by the constant (a fp64 division is ~20
// instructions; 23 of them sat in the once-per-iteration code of the dynamic model).  The VALUE
// of the model keeps the reference's divisions.
constexpr double inv_m = 1.0 / m;
constexpr double inv_Iz = 1.0 / Iz;
constexpr double g1 = 3.6 * gain1;
}  // namespace vp

// Non-trivial entries of d f / d(x, u).  Rows X+, Y+ are identity in (X, Y);
// row yaw+ is (yaw: 1, r: Ts); rows (v_x+, v_y+, r+) depend on (v_x, v_y, r, u).
struct StageLin {
  double aX[3];     // dX+/d(yaw, v_x, v_y)
  double aY[3];     // dY+/d(yaw, v_x, v_y)
  double d[3][3];   // d(v_x+, v_y+, r+)/d(v_x, v_y, r)
  double bu[3][2];  // d(v_x+, v_y+, r+)/d(thr, steer)
};
constexpr int kLinDoubles = 21;

// Packed symmetric 6x6 over (yaw, v_x, v_y, r, thr, steer): sum_k lam_k d2 f_k.
constexpr int kHmDoubles = 21;
MPCR_HD int hm_idx(int i, int j) {  // upper triangle, row-major
  if (i > j) { int t = i; i = j; j = t; }
  return i * 6 - (i * (i - 1)) / 2 + (j - i);
}

// x / y for a compile-time constant y: q0 = x * RN(1/y), one FMA residual, one FMA correction
// (Markstein).  With RN(1/y) correctly rounded this IS the correctly rounded quotient - the same
// value `x / y` gives (checked bit for bit against the division on the solver fixtures) - in 3
// instructions and without the division's special-case branch.  Used for the model VALUE, which
// has to keep the reference's arithmetic.
template <typename T = void>
MPCR_HD double div_const(double x, double y, double inv_y) {
  const double q = x * inv_y;
  const double r = fma(-q, y, x);
  return fma(r, inv_y, q);
}

MPCR_HD double Fx_force(double thr, double v_x) {
  // control/MPC.py:273-283, same operation order as the reference
  double wheel_rpm = div_const(v_x, vp::C_wheel, 1.0 / vp::C_wheel) * 60;
  double rpm = wheel_rpm * vp::R * 4.5;
  double eta = vp::eta1 * rpm + vp::eta0;
  double wheel_force = div_const(thr * eta * vp::T_max * vp::R, vp::r_wheel, 1.0 / vp::r_wheel);
  double drag_force = 0.5 * vp::rho * vp::C_d * vp::A_f * (v_x * v_x);
  double rolling = vp::C_roll * vp::m * vp::g;
  return wheel_force - drag_force - rolling;
}

MPCR_HD double steer_angle(double st, double v_x, double v_y) {
  // control/MPC.py:262-271 + control/util.py:10-11
  double vel = sqrt(v_x * v_x + v_y * v_y) * 3.6;
  double gain = vp::gain1 * vel + vp::gain0;
  return div_const(st * gain * vp::max_steer, 360.0, 1.0 / 360.0) * 2 * 3.14;
}

MPCR_HD void f_dynamic(const double* x, const double* u, double Ts, double* f) {
  // control/MPC.py:186-229
  double yaw = x[2], v_x = x[3], v_y = x[4], r = x[5];
  double fx = Fx_force(u[0], v_x);
  double delta = steer_angle(u[1], v_x, v_y);
  double thf = atan2_d(v_y + vp::lf * r, v_x + 0.1);
  double thr = atan2_d(v_y - vp::lr * r, v_x + 0.1);
  double Fyf = vp::Cf * (delta - thf);
  double Fyr = vp::Cr * (-thr);
  double sd, cd;
  sincos_d(delta, &sd, &cd);
  double vxd = div_const(fx - Fyf * sd, vp::m, vp::inv_m) + (v_y * r);
  double vyd = div_const(Fyf * cd + Fyr, vp::m, vp::inv_m) - (v_x * r);
  double rd = div_const((Fyf * cd * vp::lf) - (Fyr * vp::lr), vp::Iz, vp::inv_Iz);
  double sy, cy;
  sincos_d(yaw, &sy, &cy);
  f[0] = x[0] + (v_x * cy - v_y * sy) * Ts;
  f[1] = x[1] + (v_x * sy + v_y * cy) * Ts;
  f[2] = yaw + r * Ts;
  f[3] = v_x + vxd * Ts;
  f[4] = v_y + vyd * Ts;
  f[5] = r + rd * Ts;
}

MPCR_HD void f_kinematic(const double* x, const double* u, double Ts, double* f) {
  // control/MPC.py:231-260
  double yaw = x[2], v_x = x[3], v_y = x[4], r = x[5];
  double fx = Fx_force(u[0], v_x);
  double delta = steer_angle(u[1], v_x, v_y);
  double sy, cy;
  sincos_d(yaw, &sy, &cy);
  f[0] = x[0] + (v_x * cy - v_y * sy) * Ts;
  f[1] = x[1] + (v_x * sy + v_y * cy) * Ts;
  f[2] = yaw + r * Ts;
  f[3] = v_x + (fx / vp::m) * Ts;
  f[4] = r * vp::lr;
  f[5] = (v_x / (vp::lr + vp::lf)) * tan(delta);
}

MPCR_HD void f_model(int model, const double* x, const double* u, double Ts, double* f) {
  if (model == kDynamic) {
    f_dynamic(x, u, Ts, f);
  } else if (model == kKinematic) {
    f_kinematic(x, u, Ts, f);
  } else {
    // models/BlendedBicycleModel.py:23-26,46 applied to the two NLP models
    double fd[6], fk[6];
    f_dynamic(x, u, Ts, fd);
    f_kinematic(x, u, Ts, fk);
    double vel = sqrt(x[3] * x[3] + x[4] * x[4]);
    double lam = (vel - vp::Vblendmin) / (vp::Vblendmax - vp::Vblendmin);
    lam = fmin(fmax(lam, 0.0), 1.0);
    for (int k = 0; k < 6; ++k) f[k] = lam * fd[k] + (1 - lam) * fk[k];
  }
}

// ---------------------------------------------------------------------------
// Derivatives.  Local variable order of the 5-block: q = (v_x, v_y, r, thr, st).
// ---------------------------------------------------------------------------
struct Sym5 {            // symmetric 5x5, full storage (registers; unrolled)
  double a[5][5];
};

MPCR_HD void sym5_zero(Sym5& h) {
  for (int i = 0; i < 5; ++i)
    for (int j = 0; j < 5; ++j) h.a[i][j] = 0.0;
}

struct Common {
  double fx, gfx_vx, gfx_th, hfx_vxvx, hfx_vxth;        // F_x and derivatives
  double delta, gd_vx, gd_vy, gd_st;                    // steer angle
  double hd_vxvx, hd_vyvy, hd_vxvy, hd_vxst, hd_vyst;
};

MPCR_HD void common_eval(const double* x, const double* u, Common& c) {
  double v_x = x[3], v_y = x[4];
  double thr = u[0], st = u[1];
  double eta = vp::eta1 * vp::kr * v_x + vp::eta0;
  c.fx = thr * eta * vp::Gf - vp::kd * v_x * v_x - vp::Froll;
  c.gfx_th = eta * vp::Gf;
  c.gfx_vx = thr * vp::Gf * vp::eta1 * vp::kr - 2 * vp::kd * v_x;
  c.hfx_vxvx = -2 * vp::kd;
  c.hfx_vxth = vp::Gf * vp::eta1 * vp::kr;
  double V = sqrt(v_x * v_x + v_y * v_y);
  double iV = 1.0 / V;
  double gain = vp::gain0 + vp::g1 * V;
  c.delta = vp::cdel * st * gain;
  double k = vp::cdel * vp::g1;
  c.gd_st = vp::cdel * gain;
  c.gd_vx = k * st * v_x * iV;
  c.gd_vy = k * st * v_y * iV;
  c.hd_vxst = k * v_x * iV;
  c.hd_vyst = k * v_y * iV;
  double iV3 = iV * iV * iV;
  c.hd_vxvx = k * st * v_y * v_y * iV3;
  c.hd_vyvy = k * st * v_x * v_x * iV3;
  c.hd_vxvy = -k * st * v_x * v_y * iV3;
}

// position rows: Jacobian entries and lam-weighted Hessian over (yaw, v_x, v_y)
MPCR_HD void position_rows(const double* x, double Ts, const double* lam, StageLin& L,
                           double* hm /* may be null */) {
  double yaw = x[2], v_x = x[3], v_y = x[4];
  double s, c;
  sincos_d(yaw, &s, &c);
  L.aX[0] = (-v_x * s - v_y * c) * Ts;
  L.aX[1] = c * Ts;
  L.aX[2] = -s * Ts;
  L.aY[0] = (v_x * c - v_y * s) * Ts;
  L.aY[1] = s * Ts;
  L.aY[2] = c * Ts;
  if (hm) {
    double lX = lam[0], lY = lam[1];
    hm[hm_idx(0, 0)] += Ts * (lX * (-v_x * c + v_y * s) + lY * (-v_x * s - v_y * c));
    hm[hm_idx(0, 1)] += Ts * (-lX * s + lY * c);
    hm[hm_idx(0, 2)] += Ts * (-lX * c - lY * s);
  }
}

// Value, Jacobian and (if lam != null) lam-weighted Hessian of the dynamic model.
MPCR_HD void lin_dynamic(const double* x, const double* u, double Ts, const double* lam,
                         double* f, StageLin& L, double* hm) {
  Common c;
  common_eval(x, u, c);
  double v_x = x[3], v_y = x[4], r = x[5];
  // slip angles
  double b = v_x + 0.1;
  double af = v_y + vp::lf * r, ar = v_y - vp::lr * r;
  double Df = af * af + b * b, Dr = ar * ar + b * b;
  double iDf = 1.0 / Df, iDr = 1.0 / Dr;
  double thf = atan2_d(af, b), thr = atan2_d(ar, b);
  // gradients over q = (v_x, v_y, r, thr, st)
  double gthf[3] = {-af * iDf, b * iDf, vp::lf * b * iDf};
  double gthr[3] = {-ar * iDr, b * iDr, -vp::lr * b * iDr};
  double sd, cd;
  sincos_d(c.delta, &sd, &cd);
  double Fyf = vp::Cf * (c.delta - thf);
  double Fyr = -vp::Cr * thr;
  double gd[5] = {c.gd_vx, c.gd_vy, 0.0, 0.0, c.gd_st};
  double gFyf[5], gFyr[5];
  for (int k = 0; k < 5; ++k) {
    double gt = (k < 3) ? gthf[k] : 0.0;
    gFyf[k] = vp::Cf * (gd[k] - gt);
    gFyr[k] = (k < 3) ? -vp::Cr * gthr[k] : 0.0;
  }
  double gvx[5], gvy[5], gr[5];
  for (int k = 0; k < 5; ++k) {
    double gA = gFyf[k] * sd + Fyf * cd * gd[k];
    double gB = gFyf[k] * cd - Fyf * sd * gd[k];
    double gfx = (k == 0) ? c.gfx_vx : ((k == 3) ? c.gfx_th : 0.0);
    gvx[k] = (gfx - gA) * vp::inv_m;
    gvy[k] = (gB + gFyr[k]) * vp::inv_m;
    gr[k] = (gB * vp::lf - gFyr[k] * vp::lr) * vp::inv_Iz;
  }
  gvx[1] += r;
  gvx[2] += v_y;
  gvy[0] -= r;
  gvy[2] -= v_x;
  if (f) {
    double vxd = div_const(c.fx - Fyf * sd, vp::m, vp::inv_m) + (v_y * r);
    double vyd = div_const(Fyf * cd + Fyr, vp::m, vp::inv_m) - (v_x * r);
    double rd = div_const((Fyf * cd * vp::lf) - (Fyr * vp::lr), vp::Iz, vp::inv_Iz);
    double sy, cy;
    sincos_d(x[2], &sy, &cy);
    f[0] = x[0] + (v_x * cy - v_y * sy) * Ts;
    f[1] = x[1] + (v_x * sy + v_y * cy) * Ts;
    f[2] = x[2] + r * Ts;
    f[3] = v_x + vxd * Ts;
    f[4] = v_y + vyd * Ts;
    f[5] = r + rd * Ts;
  }
  for (int k = 0; k < 3; ++k) {
    L.d[0][k] = gvx[k] * Ts + (k == 0 ? 1.0 : 0.0);
    L.d[1][k] = gvy[k] * Ts + (k == 1 ? 1.0 : 0.0);
    L.d[2][k] = gr[k] * Ts + (k == 2 ? 1.0 : 0.0);
  }
  for (int k = 0; k < 2; ++k) {
    L.bu[0][k] = gvx[3 + k] * Ts;
    L.bu[1][k] = gvy[3 + k] * Ts;
    L.bu[2][k] = gr[3 + k] * Ts;
  }
  if (hm) {
    for (int k = 0; k < kHmDoubles; ++k) hm[k] = 0.0;
  }
  position_rows(x, Ts, lam, L, hm);
  if (!hm) return;
  // lam-weighted Hessian of the (v_x, v_y, r) rows over q
  double c1 = lam[3] * vp::inv_m;
  double c2 = lam[4] * vp::inv_m + lam[5] * vp::lf * vp::inv_Iz;
  double c3 = lam[4] * vp::inv_m - lam[5] * vp::lr * vp::inv_Iz;
  double w = -c1 * sd + c2 * cd;
  double wp = -c1 * cd - c2 * sd;
  Sym5 H;
  sym5_zero(H);
  // (Cf w + Fyf w') * Hess(delta)
  double kdlt = vp::Cf * w + Fyf * wp;
  H.a[0][0] += kdlt * c.hd_vxvx;
  H.a[1][1] += kdlt * c.hd_vyvy;
  H.a[0][1] += kdlt * c.hd_vxvy;
  H.a[0][4] += kdlt * c.hd_vxst;
  H.a[1][4] += kdlt * c.hd_vyst;
  // atan2 Hessians over (v_x, v_y, r): ga = (0, 1, l), gb = (1, 0, 0)
  {
    double kf = -vp::Cf * w;           // multiplies Hess(theta_f)
    double haa = -2 * af * b * iDf * iDf, hbb = -haa, hab = (af * af - b * b) * iDf * iDf;
    double l = vp::lf;
    H.a[0][0] += kf * hbb;
    H.a[0][1] += kf * hab;
    H.a[0][2] += kf * hab * l;
    H.a[1][1] += kf * haa;
    H.a[1][2] += kf * haa * l;
    H.a[2][2] += kf * haa * l * l;
  }
  {
    double kf = -c3 * vp::Cr;          // multiplies Hess(theta_r)
    double haa = -2 * ar * b * iDr * iDr, hbb = -haa, hab = (ar * ar - b * b) * iDr * iDr;
    double l = -vp::lr;
    H.a[0][0] += kf * hbb;
    H.a[0][1] += kf * hab;
    H.a[0][2] += kf * hab * l;
    H.a[1][1] += kf * haa;
    H.a[1][2] += kf * haa * l;
    H.a[2][2] += kf * haa * l * l;
  }
  // w' sym(gFyf, gd) - Fyf w gd gd'
  for (int i = 0; i < 5; ++i)
    for (int j = i; j < 5; ++j)
      H.a[i][j] += wp * (gFyf[i] * gd[j] + gd[i] * gFyf[j]) - Fyf * w * gd[i] * gd[j];
  // c1 Hess(Fx)
  H.a[0][0] += c1 * c.hfx_vxvx;
  H.a[0][3] += c1 * c.hfx_vxth;
  // bilinear terms lam_vx v_y r - lam_vy v_x r
  H.a[1][2] += lam[3];
  H.a[0][2] -= lam[4];
  for (int i = 0; i < 5; ++i)
    for (int j = i; j < 5; ++j) hm[hm_idx(1 + i, 1 + j)] += Ts * H.a[i][j];
}

MPCR_HD void lin_kinematic(const double* x, const double* u, double Ts, const double* lam,
                           double* f, StageLin& L, double* hm) {
  Common c;
  common_eval(x, u, c);
  double v_x = x[3], r = x[5];
  const double Lw = vp::lr + vp::lf;
  double td = tan(c.delta);
  double sec2 = 1.0 + td * td;
  if (f) {
    double sy, cy;
    sincos_d(x[2], &sy, &cy);
    f[0] = x[0] + (v_x * cy - x[4] * sy) * Ts;
    f[1] = x[1] + (v_x * sy + x[4] * cy) * Ts;
    f[2] = x[2] + r * Ts;
    f[3] = v_x + (c.fx / vp::m) * Ts;
    f[4] = r * vp::lr;
    f[5] = (v_x / Lw) * td;
  }
  double gd[5] = {c.gd_vx, c.gd_vy, 0.0, 0.0, c.gd_st};
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) L.d[i][k] = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 2; ++k) L.bu[i][k] = 0.0;
  L.d[0][0] = 1.0 + c.gfx_vx * vp::inv_m * Ts;
  L.bu[0][0] = c.gfx_th * vp::inv_m * Ts;
  L.d[1][2] = vp::lr;
  double k1 = v_x / Lw * sec2;
  L.d[2][0] = k1 * gd[0] + td / Lw;
  L.d[2][1] = k1 * gd[1];
  L.bu[2][1] = k1 * gd[4];
  if (hm) {
    for (int k = 0; k < kHmDoubles; ++k) hm[k] = 0.0;
  }
  position_rows(x, Ts, lam, L, hm);
  if (!hm) return;
  Sym5 H;
  sym5_zero(H);
  double a = lam[3] * Ts * vp::inv_m;
  H.a[0][0] += a * c.hfx_vxvx;
  H.a[0][3] += a * c.hfx_vxth;
  double l5 = lam[5];
  // Hess of (v_x / L) tan(delta)
  double k2 = v_x / Lw * 2 * sec2 * td;
  for (int i = 0; i < 5; ++i)
    for (int j = i; j < 5; ++j) {
      double e = k2 * gd[i] * gd[j];
      if (i == 0) e += sec2 / Lw * gd[j];
      if (j == 0) e += sec2 / Lw * gd[i];
      H.a[i][j] += l5 * e;
    }
  H.a[0][0] += l5 * k1 * c.hd_vxvx;
  H.a[1][1] += l5 * k1 * c.hd_vyvy;
  H.a[0][1] += l5 * k1 * c.hd_vxvy;
  H.a[0][4] += l5 * k1 * c.hd_vxst;
  H.a[1][4] += l5 * k1 * c.hd_vyst;
  for (int i = 0; i < 5; ++i)
    for (int j = i; j < 5; ++j) hm[hm_idx(1 + i, 1 + j)] += H.a[i][j];
}

MPCR_HD void lin_blended(const double* x, const double* u, double Ts, const double* lam,
                         double* f, StageLin& L, double* hm) {
  double fd[6], fk[6];
  StageLin Ld, Lk;
  double hd[kHmDoubles], hk[kHmDoubles];
  lin_dynamic(x, u, Ts, lam, fd, Ld, hm ? hd : nullptr);
  lin_kinematic(x, u, Ts, lam, fk, Lk, hm ? hk : nullptr);
  double v_x = x[3], v_y = x[4];
  double V = sqrt(v_x * v_x + v_y * v_y);
  const double span = vp::Vblendmax - vp::Vblendmin;
  double raw = (V - vp::Vblendmin) / span;
  bool inside = raw > 0.0 && raw < 1.0;
  double lm = fmin(fmax(raw, 0.0), 1.0);
  double gl[2] = {inside ? v_x / V / span : 0.0, inside ? v_y / V / span : 0.0};
  if (f)
    for (int k = 0; k < 6; ++k) f[k] = lm * fd[k] + (1 - lm) * fk[k];
  for (int k = 0; k < 3; ++k) {   // position rows are identical in both models
    L.aX[k] = Ld.aX[k];
    L.aY[k] = Ld.aY[k];
  }
  for (int i = 0; i < 3; ++i) {
    double diff = fd[3 + i] - fk[3 + i];
    for (int k = 0; k < 3; ++k)
      L.d[i][k] = lm * Ld.d[i][k] + (1 - lm) * Lk.d[i][k] + (k < 2 ? diff * gl[k] : 0.0);
    for (int k = 0; k < 2; ++k) L.bu[i][k] = lm * Ld.bu[i][k] + (1 - lm) * Lk.bu[i][k];
  }
  if (!hm) return;
  for (int k = 0; k < kHmDoubles; ++k) hm[k] = lm * hd[k] + (1 - lm) * hk[k];
  // + sym(grad lam, grad(lam' (fd - fk))) + (lam' (fd - fk)) Hess(lam)
  double ldiff = 0.0;
  double gdiff[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < 3; ++i) {
    ldiff += lam[3 + i] * (fd[3 + i] - fk[3 + i]);
    for (int k = 0; k < 3; ++k) gdiff[k] += lam[3 + i] * (Ld.d[i][k] - Lk.d[i][k]);
    for (int k = 0; k < 2; ++k) gdiff[3 + k] += lam[3 + i] * (Ld.bu[i][k] - Lk.bu[i][k]);
  }
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 5; ++j) {
      // entry (i, j) of gl gdiff' + gdiff gl' restricted to the upper triangle
      if (j >= i) {
        double e = gl[i] * gdiff[j];
        if (j < 2) e += gdiff[i] * gl[j];
        hm[hm_idx(1 + i, 1 + j)] += e;
      }
    }
  if (inside) {
    double iV3 = 1.0 / (V * V * V);
    hm[hm_idx(1, 1)] += ldiff * v_y * v_y * iV3 / span;
    hm[hm_idx(2, 2)] += ldiff * v_x * v_x * iV3 / span;
    hm[hm_idx(1, 2)] += -ldiff * v_x * v_y * iV3 / span;
  }
}

MPCR_HD void lin_model(int model, const double* x, const double* u, double Ts,
                       const double* lam, double* f, StageLin& L, double* hm) {
  if (model == kDynamic) lin_dynamic(x, u, Ts, lam, f, L, hm);
  else if (model == kKinematic) lin_kinematic(x, u, Ts, lam, f, L, hm);
  else lin_blended(x, u, Ts, lam, f, L, hm);
}

// ---------------------------------------------------------------------------
// Centerline polynomials (control/util.py:4-8: sum_k c[deg-k] s^k, monomial
// evaluation in absolute progress, lowest order first) and the stage cost.
// ---------------------------------------------------------------------------
struct Poly4 {
  double p, d1, d2, d3;
};

MPCR_HD void poly4_eval(const double* c /* highest first, 5 */, double s, int order, Poly4& o) {
  double s2 = s * s, s3 = s2 * s, s4 = s2 * s2;
  o.p = c[4] + c[3] * s + c[2] * s2 + c[1] * s3 + c[0] * s4;
  o.d1 = c[3] + 2 * c[2] * s + 3 * c[1] * s2 + 4 * c[0] * s3;
  if (order >= 2) {
    o.d2 = 2 * c[2] + 6 * c[1] * s + 12 * c[0] * s2;
    o.d3 = 6 * c[1] + 24 * c[0] * s;
  } else {
    o.d2 = 0.0;
    o.d3 = 0.0;
  }
}

MPCR_HD double ipow(double b, int e) {
  double r = 1.0;
  for (int i = 0; i < e; ++i) r *= b;
  return r;
}

struct StageCost {
  double val;
  double eC, eL;
  double g3[3];          // d/d(X, Y, s)
  double gvx, gvy;
  double h3[6];          // packed sym 3x3 over (X, Y, s): 00 01 02 11 12 22
  double hvx, hvy;
  // lane row e_C: gradient / Hessian pieces (only filled at order 2)
  double gC[3];
  double hC02, hC12, hC22;
};

// l(x, s) = q_vy v_y^2 + alpha_c eC^n + alpha_L eL^2 + exp(q_vmax (v_x - v_max))
// (control/MPC.py:86-98).  order 0: value only; 1: + gradient; 2: + Hessian.
MPCR_HD void stage_cost(const Params& p, const double* cx, const double* cy, const double* x,
                        double s, int order, StageCost& o) {
  Poly4 px, py;
  poly4_eval(cx, s, order, px);
  poly4_eval(cy, s, order, py);
  double dX = x[0] - px.p, dY = x[1] - py.p;
  double eC = py.d1 * dX - px.d1 * dY;
  double eL = -px.d1 * dX - py.d1 * dY;
  o.eC = eC;
  o.eL = eL;
  double v_x = x[3], v_y = x[4];
  double ex = exp_d(p.q_v_max * (v_x - p.v_max));
  double eCn1 = ipow(eC, p.n_exp - 1);
  o.val = p.q_v_y * v_y * v_y + p.alpha_c * eCn1 * eC + p.alpha_L * eL * eL + ex;
  if (order < 1) return;
  double gC[3] = {py.d1, -px.d1, py.d2 * dX - px.d2 * dY};
  double gL[3] = {-px.d1, -py.d1, -px.d2 * dX + px.d1 * px.d1 - py.d2 * dY + py.d1 * py.d1};
  double dC = p.alpha_c * p.n_exp * eCn1;
  double dL = p.alpha_L * 2 * eL;
  for (int k = 0; k < 3; ++k) {
    o.g3[k] = dC * gC[k] + dL * gL[k];
    o.gC[k] = gC[k];
  }
  o.gvx = p.q_v_max * ex;
  o.gvy = 2 * p.q_v_y * v_y;
  if (order < 2) return;
  double hC02 = py.d2, hC12 = -px.d2;
  double hC22 = py.d3 * dX - py.d2 * px.d1 - px.d3 * dY + px.d2 * py.d1;
  double hL02 = -px.d2, hL12 = -py.d2;
  double hL22 = -px.d3 * dX - py.d3 * dY + 3 * (px.d1 * px.d2 + py.d1 * py.d2);
  double ddC = (p.n_exp >= 2) ? p.alpha_c * p.n_exp * (p.n_exp - 1) * ipow(eC, p.n_exp - 2) : 0.0;
  double aL2 = 2 * p.alpha_L;
  o.h3[0] = ddC * gC[0] * gC[0] + aL2 * gL[0] * gL[0];
  o.h3[1] = ddC * gC[0] * gC[1] + aL2 * gL[0] * gL[1];
  o.h3[2] = ddC * gC[0] * gC[2] + aL2 * gL[0] * gL[2] + dC * hC02 + dL * hL02;
  o.h3[3] = ddC * gC[1] * gC[1] + aL2 * gL[1] * gL[1];
  o.h3[4] = ddC * gC[1] * gC[2] + aL2 * gL[1] * gL[2] + dC * hC12 + dL * hL12;
  o.h3[5] = ddC * gC[2] * gC[2] + aL2 * gL[2] * gL[2] + dC * hC22 + dL * hL22;
  o.hvx = p.q_v_max * p.q_v_max * ex;
  o.hvy = 2 * p.q_v_y;
  o.hC02 = hC02;
  o.hC12 = hC12;
  o.hC22 = hC22;
}

}  // namespace mpcr
