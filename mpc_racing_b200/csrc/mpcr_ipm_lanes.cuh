// One-problem-per-LANE variant of the interior-point solver of mpcr_ipm.cuh: the
// throughput kernel for large batches (>= a few thousand problems per GPU).
//
// Same algorithm, constants and phase loop as the warp-per-problem solver (and as
// oracle/ipm.py): IPOPT's filter line-search IPM on the NLP of control/MPC.py:57-149,
// KKT systems solved by a stage-wise Riccati recursion + rank-2 Woodbury correction for
// the wrap rows.  What changes is the mapping to the machine:
//
//   * every thread owns ONE problem and walks its N stages sequentially, so all 32
//     lanes of a warp do useful fp64 work in every instruction (the warp-per-problem
//     kernel keeps 9-12 of 32 lanes busy inside the Riccati sweeps) and no shuffles or
//     warp reductions are needed;
//   * the whole per-problem state (iterate, multipliers, slack rows, stage
//     linearisations, Riccati gains: kLaneWsDoubles(N) doubles, ~78 kB at N = 30) lives in a
//     global-memory workspace laid out [element][lane]: lane l of a warp reads element e
//     at base + e * 32 + l, so every access of a warp is one fully coalesced 256-byte
//     transaction.  The working set streams through L2/HBM once per pass; the passes
//     are fused so that an IPM iteration touches each stage record a handful of times;
//   * the 9x9 cost-to-go matrix and the per-stage blocks are register/local resident.
//
// Threads of a warp run their phase loops independently (each has its own phase,
// step length, barrier parameter ...); the hardware serialises divergent phases.
//
// The file compiles for the host as well (stride 1): tests/emu runs the identical code
// on the CPU as test infrastructure - the product path is the CUDA build.
#pragma once
#include "mpcr_ipm.cuh"
#include "mpcr_rk4.cuh"

namespace mpcr {

#if defined(__CUDACC__)
constexpr int kLaneStride = 32;
#else
constexpr int kLaneStride = 1;
#endif

// ---------------------------------------------------------------------------
// workspace layout (element indices; one "element" = one double per lane)
// ---------------------------------------------------------------------------
namespace lw {
// per-stage record (stage k = transition k: input u_k, successor state x_{k+1})
constexpr int X = 0;        // x_{k+1} (6)
constexpr int SH = 6;       // s_hat_{k+1}
constexpr int U = 7;        // u_k (2)
constexpr int LAM = 9;      // multipliers of x_{k+1} = f(x_k, u_k) (6)
constexpr int CRES = 15;    // x_{k+1} - f(x_k, u_k) (6)
constexpr int DCS = 21;     // row scaling of the dynamics rows (6)
constexpr int DX = 27;      // search direction: dx (6), dsh, du (2)
constexpr int NLAM = 36;    // new multipliers of the linearised problem (6)
constexpr int TC = 42;      // trial point: constraint residual (6)
constexpr int RS = 48;      // slack rows [kNumRows] each: s, yd, vL, vU, rd, dsr, trial rd
constexpr int RYD = RS + kNumRows;
constexpr int RVL = RYD + kNumRows;
constexpr int RVU = RVL + kNumRows;
constexpr int RRD = RVU + kNumRows;
constexpr int RDSR = RRD + kNumRows;
constexpr int RTRD = RDSR + kNumRows;
constexpr int RDYD = RTRD + kNumRows;   // dual step of the row multiplier (current direction)
constexpr int A = RDYD + kNumRows;      // stage linearisation: 21 (Euler) / 32 (RK4, see lane_lin_rk4)
constexpr int HSD = A + 32;             // derivative part of the stage Hessian: (X,Y,s) block (6) + model block (21)
constexpr int GC = HSD + 27;            // df * grad cost of state k+1 over (X, Y, s, vx, vy) (5)
constexpr int GCZ = GC + 5;             // ddl * grad e_C of state k+1 over (X, Y, s) (3)
constexpr int SIGC = GCZ + 3;           // steer_k - steer_{k-1} (0 for k = 0)
constexpr int HS = SIGC + 1;            // assembled packed stage Hessian (32)
constexpr int GS = HS + 32;             // assembled stage gradient (12)
constexpr int K = GS + 12;              // feedback gain (3 x 9)
constexpr int KF = K + 27;              // feed-forward of the 3 right-hand sides (3 x 3)
constexpr int Z3 = KF + 9;              // forward-sweep results (3 x 9): dx(6), dsh, dthr, dst
constexpr int CSOC = Z3 + 27;           // second-order-correction right-hand sides
constexpr int DSOC = CSOC + 6;
constexpr int SVD = DSOC + kNumRows;    // saved plain step (9), its slack steps, multipliers, dyd
constexpr int SVDSR = SVD + 9;
constexpr int SVLAM = SVDSR + kNumRows;
constexpr int SVDYD = SVLAM + 6;
constexpr int XR = SVDYD + kNumRows;    // restoration: reference point x_R (x (6), s_hat, u (2))
constexpr int PG = XR + 9;              // restoration: gradient of the proximity term over the same 9 variables
constexpr int REC = PG + 9;
// per-problem tail after the N stage records
constexpr int T_HSD = 0;                // terminal stage: derivative part (27)
constexpr int T_HS = 27;                // assembled (32)
constexpr int T_GS = 59;                // (12)
constexpr int T_FTH = 71;               // filter (32 + 32)
constexpr int T_FPH = 103;
constexpr int T_FTH2 = 135;             // the restoration phase's own filter
constexpr int T_FPH2 = 167;
constexpr int TAIL = 199;
}  // namespace lw

MPCR_HD int lane_ws_doubles(int N) { return N * lw::REC + lw::TAIL; }

struct LaneWs {
  double* p;     // already offset by the lane
  MPCR_HD double& operator()(int e) const { return p[(size_t)e * kLaneStride]; }
  MPCR_HD double& s(int k, int f) const { return p[(size_t)(k * lw::REC + f) * kLaneStride]; }
  MPCR_HD double& t(int N, int f) const { return p[(size_t)(N * lw::REC + f) * kLaneStride]; }
  // Prefetch elements [f, f + n) of stage k into L1 (a warp-wide prefetch covers the element's
  // 256-byte row).  The lane kernel runs ~1 warp per scheduler, so nothing else hides the HBM
  // latency of its workspace: every pass asks for the NEXT stage's inputs before it starts on the
  // current one.
  MPCR_HD void pf(int k, int f, int n) const {
#if defined(__CUDA_ARCH__) && !defined(MPCR_LANE_NO_PREFETCH)
#pragma unroll 4
    for (int i = 0; i < n; ++i)
      asm volatile("prefetch.global.L1 [%0];" ::"l"(p + (size_t)(k * lw::REC + f + i) * kLaneStride));
#else
    (void)k; (void)f; (void)n;
#endif
  }
  MPCR_HD void pf_rows(int k, int first_array, int n_arrays) const {   // rows 0..6 of consecutive row arrays
    for (int a = 0; a < n_arrays; ++a) pf(k, first_array + a * kNumRows, 7);
  }
};

// packed symmetric 9x9 (upper triangle, row-major)
MPCR_HD constexpr int p9(int i, int j) {
  return (i <= j) ? (i * 9 - (i * (i - 1)) / 2 + (j - i)) : (j * 9 - (j * (j - 1)) / 2 + (i - j));
}

// Stage linearisation as the lane solver stores it: entries 0..20 are StageLin (aX[3], aY[3],
// d[3][3], bu[3][2]); with the RK4 integrator the map's Jacobian has 11 more non-trivial entries:
//   21..23  dX+/d(r, thr, st)      24..26  dY+/d(r, thr, st)
//   27..30  dyaw+/d(v_x, v_y, thr, st)      31  dyaw+/dr  (= Ts for Euler)
constexpr int kLinEuler = 21, kLinRk4 = 32;
template <int IG> MPCR_HD constexpr int lin_count() { return IG ? kLinRk4 : kLinEuler; }

// out = M' v with M = [A B] (9 x 12), cf. mulM
template <int IG>
MPCR_D void mulMx(const double* v, const double* A, double Ts, double* out) {
  if (!IG) { mulM(v, A, Ts, out); return; }
  const double *aX = A, *aY = A + 3, *d = A + 6, *bu = A + 15, *xe = A + 21, *ye = A + 24, *aw = A + 27;
  const double tsr = A[31];
  out[0] = v[0];
  out[1] = v[1];
  out[2] = aX[0] * v[0] + aY[0] * v[1] + v[2];
  out[3] = aX[1] * v[0] + aY[1] * v[1] + aw[0] * v[2] + d[0] * v[3] + d[3] * v[4] + d[6] * v[5];
  out[4] = aX[2] * v[0] + aY[2] * v[1] + aw[1] * v[2] + d[1] * v[3] + d[4] * v[4] + d[7] * v[5];
  out[5] = xe[0] * v[0] + ye[0] * v[1] + tsr * v[2] + d[2] * v[3] + d[5] * v[4] + d[8] * v[5];
  out[6] = v[6];
  out[7] = 0.0;
  out[8] = 0.0;
  out[9] = xe[1] * v[0] + ye[1] * v[1] + aw[2] * v[2] + bu[0] * v[3] + bu[2] * v[4] + bu[4] * v[5] + v[7];
  out[10] = xe[2] * v[0] + ye[2] * v[1] + aw[3] * v[2] + bu[1] * v[3] + bu[3] * v[4] + bu[5] * v[5] + v[8];
  out[11] = v[6];
}

// zn = M [z; w], cf. applyM
template <int IG>
MPCR_D void applyMx(const double* z, const double* w, const double* A, double Ts, double* zn) {
  applyM(z, w, A, Ts, zn);
  if (IG) {
    const double *xe = A + 21, *ye = A + 24, *aw = A + 27;
    zn[0] += xe[0] * z[5] + xe[1] * w[0] + xe[2] * w[1];
    zn[1] += ye[0] * z[5] + ye[1] * w[0] + ye[2] * w[1];
    zn[2] = z[2] + aw[0] * z[3] + aw[1] * z[4] + A[31] * z[5] + aw[2] * w[0] + aw[3] * w[1];
  }
}

// t[k] = sum_j P[k][j] M[j][c]  (column c of M = [A B]), P packed symmetric
template <int C, int IG>
MPCR_D void p_times_mcol(const double* P, const double* A, double Ts, double* t) {
  const double *aX = A, *aY = A + 3, *d = A + 6, *bu = A + 15, *xe = A + 21, *ye = A + 24, *aw = A + 27;
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    double v;
    if (C == 0) v = P[p9(k, 0)];
    else if (C == 1) v = P[p9(k, 1)];
    else if (C == 2) v = aX[0] * P[p9(k, 0)] + aY[0] * P[p9(k, 1)] + P[p9(k, 2)];
    else if (C == 3) v = aX[1] * P[p9(k, 0)] + aY[1] * P[p9(k, 1)] + d[0] * P[p9(k, 3)] + d[3] * P[p9(k, 4)] + d[6] * P[p9(k, 5)] + (IG ? aw[0] * P[p9(k, 2)] : 0.0);
    else if (C == 4) v = aX[2] * P[p9(k, 0)] + aY[2] * P[p9(k, 1)] + d[1] * P[p9(k, 3)] + d[4] * P[p9(k, 4)] + d[7] * P[p9(k, 5)] + (IG ? aw[1] * P[p9(k, 2)] : 0.0);
    else if (C == 5) v = (IG ? A[31] : Ts) * P[p9(k, 2)] + d[2] * P[p9(k, 3)] + d[5] * P[p9(k, 4)] + d[8] * P[p9(k, 5)] + (IG ? xe[0] * P[p9(k, 0)] + ye[0] * P[p9(k, 1)] : 0.0);
    else if (C == 6) v = P[p9(k, 6)];
    else if (C == 9) v = bu[0] * P[p9(k, 3)] + bu[2] * P[p9(k, 4)] + bu[4] * P[p9(k, 5)] + P[p9(k, 7)] + (IG ? xe[1] * P[p9(k, 0)] + ye[1] * P[p9(k, 1)] + aw[2] * P[p9(k, 2)] : 0.0);
    else if (C == 10) v = bu[1] * P[p9(k, 3)] + bu[3] * P[p9(k, 4)] + bu[5] * P[p9(k, 5)] + P[p9(k, 8)] + (IG ? xe[2] * P[p9(k, 0)] + ye[2] * P[p9(k, 1)] + aw[3] * P[p9(k, 2)] : 0.0);
    else if (C == 11) v = P[p9(k, 6)];
    else v = 0.0;   // columns 7, 8 of M are zero
    t[k] = v;
  }
}

// entry r of M' t (row r of mulMx's output), compile-time r
template <int R, int IG>
MPCR_D double mt_entry(const double* v, const double* A, double Ts) {
  const double *aX = A, *aY = A + 3, *d = A + 6, *bu = A + 15, *xe = A + 21, *ye = A + 24, *aw = A + 27;
  if (R == 0) return v[0];
  if (R == 1) return v[1];
  if (R == 2) return aX[0] * v[0] + aY[0] * v[1] + v[2];
  if (R == 3) return aX[1] * v[0] + aY[1] * v[1] + d[0] * v[3] + d[3] * v[4] + d[6] * v[5] + (IG ? aw[0] * v[2] : 0.0);
  if (R == 4) return aX[2] * v[0] + aY[2] * v[1] + d[1] * v[3] + d[4] * v[4] + d[7] * v[5] + (IG ? aw[1] * v[2] : 0.0);
  if (R == 5) return (IG ? A[31] : Ts) * v[2] + d[2] * v[3] + d[5] * v[4] + d[8] * v[5] + (IG ? xe[0] * v[0] + ye[0] * v[1] : 0.0);
  if (R == 6) return v[6];
  if (R == 9) return bu[0] * v[3] + bu[2] * v[4] + bu[4] * v[5] + v[7] + (IG ? xe[1] * v[0] + ye[1] * v[1] + aw[2] * v[2] : 0.0);
  if (R == 10) return bu[1] * v[3] + bu[3] * v[4] + bu[5] * v[5] + v[8] + (IG ? xe[2] * v[0] + ye[2] * v[1] + aw[3] * v[2] : 0.0);
  if (R == 11) return v[6];
  return 0.0;
}

// value + linearisation (+ lam-weighted Hessian) of one stage in the lane solver's layout
template <int MODEL, int IG>
MPCR_D void lane_lin(const double* x, const double* u, double Ts, const double* lam, double* A,
                     double* hm) {
  if (!IG) {
    StageLin lin;
    lin_model(MODEL, x, u, Ts, lam, nullptr, lin, hm);
    for (int a = 0; a < 3; ++a) { A[a] = lin.aX[a]; A[3 + a] = lin.aY[a]; }
    for (int i = 0; i < 3; ++i)
      for (int a = 0; a < 3; ++a) A[6 + 3 * i + a] = lin.d[i][a];
    for (int i = 0; i < 3; ++i)
      for (int a = 0; a < 2; ++a) A[15 + 2 * i + a] = lin.bu[i][a];
  } else {
    double f[6], J[6][6];
    rk4_model(MODEL, x, u, Ts, lam, f, J, hm);
    for (int a = 0; a < 3; ++a) { A[a] = J[0][a]; A[3 + a] = J[1][a]; A[21 + a] = J[0][3 + a]; A[24 + a] = J[1][3 + a]; }
    for (int i = 0; i < 3; ++i) {
      for (int a = 0; a < 3; ++a) A[6 + 3 * i + a] = J[3 + i][1 + a];
      for (int a = 0; a < 2; ++a) A[15 + 2 * i + a] = J[3 + i][4 + a];
    }
    A[27] = J[2][1];
    A[28] = J[2][2];
    A[29] = J[2][4];
    A[30] = J[2][5];
    A[31] = J[2][3];
  }
}

template <int MODEL, int IG>
MPCR_D void lane_step(const double* x, const double* u, double Ts, double* f) {
  if (!IG) f_model(MODEL, x, u, Ts, f);
  else rk4_model(MODEL, x, u, Ts, nullptr, f, nullptr, nullptr);
}

// dense value of the packed stage Hessian at (r, c) from a register copy hs[32]
MPCR_HD double hs_at(const double* hs, int r, int c) {
  const int sl = hs_slot(r, c);
  return sl < 0 ? 0.0 : hs[sl];
}
MPCR_HD constexpr bool dw_diag(int r) { return r <= 6 || r == 9 || r == 10; }

// ---------------------------------------------------------------------------
template <int MODEL, int IG = 0>
struct LaneSolver {
  static constexpr int NA = IG ? kLinRk4 : kLinEuler;
  const Config& cfg;
  const Params& p;
  const IpmOptions& o;
  const ProblemView& pb;
  const LaneWs w;
  const int N;
  // multipliers of the initial-condition rows, problem-wide scalars
  double y0[7], newy0[7], sv_y0[7];
  double df, ddl, mu, tau;
  double lo_[kNumRows], hi_[kNumRows];
  double fval, barB;
  int n_eq_ineq, n_bounds;
  double Dw[2];
  int filt_n;
  // feasibility restoration (quadratic-penalty variant): while `resto` is set the
  // objective is  wq/2 |d(x) - s|^2 + zeta/2 |D_R (x - x_R)|^2, the rows' multipliers are yd = wq (d - s)
  bool resto;
  double wq, zeta;
  int fbase_th, fbase_ph;   // filter arrays in use (outer / restoration)

  MPCR_D LaneSolver(const Config& c, const Params& prm, const ProblemView& pv, double* ws)
      : cfg(c), p(prm), o(c.o), pb(pv), w{ws}, N(c.N) {}

  MPCR_D unsigned rowmask(int k) const {
    unsigned m = 0x4F;                                   // rows 0-3, 6
    if (k > 0 || cfg.wrap) m |= (1u << kDThr) | (1u << kDSt);
    if (k == 0 && pb.has_prev_thr) m |= 1u << kDThr0;
    if (k == 0 && pb.has_prev_st) m |= 1u << kDSt0;
    if (cfg.lane_con) m |= 1u << kLane;
    return m;
  }
  MPCR_D static bool on(unsigned m, int r) { return (m >> r) & 1u; }
  MPCR_D double lo(int r) const { return lo_[r]; }
  MPCR_D double hi(int r) const { return hi_[r]; }

  MPCR_D void row_bounds(int r, double& l, double& h) const {
    const double inf = 1.0e300;
    switch (r) {
      case kThrUb: l = -inf; h = p.d_max; break;
      case kThrLb: l = p.min_throttle; h = inf; break;
      case kStUb: l = -inf; h = p.max_steer; break;
      case kStLb: l = p.min_steer; h = inf; break;
      case kDThr: case kDThr0: l = p.min_throttle_delta; h = p.max_throttle_delta; break;
      case kDSt: case kDSt0: l = p.min_steer_delta; h = p.max_steer_delta; break;
      case kProg: l = p.min_s_delta; h = cfg.Ts * p.v_max; break;
      default: l = -pb.max_err; h = pb.max_err; break;
    }
  }
  MPCR_D double row_value(int r, const double* uu, const double* up, double shh, double shp,
                          double eC) const {
    switch (r) {
      case kThrUb: case kThrLb: return uu[0];
      case kStUb: case kStLb: return uu[1];
      case kDThr: return uu[0] - up[0];
      case kDSt: return uu[1] - up[1];
      case kProg: return shh - shp;
      case kDThr0: return uu[0] - pb.u_prev[0];
      case kDSt0: return uu[1] - pb.u_prev[1];
      default: return ddl * eC;
    }
  }
  MPCR_D static double row_lin(int r, const double* duu, const double* dup, double dshh,
                               double dshp, const double* dxy, const double* gC) {
    switch (r) {
      case kThrUb: case kThrLb: case kDThr0: return duu[0];
      case kStUb: case kStLb: case kDSt0: return duu[1];
      case kDThr: return duu[0] - dup[0];
      case kDSt: return duu[1] - dup[1];
      case kProg: return dshh - dshp;
      default: return gC[0] * dxy[0] + gC[1] * dxy[1] + gC[2] * dshh;
    }
  }
  // Sigma_r and d(barrier objective)/d(slack r) from the slack and its bound multipliers
  MPCR_D void sigma_gsb(int r, double s, double vL, double vU, double& sg, double& g) const {
    sg = 0.0;
    g = 0.0;
    if (row_hasL(r)) {
      const double isl = 1.0 / (s - lo(r));
      sg += vL * isl;
      g -= mu * isl;
      if (!row_hasU(r)) g += o.kappa_d * mu;
    }
    if (row_hasU(r)) {
      const double isu = 1.0 / (hi(r) - s);
      sg += vU * isu;
      g += mu * isu;
      if (!row_hasL(r)) g -= o.kappa_d * mu;
    }
  }

  // -------------------------------------------------------------------------
  // initial guess (control/MPC.py:109-131) and zeroed multipliers
  // -------------------------------------------------------------------------
  MPCR_D void init_iterate() {
    double xr[6];
    for (int a = 0; a < 6; ++a) xr[a] = pb.x0[a];
    int nr = 0, nb = 0;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      double uk[2] = {pb.u_prev[0], pb.u_prev[1]};
      if (pb.u_warm) {
        const int src = (k + 1 < N) ? k + 1 : N - 1;     // shift by one, repeat the last
        uk[0] = pb.u_warm[src];
        uk[1] = pb.u_warm[N + src];
      }
      double fn[6];
      lane_step<kDynamic, IG>(xr, uk, cfg.Ts, fn);       // always f_vehicle (MPC.py:129)
      for (int a = 0; a < 6; ++a) {
        xr[a] = fn[a];
        w.s(k, lw::X + a) = fn[a];
        w.s(k, lw::LAM + a) = 0.0;
        w.s(k, lw::CRES + a) = 0.0;
        w.s(k, lw::DCS + a) = 1.0;
        w.s(k, lw::DX + a) = 0.0;
      }
      w.s(k, lw::SH) = pb.s0 + (k + 1) * cfg.Ts * p.v_max;
      w.s(k, lw::U) = uk[0];
      w.s(k, lw::U + 1) = uk[1];
      w.s(k, lw::DX + 6) = 0.0;
      w.s(k, lw::DX + 7) = 0.0;
      w.s(k, lw::DX + 8) = 0.0;
      const unsigned m = rowmask(k);
#pragma unroll 1
      for (int r = 0; r < kNumRows; ++r) {
        w.s(k, lw::RS + r) = 0.0;
        w.s(k, lw::RYD + r) = 0.0;
        w.s(k, lw::RRD + r) = 0.0;
        w.s(k, lw::RDSR + r) = 0.0;
        w.s(k, lw::RDYD + r) = 0.0;
        w.s(k, lw::RVL + r) = (on(m, r) && row_hasL(r)) ? o.bound_mult_init_val : 0.0;
        w.s(k, lw::RVU + r) = (on(m, r) && row_hasU(r)) ? o.bound_mult_init_val : 0.0;
        if (on(m, r)) { nr += 1; nb += (row_hasL(r) ? 1 : 0) + (row_hasU(r) ? 1 : 0); }
      }
    }
    n_eq_ineq = nr + 6 * N + 7;
    n_bounds = nb;
    for (int a = 0; a < 7; ++a) { y0[a] = 0.0; newy0[a] = 0.0; sv_y0[a] = 0.0; }
  }

  // -------------------------------------------------------------------------
  // Optimality error ingredients (see WarpSolver::Errors)
  // -------------------------------------------------------------------------
  struct Errors {
    double dual_inf, pr_inf, prod_max, prod_min, s_d, s_c, viol;
    MPCR_D double compl_inf(double mu_t) const {
      return fmax(fabs(prod_max - mu_t), fabs(prod_min - mu_t));
    }
    MPCR_D double E(double mu_t) const {
      return fmax(dual_inf / s_d, fmax(pr_inf, compl_inf(mu_t) / s_c));
    }
  };

  // -------------------------------------------------------------------------
  // Pass D (ascending): derivatives at the current iterate -> stage records; in Newton
  // mode also the optimality error; in Init mode the gradient-based scaling factors.
  //   mode 0: Newton (exact Hessian), mode 1: least-squares multiplier init
  // -------------------------------------------------------------------------
  // keep_scaling: mode-1 evaluation that leaves the row scaling untouched (multiplier re-init after restoration)
  MPCR_D void pass_derivs(int mode, Errors& er, double& gmax_out, double& ddl_out, bool keep_scaling) {
    const double Ts = cfg.Ts;
    double xp[6], shp = pb.s0, up[2];
    for (int a = 0; a < 6; ++a) xp[a] = pb.x0[a];
    up[0] = w.s(N - 1, lw::U);
    up[1] = w.s(N - 1, lw::U + 1);
    double hcp[10];
    for (int a = 0; a < 10; ++a) hcp[a] = 0.0;
    // carried quantities of stage k-1 for the dual infeasibility of state k / input k-1
    double gxp[6], lamp[6], gsp = 0.0, yprog_p = 0.0, ru0p = 0.0, ru1p = 0.0, sigc_p = 0.0;
    for (int a = 0; a < 6; ++a) { gxp[a] = 0.0; lamp[a] = 0.0; }
    double wrap_ydthr = 0.0, wrap_ydst = 0.0;
    double dinf = 0.0, prinf = 0.0, pmax = -1.0e300, pmin = 1.0e300, sum_y = 0.0, sum_v = 0.0;
    double vio = 0.0;
    double gmax = 0.0, ddlmin = 1.0;
    double sig_prev_for_scale = 0.0;
    const double bd2 = 2.0 * p.beta_delta;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      if (k + 1 < N) { w.pf(k + 1, lw::X, 27); w.pf_rows(k + 1, lw::RS, 5); }
      const unsigned m = rowmask(k);
      double xk[6], uk[2], lamk[6];
      for (int a = 0; a < 6; ++a) xk[a] = w.s(k, lw::X + a);
      const double shk = w.s(k, lw::SH);
      uk[0] = w.s(k, lw::U);
      uk[1] = w.s(k, lw::U + 1);
      for (int a = 0; a < 6; ++a) lamk[a] = (mode == 0) ? w.s(k, lw::LAM + a) : 0.0;
      double A[kLinRk4];
      double hm[kHmDoubles];
      lane_lin<MODEL, IG>(xp, uk, Ts, lamk, A, hm);
      StageCost sc;
      stage_cost(p, pb.cx, pb.cy, xk, shk, 2, sc);
      double gcost[5], gC[3];
      gcost[0] = df * sc.g3[0];
      gcost[1] = df * sc.g3[1];
      gcost[2] = df * sc.g3[2] - ((k == N - 1) ? df * p.lambda_s : 0.0);
      gcost[3] = df * sc.gvx;
      gcost[4] = df * sc.gvy;
      for (int a = 0; a < 3; ++a) gC[a] = ddl * sc.gC[a];
      const double sigc = (k >= 1 && !resto) ? (uk[1] - up[1]) : 0.0;
      double hc[10];            // cost Hessian of state k+1: (X,Y,s) block (6), v_x, v_y, [restoration: yaw, r]
      for (int a = 0; a < 10; ++a) hc[a] = 0.0;
      const double ydl = on(m, kLane) ? w.s(k, lw::RYD + kLane) : 0.0;
      double pgr[9], pdu[2] = {0.0, 0.0};
      for (int a = 0; a < 9; ++a) pgr[a] = 0.0;
      if (resto) {
        // proximity term zeta/2 |D_R (v - v_R)|^2 over (x_{k+1}, s_{k+1}, u_k) replaces the cost
        const double v9[9] = {xk[0], xk[1], xk[2], xk[3], xk[4], xk[5], shk, uk[0], uk[1]};
        double pd9[9];
        for (int a = 0; a < 9; ++a) {
          const double ref = w.s(k, lw::XR + a);
          const double dr = 1.0 / fmax(1.0, fabs(ref));
          pd9[a] = zeta * dr * dr;
          pgr[a] = pd9[a] * (v9[a] - ref);
          w.s(k, lw::PG + a) = pgr[a];
        }
        gcost[0] = pgr[0]; gcost[1] = pgr[1]; gcost[2] = pgr[6]; gcost[3] = pgr[3]; gcost[4] = pgr[4];
        hc[0] = pd9[0]; hc[3] = pd9[1]; hc[5] = pd9[6]; hc[6] = pd9[3]; hc[7] = pd9[4];
        hc[8] = pd9[2]; hc[9] = pd9[5];
        pdu[0] = pd9[7]; pdu[1] = pd9[8];
        if (cfg.lane_con && on(m, kLane)) {
          const double yl = ydl * ddl;
          hc[2] += yl * sc.hC02;
          hc[4] += yl * sc.hC12;
          hc[5] += yl * sc.hC22;
        }
      } else if (mode == 0) {
        for (int a = 0; a < 6; ++a) hc[a] = df * sc.h3[a];
        hc[6] = df * sc.hvx;
        hc[7] = df * sc.hvy;
        if (cfg.lane_con && on(m, kLane)) {
          const double yl = ydl * ddl;           // + yd * Hess(e_C); the Sigma part is a row term
          hc[2] += yl * sc.hC02;
          hc[4] += yl * sc.hC12;
          hc[5] += yl * sc.hC22;
        }
      }
      // ---- stage record ------------------------------------------------------------
      for (int a = 0; a < NA; ++a) w.s(k, lw::A + a) = A[a];
      for (int a = 0; a < 6; ++a) w.s(k, lw::HSD + a) = hcp[a];
      for (int a = 0; a < kHmDoubles; ++a) {
        double v = (mode == 0) ? -hm[a] : 0.0;
        if (a == hm_idx(1, 1)) v += hcp[6];
        if (a == hm_idx(2, 2)) v += hcp[7];
        if (a == hm_idx(0, 0)) v += hcp[8];
        if (a == hm_idx(3, 3)) v += hcp[9];
        if (a == hm_idx(4, 4)) v += pdu[0];
        if (a == hm_idx(5, 5)) v += pdu[1];
        w.s(k, lw::HSD + 6 + a) = v;
      }
      for (int a = 0; a < 5; ++a) w.s(k, lw::GC + a) = gcost[a];
      for (int a = 0; a < 3; ++a) w.s(k, lw::GCZ + a) = gC[a];
      w.s(k, lw::SIGC) = sigc;
      if (k == N - 1) {
        for (int a = 0; a < 6; ++a) w.t(N, lw::T_HSD + a) = hc[a];
        for (int a = 0; a < kHmDoubles; ++a) {
          double v = 0.0;
          if (a == hm_idx(1, 1)) v = hc[6];
          if (a == hm_idx(2, 2)) v = hc[7];
          if (a == hm_idx(0, 0)) v = hc[8];
          if (a == hm_idx(3, 3)) v = hc[9];
          w.t(N, lw::T_HSD + 6 + a) = v;
        }
      }
      if (mode == 1 && keep_scaling) {
      } else if (mode == 1) {
        // ---- gradient-based scaling at the starting point (eval ran with df = ddl = 1) ----
        for (int a = 0; a < 5; ++a) gmax = fmax(gmax, fabs(gcost[a]));
        if (k >= 1) gmax = fmax(gmax, fabs(bd2 * (sig_prev_for_scale - sigc)));
        sig_prev_for_scale = sigc;
        double rm[6];
        rm[0] = fmax(1.0, fmax(fabs(A[0]), fmax(fabs(A[1]), fabs(A[2]))));
        rm[1] = fmax(1.0, fmax(fabs(A[3]), fmax(fabs(A[4]), fabs(A[5]))));
        rm[2] = fmax(1.0, fabs(Ts));
        if (IG) {
          for (int a = 0; a < 3; ++a) { rm[0] = fmax(rm[0], fabs(A[21 + a])); rm[1] = fmax(rm[1], fabs(A[24 + a])); }
          rm[2] = fmax(1.0, fabs(A[31]));
          for (int a = 0; a < 4; ++a) rm[2] = fmax(rm[2], fabs(A[27 + a]));
        }
        for (int i = 0; i < 3; ++i) {
          double mm = 1.0;
          for (int a = 0; a < 3; ++a) mm = fmax(mm, fabs(A[6 + 3 * i + a]));
          for (int a = 0; a < 2; ++a) mm = fmax(mm, fabs(A[15 + 2 * i + a]));
          rm[3 + i] = mm;
        }
        for (int a = 0; a < 6; ++a)
          w.s(k, lw::DCS + a) = (rm[a] > o.nlp_scaling_max_gradient)
                                    ? fmax(o.nlp_scaling_max_gradient / rm[a], o.nlp_scaling_min_value) : 1.0;
        const double gl = fmax(fabs(gC[0]), fmax(fabs(gC[1]), fabs(gC[2])));
        double dd = 1.0;
        if (cfg.lane_con && gl > o.nlp_scaling_max_gradient)
          dd = fmax(o.nlp_scaling_max_gradient / gl, o.nlp_scaling_min_value);
        ddlmin = fmin(ddlmin, dd);
      } else {
        // ---- optimality error (WarpSolver::compute_errors) -------------------------------
        double lv[9], out[12];
        for (int a = 0; a < 6; ++a) lv[a] = lamk[a];
        lv[6] = lv[7] = lv[8] = 0.0;
        mulMx<IG>(lv, A, Ts, out);
        const double yprog = w.s(k, lw::RYD + kProg);
        const double ydthr = on(m, kDThr) ? w.s(k, lw::RYD + kDThr) : 0.0;
        const double ydst = on(m, kDSt) ? w.s(k, lw::RYD + kDSt) : 0.0;
        if (k == 0) {
          for (int a = 0; a < 6; ++a) dinf = fmax(dinf, fabs(y0[a] - out[a]));
          dinf = fmax(dinf, fabs(y0[6] - yprog));
          for (int a = 0; a < 7; ++a) sum_y += fabs(y0[a]);
          wrap_ydthr = ydthr;
          wrap_ydst = ydst;
        } else {
          for (int a = 0; a < 6; ++a) dinf = fmax(dinf, fabs(gxp[a] + lamp[a] - out[a]));
          dinf = fmax(dinf, fabs(gsp + yprog_p - yprog));
          const double r0 = ru0p - ydthr;
          const double r1 = ru1p - ydst + bd2 * df * (sigc_p - sigc);
          dinf = fmax(dinf, fmax(fabs(r0), fabs(r1)));
        }
        gxp[0] = gcost[0] + ydl * gC[0];
        gxp[1] = gcost[1] + ydl * gC[1];
        gxp[2] = pgr[2];
        gxp[3] = gcost[3];
        gxp[4] = gcost[4];
        gxp[5] = pgr[5];
        gsp = gcost[2] + ydl * gC[2];
        for (int a = 0; a < 6; ++a) lamp[a] = lamk[a];
        yprog_p = yprog;
        ru0p = -out[9] + w.s(k, lw::RYD + kThrUb) + w.s(k, lw::RYD + kThrLb) + ydthr +
               (on(m, kDThr0) ? w.s(k, lw::RYD + kDThr0) : 0.0) + pgr[7];
        ru1p = -out[10] + w.s(k, lw::RYD + kStUb) + w.s(k, lw::RYD + kStLb) + ydst +
               (on(m, kDSt0) ? w.s(k, lw::RYD + kDSt0) : 0.0) + pgr[8];
        sigc_p = sigc;
        for (int a = 0; a < 6; ++a) {
          const double c = w.s(k, lw::CRES + a), dc = w.s(k, lw::DCS + a);
          prinf = fmax(prinf, fabs(dc * c));
          vio = fmax(vio, fabs(c));
          sum_y += fabs(lamk[a] / dc);
        }
#pragma unroll 1
        for (int r = 0; r < kNumRows; ++r) {
          if (!on(m, r)) continue;
          const double s = w.s(k, lw::RS + r), yd = w.s(k, lw::RYD + r), rd = w.s(k, lw::RRD + r);
          double rs = -yd;
          if (row_hasL(r)) {
            const double vL = w.s(k, lw::RVL + r);
            rs -= vL;
            const double pr_ = (s - lo(r)) * vL;
            pmax = fmax(pmax, pr_);
            pmin = fmin(pmin, pr_);
            sum_v += fabs(vL);
          }
          if (row_hasU(r)) {
            const double vU = w.s(k, lw::RVU + r);
            rs += vU;
            const double pr_ = (hi(r) - s) * vU;
            pmax = fmax(pmax, pr_);
            pmin = fmin(pmin, pr_);
            sum_v += fabs(vU);
          }
          dinf = fmax(dinf, fabs(rs));
          if (!resto) prinf = fmax(prinf, fabs(rd));   // restoration: the rows sit in the objective
          sum_y += fabs(yd);
          double bl, bh;
          row_bounds(r, bl, bh);
          double val = rd + s;
          if (r == kLane) val /= ddl;
          if (row_hasL(r)) vio = fmax(vio, bl - val);
          if (row_hasU(r)) vio = fmax(vio, val - bh);
        }
      }
      for (int a = 0; a < 6; ++a) xp[a] = xk[a];
      shp = shk;
      up[0] = uk[0];
      up[1] = uk[1];
      for (int a = 0; a < 10; ++a) hcp[a] = hc[a];
    }
    (void)shp;
    if (mode == 1) gmax = fmax(gmax, fabs(bd2 * sig_prev_for_scale));   // stage N-1: no successor
    if (mode == 0) {
      // state N / input N-1: no successor stage; the wrap rows couple u_{N-1} with stage 0
      for (int a = 0; a < 6; ++a) dinf = fmax(dinf, fabs(gxp[a] + lamp[a]));
      dinf = fmax(dinf, fabs(gsp + yprog_p));
      const double r0 = ru0p - wrap_ydthr;
      const double r1 = ru1p - wrap_ydst + bd2 * df * sigc_p;
      dinf = fmax(dinf, fmax(fabs(r0), fabs(r1)));
      er.dual_inf = dinf;
      er.pr_inf = prinf;
      er.viol = vio;
      er.prod_max = pmax;
      er.prod_min = pmin;
      const double smax = o.s_max;
      er.s_d = fmax(smax, (sum_y + sum_v) / (double)(n_eq_ineq + n_bounds)) / smax;
      er.s_c = fmax(smax, sum_v / (double)n_bounds) / smax;
    }
    gmax_out = gmax;
    ddl_out = ddlmin;
  }

  // rescale the stored cost / lane-row gradients after the scaling factors are known
  MPCR_D void pass_rescale() {
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      for (int a = 0; a < 5; ++a) w.s(k, lw::GC + a) *= df;
      for (int a = 0; a < 3; ++a) w.s(k, lw::GCZ + a) *= ddl;
    }
  }

  // -------------------------------------------------------------------------
  // Row terms of one stage (WarpSolver::write_row_terms): D_r and e_r by role.
  // -------------------------------------------------------------------------
  struct RowAcc {
    double Rtt, Rss, eth, est, Dthr, ethr, Dst, esr, Dprog, eprog, Dlane, elane;
  };
  // rsel: 0 = zero right-hand side, 1 = current residual (RRD), 2 = SOC (DSOC)
  MPCR_D void row_acc(int k, int mode, double dw, int rsel, RowAcc& q) const {
    q.Rtt = q.Rss = q.eth = q.est = q.Dthr = q.ethr = q.Dst = q.esr = 0.0;
    q.Dprog = q.eprog = q.Dlane = q.elane = 0.0;
    const unsigned m = rowmask(k);
#pragma unroll 1
    for (int r = 0; r < kNumRows; ++r) {
      double Dr = 0.0, er = 0.0;
      if (on(m, r)) {
        const double vL = w.s(k, lw::RVL + r), vU = w.s(k, lw::RVU + r);
        if (mode == 0) {
          double sg, g;
          sigma_gsb(r, w.s(k, lw::RS + r), vL, vU, sg, g);
          Dr = sg + dw;
          const double rdv = (rsel == 0) ? 0.0 : w.s(k, (rsel == 1 ? lw::RRD : lw::DSOC) + r);
          er = Dr * rdv + g;
          if (resto) {
            // row in the objective: min over ds of wq/2 (J dx - ds + rd)^2 + Sigma/2 ds^2 + g ds
            const double kap = wq / (wq + Dr);
            Dr *= kap;
            er *= kap;
          }
        } else {
          Dr = 1.0;
          er = -((row_hasU(r) ? vU : 0.0) - (row_hasL(r) ? vL : 0.0));
        }
      }
      switch (r) {
        case kThrUb: case kThrLb: case kDThr0: q.Rtt += Dr; q.eth += er; break;
        case kStUb: case kStLb: case kDSt0: q.Rss += Dr; q.est += er; break;
        case kDThr: q.Dthr = Dr; q.ethr = er; break;
        case kDSt: q.Dst = Dr; q.esr = er; break;
        case kProg: q.Dprog = Dr; q.eprog = er; break;
        default: q.Dlane = Dr; q.elane = er; break;
      }
    }
  }

  // -------------------------------------------------------------------------
  // Pass R (descending): row terms + stage assembly + Riccati factorisation with 3
  // right-hand sides.  Returns false when a pivot block is not positive definite.
  // -------------------------------------------------------------------------
  MPCR_D bool pass_riccati(int mode, double dw, int rsel, int csel) {
    const double Ts = cfg.Ts;
    const int nrhs = cfg.wrap ? kRhs : 1;
    const double sgn = (mode == 0) ? 1.0 : -1.0;   // LSQ uses the negated gradient
    // wrap rows (kDThr, kDSt of stage 0)
    RowAcc q0;
    row_acc(0, mode, dw, rsel, q0);
    const double Dw0 = q0.Dthr, Dw1 = q0.Dst, ew0 = q0.ethr, ew1 = q0.esr;
    Dw[0] = Dw0;
    Dw[1] = Dw1;
    double P[45], pv[kRhs][9];
    RowAcc qn;                      // rows of the stage above (k + 1); first the terminal's owner N-1
    RowAcc q;
    row_acc(N - 1, mode, dw, rsel, q);
    {
      // ---- terminal stage: cost of state N (+ lane row of stage N-1) ------------------------
      double gc[5], gCn[3], hsd[27];
      for (int a = 0; a < 5; ++a) gc[a] = w.s(N - 1, lw::GC + a);
      for (int a = 0; a < 3; ++a) gCn[a] = w.s(N - 1, lw::GCZ + a);
      for (int a = 0; a < 27; ++a) hsd[a] = w.t(N, lw::T_HSD + a);
      double hN[32], gN[12];
      hN[0] = hsd[0] + q.Dlane * gCn[0] * gCn[0];
      hN[1] = hsd[1] + q.Dlane * gCn[0] * gCn[1];
      hN[2] = hsd[2] + q.Dlane * gCn[0] * gCn[2];
      hN[3] = hsd[3] + q.Dlane * gCn[1] * gCn[1];
      hN[4] = hsd[4] + q.Dlane * gCn[1] * gCn[2];
      hN[5] = hsd[5] + q.Dlane * gCn[2] * gCn[2];
      for (int a = 0; a < 21; ++a) hN[6 + a] = hsd[6 + a];
      for (int a = 27; a < 32; ++a) hN[a] = 0.0;
      gN[0] = sgn * gc[0] + q.elane * gCn[0];
      gN[1] = sgn * gc[1] + q.elane * gCn[1];
      gN[2] = resto ? w.s(N - 1, lw::PG + 2) : 0.0;
      gN[3] = sgn * gc[3];
      gN[4] = sgn * gc[4];
      gN[5] = resto ? w.s(N - 1, lw::PG + 5) : 0.0;
      gN[6] = sgn * gc[2] + q.elane * gCn[2];
      for (int a = 7; a < 12; ++a) gN[a] = 0.0;
      for (int a = 0; a < 32; ++a) w.t(N, lw::T_HS + a) = hN[a];
      for (int a = 0; a < 12; ++a) w.t(N, lw::T_GS + a) = gN[a];
#pragma unroll
      for (int r = 0; r < 9; ++r)
#pragma unroll
        for (int c = r; c < 9; ++c)
          P[p9(r, c)] = hs_at(hN, r, c) + ((r == c && dw_diag(r)) ? dw : 0.0);
      for (int a = 0; a < 9; ++a) { pv[0][a] = gN[a]; pv[1][a] = 0.0; pv[2][a] = 0.0; }
    }
    bool ok = true;
#pragma unroll 1
    for (int k = N - 1; k >= 0; --k) {
      if (k >= 1) { w.pf(k - 1, lw::CRES, 6); w.pf(k - 1, lw::A, NA); w.pf(k - 1, lw::HSD, 27 + 9); }
      if (k >= 2) { w.pf_rows(k - 2, lw::RS, 1); w.pf_rows(k - 2, lw::RVL, 3); w.pf(k - 2, lw::GC, 9); }
      // rows of this stage are in q; the lane row of state k belongs to stage k-1
      double Dlane_p = 0.0, elane_p = 0.0;
      double gcp[5] = {0, 0, 0, 0, 0}, gCp[3] = {0, 0, 0};
      if (k >= 1) {
        row_acc(k - 1, mode, dw, rsel, qn);
        Dlane_p = qn.Dlane;
        elane_p = qn.elane;
        for (int a = 0; a < 5; ++a) gcp[a] = w.s(k - 1, lw::GC + a);
        for (int a = 0; a < 3; ++a) gCp[a] = w.s(k - 1, lw::GCZ + a);
      }
      const bool wrap0 = (k == 0);
      const double bsig = (k >= 1 && !resto) ? 2.0 * p.beta_delta * df : 0.0;
      const double bq = (mode == 0) ? bsig : 0.0;
      const double pg_yaw = (resto && k >= 1) ? w.s(k - 1, lw::PG + 2) : 0.0;
      const double pg_r = (resto && k >= 1) ? w.s(k - 1, lw::PG + 5) : 0.0;
      const double pg_thr = resto ? w.s(k, lw::PG + 7) : 0.0;
      const double pg_st = resto ? w.s(k, lw::PG + 8) : 0.0;
      const double dthr = wrap0 ? 0.0 : q.Dthr;
      const double dst = wrap0 ? 0.0 : q.Dst;
      double hs[32], gs[12], A[kLinRk4], b[6];
      for (int a = 0; a < 6; ++a) hs[a] = w.s(k, lw::HSD + a);
      for (int a = 0; a < 21; ++a) hs[6 + a] = w.s(k, lw::HSD + 6 + a);
      hs[0] += Dlane_p * gCp[0] * gCp[0];
      hs[1] += Dlane_p * gCp[0] * gCp[1];
      hs[2] += Dlane_p * gCp[0] * gCp[2];
      hs[3] += Dlane_p * gCp[1] * gCp[1];
      hs[4] += Dlane_p * gCp[1] * gCp[2];
      hs[5] += Dlane_p * gCp[2] * gCp[2];
      hs[6 + hm_idx(4, 4)] += q.Rtt + dthr;
      hs[6 + hm_idx(5, 5)] += q.Rss + dst + bq;
      hs[27] = q.Dprog;
      hs[28] = dthr;
      hs[29] = dst + bq;
      hs[30] = -dthr;
      hs[31] = -(dst + bq);
      double ethk = q.eth + (wrap0 ? 0.0 : q.ethr);
      double estk = q.est + (wrap0 ? 0.0 : q.esr);
      if (cfg.wrap) {
        if (k == 0) { ethk += ew0; estk += ew1; }
        if (k == N - 1) { ethk -= ew0; estk -= ew1; }
      }
      const double bg = sgn * bsig * w.s(k, lw::SIGC);
      gs[0] = sgn * gcp[0] + elane_p * gCp[0];
      gs[1] = sgn * gcp[1] + elane_p * gCp[1];
      gs[2] = pg_yaw;
      gs[3] = sgn * gcp[3];
      gs[4] = sgn * gcp[4];
      gs[5] = pg_r;
      gs[6] = sgn * gcp[2] + elane_p * gCp[2];
      gs[7] = wrap0 ? 0.0 : -q.ethr;
      gs[8] = (wrap0 ? 0.0 : -q.esr) - bg;
      gs[9] = ethk + pg_thr;
      gs[10] = estk + bg + pg_st;
      gs[11] = q.eprog;
      for (int a = 0; a < 32; ++a) w.s(k, lw::HS + a) = hs[a];
      for (int a = 0; a < 12; ++a) w.s(k, lw::GS + a) = gs[a];
      for (int a = 0; a < NA; ++a) A[a] = w.s(k, lw::A + a);
      for (int a = 0; a < 6; ++a)
        b[a] = (csel == 0) ? 0.0 : -w.s(k, (csel == 1 ? lw::CRES : lw::CSOC) + a);
      // ---- G columns of the inputs w = (thr, st, ds): Gw[i][r] = G[r][9 + i] ----------------
      double Gw[3][12];
      {
        double t[9];
        p_times_mcol<9, IG>(P, A, Ts, t);
        Gw[0][0] = mt_entry<0, IG>(t, A, Ts); Gw[0][1] = mt_entry<1, IG>(t, A, Ts); Gw[0][2] = mt_entry<2, IG>(t, A, Ts);
        Gw[0][3] = mt_entry<3, IG>(t, A, Ts); Gw[0][4] = mt_entry<4, IG>(t, A, Ts); Gw[0][5] = mt_entry<5, IG>(t, A, Ts);
        Gw[0][6] = mt_entry<6, IG>(t, A, Ts); Gw[0][7] = 0.0; Gw[0][8] = 0.0;
        Gw[0][9] = mt_entry<9, IG>(t, A, Ts); Gw[0][10] = mt_entry<10, IG>(t, A, Ts); Gw[0][11] = mt_entry<11, IG>(t, A, Ts);
        p_times_mcol<10, IG>(P, A, Ts, t);
        Gw[1][0] = mt_entry<0, IG>(t, A, Ts); Gw[1][1] = mt_entry<1, IG>(t, A, Ts); Gw[1][2] = mt_entry<2, IG>(t, A, Ts);
        Gw[1][3] = mt_entry<3, IG>(t, A, Ts); Gw[1][4] = mt_entry<4, IG>(t, A, Ts); Gw[1][5] = mt_entry<5, IG>(t, A, Ts);
        Gw[1][6] = mt_entry<6, IG>(t, A, Ts); Gw[1][7] = 0.0; Gw[1][8] = 0.0;
        Gw[1][9] = Gw[0][10]; Gw[1][10] = mt_entry<10, IG>(t, A, Ts); Gw[1][11] = mt_entry<11, IG>(t, A, Ts);
        p_times_mcol<11, IG>(P, A, Ts, t);
        Gw[2][0] = mt_entry<0, IG>(t, A, Ts); Gw[2][1] = mt_entry<1, IG>(t, A, Ts); Gw[2][2] = mt_entry<2, IG>(t, A, Ts);
        Gw[2][3] = mt_entry<3, IG>(t, A, Ts); Gw[2][4] = mt_entry<4, IG>(t, A, Ts); Gw[2][5] = mt_entry<5, IG>(t, A, Ts);
        Gw[2][6] = mt_entry<6, IG>(t, A, Ts); Gw[2][7] = 0.0; Gw[2][8] = 0.0;
        Gw[2][9] = Gw[0][11]; Gw[2][10] = Gw[1][11]; Gw[2][11] = mt_entry<11, IG>(t, A, Ts);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int r = 0; r < 12; ++r) Gw[i][r] += hs_at(hs, r, 9 + i);
        Gw[0][9] += dw;       // (thr, thr) and (st, st) receive delta_w, (ds, ds) does not
        Gw[1][10] += dw;
      }
      // ---- LDL' of H_ww ---------------------------------------------------------------------
      const double h00 = Gw[0][9], h10 = Gw[0][10], h20 = Gw[0][11];
      const double h11 = Gw[1][10], h21 = Gw[1][11], h22 = Gw[2][11];
      const double d0 = h00;
      const double id0 = 1.0 / d0;
      const double l10 = h10 * id0, l20 = h20 * id0;
      const double d1 = h11 - l10 * h10;
      const double id1 = 1.0 / d1;
      const double m21 = h21 - l20 * h10;
      const double l21 = m21 * id1;
      const double d2 = h22 - l20 * h20 - l21 * m21;
      const double id2 = 1.0 / d2;
      if (!(d0 > 0.0 && d1 > 0.0 && d2 > 0.0)) { ok = false; break; }
      // ---- right-hand sides -----------------------------------------------------------------
      double pvn[kRhs][9];
#pragma unroll
      for (int j = 0; j < kRhs; ++j) {
        if (j < nrhs) {
          double tv[9];
#pragma unroll
          for (int r = 0; r < 9; ++r) {
            double t0 = pv[j][r];
            if (j == 0) {
#pragma unroll
              for (int c = 0; c < 6; ++c) t0 += P[p9(r, c)] * b[c];
            }
            tv[r] = t0;
          }
          double g[12];
          mulMx<IG>(tv, A, Ts, g);
          if (j == 0) {
#pragma unroll
            for (int r = 0; r < 12; ++r) g[r] += gs[r];
          } else {
            const double a = (k == 0 ? 1.0 : 0.0) - (k == N - 1 ? 1.0 : 0.0);
            g[9 + (j - 1)] += a;
          }
          const double f0 = g[9], f1 = g[10] - l10 * f0, f2 = g[11] - l20 * f0 - l21 * f1;
          const double z2 = f2 * id2;
          const double z1 = f1 * id1 - l21 * z2;
          const double z0 = f0 * id0 - l10 * z1 - l20 * z2;
          const double k0 = -z0, k1 = -z1, k2 = -z2;
          w.s(k, lw::KF + j * 3 + 0) = k0;
          w.s(k, lw::KF + j * 3 + 1) = k1;
          w.s(k, lw::KF + j * 3 + 2) = k2;
#pragma unroll
          for (int a = 0; a < 9; ++a) pvn[j][a] = g[a] + Gw[0][a] * k0 + Gw[1][a] * k1 + Gw[2][a] * k2;
        } else {
#pragma unroll
          for (int a = 0; a < 9; ++a) pvn[j][a] = 0.0;
        }
      }
      // ---- gains and cost-to-go, column by column (upper triangle) ---------------------------
      double Pn[45];
      riccati_col<0>(P, Pn, A, hs, Gw, Ts, dw, id0, id1, id2, l10, l20, l21, k);
      riccati_col<1>(P, Pn, A, hs, Gw, Ts, dw, id0, id1, id2, l10, l20, l21, k);
      riccati_col<2>(P, Pn, A, hs, Gw, Ts, dw, id0, id1, id2, l10, l20, l21, k);
      riccati_col<3>(P, Pn, A, hs, Gw, Ts, dw, id0, id1, id2, l10, l20, l21, k);
      riccati_col<4>(P, Pn, A, hs, Gw, Ts, dw, id0, id1, id2, l10, l20, l21, k);
      riccati_col<5>(P, Pn, A, hs, Gw, Ts, dw, id0, id1, id2, l10, l20, l21, k);
      riccati_col<6>(P, Pn, A, hs, Gw, Ts, dw, id0, id1, id2, l10, l20, l21, k);
      riccati_col<7>(P, Pn, A, hs, Gw, Ts, dw, id0, id1, id2, l10, l20, l21, k);
      riccati_col<8>(P, Pn, A, hs, Gw, Ts, dw, id0, id1, id2, l10, l20, l21, k);
#pragma unroll
      for (int a = 0; a < 45; ++a) P[a] = Pn[a];
#pragma unroll
      for (int j = 0; j < kRhs; ++j)
#pragma unroll
        for (int a = 0; a < 9; ++a) pv[j][a] = pvn[j][a];
      q = qn;
    }
    return ok;
  }

  // gain column c and column c (rows <= c) of the new cost-to-go
  template <int C>
  MPCR_D void riccati_col(const double* P, double* Pn, const double* A, const double* hs,
                          const double (*Gw)[12], double Ts, double dw, double id0, double id1,
                          double id2, double l10, double l20, double l21, int k) const {
    const double r0 = Gw[0][C], r1 = Gw[1][C], r2 = Gw[2][C];
    const double f0 = r0, f1 = r1 - l10 * f0, f2 = r2 - l20 * f0 - l21 * f1;
    const double z2 = f2 * id2;
    const double z1 = f1 * id1 - l21 * z2;
    const double z0 = f0 * id0 - l10 * z1 - l20 * z2;
    const double k0 = -z0, k1 = -z1, k2 = -z2;
    w.s(k, lw::K + 0 * 9 + C) = k0;
    w.s(k, lw::K + 1 * 9 + C) = k1;
    w.s(k, lw::K + 2 * 9 + C) = k2;
    double t[9];
    p_times_mcol<C, IG>(P, A, Ts, t);
    double g[9];
    g[0] = mt_entry<0, IG>(t, A, Ts);
    if (C >= 1) g[1] = mt_entry<1, IG>(t, A, Ts);
    if (C >= 2) g[2] = mt_entry<2, IG>(t, A, Ts);
    if (C >= 3) g[3] = mt_entry<3, IG>(t, A, Ts);
    if (C >= 4) g[4] = mt_entry<4, IG>(t, A, Ts);
    if (C >= 5) g[5] = mt_entry<5, IG>(t, A, Ts);
    if (C >= 6) g[6] = mt_entry<6, IG>(t, A, Ts);
    if (C >= 7) g[7] = 0.0;
    if (C >= 8) g[8] = 0.0;
#pragma unroll
    for (int a = 0; a <= C; ++a) {
      double v = g[a] + hs_at(hs, a, C);
      if (a == C && dw_diag(C)) v += dw;
      Pn[p9(a, C)] = v + Gw[0][a] * k0 + Gw[1][a] * k1 + Gw[2][a] * k2;
    }
  }

  // -------------------------------------------------------------------------
  // Pass F (ascending): roll the 3 right-hand sides forward; results into Z3.
  // mwa / mwb: (dthr, dst) of each right-hand side at stage 0 / stage N-1.
  // -------------------------------------------------------------------------
  MPCR_D void pass_forward(int csel, double (*mwa)[2], double (*mwb)[2]) {
    const double Ts = cfg.Ts;
    const int nrhs = cfg.wrap ? kRhs : 1;
    double z[kRhs][9];
    for (int j = 0; j < kRhs; ++j)
      for (int a = 0; a < 9; ++a) z[j][a] = 0.0;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      if (k + 1 < N) { w.pf(k + 1, lw::CRES, 6); w.pf(k + 1, lw::A, NA); w.pf(k + 1, lw::K, 36); }
      double A[kLinRk4], Kk[27];
      for (int a = 0; a < NA; ++a) A[a] = w.s(k, lw::A + a);
      for (int a = 0; a < 27; ++a) Kk[a] = w.s(k, lw::K + a);
#pragma unroll
      for (int j = 0; j < kRhs; ++j) {
        if (j < nrhs) {
          double ww[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            double acc = w.s(k, lw::KF + j * 3 + i);
#pragma unroll
            for (int a = 0; a < 9; ++a) acc += Kk[i * 9 + a] * z[j][a];
            ww[i] = acc;
          }
          double zn[9];
          applyMx<IG>(z[j], ww, A, Ts, zn);
          if (j == 0 && csel != 0) {
#pragma unroll
            for (int a = 0; a < 6; ++a) zn[a] -= w.s(k, (csel == 1 ? lw::CRES : lw::CSOC) + a);
          }
#pragma unroll
          for (int a = 0; a < 9; ++a) z[j][a] = zn[a];
#pragma unroll
          for (int a = 0; a < 7; ++a) w.s(k, lw::Z3 + j * 9 + a) = zn[a];
          w.s(k, lw::Z3 + j * 9 + 7) = ww[0];
          w.s(k, lw::Z3 + j * 9 + 8) = ww[1];
          if (k == 0) { mwa[j][0] = ww[0]; mwa[j][1] = ww[1]; }
          if (k == N - 1) { mwb[j][0] = ww[0]; mwb[j][1] = ww[1]; }
        }
      }
    }
  }

  // -------------------------------------------------------------------------
  // Pass C (descending): Woodbury combination of the 3 roll-outs into the step, costate
  // recursion for the new equality multipliers, slack steps + fraction-to-the-boundary
  // step, dual steps of the rows, directional derivative and theta of the current point.
  // -------------------------------------------------------------------------
  struct StepInfo {
    double a_pr;      // fraction-to-the-boundary primal step
    double a_du;      // same for the bound multipliers
    double gs_ds;     // grad_s(phi)' ds
    double gf_d;      // grad_x(f)' dx (scaled objective)
    double theta;     // |c|_1 + |d - s|_1 at the current point
  };

  MPCR_D void load_combined(int k, double c1, double c2, double* d9) const {
    const int nrhs = cfg.wrap ? kRhs : 1;
#pragma unroll
    for (int a = 0; a < 9; ++a) {
      double v = w.s(k, lw::Z3 + a);
      if (nrhs > 1) v += c1 * w.s(k, lw::Z3 + 9 + a) + c2 * w.s(k, lw::Z3 + 18 + a);
      d9[a] = v;
    }
  }

  MPCR_D void pass_combine(double c1, double c2, double dw, int rsel, StepInfo& si) {
    const double Ts = cfg.Ts;
    double dcur[9], dprev[9], dlast_u[2];
    load_combined(N - 1, c1, c2, dcur);
    dlast_u[0] = dcur[7];
    dlast_u[1] = dcur[8];
    // terminal costate
    double pi[7];
    {
      double hN[32];
      for (int a = 0; a < 32; ++a) hN[a] = w.t(N, lw::T_HS + a);
#pragma unroll
      for (int a = 0; a < 7; ++a) {
        double acc = w.t(N, lw::T_GS + a);
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          if (hs_slot(a, c) < 0) continue;
          double h = hs_at(hN, a, c);
          if (a == c) h += dw;
          acc += h * dcur[c];
        }
        pi[a] = acc;
      }
    }
    double a_pr = 1.0, a_du = 1.0, gsds = 0.0, gfd = 0.0, th = 0.0;
#pragma unroll 1
    for (int k = N - 1; k >= 0; --k) {
      if (k >= 1) {
        w.pf(k - 1, lw::HS, 32 + 7); w.pf(k - 1, lw::A, NA); w.pf(k - 1, lw::GC, 9); w.pf(k - 1, lw::CRES, 12);
        w.pf_rows(k - 1, lw::RS, 5);
      }
      if (k >= 2) w.pf(k - 2, lw::Z3, 27);
      const unsigned m = rowmask(k);
      // new multipliers of the dynamics rows of stage k: -pi_{k+1}
      for (int a = 0; a < 6; ++a) w.s(k, lw::NLAM + a) = -pi[a];
      if (k >= 1) load_combined(k - 1, c1, c2, dprev);
      else
        for (int a = 0; a < 9; ++a) dprev[a] = 0.0;
      // ---- store the step of this stage ---------------------------------------------------
      for (int a = 0; a < 9; ++a) w.s(k, lw::DX + a) = dcur[a];
      // ---- costate recursion: pi_k = loc_k + A_k' pi_{k+1} -----------------------------------
      {
        double hs[32], A[kLinRk4];
        for (int a = 0; a < 32; ++a) hs[a] = w.s(k, lw::HS + a);
        for (int a = 0; a < NA; ++a) A[a] = w.s(k, lw::A + a);
        const double v12[12] = {dprev[0], dprev[1], dprev[2], dprev[3], dprev[4], dprev[5], dprev[6],
                                0.0, 0.0, dcur[7], dcur[8], 0.0};
        double qv[12];
        {
          const double pi9[9] = {pi[0], pi[1], pi[2], pi[3], pi[4], pi[5], pi[6], 0.0, 0.0};
          mulMx<IG>(pi9, A, Ts, qv);       // A_k' pi_{k+1}; entries 0..6 are used
        }
#pragma unroll
        for (int a = 0; a < 7; ++a) {
          double acc = w.s(k, lw::GS + a);
#pragma unroll
          for (int c = 0; c < 11; ++c) {
            if (c == 7 || c == 8) continue;
            if (hs_slot(a, c) < 0) continue;
            double h = hs_at(hs, a, c);
            if (a == c) h += dw;       // a <= 6: flagged diagonal
            acc += h * v12[c];
          }
          pi[a] = acc + qv[a];
        }
      }
      // ---- slack steps, fraction to the boundary, dual steps ---------------------------------
      {
        double dup[2];
        if (k >= 1) { dup[0] = dprev[7]; dup[1] = dprev[8]; }
        else { dup[0] = dlast_u[0]; dup[1] = dlast_u[1]; }
        const double dshp = (k >= 1) ? dprev[6] : 0.0;
        double gC[3];
        for (int a = 0; a < 3; ++a) gC[a] = w.s(k, lw::GCZ + a);
        for (int a = 0; a < 6; ++a) th += fabs(w.s(k, lw::DCS + a) * w.s(k, lw::CRES + a));
#pragma unroll 1
        for (int r = 0; r < kNumRows; ++r) {
          double dsr = 0.0;
          if (on(m, r)) {
            const double rdv = (rsel == 0) ? 0.0 : w.s(k, (rsel == 1 ? lw::RRD : lw::DSOC) + r);
            const double jdx = row_lin(r, dcur + 7, dup, dcur[6], dshp, dcur, gC);
            const double s = w.s(k, lw::RS + r);
            const double vL = row_hasL(r) ? w.s(k, lw::RVL + r) : 0.0;
            const double vU = row_hasU(r) ? w.s(k, lw::RVU + r) : 0.0;
            double sg, gb;
            sigma_gsb(r, s, vL, vU, sg, gb);
            dsr = jdx + rdv;
            if (resto) {
              const double den = wq + sg + dw;
              dsr = (wq * (jdx + rdv) - gb) / den;
              // directional derivative of wq/2 (d - s)^2 along (dx, ds)
              gfd += wq * w.s(k, lw::RRD + r) * (jdx - dsr);
            } else {
              th += fabs(w.s(k, lw::RRD + r));
            }
            if (row_hasL(r)) {
              const double sl = s - lo(r);
              if (dsr < 0.0) a_pr = fmin(a_pr, -tau * sl / dsr);
              gsds += (-mu / sl + (row_hasU(r) ? 0.0 : o.kappa_d * mu)) * dsr;
              const double dv = mu / sl - vL - vL / sl * dsr;
              if (dv < 0.0) a_du = fmin(a_du, -tau * vL / dv);
            }
            if (row_hasU(r)) {
              const double su = hi(r) - s;
              if (dsr > 0.0) a_pr = fmin(a_pr, tau * su / dsr);
              gsds += (mu / su - (row_hasL(r) ? 0.0 : o.kappa_d * mu)) * dsr;
              const double dv = mu / su - vU + vU / su * dsr;
              if (dv < 0.0) a_du = fmin(a_du, -tau * vU / dv);
            }
            w.s(k, lw::RDYD + r) = resto ? 0.0 : ((sg + dw) * dsr + gb - w.s(k, lw::RYD + r));
          }
          w.s(k, lw::RDSR + r) = dsr;
        }
        // grad f' d
        if (resto) {
          for (int a = 0; a < 9; ++a) gfd += w.s(k, lw::PG + a) * dcur[a];   // (x, s_hat, u) order = step order
        } else {
          gfd += w.s(k, lw::GC + 0) * dcur[0] + w.s(k, lw::GC + 1) * dcur[1] + w.s(k, lw::GC + 2) * dcur[6] +
                 w.s(k, lw::GC + 3) * dcur[3] + w.s(k, lw::GC + 4) * dcur[4];
          if (k >= 1) gfd += 2.0 * p.beta_delta * df * w.s(k, lw::SIGC) * (dcur[8] - dprev[8]);
        }
      }
      for (int a = 0; a < 9; ++a) dcur[a] = dprev[a];
    }
    for (int a = 0; a < 7; ++a) newy0[a] = -pi[a];
    si.a_pr = a_pr;
    si.a_du = a_du;
    si.gs_ds = gsds;
    si.gf_d = gfd;
    si.theta = th;
  }

  // -------------------------------------------------------------------------
  // Pass T (ascending): objective / constraints at x + at * d (WarpSolver::eval_point);
  // the trial residuals go to TC / RTRD.
  // -------------------------------------------------------------------------
  struct Trial { double f, theta, barB; bool finite; };

  // inner: restoration objective (rows + proximity term) and theta without the rows
  MPCR_D void pass_trial(double at, Trial& t, bool inner, double dfu) {
    const double Ts = cfg.Ts;
    double xp[6], shp = pb.s0, up[2];
    for (int a = 0; a < 6; ++a) xp[a] = pb.x0[a];
    up[0] = w.s(N - 1, lw::U) + at * w.s(N - 1, lw::DX + 7);
    up[1] = w.s(N - 1, lw::U + 1) + at * w.s(N - 1, lw::DX + 8);
    double fsum = 0.0, th = 0.0, bar = 0.0;
    bool fin = true;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      if (k + 1 < N) { w.pf(k + 1, lw::X, 9); w.pf(k + 1, lw::DCS, 15); w.pf_rows(k + 1, lw::RS, 1); w.pf_rows(k + 1, lw::RDSR, 1); }
      const unsigned m = rowmask(k);
      double tx[6], tu[2];
      for (int a = 0; a < 6; ++a) tx[a] = w.s(k, lw::X + a) + at * w.s(k, lw::DX + a);
      const double tsh = w.s(k, lw::SH) + at * w.s(k, lw::DX + 6);
      tu[0] = w.s(k, lw::U) + at * w.s(k, lw::DX + 7);
      tu[1] = w.s(k, lw::U + 1) + at * w.s(k, lw::DX + 8);
      double f[6];
      lane_step<MODEL, IG>(xp, tu, Ts, f);
      StageCost sc;
      stage_cost(p, pb.cx, pb.cy, tx, tsh, 0, sc);
      double fl = 0.0;
      if (!inner) {
        fl = sc.val;
        if (k == N - 1) fl -= p.lambda_s * tsh;
        if (k >= 1) {
          const double dsg = tu[1] - up[1];
          fl += p.beta_delta * dsg * dsg;
        }
      } else {
        const double v9[9] = {tx[0], tx[1], tx[2], tx[3], tx[4], tx[5], tsh, tu[0], tu[1]};
        for (int a = 0; a < 9; ++a) {
          const double ref = w.s(k, lw::XR + a);
          const double dr = 1.0 / fmax(1.0, fabs(ref));
          fl += 0.5 * zeta * dr * dr * (v9[a] - ref) * (v9[a] - ref);
        }
      }
      double thk = 0.0;
      for (int a = 0; a < 6; ++a) {
        const double c = tx[a] - f[a];
        w.s(k, lw::TC + a) = c;
        thk += fabs(w.s(k, lw::DCS + a) * c);
      }
      double prod = 1.0, lin = 0.0;
#pragma unroll 1
      for (int r = 0; r < kNumRows; ++r) {
        double rdv = 0.0;
        if (on(m, r)) {
          const double sv = w.s(k, lw::RS + r) + at * w.s(k, lw::RDSR + r);
          rdv = row_value(r, tu, up, tsh, shp, sc.eC) - sv;
          if (inner) fl += 0.5 * wq * rdv * rdv;
          else thk += fabs(rdv);
          if (row_hasL(r)) {
            const double sl = sv - lo(r);
            prod *= sl;
            if (!row_hasU(r)) lin += sl;
            if (!(sl > 0.0)) fin = false;
          }
          if (row_hasU(r)) {
            const double su = hi(r) - sv;
            prod *= su;
            if (!row_hasL(r)) lin += su;
            if (!(su > 0.0)) fin = false;
          }
        }
        w.s(k, lw::RTRD + r) = rdv;
      }
      const double bk = -log_d(prod) + o.kappa_d * lin;
      fin = fin && finite_d(fl) && finite_d(thk) && finite_d(bk);
      fsum += fl;
      th += thk;
      bar += bk;
      for (int a = 0; a < 6; ++a) xp[a] = tx[a];
      shp = tsh;
      up[0] = tu[0];
      up[1] = tu[1];
    }
    t.f = dfu * fsum;
    t.theta = th;
    t.barB = bar;
    t.finite = fin;
  }

  // filter ----------------------------------------------------------------------
  MPCR_D bool filter_ok(double th, double ph) const {
    const int n = filt_n < 32 ? filt_n : 32;
    bool ok = true;
#pragma unroll 1
    for (int i = 0; i < n; ++i)
      ok = ok && ((th <= w.t(N, fbase_th + i)) || (ph <= w.t(N, fbase_ph + i)));
    return ok;
  }
  MPCR_D void filter_add(double th, double ph) {
    w.t(N, fbase_th + (filt_n & 31)) = th;
    w.t(N, fbase_ph + (filt_n & 31)) = ph;
    filt_n++;
  }
  MPCR_D static bool cmp_le(double a, double b, double ref) {
    return a - b <= 10.0 * 2.220446049250313e-16 * fabs(ref);
  }

  // save / restore the plain step around a second-order correction
  MPCR_D void save_step() {
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      for (int a = 0; a < 9; ++a) w.s(k, lw::SVD + a) = w.s(k, lw::DX + a);
      for (int a = 0; a < 6; ++a) w.s(k, lw::SVLAM + a) = w.s(k, lw::NLAM + a);
#pragma unroll 1
      for (int r = 0; r < kNumRows; ++r) {
        w.s(k, lw::SVDSR + r) = w.s(k, lw::RDSR + r);
        w.s(k, lw::SVDYD + r) = w.s(k, lw::RDYD + r);
      }
    }
    for (int a = 0; a < 7; ++a) sv_y0[a] = newy0[a];
  }
  MPCR_D void restore_step() {
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      for (int a = 0; a < 9; ++a) w.s(k, lw::DX + a) = w.s(k, lw::SVD + a);
      for (int a = 0; a < 6; ++a) w.s(k, lw::NLAM + a) = w.s(k, lw::SVLAM + a);
#pragma unroll 1
      for (int r = 0; r < kNumRows; ++r) {
        w.s(k, lw::RDSR + r) = w.s(k, lw::SVDSR + r);
        w.s(k, lw::RDYD + r) = w.s(k, lw::SVDYD + r);
      }
    }
    for (int a = 0; a < 7; ++a) newy0[a] = sv_y0[a];
  }

  // -------------------------------------------------------------------------
  MPCR_D void solve(const SolveOut& out) {
    const double inf = 1.0e300;
    init_iterate();
    df = 1.0;
    ddl = 1.0;
    mu = o.mu_init;
    tau = fmax(o.tau_min, 1.0 - mu);
    filt_n = 0;
    fval = 0.0;
    barB = 0.0;
    resto = false;
    wq = o.resto_weight;
    zeta = 0.0;
    fbase_th = lw::T_FTH;
    fbase_ph = lw::T_FPH;
    Dw[0] = Dw[1] = 0.0;
    for (int r = 0; r < kNumRows; ++r) { lo_[r] = -inf; hi_[r] = inf; }

    int phase = kPhInit;
    int status = kMaxIterExceeded;
    int it = 0, acc_count = 0, n_ls = 0, soc_count = 0, ls_fail_mode = 0;
    bool cur_finite = true, soc_done = false, f_type = false;
    double theta_max = 0.0, theta_min = 0.0, mu_floor = 0.0;
    double dw = 0.0, dw_last = 0.0, last_alpha = 0.0, last_dw = 0.0;
    double alpha = 0.0, a_soc = 0.0, amin = 0.0, gpd = 0.0, theta = 0.0, phi = 0.0, th_old = 0.0;
    double a_du = 1.0, sv_a_du = 1.0;      // dual fraction-to-the-boundary step of the current / saved direction
    Trial tp;
    tp.f = tp.theta = tp.barB = 0.0;
    tp.finite = true;
    // restoration bookkeeping: the outer problem's scalars while the restoration phase runs
    bool want_resto = false, reinit = false;
    int resto_it = 0, filt_n_o = 0, acc_o = 0;
    double mu_o = 0.0, tau_o = 0.0, df_o = 1.0, theta_max_o = 0.0, theta_min_o = 0.0, mu_floor_o = 0.0;
    double theta_R = 0.0;

    while (true) {
      if (want_resto) {
        // ---- enter the restoration phase at the current iterate (x_R, s_R) ---------------------
        want_resto = false;
        theta_R = theta;
        filter_add((1.0 - o.gamma_theta) * theta, phi - o.gamma_phi * theta);   // IPOPT: the start point enters the filter
        mu_o = mu; tau_o = tau; df_o = df; theta_max_o = theta_max; theta_min_o = theta_min;
        mu_floor_o = mu_floor; filt_n_o = filt_n; acc_o = acc_count;
        double vmax = 0.0;
#pragma unroll 1
        for (int k = 0; k < N; ++k) {
          const unsigned m = rowmask(k);
          for (int a = 0; a < 6; ++a) {
            w.s(k, lw::XR + a) = w.s(k, lw::X + a);
            vmax = fmax(vmax, fabs(w.s(k, lw::DCS + a) * w.s(k, lw::CRES + a)));
            w.s(k, lw::LAM + a) = 0.0;
          }
          w.s(k, lw::XR + 6) = w.s(k, lw::SH);
          w.s(k, lw::XR + 7) = w.s(k, lw::U);
          w.s(k, lw::XR + 8) = w.s(k, lw::U + 1);
          for (int a = 0; a < 9; ++a) w.s(k, lw::DX + a) = 0.0;
#pragma unroll 1
          for (int r = 0; r < kNumRows; ++r) {
            w.s(k, lw::RDSR + r) = 0.0;
            if (!on(m, r)) continue;
            const double rd0 = w.s(k, lw::RRD + r);
            vmax = fmax(vmax, fabs(rd0));
            const double dval = rd0 + w.s(k, lw::RS + r);
            double v = w.s(k, lw::RS + r);
            const double l = lo(r), h = hi(r);
            double pl, pu;
            if (row_hasL(r) && row_hasU(r)) {
              pl = fmin(o.bound_push * fmax(1.0, fabs(l)), o.bound_frac * (h - l));
              pu = fmin(o.bound_push * fmax(1.0, fabs(h)), o.bound_frac * (h - l));
            } else {
              pl = o.bound_push * fmax(1.0, fabs(l));
              pu = o.bound_push * fmax(1.0, fabs(h));
            }
            if (row_hasL(r)) v = fmax(v, l + pl);
            if (row_hasU(r)) v = fmin(v, h - pu);
            w.s(k, lw::RS + r) = v;
            w.s(k, lw::RRD + r) = dval - v;
            w.s(k, lw::RYD + r) = wq * (dval - v);
            w.s(k, lw::RVL + r) = row_hasL(r) ? o.bound_mult_init_val : 0.0;
            w.s(k, lw::RVU + r) = row_hasU(r) ? o.bound_mult_init_val : 0.0;
          }
        }
        for (int a = 0; a < 7; ++a) y0[a] = 0.0;
        mu = fmax(mu_o, vmax);
        tau = fmax(o.tau_min, 1.0 - mu);
        zeta = sqrt(mu);
        df = 1.0;
        mu_floor = fmax(o.mu_min, fmin(o.tol, o.compl_inf_tol) / (o.barrier_tol_factor + 1.0));
        fbase_th = lw::T_FTH2;
        fbase_ph = lw::T_FPH2;
        filt_n = 0;
        acc_count = 0;
        resto = true;
        resto_it = 0;
        alpha = 0.0;
        phase = kPhInit2;          // evaluate the start point with the restoration objective
      }
      // ================= A. derivatives, convergence tests, barrier update ===============
      if (resto && phase == kPhNewton && resto_it >= 1) {
        // ---- leave the restoration phase?  The point must be acceptable to the ORIGINAL filter and
        //      have reduced the original infeasibility by kappa_resto ------------------------------
        Trial to;
        pass_trial(0.0, to, false, df_o);
        bool leave = to.finite && to.theta <= o.kappa_resto * theta_R && to.theta <= theta_max_o;
        if (leave) {
          const double ph_o = to.f + mu_o * to.barB;
          const int n = filt_n_o < 32 ? filt_n_o : 32;
#pragma unroll 1
          for (int i = 0; i < n; ++i)
            leave = leave && ((to.theta <= w.t(N, lw::T_FTH + i)) || (ph_o <= w.t(N, lw::T_FPH + i)));
        }
        if (leave) {
          resto = false;
          mu = mu_o; tau = tau_o; df = df_o; theta_max = theta_max_o; theta_min = theta_min_o;
          mu_floor = mu_floor_o; filt_n = filt_n_o; acc_count = acc_o;
          fbase_th = lw::T_FTH;
          fbase_ph = lw::T_FPH;
#pragma unroll 1
          for (int k = 0; k < N; ++k) {
            const unsigned m = rowmask(k);
#pragma unroll 1
            for (int r = 0; r < kNumRows; ++r) {
              w.s(k, lw::RVL + r) = (on(m, r) && row_hasL(r)) ? o.bound_mult_init_val : 0.0;
              w.s(k, lw::RVU + r) = (on(m, r) && row_hasU(r)) ? o.bound_mult_init_val : 0.0;
            }
          }
          reinit = true;           // least-squares multipliers at the restored point, then resume
          phase = kPhInit;
        } else if (resto_it >= o.resto_max_iter) {
          status = kRestorationNeeded;
          break;
        }
      }
      if (phase == kPhInit || phase == kPhNewton) {
        Errors e;
        double gmax = 0.0, ddlmin = 1.0;
        pass_derivs(phase == kPhInit ? 1 : 0, e, gmax, ddlmin, reinit);
        if (phase == kPhInit && reinit) {
          // scaling, bounds and mu of the outer problem stay as they are
        } else if (phase == kPhInit) {
          if (gmax > o.nlp_scaling_max_gradient)
            df = fmax(o.nlp_scaling_max_gradient / gmax, o.nlp_scaling_min_value);
          ddl = ddlmin;
          pass_rescale();
          mu_floor = fmax(o.mu_min, fmin(o.tol, o.compl_inf_tol * df) / (o.barrier_tol_factor + 1.0));
          for (int r = 0; r < kNumRows; ++r) {
            double l, h;
            row_bounds(r, l, h);
            if (r == kLane) { l *= ddl; h *= ddl; }
            if (row_hasL(r)) l -= fmin(o.constr_viol_tol, o.bound_relax_factor * fmax(1.0, fabs(l)));
            else l = -inf;
            if (row_hasU(r)) h += fmin(o.constr_viol_tol, o.bound_relax_factor * fmax(1.0, fabs(h)));
            else h = inf;
            lo_[r] = l;
            hi_[r] = h;
          }
        } else {
          if (!cur_finite) { status = kInvalidNumber; break; }
          bool done = false;
          const double E0 = e.E(0.0), cmp0 = e.compl_inf(0.0);
          const double viol = e.viol;
          if (out.trace) {
            double* tr = out.trace + (size_t)it * kTraceCols;
            tr[0] = fval / df; tr[1] = E0; tr[2] = e.pr_inf; tr[3] = e.dual_inf;
            tr[4] = cmp0; tr[5] = mu; tr[6] = last_alpha; tr[7] = last_dw;
          }
          if (!finite_d(E0)) {
            status = kInvalidNumber;
            done = true;
          } else if (E0 <= o.tol && e.dual_inf / df <= o.dual_inf_tol && (resto || viol <= o.constr_viol_tol) &&
                     cmp0 / df <= o.compl_inf_tol) {
            // restoration: a stationary point of the infeasibility measure that the filter rejects
            status = resto ? kRestorationNeeded : kSolveSucceeded;
            done = true;
          } else {
            if (!resto && E0 <= o.acceptable_tol && e.dual_inf / df <= o.acceptable_dual_inf_tol &&
                viol <= o.acceptable_constr_viol_tol && cmp0 / df <= o.acceptable_compl_inf_tol) {
              acc_count++;
              if (acc_count >= o.acceptable_iter) {
                status = kSolvedToAcceptableLevel;
                done = true;
              }
            } else {
              acc_count = 0;
            }
            // almost feasible: no restoration phase; the acceptable point (this iterate) or failure
            ls_fail_mode = (!resto && e.pr_inf <= 1e-2 * o.tol) ? (acc_count > 0 ? 1 : 2) : 0;
            if (!done && it >= o.max_iter) { status = kMaxIterExceeded; done = true; }
          }
          while (!done && e.E(mu) <= o.barrier_tol_factor * mu && mu > mu_floor) {
            mu = fmax(mu_floor, fmin(o.mu_linear_decrease_factor * mu,
                                     powd(mu, o.mu_superlinear_decrease_power)));
            tau = fmax(o.tau_min, 1.0 - mu);
            filt_n = 0;
          }
          if (done) break;
          dw = 0.0;
        }
      }
      // ================= B. linear solve (single call site) ===================================
      if (phase == kPhInit || phase == kPhNewton || phase == kPhSoc) {
        const int mode = (phase == kPhInit) ? 1 : 0;
        const int sel = (phase == kPhInit) ? 0 : ((phase == kPhSoc) ? 2 : 1);
        double dwu = (phase == kPhInit) ? 1.0 : dw;
        bool first_try = true, ok = false, failed = false;
#pragma unroll 1
        while (true) {
          ok = pass_riccati(mode, dwu, sel, sel);
          if (ok || phase != kPhNewton) break;
          if (first_try) {
            dw = (dw_last == 0.0) ? o.delta_w_init : fmax(o.delta_w_min, dw_last * o.kappa_w_minus);
            first_try = false;
          } else {
            dw *= (dw_last == 0.0 || 1.0e5 * dw_last < dw) ? o.kappa_w_plus_first : o.kappa_w_plus;
          }
          if (dw > o.delta_w_max) { failed = true; break; }
          dwu = dw;
        }
        if (failed) { status = kErrorInStepComputation; break; }
        StepInfo si;
        si.a_pr = 1.0; si.a_du = 1.0; si.gs_ds = 0.0; si.gf_d = 0.0; si.theta = 0.0;
        if (ok) {
          double mwa[kRhs][2], mwb[kRhs][2];
          for (int j = 0; j < kRhs; ++j) { mwa[j][0] = mwa[j][1] = mwb[j][0] = mwb[j][1] = 0.0; }
          pass_forward(sel, mwa, mwb);
          double c1 = 0.0, c2 = 0.0;
          if (cfg.wrap) {
            double t0[2], Mm[2][2];
            for (int i = 0; i < 2; ++i) {
              t0[i] = mwa[0][i] - mwb[0][i];
              for (int kk = 0; kk < 2; ++kk) Mm[i][kk] = mwa[1 + kk][i] - mwb[1 + kk][i];
            }
            const double a00 = 1.0 - Mm[0][0] * Dw[0], a01 = -Mm[0][1] * Dw[1];
            const double a10 = -Mm[1][0] * Dw[0], a11 = 1.0 - Mm[1][1] * Dw[1];
            const double det = a00 * a11 - a01 * a10;
            const double t1 = (a11 * t0[0] - a01 * t0[1]) / det;
            const double t2 = (-a10 * t0[0] + a00 * t0[1]) / det;
            c1 = Dw[0] * t1;
            c2 = Dw[1] * t2;
          }
          pass_combine(c1, c2, dwu, sel, si);
        }
        if (phase == kPhInit) {
          // ---- least-squares multipliers: y = -(QP multipliers), yd = (vU - vL) - J_d dx ----
          double ymax = 0.0;
          if (ok) {
#pragma unroll 1
            for (int k = 0; k < N; ++k) {
              const unsigned m = rowmask(k);
              for (int a = 0; a < 6; ++a) {
                const double l = -w.s(k, lw::NLAM + a);
                w.s(k, lw::LAM + a) = l;
                ymax = fmax(ymax, fabs(l / w.s(k, lw::DCS + a)));
              }
#pragma unroll 1
              for (int r = 0; r < kNumRows; ++r) {
                if (!on(m, r)) continue;
                const double yd = ((row_hasU(r) ? w.s(k, lw::RVU + r) : 0.0) -
                                   (row_hasL(r) ? w.s(k, lw::RVL + r) : 0.0)) - w.s(k, lw::RDSR + r);
                w.s(k, lw::RYD + r) = yd;
                ymax = fmax(ymax, fabs(yd));
              }
            }
            for (int a = 0; a < 7; ++a) { y0[a] = -newy0[a]; ymax = fmax(ymax, fabs(y0[a])); }
          }
          const bool discard = !ok || !(ymax <= o.constr_mult_init_max);
          if (discard)
            for (int a = 0; a < 7; ++a) y0[a] = 0.0;
#pragma unroll 1
          for (int k = 0; k < N; ++k) {
            for (int a = 0; a < 9; ++a) w.s(k, lw::DX + a) = 0.0;   // the init phases evaluate the point itself
#pragma unroll 1
            for (int r = 0; r < kNumRows; ++r) {
              w.s(k, lw::RDSR + r) = 0.0;
              if (discard) w.s(k, lw::RYD + r) = 0.0;
            }
            if (discard)
              for (int a = 0; a < 6; ++a) w.s(k, lw::LAM + a) = 0.0;
          }
          alpha = 0.0;
        } else if (phase == kPhNewton) {
          if (dw > 0.0) dw_last = dw;
          last_dw = dw;
          alpha = si.a_pr;
          a_du = si.a_du;
          gpd = si.gs_ds + si.gf_d;
          theta = si.theta;
          phi = fval + mu * barB;
          if (gpd < 0.0) {
            amin = fmin(o.gamma_theta, o.gamma_phi * theta / (-gpd));
            if (theta <= theta_min)
              amin = fmin(amin, o.delta * powd(theta, o.s_theta) / powd(-gpd, o.s_phi));
          } else {
            amin = o.gamma_theta;
          }
          amin *= o.alpha_min_frac;
          n_ls = 0;
          soc_done = false;
        } else {   // kPhSoc
          if (!ok) {
            restore_step();
            a_du = sv_a_du;
            alpha *= o.alpha_red_factor;
            phase = kPhBacktrack;
            if (!(alpha >= amin && n_ls < 60)) {
              if (o.restoration && !resto && ls_fail_mode == 0) { want_resto = true; continue; }
              status = ls_fail_mode == 1 ? kSolvedToAcceptableLevel : kRestorationNeeded;
              break;
            }
          } else {
            a_soc = si.a_pr;
            a_du = si.a_du;
          }
        }
      }
      // ================= C. trial point (single call site) =======================================
      pass_trial((phase == kPhSoc) ? a_soc : alpha, tp, resto, df);
      // ================= D. decisions ===========================================================
      if (phase == kPhInit && reinit) {
        phase = kPhInit2;          // the slacks of the restored point are kept
        continue;
      }
      if (phase == kPhInit) {
        // slacks pushed into their (relaxed) bounds; RTRD holds the row values (s was 0)
#pragma unroll 1
        for (int k = 0; k < N; ++k) {
          const unsigned m = rowmask(k);
#pragma unroll 1
          for (int r = 0; r < kNumRows; ++r) {
            if (!on(m, r)) { w.s(k, lw::RS + r) = 0.0; continue; }
            double v = w.s(k, lw::RTRD + r);
            const double l = lo(r), h = hi(r);
            double pl, pu;
            if (row_hasL(r) && row_hasU(r)) {
              pl = fmin(o.bound_push * fmax(1.0, fabs(l)), o.bound_frac * (h - l));
              pu = fmin(o.bound_push * fmax(1.0, fabs(h)), o.bound_frac * (h - l));
            } else {
              pl = o.bound_push * fmax(1.0, fabs(l));
              pu = o.bound_push * fmax(1.0, fabs(h));
            }
            if (row_hasL(r)) v = fmax(v, l + pl);
            if (row_hasU(r)) v = fmin(v, h - pu);
            w.s(k, lw::RS + r) = v;
          }
        }
        phase = kPhInit2;
        continue;
      }
      if (phase == kPhInit2) {
#pragma unroll 1
        for (int k = 0; k < N; ++k) {
          for (int a = 0; a < 6; ++a) w.s(k, lw::CRES + a) = w.s(k, lw::TC + a);
#pragma unroll 1
          for (int r = 0; r < kNumRows; ++r) w.s(k, lw::RRD + r) = w.s(k, lw::RTRD + r);
        }
        fval = tp.f;
        barB = tp.barB;
        cur_finite = tp.finite;
        if (!reinit) {
          theta_max = o.theta_max_fact * fmax(1.0, tp.theta);
          theta_min = o.theta_min_fact * fmax(1.0, tp.theta);
        }
        reinit = false;
        phase = kPhNewton;
        continue;
      }
      // ---- acceptance test of the trial point (plain or SOC step) --------------------------------
      n_ls += (phase == kPhSoc) ? 0 : 1;
      const double tphi = tp.f + mu * tp.barB;
      bool ok = tp.finite && tp.theta <= theta_max;
      if (ok) {
        const bool switching = gpd < 0.0 &&
            alpha * powd(-gpd, o.s_phi) > o.delta * powd(theta, o.s_theta);
        if (theta <= theta_min && switching) {
          ok = cmp_le(tphi - phi, o.eta_phi * alpha * gpd, phi);
          f_type = true;
        } else {
          ok = cmp_le(tp.theta, (1.0 - o.gamma_theta) * theta, theta) ||
               cmp_le(tphi - phi, -o.gamma_phi * theta, phi);
          f_type = false;
        }
        if (ok) ok = filter_ok(tp.theta, tphi);
      }
      if (!ok) {
        if (phase == kPhSoc) {
          soc_count++;
          if (!tp.finite || tp.theta > theta_max || tp.theta > o.kappa_soc * th_old ||
              soc_count >= o.max_soc) {
            restore_step();
            a_du = sv_a_du;
            alpha *= o.alpha_red_factor;
            phase = kPhBacktrack;
            if (!(alpha >= amin && n_ls < 60)) {
              if (o.restoration && !resto && ls_fail_mode == 0) { want_resto = true; continue; }
              status = ls_fail_mode == 1 ? kSolvedToAcceptableLevel : kRestorationNeeded;
              break;
            }
          } else {
            th_old = tp.theta;
#pragma unroll 1
            for (int k = 0; k < N; ++k) {
              for (int a = 0; a < 6; ++a)
                w.s(k, lw::CSOC + a) = a_soc * w.s(k, lw::CSOC + a) + w.s(k, lw::TC + a);
#pragma unroll 1
              for (int r = 0; r < kNumRows; ++r)
                w.s(k, lw::DSOC + r) = a_soc * w.s(k, lw::DSOC + r) + w.s(k, lw::RTRD + r);
            }
          }
          continue;
        }
        if (!soc_done && n_ls == 1 && tp.finite && tp.theta >= theta && o.max_soc > 0) {
          soc_done = true;
          soc_count = 0;
          th_old = theta;
#pragma unroll 1
          for (int k = 0; k < N; ++k) {
            for (int a = 0; a < 6; ++a)
              w.s(k, lw::CSOC + a) = alpha * w.s(k, lw::CRES + a) + w.s(k, lw::TC + a);
#pragma unroll 1
            for (int r = 0; r < kNumRows; ++r)
              w.s(k, lw::DSOC + r) = alpha * w.s(k, lw::RRD + r) + w.s(k, lw::RTRD + r);
          }
          save_step();
          sv_a_du = a_du;
          phase = kPhSoc;
          continue;
        }
        alpha *= o.alpha_red_factor;
        phase = kPhBacktrack;
        if (!(alpha >= amin && n_ls < 60)) {
              if (o.restoration && !resto && ls_fail_mode == 0) { want_resto = true; continue; }
              status = ls_fail_mode == 1 ? kSolvedToAcceptableLevel : kRestorationNeeded;
              break;
            }
        continue;
      }
      // ================= E. accept the trial point ===============================================
      if (!f_type) filter_add((1.0 - o.gamma_theta) * theta, phi - o.gamma_phi * theta);
      const double astep = (phase == kPhSoc) ? a_soc : alpha;
      const double az = a_du;
#pragma unroll 1
      for (int k = 0; k < N; ++k) {
        if (k + 1 < N) { w.pf(k + 1, lw::X, 48); w.pf_rows(k + 1, lw::RS, 4); w.pf_rows(k + 1, lw::RDSR, 3); }
        const unsigned m = rowmask(k);
        for (int a = 0; a < 6; ++a) {
          w.s(k, lw::X + a) += astep * w.s(k, lw::DX + a);
          w.s(k, lw::CRES + a) = w.s(k, lw::TC + a);
          const double l = w.s(k, lw::LAM + a);
          w.s(k, lw::LAM + a) = l + astep * (w.s(k, lw::NLAM + a) - l);
        }
        w.s(k, lw::SH) += astep * w.s(k, lw::DX + 6);
        w.s(k, lw::U) += astep * w.s(k, lw::DX + 7);
        w.s(k, lw::U + 1) += astep * w.s(k, lw::DX + 8);
#pragma unroll 1
        for (int r = 0; r < kNumRows; ++r) {
          if (!on(m, r)) continue;
          const double dsr = w.s(k, lw::RDSR + r);
          w.s(k, lw::RYD + r) += astep * w.s(k, lw::RDYD + r);
          const double sold = w.s(k, lw::RS + r);
          const double snew = sold + astep * dsr;
          if (row_hasL(r)) {
            const double slo = sold - lo(r), sl = snew - lo(r);
            const double vL = w.s(k, lw::RVL + r);
            const double dv = mu / slo - vL - vL / slo * dsr;
            const double v = vL + az * dv;
            w.s(k, lw::RVL + r) = fmax(fmin(v, o.kappa_sigma * mu / sl), mu / (o.kappa_sigma * sl));
          }
          if (row_hasU(r)) {
            const double suo = hi(r) - sold, su = hi(r) - snew;
            const double vU = w.s(k, lw::RVU + r);
            const double dv = mu / suo - vU + vU / suo * dsr;
            const double v = vU + az * dv;
            w.s(k, lw::RVU + r) = fmax(fmin(v, o.kappa_sigma * mu / su), mu / (o.kappa_sigma * su));
          }
          w.s(k, lw::RS + r) = snew;
          w.s(k, lw::RRD + r) = w.s(k, lw::RTRD + r);
          if (resto) w.s(k, lw::RYD + r) = wq * w.s(k, lw::RTRD + r);
        }
      }
      for (int a = 0; a < 7; ++a) y0[a] += astep * (newy0[a] - y0[a]);
      fval = tp.f;
      barB = tp.barB;
      cur_finite = tp.finite;
      last_alpha = astep;
      it++;
      if (resto) resto_it++;
      phase = kPhNewton;
    }

    // ---- outputs ---------------------------------------------------------------------------
    {
      double xp[6], shp = pb.s0;
      for (int a = 0; a < 6; ++a) { xp[a] = pb.x0[a]; out.X[a * (N + 1)] = pb.x0[a]; }
      out.S[0] = pb.s0;
      out.obj[0] = fval / df;
      out.status[0] = status;
      out.iters[0] = it;
      const int bs = 7 + (cfg.lane_con ? 1 : 0);
      const int base2 = 7 + N * bs;
      const double idf = 1.0 / df;
      if (out.lam_g) {
        out.lam_g[0] = y0[6] * idf;
        for (int a = 0; a < 6; ++a) out.lam_g[1 + a] = y0[a] * idf;
      }
#pragma unroll 1
      for (int k = 0; k < N; ++k) {
        const unsigned m = rowmask(k);
        Poly4 px, py;
        poly4_eval(pb.cx, shp, 1, px);
        poly4_eval(pb.cy, shp, 1, py);
        const double dX = xp[0] - px.p, dY = xp[1] - py.p;
        out.eC[k] = py.d1 * dX - px.d1 * dY;      // e_hat_C(S_hat[i], States[:, i]), i < N
        out.eL[k] = -px.d1 * dX - py.d1 * dY;
        out.U[k] = w.s(k, lw::U);
        out.U[N + k] = w.s(k, lw::U + 1);
        for (int a = 0; a < 6; ++a) {
          xp[a] = w.s(k, lw::X + a);
          out.X[a * (N + 1) + k + 1] = xp[a];
        }
        shp = w.s(k, lw::SH);
        out.S[k + 1] = shp;
        if (out.lam_g) {
          double* l1 = out.lam_g + 7 + k * bs;     // rows of stage i = k + 1
          for (int a = 0; a < 6; ++a) l1[a] = w.s(k, lw::LAM + a) * idf;
          l1[6] = w.s(k, lw::RYD + kProg) * idf;
          if (cfg.lane_con) l1[7] = w.s(k, lw::RYD + kLane) * ddl * idf;
          const int off = cfg.wrap ? 6 * k : (k == 0 ? 0 : 4 + 6 * (k - 1));
          double* l2 = out.lam_g + base2 + off;
          l2[0] = w.s(k, lw::RYD + kThrUb) * idf;
          l2[1] = w.s(k, lw::RYD + kThrLb) * idf;
          l2[2] = w.s(k, lw::RYD + kStUb) * idf;
          l2[3] = w.s(k, lw::RYD + kStLb) * idf;
          if (on(m, kDThr)) { l2[4] = w.s(k, lw::RYD + kDThr) * idf; l2[5] = w.s(k, lw::RYD + kDSt) * idf; }
          if (k == 0) {
            double* l3 = out.lam_g + base2 + (cfg.wrap ? 6 * N : 6 * N - 2);
            int qn = 0;
            if (on(m, kDThr0)) l3[qn++] = w.s(k, lw::RYD + kDThr0) * idf;
            if (on(m, kDSt0)) l3[qn++] = w.s(k, lw::RYD + kDSt0) * idf;
          }
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------
// One problem per calling thread: load inputs, run the solver, store the outputs.
// ---------------------------------------------------------------------------
template <int MODEL, int IG = 0>
MPCR_D void lane_solve_problem(const Config& cfg, const BatchIn& in, const BatchOut& out, int b,
                               int trace_rows, double* ws) {
  const int N = cfg.N;
  ProblemView pv;
  for (int k = 0; k < 6; ++k) pv.x0[k] = in.x0[(size_t)b * 6 + k];
  pv.u_prev[0] = in.u_prev[(size_t)b * 2 + 0];
  pv.u_prev[1] = in.u_prev[(size_t)b * 2 + 1];
  pv.has_prev_thr = pv.u_prev[0] == pv.u_prev[0];   // NaN = None (control/MPC.py:145-149)
  pv.has_prev_st = pv.u_prev[1] == pv.u_prev[1];
  pv.s0 = in.s0[b];
  for (int k = 0; k < 5; ++k) { pv.cx[k] = in.cx[(size_t)b * 5 + k]; pv.cy[k] = in.cy[(size_t)b * 5 + k]; }
  pv.max_err = in.max_err ? in.max_err[b] : 0.0;
  pv.u_warm = in.u_warm ? in.u_warm + (size_t)b * 2 * N : nullptr;
  if (!pv.has_prev_thr) pv.u_prev[0] = 0.0;
  if (!pv.has_prev_st) pv.u_prev[1] = 0.0;
  Params prm = cfg.p;
  if (in.params) {
    const double* q = in.params + (size_t)b * 5;
    prm.alpha_c = q[0];
    prm.d_max = q[1];
    prm.q_v_y = q[2];
    prm.n_exp = (int)(q[3] + 0.5);
    prm.beta_delta = q[4];
  }
  SolveOut so;
  so.U = out.U + (size_t)b * 2 * N;
  so.X = out.X + (size_t)b * 6 * (N + 1);
  so.S = out.S + (size_t)b * (N + 1);
  so.eC = out.eC + (size_t)b * N;
  so.eL = out.eL + (size_t)b * N;
  so.lam_g = out.lam_g ? out.lam_g + (size_t)b * lam_g_stride(cfg) : nullptr;
  so.obj = out.obj + b;
  so.status = out.status + b;
  so.iters = out.iters + b;
  so.trace = out.trace ? out.trace + (size_t)b * trace_rows * kTraceCols : nullptr;
  if (so.lam_g) {
    const int ls = lam_g_stride(cfg);
    for (int k = 0; k < ls; ++k) so.lam_g[k] = 0.0;
  }
  LaneSolver<MODEL, IG> solver(cfg, prm, pv, ws);
  solver.solve(so);
}

MPCR_D void lane_solve_problem_any(const Config& cfg, const BatchIn& in, const BatchOut& out, int b,
                                   int trace_rows, double* ws) {
  if (cfg.integrator == 0) {
    if (cfg.model == kDynamic) lane_solve_problem<kDynamic, 0>(cfg, in, out, b, trace_rows, ws);
    else if (cfg.model == kKinematic) lane_solve_problem<kKinematic, 0>(cfg, in, out, b, trace_rows, ws);
    else lane_solve_problem<kBlended, 0>(cfg, in, out, b, trace_rows, ws);
  } else {
    if (cfg.model == kDynamic) lane_solve_problem<kDynamic, 1>(cfg, in, out, b, trace_rows, ws);
    else if (cfg.model == kKinematic) lane_solve_problem<kKinematic, 1>(cfg, in, out, b, trace_rows, ws);
    else lane_solve_problem<kBlended, 1>(cfg, in, out, b, trace_rows, ws);
  }
}

}  // namespace mpcr
