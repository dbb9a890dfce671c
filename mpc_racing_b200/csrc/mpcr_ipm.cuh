// One-problem-per-warp primal-dual interior-point solver for the racing NLP of
// control/MPC.py (reference), replacing CasADi Opti + IPOPT + MUMPS.
//
//   * lane k (0..N-1) owns stage transition k: input u_k, successor state
//     (x_{k+1}, s_{k+1}), the dynamics multipliers of rows x_{k+1} = f(x_k,u_k),
//     and the slack rows of u_k / of the progress increment s_{k+1} - s_k;
//   * the KKT system of every iteration is solved by a stage-wise Riccati
//     recursion on the augmented state z = (x, s_hat, u_prev) in R^9 with input
//     w = (thr, steer, ds) in R^3; the two wrap-around rate rows of
//     control/MPC.py:142-143 (i = 0) couple stage 0 with stage N-1 and are
//     handled as a rank-2 (Woodbury) correction with two extra right-hand sides;
//   * algorithm, constants and option defaults: IPOPT (Waechter & Biegler 2006),
//     the same restatement as oracle/ipm.py.
//
// The file is written against a tiny set of warp primitives (lane_id, shfl,
// syncwarp) so that tests/emu can run the identical code on the CPU with a
// fiber-based SIMT emulator (test infrastructure only - the product path is the
// CUDA build).
#pragma once
#include "mpcr_model.cuh"
#include "mpcr_rk4.cuh"

namespace mpcr {

// ---------------------------------------------------------------------------
// warp primitives
// ---------------------------------------------------------------------------
#if defined(MPCR_EMU)
MPCR_D int lane_id() { return emu::lane(); }
MPCR_D double shfl(double v, int src) { return emu::shfl_f64(v, src); }
MPCR_D int shfl_i(int v, int src) { return emu::shfl_i32(v, src); }
MPCR_D void syncwarp() { emu::syncwarp(); }
MPCR_D double shfl_xor(double v, int m) { return emu::shfl_f64(v, emu::lane() ^ m); }
MPCR_D int shfl_xor_i(int v, int m) { return emu::shfl_i32(v, emu::lane() ^ m); }
#else
MPCR_D int lane_id() { return threadIdx.x & 31; }
MPCR_D double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
MPCR_D int shfl_i(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
MPCR_D void syncwarp() { __syncwarp(); }
MPCR_D double shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
MPCR_D int shfl_xor_i(int v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
#endif

#if defined(__CUDACC__)
#define MPCR_NOINLINE static __device__ __noinline__
#else
#define MPCR_NOINLINE inline
#endif

MPCR_NOINLINE double powd(double x, double y) { return pow(x, y); }

// Reciprocal of a Riccati pivot / of a slack: hardware seed (rcp.approx.ftz.f64, 2^-23) + two Newton steps, about
// 1 ulp, 5 instructions and no special-case branch (1.0 / x is ~10 instructions with a slow-path
// call; three of them sit on the serial critical path of every stage).  Pivots that are not
// positive are rejected before any of these values is used.
MPCR_D double rcp_pivot(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  return r;
#else
  return 1.0 / x;
#endif
}

MPCR_NOINLINE double warp_sum(double v) {
  for (int m = 16; m > 0; m >>= 1) v += shfl_xor(v, m);
  return v;
}
MPCR_NOINLINE double warp_max(double v) {
  for (int m = 16; m > 0; m >>= 1) v = fmax(v, shfl_xor(v, m));
  return v;
}
MPCR_NOINLINE double warp_min(double v) {
  for (int m = 16; m > 0; m >>= 1) v = fmin(v, shfl_xor(v, m));
  return v;
}
MPCR_D bool warp_all(bool p) {
#if defined(__CUDA_ARCH__)
  return __all_sync(0xffffffffu, p) != 0;      // one VOTE instead of a called 5-step butterfly
#else
  int v = p ? 1 : 0;
  for (int m = 16; m > 0; m >>= 1) v &= shfl_xor_i(v, m);
  return v != 0;
#endif
}
MPCR_D bool warp_any(bool p) { return !warp_all(!p); }
// NaN-propagating max for finiteness checks
MPCR_D bool finite_d(double v) { return v == v && fabs(v) < 1.0e300; }

// ---------------------------------------------------------------------------
// Per-warp workspace (shared memory)
// ---------------------------------------------------------------------------
// stage storage layout (doubles)
// Every block starts at an even offset and the stride is even, so that (with the 16-byte aligned
// dynamic shared array) contiguous runs are loaded / stored as LDS.128 / STS.128.
constexpr int kOffA = 0;        // StageLin (21)
constexpr int kOffBase0 = 21;   // derivative part of (thr,thr) (the entry also carries row terms)
constexpr int kOffB = 22;       // b = -c_{k+1}  (6)
constexpr int kOffHs = 28;      // packed sparse stage Hessian (32)
constexpr int kOffGs = 60;      // dense stage gradient over (z, w) (12)
constexpr int kOffZero = 109;   // pad slot, kept at 0 (gs[2] is NOT zero in the restoration phase)
constexpr int kZeroSlot = kOffZero - 28;   // the always-zero slot relative to kOffHs
constexpr int kOffK = 72;       // feedback gain (3 x 9); reused for the forward-sweep results
constexpr int kOffBase1 = 99;   // derivative part of (st,st)
constexpr int kOffKf = 100;     // feed-forward, 3 right-hand sides (3 x 3)
constexpr int kStrideEuler = 110;   // even (alignment); lane-parallel writes then see 4-way bank conflicts,
                                    // ~120 stores per iteration - negligible next to the Riccati loads
// RK4 integrator (IG = 1): the map's Jacobian has 11 more non-trivial entries, stored behind the
// Euler record so that every Euler offset stays where it is:
//   +0..2  dX+/d(r, thr, st)    +3..5  dY+/d(r, thr, st)    +6..9  dyaw+/d(v_x, v_y, thr, st)
//   +10    dyaw+/dr (= Ts for Euler)                         +11    pad
constexpr int kOffAx = kStrideEuler;
constexpr int kLinEuler = 21, kLinRk4 = 32;
MPCR_HD constexpr int stage_stride(int ig) { return ig ? kStrideEuler + 12 : kStrideEuler; }
constexpr int kNz = 9, kNw = 3, kNzw = 12, kRhs = 3;
constexpr int kPS = 10;         // row stride of the 9x9 cost-to-go matrix (even: aligned rows)

constexpr int kMisc = 24;       // relaxed row bounds lo[10], hi[10] + spare
constexpr int kPrmSlots = 20;   // >= sizeof(Params) / 8, even (static_assert below)
constexpr int kPbSlots = 24;    // >= sizeof(ProblemView) / 8, even
constexpr int kTS = 18;         // row stride of T [9][15 used]: even (aligned rows), 9 row writers on distinct banks
constexpr int kGS = 16;         // row stride of G [12][15 used]
constexpr int kScrT = 0, kScrG = 9 * kTS, kScrMisc = kScrG + 12 * kGS;   // T and G contiguous (kkt_solve scratch)

// On the device every shared-memory pointer is re-derived from the dynamic shared array at the
// point of use, so that the compiler knows the address space and emits LDS/STS (pointers kept
// in the solver object lose their provenance and degrade to generic LD/ST - profiles/r1f).
template <int IG>
struct WarpWs {
  static constexpr int kStride = stage_stride(IG);
  double* base_;   // host build (emulation): start of the workspace; unused on the device
  int shalf_;      // HALF the offset of the scratch area inside the shared/host workspace: written as
                   // 2 * shalf_ the compiler sees a 16-byte aligned scratch base (vector LDS/STS)
  MPCR_D double* sm() const {
#if defined(__CUDA_ARCH__)
    extern __shared__ __align__(16) double mpcr_dyn_smem[];
    return mpcr_dyn_smem;
#else
    return base_;
#endif
  }
  MPCR_HD void bind(double* base, int N) {
    base_ = base;
    shalf_ = (N + 1) * (kStride / 2);
  }
  MPCR_D double* T() const { return sm() + 2 * shalf_ + kScrT; }       // [9][kTS]
  MPCR_D double* G() const { return sm() + 2 * shalf_ + kScrG; }       // [12][kGS]
  MPCR_D double* misc() const { return sm() + 2 * shalf_ + kScrMisc; }   // [kMisc]
  MPCR_D double* prm_mem() const { return sm() + 2 * shalf_ + kScrMisc + kMisc; }        // Params
  MPCR_D double* pb_mem() const { return sm() + 2 * shalf_ + kScrMisc + kMisc + kPrmSlots; }   // ProblemView
  MPCR_D double* st(int k) const { return sm() + k * kStride; }   // [(N+1)][kStride]
};

// packed slot of the sparse stage Hessian for entry (row, col) of the 12x12
// matrix over (X,Y,yaw,vx,vy,r,s,p_thr,p_st | thr,st,ds); -1 = structurally zero
MPCR_HD int hs_slot(int r, int c) {
  if (r > c) { int t = r; r = c; c = t; }
  // (X, Y, s) block
  if (r == 0 && c == 0) return 0;
  if (r == 0 && c == 1) return 1;
  if (r == 0 && c == 6) return 2;
  if (r == 1 && c == 1) return 3;
  if (r == 1 && c == 6) return 4;
  if (r == 6 && c == 6) return 5;
  // model block over (yaw, vx, vy, r, thr, st) -> indices 2,3,4,5,9,10
  int mr = (r >= 2 && r <= 5) ? r - 2 : ((r == 9 || r == 10) ? r - 5 : -1);
  int mc = (c >= 2 && c <= 5) ? c - 2 : ((c == 9 || c == 10) ? c - 5 : -1);
  if (mr >= 0 && mc >= 0) return 6 + hm_idx(mr, mc);
  if (r == 11 && c == 11) return 27;
  if (r == 7 && c == 7) return 28;
  if (r == 8 && c == 8) return 29;
  if (r == 7 && c == 9) return 30;
  if (r == 8 && c == 10) return 31;
  return -1;
}

// out[c] = sum_k v[k] M[k][c],  M = [A B] (9 x 12) of one stage
MPCR_D void mulM(const double* v, const double* A, double Ts, double* out) {
  const double *aX = A, *aY = A + 3, *d = A + 6, *bu = A + 15;
  out[0] = v[0];
  out[1] = v[1];
  out[2] = aX[0] * v[0] + aY[0] * v[1] + v[2];
  out[3] = aX[1] * v[0] + aY[1] * v[1] + d[0] * v[3] + d[3] * v[4] + d[6] * v[5];
  out[4] = aX[2] * v[0] + aY[2] * v[1] + d[1] * v[3] + d[4] * v[4] + d[7] * v[5];
  out[5] = Ts * v[2] + d[2] * v[3] + d[5] * v[4] + d[8] * v[5];
  out[6] = v[6];
  out[7] = 0.0;
  out[8] = 0.0;
  out[9] = bu[0] * v[3] + bu[2] * v[4] + bu[4] * v[5] + v[7];
  out[10] = bu[1] * v[3] + bu[3] * v[4] + bu[5] * v[5] + v[8];
  out[11] = v[6];
}

// zn = M [z; w]
MPCR_D void applyM(const double* z, const double* w, const double* A, double Ts, double* zn) {
  const double *aX = A, *aY = A + 3, *d = A + 6, *bu = A + 15;
  zn[0] = z[0] + aX[0] * z[2] + aX[1] * z[3] + aX[2] * z[4];
  zn[1] = z[1] + aY[0] * z[2] + aY[1] * z[3] + aY[2] * z[4];
  zn[2] = z[2] + Ts * z[5];
  for (int i = 0; i < 3; ++i)
    zn[3 + i] = d[3 * i] * z[3] + d[3 * i + 1] * z[4] + d[3 * i + 2] * z[5] + bu[2 * i] * w[0] +
                bu[2 * i + 1] * w[1];
  zn[6] = z[6] + w[2];
  zn[7] = w[0];
  zn[8] = w[1];
}

// RK4 variants: A[0..20] as above, Ax[0..10] = the 11 extra entries (kOffAx)
template <int IG>
MPCR_D void mulMx(const double* v, const double* A, const double* Ax, double Ts, double* out) {
  mulM(v, A, Ts, out);
  if (IG) {
    const double *xe = Ax, *ye = Ax + 3, *aw = Ax + 6;
    out[3] += aw[0] * v[2];
    out[4] += aw[1] * v[2];
    out[5] += xe[0] * v[0] + ye[0] * v[1] + (Ax[10] - Ts) * v[2];
    out[9] += xe[1] * v[0] + ye[1] * v[1] + aw[2] * v[2];
    out[10] += xe[2] * v[0] + ye[2] * v[1] + aw[3] * v[2];
  }
}
template <int IG>
MPCR_D void applyMx(const double* z, const double* w, const double* A, const double* Ax, double Ts, double* zn) {
  applyM(z, w, A, Ts, zn);
  if (IG) {
    const double *xe = Ax, *ye = Ax + 3, *aw = Ax + 6;
    zn[0] += xe[0] * z[5] + xe[1] * w[0] + xe[2] * w[1];
    zn[1] += ye[0] * z[5] + ye[1] * w[0] + ye[2] * w[1];
    zn[2] = z[2] + aw[0] * z[3] + aw[1] * z[4] + Ax[10] * z[5] + aw[2] * w[0] + aw[3] * w[1];
  }
}

// slack rows owned by lane k
enum Row : int {
  kThrUb = 0, kThrLb = 1, kStUb = 2, kStLb = 3, kDThr = 4, kDSt = 5, kProg = 6,
  kDThr0 = 7, kDSt0 = 8, kLane = 9
};
constexpr unsigned kHasLMask = 0x3FAu;   // every row except thr_ub, st_ub has a lower bound
constexpr unsigned kHasUMask = 0x3F5u;   // every row except thr_lb, st_lb has an upper bound
MPCR_HD bool row_hasL(int r) { return (kHasLMask >> r) & 1u; }
MPCR_HD bool row_hasU(int r) { return (kHasUMask >> r) & 1u; }

struct ProblemView {     // uniform over the warp
  double x0[6];
  double u_prev[2];
  bool has_prev_thr, has_prev_st;
  double s0;
  double cx[5], cy[5];
  double max_err;
  const double* u_warm;  // [2][N] or null
};

static_assert(sizeof(Params) <= kPrmSlots * sizeof(double), "kPrmSlots too small");
static_assert(sizeof(ProblemView) <= kPbSlots * sizeof(double), "kPbSlots too small");
constexpr int kScratch = kScrMisc + kMisc + kPrmSlots + kPbSlots;
MPCR_HD int ws_doubles(int N, int ig = 0) { return (N + 1) * stage_stride(ig) + kScratch; }

struct SolveOut {        // pointers already offset to this problem
  double* U; double* X; double* S; double* eC; double* eL; double* lam_g;
  double* obj; int32_t* status; int32_t* iters; double* trace;
};

// solver phases: every heavy routine has exactly ONE call site inside the phase loop
// (keeps the kernel's instruction footprint inside the SM instruction cache)
enum Phase : int { kPhInit = 0, kPhInit2 = 1, kPhNewton = 2, kPhBacktrack = 3, kPhSoc = 4,
                   kPhRestoCheck = 5 };   // restoration: evaluate the ORIGINAL merit at the new point

// ---------------------------------------------------------------------------
// The solver.  All members are per-lane unless noted "uniform".
// ---------------------------------------------------------------------------
// Row (slack) state of one lane.  These arrays are indexed with RUNTIME row numbers (the row loops
// are deliberately not unrolled), so they live in local memory; keeping them in their own object
// lets the compiler scalarise every other member of WarpSolver into registers.
struct RowState {
  double s[kNumRows], yd[kNumRows], vL[kNumRows], vU[kNumRows];
  double rd[kNumRows];      // d_r(x) - s_r at the current iterate
  double dsr[kNumRows];     // slack step of the current direction
  // feasibility restoration: reference point (x, s_hat, u) of this lane's stage and the gradient
  // of the proximity term over the same 9 variables
  double xr[9], pg[9];
  double osave[12];         // the outer problem's scalars while the restoration phase runs
};

// RESTO: compile the feasibility-restoration phase in (mpcr_config.restoration); the RESTO = false
//        instantiation is the plain solver without any of its code or registers.
// IG:    integrator of the dynamics rows, 0 = explicit Euler (reference), 1 = RK4 (32-entry stage
//        linearisation).
template <int MODEL, bool RESTO = false, int IG = 0>
struct WarpSolver {
  const Config& cfg;
  const IpmOptions& o;
  WarpWs<IG> ws;
  // per-problem parameters and inputs (uniform over the warp): ONE copy in shared memory, written
  // by solve_problem, instead of 32 per-lane copies in local memory behind L2 latency
  MPCR_D const Params& prm() const { return *reinterpret_cast<const Params*>(ws.prm_mem()); }
  MPCR_D const ProblemView& pbv() const { return *reinterpret_cast<const ProblemView*>(ws.pb_mem()); }
  int lane, N;
  bool act;
  // primal iterate
  double x[6], sh, u[2];
  // multipliers (dynamics rows carry lam = dc * y_scaled; y0 uniform: x0 rows (6) + s0 row)
  double lam[6], y0[7];
  double dcs[6];            // gradient-based row scaling of the dynamics rows
  double ddl;               // same for the lane row
  double df;                // objective scaling (uniform)
  // slack rows (runtime-indexed -> local memory)
  unsigned rowmask;
  RowState& R;              // runtime-indexed row arrays live OUTSIDE this object (see RowState)
  double cres[6];           // x_{k+1} - f(x_k, u_k) at the current iterate
  double fval, barB;        // df * f(x); barrier sum B(s): phi = fval + mu * barB (uniform)
  // search direction
  double dx[6], dsh, du[2];
  double mu, tau;
  // filter: one entry per lane
  double filt_th, filt_ph;
  int filt_n;
  // feasibility restoration (quadratic-penalty variant, same text as LaneSolver): while `in_resto()` is set
  // the objective is wq/2 |d(x) - s|^2 + zeta/2 |D_R (x - x_R)|^2 and the rows' multipliers are o.resto_weight (d - s)
  bool resto_;
  MPCR_D bool in_resto() const { return RESTO && resto_; }
  // plain build with mpcr_config.restoration set: the line search gave up where IPOPT would start
  // its restoration phase - the caller re-solves this problem with the RESTO build (uniform)
  bool wanted_resto;
  // packed slots of the stage-Hessian entries (r, lane), r = 0..11, one byte each (registers,
  // no runtime-indexed local array): rows 0-7 in colslot_lo, rows 8-11 in colslot_hi
  unsigned long long colslot_lo;
  unsigned colslot_hi;
  double dwflag;            // 1 if diagonal entry (lane, lane) receives delta_w

  MPCR_D WarpSolver(const Config& c, RowState& rows, double* wsbase)
      : cfg(c), o(c.o), R(rows) {
    lane = lane_id();
    N = c.N;
    act = lane < N;
    ws.bind(wsbase, N);
    colslot_lo = 0ull;
    colslot_hi = 0u;
    for (int r = 0; r < 12; ++r) {
      // lanes 0-11: column `lane` of the stage Hessian; lane 12: the stage gradient (right-hand
      // side 0, stored right behind the Hessian slots); lanes 13, 14 (wrap vectors): zero
      const unsigned long long v = (lane < 12) ? (unsigned long long)(slot_of(r, lane) & 0xff)
                                   : (lane == 12 ? (unsigned long long)(kOffGs - kOffHs + r) : kZeroSlot);
      if (r < 8) colslot_lo |= v << (8 * r);
      else colslot_hi |= (unsigned)v << (8 * (r - 8));
    }
    dwflag = (lane <= 6 || lane == 9 || lane == 10) ? 1.0 : 0.0;
  }

  MPCR_D bool on(int r) const { return (rowmask >> r) & 1u; }
  MPCR_D int prev_lane() const { return lane == 0 ? N - 1 : lane - 1; }
  MPCR_D int next_lane() const { return lane >= N - 1 ? 0 : lane + 1; }
  MPCR_D double lo(int r) const { return ws.misc()[r]; }
  MPCR_D double hi(int r) const { return ws.misc()[kNumRows + r]; }

  // original (unrelaxed, unscaled-row) bounds of row r
  MPCR_D void row_bounds(int r, double& l, double& h) const {
    const double inf = 1.0e300;
    switch (r) {
      case kThrUb: l = -inf; h = prm().d_max; break;
      case kThrLb: l = prm().min_throttle; h = inf; break;
      case kStUb: l = -inf; h = prm().max_steer; break;
      case kStLb: l = prm().min_steer; h = inf; break;
      case kDThr: case kDThr0: l = prm().min_throttle_delta; h = prm().max_throttle_delta; break;
      case kDSt: case kDSt0: l = prm().min_steer_delta; h = prm().max_steer_delta; break;
      case kProg: l = prm().min_s_delta; h = cfg.Ts * prm().v_max; break;
      default: l = -pbv().max_err; h = pbv().max_err; break;
    }
  }

  // value of row r (scaled) at a point
  MPCR_D double row_value(int r, const double* uu, const double* up, double shh, double shp,
                          double eC) const {
    switch (r) {
      case kThrUb: case kThrLb: return uu[0];
      case kStUb: case kStLb: return uu[1];
      case kDThr: return uu[0] - up[0];
      case kDSt: return uu[1] - up[1];
      case kProg: return shh - shp;
      case kDThr0: return uu[0] - pbv().u_prev[0];
      case kDSt0: return uu[1] - pbv().u_prev[1];
      default: return ddl * eC;
    }
  }
  // J_r * (step): the row's linear form applied to a direction
  MPCR_D double row_lin(int r, const double* duu, const double* dup, double dshh, double dshp,
                        const double* dxy) const {
    switch (r) {
      case kThrUb: case kThrLb: case kDThr0: return duu[0];
      case kStUb: case kStLb: case kDSt0: return duu[1];
      case kDThr: return duu[0] - dup[0];
      case kDSt: return duu[1] - dup[1];
      case kProg: return dshh - dshp;
      default: return gC[0] * dxy[0] + gC[1] * dxy[1] + gC[2] * dshh;
    }
  }

  // previous-stage quantities via shuffles (lane 0: fixed initial state / wrap)
  MPCR_D void gather_prev(const double* xx, double shh, const double* uu, double* xp,
                          double& shp, double* up) const {
    int src = prev_lane();
    for (int k = 0; k < 6; ++k) {
      double t = shfl(xx[k], src);
      xp[k] = (lane == 0) ? pbv().x0[k] : t;
    }
    double t = shfl(shh, src);
    shp = (lane == 0) ? pbv().s0 : t;
    up[0] = shfl(uu[0], src);   // lane 0 receives u_{N-1} (wrap rows)
    up[1] = shfl(uu[1], src);
  }

  // -------------------------------------------------------------------------
  // objective / constraints at a point (line search trial or current iterate)
  // -------------------------------------------------------------------------
  struct Point {
    double x[6], sh, u[2];
    double c[6];
    double f, theta, barB;  // uniform
    bool finite;            // uniform
  };

  // slacks of the point are s + at * dsr (not stored)
  // inner: restoration objective and theta without the rows; dfu: objective scaling to apply
  MPCR_D void eval_point(Point& t, double* trd /* [kNumRows] out: d_r - s_r */, double at, bool inner,
                         double dfu) const {
    double xp[6], shp, up[2];
    gather_prev(t.x, t.sh, t.u, xp, shp, up);
    double f[6];
    model_step(xp, t.u, f);
    StageCost sc;
    stage_cost(prm(), pbv().cx, pbv().cy, t.x, t.sh, 0, sc);
    double fl = 0.0, th = 0.0;
    double prod = 1.0, lin = 0.0;       // barrier: -log(prod of slack distances) + kappa_d * lin
    bool fin = true;
    for (int k = 0; k < 6; ++k) t.c[k] = 0.0;
    if (act) {
      if (!inner) {
        fl = sc.val;
        if (lane == N - 1) fl -= prm().lambda_s * t.sh;
        if (lane >= 1) {
          double dsg = t.u[1] - up[1];
          fl += prm().beta_delta * dsg * dsg;
        }
      } else {
        const double v9[9] = {t.x[0], t.x[1], t.x[2], t.x[3], t.x[4], t.x[5], t.sh, t.u[0], t.u[1]};
#pragma unroll 1
        for (int a = 0; a < 9; ++a) {
          const double ref = R.xr[a];
          const double dr = 1.0 / fmax(1.0, fabs(ref));
          fl += 0.5 * R.osave[10] * dr * dr * (v9[a] - ref) * (v9[a] - ref);
        }
      }
      for (int k = 0; k < 6; ++k) {
        t.c[k] = t.x[k] - f[k];
        th += fabs(dcs[k] * t.c[k]);
      }
    }
    RowRegs nx{};
    row_fetch(0, nx);
#pragma unroll 1
    for (int r = 0; r < kNumRows; ++r) {
      const RowRegs c = nx;
      row_fetch(r + 1, nx);
      double rdv = 0.0;
      if (on(r)) {
        const double sv = c.s + at * c.dsr;
        rdv = row_value(r, t.u, up, t.sh, shp, sc.eC) - sv;
        if (inner) fl += 0.5 * o.resto_weight * rdv * rdv;
        else th += fabs(rdv);
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
      trd[r] = rdv;
    }
    double bar = act ? (-log_d(prod) + o.kappa_d * lin) : 0.0;
    fin = fin && finite_d(fl) && finite_d(th) && finite_d(bar);
    t.finite = warp_all(fin);
    t.f = dfu * warp_sum(fl);
    t.theta = warp_sum(th);
    t.barB = warp_sum(bar);
  }

  // -------------------------------------------------------------------------
  // Derivative evaluation at the current iterate; fills the stage storage.
  // mode 0: Newton step (exact Hessian), mode 1: least-squares multiplier init
  // -------------------------------------------------------------------------
  StageLin lin;           // Jacobian blocks of stage `lane`
  double linx[IG ? 11 : 1];   // RK4: the 11 extra entries of the stage Jacobian (kOffAx)

  // x+ of the NLP's dynamics rows: explicit Euler (reference) or RK4 of the same vector field
  MPCR_D void model_step(const double* xp, const double* uu, double* f) const {
    if (IG) rk4_model(MODEL, xp, uu, cfg.Ts, nullptr, f, nullptr, nullptr);
    else f_model(MODEL, xp, uu, cfg.Ts, f);
  }
  // this lane's stage Jacobian in the solver's layout: A[0..20] = StageLin, Ax[0..10] = RK4 extras
  MPCR_D void lin_arrays(double* A, double* Ax) const {
    for (int k = 0; k < 3; ++k) { A[k] = lin.aX[k]; A[3 + k] = lin.aY[k]; }
    for (int i = 0; i < 3; ++i)
      for (int k = 0; k < 3; ++k) A[6 + 3 * i + k] = lin.d[i][k];
    for (int i = 0; i < 3; ++i)
      for (int k = 0; k < 2; ++k) A[15 + 2 * i + k] = lin.bu[i][k];
    if (IG)
      for (int k = 0; k < 11; ++k) Ax[k] = linx[k];
  }
  double gcost[5];        // df * grad cost of state lane+1 over (X, Y, s, vx, vy) (incl -lambda_s)
  double gC[3];           // ddl * grad e_C over (X, Y, s) at state lane+1 (lane row)
  double sigc_prev;       // sigma_k - sigma_{k-1} (0 for lane 0)
  double hcp6[6];         // df * cost Hessian over (X, Y, s) of state `lane` (+ lane-row curvature)

  MPCR_D void eval_derivs(int mode) {
    double xp[6], shp, up[2];
    gather_prev(x, sh, u, xp, shp, up);
    double hm[kHmDoubles];
    double lz[6];
    for (int k = 0; k < 6; ++k) lz[k] = (mode == 0) ? lam[k] : 0.0;
    if (!IG) {
      lin_model(MODEL, xp, u, cfg.Ts, lz, nullptr, lin, hm);
    } else {
      double f[6], J[6][6];
      rk4_model(MODEL, xp, u, cfg.Ts, lz, f, J, hm);
      for (int a = 0; a < 3; ++a) { lin.aX[a] = J[0][a]; lin.aY[a] = J[1][a]; linx[a] = J[0][3 + a]; linx[3 + a] = J[1][3 + a]; }
      for (int i = 0; i < 3; ++i) {
        for (int a = 0; a < 3; ++a) lin.d[i][a] = J[3 + i][1 + a];
        for (int a = 0; a < 2; ++a) lin.bu[i][a] = J[3 + i][4 + a];
      }
      linx[6] = J[2][1];
      linx[7] = J[2][2];
      linx[8] = J[2][4];
      linx[9] = J[2][5];
      linx[10] = J[2][3];
    }
    StageCost sc;
    stage_cost(prm(), pbv().cx, pbv().cy, x, sh, 2, sc);
    gcost[0] = df * sc.g3[0];
    gcost[1] = df * sc.g3[1];
    gcost[2] = df * sc.g3[2] - ((lane == N - 1) ? df * prm().lambda_s : 0.0);
    gcost[3] = df * sc.gvx;
    gcost[4] = df * sc.gvy;
    for (int k = 0; k < 3; ++k) gC[k] = ddl * sc.gC[k];
    sigc_prev = (lane >= 1 && act && !in_resto()) ? (u[1] - up[1]) : 0.0;
    // cost Hessian of state lane+1 -> owner of stage lane+1 (shuffle up) / terminal slot
    double hc[8];
    for (int k = 0; k < 8; ++k) hc[k] = 0.0;
    double hyr[2] = {0.0, 0.0}, pdu[2] = {0.0, 0.0};   // restoration: (yaw, r) of state lane+1, (thr, st) of this stage
    if (in_resto()) {
      // proximity term zeta/2 |D_R (v - v_R)|^2 over (x_{k+1}, s_{k+1}, u_k) replaces the cost
      const double v9[9] = {x[0], x[1], x[2], x[3], x[4], x[5], sh, u[0], u[1]};
      double pd9[9];
#pragma unroll 1
      for (int a = 0; a < 9; ++a) {
        const double ref = R.xr[a];
        const double dr = 1.0 / fmax(1.0, fabs(ref));
        pd9[a] = act ? R.osave[10] * dr * dr : 0.0;
        R.pg[a] = pd9[a] * (v9[a] - ref);
      }
      gcost[0] = R.pg[0]; gcost[1] = R.pg[1]; gcost[2] = R.pg[6]; gcost[3] = R.pg[3]; gcost[4] = R.pg[4];
      hc[0] = pd9[0]; hc[3] = pd9[1]; hc[5] = pd9[6]; hc[6] = pd9[3]; hc[7] = pd9[4];
      hyr[0] = pd9[2]; hyr[1] = pd9[5];
      pdu[0] = pd9[7]; pdu[1] = pd9[8];
      if (act && cfg.lane_con && on(kLane)) {
        double yl = R.yd[kLane] * ddl;
        hc[2] += yl * sc.hC02;
        hc[4] += yl * sc.hC12;
        hc[5] += yl * sc.hC22;
      }
    } else if (mode == 0 && act) {
      for (int k = 0; k < 6; ++k) hc[k] = df * sc.h3[k];
      hc[6] = df * sc.hvx;
      hc[7] = df * sc.hvy;
      if (cfg.lane_con && on(kLane)) {
        double yl = R.yd[kLane] * ddl;     // + yd * Hess(e_C); the Sigma part is a row term
        hc[2] += yl * sc.hC02;
        hc[4] += yl * sc.hC12;
        hc[5] += yl * sc.hC22;
      }
    }
    double hcp[8];
    for (int k = 0; k < 8; ++k) {
      double t = shfl(hc[k], prev_lane());
      hcp[k] = (lane == 0) ? 0.0 : t;
    }
    double hyrp[2] = {0.0, 0.0};
    if (in_resto()) {
      for (int k = 0; k < 2; ++k) {
        double t = shfl(hyr[k], prev_lane());
        hyrp[k] = (lane == 0) ? 0.0 : t;
      }
    }
    if (act) {
      double* st = ws.st(lane);
      for (int k = 0; k < 3; ++k) { st[kOffA + k] = lin.aX[k]; st[kOffA + 3 + k] = lin.aY[k]; }
      for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) st[kOffA + 6 + 3 * i + k] = lin.d[i][k];
      for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 2; ++k) st[kOffA + 15 + 2 * i + k] = lin.bu[i][k];
      if (IG)
        for (int k = 0; k < 11; ++k) st[kOffAx + k] = linx[k];
      double* hs = st + kOffHs;
      // (thr,thr), (st,st) also receive row terms: their derivative part is kept in `base`;
      // the (X,Y,s) block's derivative part stays in registers (hcp6)
      for (int k = 0; k < 6; ++k) hcp6[k] = hcp[k];
      for (int k = 0; k < kHmDoubles; ++k) hs[6 + k] = (mode == 0) ? -hm[k] : 0.0;
      hs[6 + hm_idx(1, 1)] += hcp[6];
      hs[6 + hm_idx(2, 2)] += hcp[7];
      if (in_resto()) {
        hs[6 + hm_idx(0, 0)] += hyrp[0];
        hs[6 + hm_idx(3, 3)] += hyrp[1];
        hs[6 + hm_idx(4, 4)] += pdu[0];
        hs[6 + hm_idx(5, 5)] += pdu[1];
      }
      st[kOffBase0] = hs[6 + hm_idx(4, 4)];
      st[kOffBase1] = hs[6 + hm_idx(5, 5)];
      if (lane == N - 1) {
        double* hN = ws.st(N) + kOffHs;
        double* bN = ws.st(N) + kOffK;          // the terminal stage has no gain: free space
        for (int k = 0; k < 6; ++k) bN[k] = hc[k];
        for (int k = 0; k < kHmDoubles; ++k) hN[6 + k] = 0.0;
        hN[6 + hm_idx(1, 1)] = hc[6];
        hN[6 + hm_idx(2, 2)] = hc[7];
        hN[6 + hm_idx(0, 0)] = hyr[0];
        hN[6 + hm_idx(3, 3)] = hyr[1];
        for (int k = 27; k < 32; ++k) hN[k] = 0.0;
      }
    }
  }

  // -------------------------------------------------------------------------
  // Row-dependent stage terms (depend on Sigma, delta_w and the right-hand side).
  //   e_r = D_r * rdv_r + gsb_r  is the condensed slack contribution of row r.
  //   mode 0: Newton/SOC (D = Sigma + dw, gsb = barrier gradient)
  //   mode 1: LSQ init    (D = 1, gradient = -(grad f + Jd'(vU - vL)))
  // -------------------------------------------------------------------------
  double Dw[2];   // D of the two wrap rows (uniform, from lane 0)

  // The row arrays live in local memory and mostly miss L1 (shared memory takes the carve-out), and
  // the row loops are not unrolled: every iteration would wait a full L2 round trip for its own
  // loads.  The loops therefore fetch row r + 1 into registers while row r is being processed.
  struct RowRegs { double s, yd, vL, vU, rd, dsr; };
  MPCR_D void row_fetch(int r, RowRegs& q) const {
    if (r < kNumRows && on(r)) {
      q.s = R.s[r];
      q.yd = R.yd[r];
      q.rd = R.rd[r];
      q.dsr = R.dsr[r];
      if (row_hasL(r)) q.vL = R.vL[r];
      if (row_hasU(r)) q.vU = R.vU[r];
    }
  }

  // Sigma_r and d(barrier objective)/d(slack r)
  MPCR_D void sigma_gsb(int r, const RowRegs& c, double& sg, double& g) const {
    double isl, isu;
    sigma_gsb(r, c, sg, g, isl, isu);
  }
  // ... also hands out the reciprocal slacks 1 / (s - lo), 1 / (hi - s) it computed
  MPCR_D void sigma_gsb(int r, const RowRegs& c, double& sg, double& g, double& isl, double& isu) const {
    sg = 0.0;
    g = 0.0;
    isl = 0.0;
    isu = 0.0;
    if (row_hasL(r)) {
      isl = rcp_pivot(c.s - lo(r));        // slacks are strictly inside their bounds
      sg += c.vL * isl;
      g -= mu * isl;
      if (!row_hasU(r)) g += o.kappa_d * mu;
    }
    if (row_hasU(r)) {
      isu = rcp_pivot(hi(r) - c.s);
      sg += c.vU * isu;
      g += mu * isu;
      if (!row_hasL(r)) g -= o.kappa_d * mu;
    }
  }

  // rdv / cv may be null (= zero right-hand sides)
  MPCR_D void write_row_terms(int mode, double dw, const double* rdv /* [kNumRows] */,
                              const double* cv /* [6] dynamics rhs */) {
    // D and e of the individual rows, accumulated by role (no per-row arrays)
    double Rtt = 0.0, Rss = 0.0, eth = 0.0, est = 0.0;     // rows acting on (thr_k, st_k) alone
    double Dthr = 0.0, ethr = 0.0, Dst = 0.0, esr = 0.0;   // rate rows (kDThr, kDSt)
    double Dprog = 0.0, eprog = 0.0, Dlane = 0.0, elane = 0.0;
    RowRegs nx{};
    row_fetch(0, nx);
    double nrdv = (rdv && on(0)) ? rdv[0] : 0.0;
#pragma unroll 1
    for (int r = 0; r < kNumRows; ++r) {
      const RowRegs c = nx;
      const double crdv = nrdv;
      row_fetch(r + 1, nx);
      if (rdv && r + 1 < kNumRows && on(r + 1)) nrdv = rdv[r + 1];
      double Dr = 0.0, er = 0.0;
      if (on(r)) {
        if (mode == 0) {
          double sg, g;
          sigma_gsb(r, c, sg, g);
          Dr = sg + dw;
          er = Dr * (rdv ? crdv : 0.0) + g;
          if (in_resto()) {
            // row in the objective: min over ds of wq/2 (J dx - ds + rd)^2 + Sigma/2 ds^2 + g ds
            const double kap = o.resto_weight / (o.resto_weight + Dr);
            Dr *= kap;
            er *= kap;
          }
        } else {
          Dr = 1.0;
          er = -((row_hasU(r) ? c.vU : 0.0) - (row_hasL(r) ? c.vL : 0.0));
        }
      }
      switch (r) {
        case kThrUb: case kThrLb: case kDThr0: Rtt += Dr; eth += er; break;
        case kStUb: case kStLb: case kDSt0: Rss += Dr; est += er; break;
        case kDThr: Dthr = Dr; ethr = er; break;
        case kDSt: Dst = Dr; esr = er; break;
        case kProg: Dprog = Dr; eprog = er; break;
        default: Dlane = Dr; elane = er; break;
      }
    }
    const bool wrap0 = (lane == 0);   // rows kDThr/kDSt of lane 0 are the non-local wrap rows
    double Dw0 = shfl(Dthr, 0), Dw1 = shfl(Dst, 0);
    double ew0 = shfl(ethr, 0), ew1 = shfl(esr, 0);
    Dw[0] = Dw0;
    Dw[1] = Dw1;
    // cost gradient / lane-row data of state `lane` live in the previous lane
    double gcp[5], gCp[3];
    for (int k = 0; k < 5; ++k) {
      double t = shfl(gcost[k], prev_lane());
      gcp[k] = (lane == 0) ? 0.0 : t;
    }
    double elane_p = shfl(elane, prev_lane()), Dlane_p = shfl(Dlane, prev_lane());
    if (lane == 0) { elane_p = 0.0; Dlane_p = 0.0; }
    for (int k = 0; k < 3; ++k) {
      double t = shfl(gC[k], prev_lane());
      gCp[k] = (lane == 0) ? 0.0 : t;
    }
    // restoration: proximity gradient over (yaw, r) of state `lane` (previous lane) and over this
    // stage's inputs
    double pgy = 0.0, pgr = 0.0, pgt = 0.0, pgs = 0.0;
    if (in_resto()) {
      const double t0 = shfl(R.pg[2], prev_lane()), t1 = shfl(R.pg[5], prev_lane());
      pgy = (lane == 0) ? 0.0 : t0;
      pgr = (lane == 0) ? 0.0 : t1;
      pgt = R.pg[7];
      pgs = R.pg[8];
    }
    if (!act) return;
    // steer-rate cost beta (sigma_k - sigma_{k-1})^2, k = 1..N-1 (control/MPC.py:97)
    const double bsig = (lane >= 1 && !in_resto()) ? 2.0 * prm().beta_delta * df : 0.0;
    const double sgn = (mode == 0) ? 1.0 : -1.0;   // LSQ uses the negated gradient
    const double bq = (mode == 0) ? bsig : 0.0;
    double* st = ws.st(lane);
    double* hs = st + kOffHs;
    double* gs = st + kOffGs;
    const double dthr = wrap0 ? 0.0 : Dthr;
    const double dst = wrap0 ? 0.0 : Dst;
    // (X, Y, s) block of state `lane`: derivative part + lane-row Sigma term D gC gC'
    hs[0] = hcp6[0] + Dlane_p * gCp[0] * gCp[0];
    hs[1] = hcp6[1] + Dlane_p * gCp[0] * gCp[1];
    hs[2] = hcp6[2] + Dlane_p * gCp[0] * gCp[2];
    hs[3] = hcp6[3] + Dlane_p * gCp[1] * gCp[1];
    hs[4] = hcp6[4] + Dlane_p * gCp[1] * gCp[2];
    hs[5] = hcp6[5] + Dlane_p * gCp[2] * gCp[2];
    hs[6 + hm_idx(4, 4)] = st[kOffBase0] + Rtt + dthr;
    hs[6 + hm_idx(5, 5)] = st[kOffBase1] + Rss + dst + bq;
    hs[27] = Dprog;
    hs[28] = dthr;
    hs[29] = dst + bq;
    hs[30] = -dthr;
    hs[31] = -(dst + bq);
    double ethk = eth + (wrap0 ? 0.0 : ethr);
    double estk = est + (wrap0 ? 0.0 : esr);
    if (cfg.wrap) {
      if (lane == 0) { ethk += ew0; estk += ew1; }
      if (lane == N - 1) { ethk -= ew0; estk -= ew1; }
    }
    const double bg = sgn * bsig * sigc_prev;
    gs[0] = sgn * gcp[0] + elane_p * gCp[0];
    gs[1] = sgn * gcp[1] + elane_p * gCp[1];
    gs[2] = pgy;
    gs[3] = sgn * gcp[3];
    gs[4] = sgn * gcp[4];
    gs[5] = pgr;
    gs[6] = sgn * gcp[2] + elane_p * gCp[2];
    gs[7] = wrap0 ? 0.0 : -ethr;
    gs[8] = (wrap0 ? 0.0 : -esr) - bg;
    gs[9] = ethk + pgt;
    gs[10] = estk + bg + pgs;
    gs[11] = eprog;
    for (int k = 0; k < 6; ++k) st[kOffB + k] = cv ? -cv[k] : 0.0;
    if (lane == N - 1) {
      double* gN = ws.st(N) + kOffGs;
      double* hN = ws.st(N) + kOffHs;
      const double* bN = ws.st(N) + kOffK;
      gN[0] = sgn * gcost[0] + elane * gC[0];
      gN[1] = sgn * gcost[1] + elane * gC[1];
      gN[2] = in_resto() ? R.pg[2] : 0.0;
      gN[3] = sgn * gcost[3];
      gN[4] = sgn * gcost[4];
      gN[5] = in_resto() ? R.pg[5] : 0.0;
      gN[6] = sgn * gcost[2] + elane * gC[2];
      for (int k = 7; k < 12; ++k) gN[k] = 0.0;
      hN[0] = bN[0] + Dlane * gC[0] * gC[0];
      hN[1] = bN[1] + Dlane * gC[0] * gC[1];
      hN[2] = bN[2] + Dlane * gC[0] * gC[2];
      hN[3] = bN[3] + Dlane * gC[1] * gC[1];
      hN[4] = bN[4] + Dlane * gC[1] * gC[2];
      hN[5] = bN[5] + Dlane * gC[2] * gC[2];
    }
  }

  // value of the assembled stage Hessian at a packed slot (+ regularisation flag)
  MPCR_D double stage_entry(const double* st, int slot, double dw) const {
    double v = st[kOffHs + (slot & 0xff)];
    if (slot & 0x100) v += dw;
    return v;
  }
  MPCR_D static int slot_of(int r, int c) {
    int sl = hs_slot(r, c);
    int slot = (sl < 0) ? kZeroSlot : sl;
    if (r == c && (r <= 6 || r == 9 || r == 10)) slot |= 0x100;
    return slot;
  }

  // -------------------------------------------------------------------------
  // Riccati factorisation + solve for 3 right-hand sides.
  //   rhs 0: the Newton (or LSQ / SOC) right-hand side stored in the stages
  //   rhs 1, 2: the wrap vectors a_1 = e(thr_0) - e(thr_{N-1}), a_2 (steer)
  // Returns false when a pivot block is not positive definite (wrong inertia).
  // -------------------------------------------------------------------------

  int ric_nneg;   // negative pivots of the last sweep (uniform)

  MPCR_D bool riccati(double dw) {
    const double Ts = cfg.Ts;
    double* T = ws.T();   // [9][kTS]:  columns 0-11 = P M,           12-14 = t_j = p_j + P b_j
    double* G = ws.G();   // [12][kGS]: columns 0-11 = M' P M + H,     12-14 = g_j = grad_j + M' t_j
    const int nrhs = cfg.wrap ? kRhs : 1;
    // Lane c < 9 carries column c of the cost-to-go matrix P (= its row c: P is symmetric) and
    // entry c of the three cost-to-go vectors in REGISTERS from one stage to the next, so a stage
    // costs two shared-memory round trips (T, G) instead of three.
    double pr[9], pj[kRhs];
    {
      const double* stN = ws.st(N);
      double* P0 = G;       // staging of the terminal matrix, stride kPS (G is free until round 2)
      for (int j = 0; j < 3; ++j) {
        int e = lane + 32 * j;
        if (e < 81) {
          int r = e / 9, c = e % 9;
          P0[r * kPS + c] = stage_entry(stN, slot_of(r, c), dw);
        }
      }
      syncwarp();
      const int c = lane < 9 ? lane : 0;
      for (int a = 0; a < 9; ++a) pr[a] = P0[a * kPS + c];
      pj[0] = stN[kOffGs + c];
      pj[1] = 0.0;
      pj[2] = 0.0;
    }
    syncwarp();
    bool ok = true;
    ric_nneg = 0;
    const double dwc = dwflag * dw;
#pragma unroll 1
    for (int k = N - 1; k >= 0; --k) {
      double* st = ws.st(k);
      // ---- round 1: row `lane` of T = P M, and t_j ---------------------------------------
      if (lane < 9) {
        double A[kLinEuler], Ax[12], out[12];
        for (int c = 0; c < kLinEuler; ++c) A[c] = st[kOffA + c];
        if (IG)
          for (int c = 0; c < 12; ++c) Ax[c] = st[kOffAx + c];
        mulMx<IG>(pr, A, Ax, Ts, out);
        for (int c = 0; c < 12; ++c) T[lane * kTS + c] = out[c];
        double t0 = pj[0];
        for (int c = 0; c < 6; ++c) t0 += pr[c] * st[kOffB + c];
        T[lane * kTS + 12] = t0;
        T[lane * kTS + 13] = pj[1];
        T[lane * kTS + 14] = pj[2];
      }
      syncwarp();
      // ---- round 2: column `lane` of G = M' T + (stage Hessian | gradient | wrap vector) -----
      if (lane < 12 + nrhs) {
        double tc[9], A[kLinEuler], Ax[12], out[12];
        for (int r = 0; r < 9; ++r) tc[r] = T[r * kTS + lane];
        for (int c = 0; c < kLinEuler; ++c) A[c] = st[kOffA + c];
        if (IG)
          for (int c = 0; c < 12; ++c) Ax[c] = st[kOffAx + c];
        mulMx<IG>(tc, A, Ax, Ts, out);
#pragma unroll
        for (int r = 0; r < 12; ++r) {
          const int slot = (r < 8) ? (int)((colslot_lo >> (8 * r)) & 0xff) : (int)((colslot_hi >> (8 * (r - 8))) & 0xff);
          G[r * kGS + lane] = out[r] + st[kOffHs + slot];
        }
        if (lane < 12) {
          G[lane * kGS + lane] += dwc;
        } else if (lane > 12) {   // wrap vector: +1 on u_0, -1 on u_{N-1}
          const double a = (k == 0 ? 1.0 : 0.0) - (k == N - 1 ? 1.0 : 0.0);
          G[(9 + (lane - 13)) * kGS + lane] += a;
        }
      }
      syncwarp();
      // ---- round 3: LDL' of H_ww (every lane), gains, cost-to-go update in registers -------
      const double h00 = G[9 * kGS + 9], h10 = G[10 * kGS + 9], h20 = G[11 * kGS + 9];
      const double h11 = G[10 * kGS + 10], h21 = G[11 * kGS + 10], h22 = G[11 * kGS + 11];
      const double d0 = h00;
      const double id0 = rcp_pivot(d0);
      const double l10 = h10 * id0, l20 = h20 * id0;
      const double d1 = h11 - l10 * h10;
      const double id1 = rcp_pivot(d1);
      const double m21 = h21 - l20 * h10;
      const double l21 = m21 * id1;
      const double d2 = h22 - l20 * h20 - l21 * m21;
      const double id2 = rcp_pivot(d2);
      // Inertia.  The sweep factorises K0 = the KKT matrix WITHOUT the two wrap rows' terms
      // Dw a a' (they are added by the Woodbury step of kkt_solve).  IPOPT tests the inertia of
      // the FULL matrix, and K0 may have up to two negative reduced-Hessian eigenvalues that the
      // rank-2 positive semidefinite wrap term lifts (the wrap rows carry most of the curvature of
      // the last throttle, DESIGN.md "Degenerate component").  So negative pivots do not abort the
      // sweep: they are counted (Sylvester: signs of the pivots = inertia of the reduced Hessian
      // of K0) and kkt_solve decides with the 2x2 capacitance matrix.  Zero / non-finite pivots and
      // more negative pivots than the wrap term can lift are wrong inertia right away.
      // Identical in every lane (same inputs, same arithmetic) -> uniform branch.
      ric_nneg += (d0 < 0.0 ? 1 : 0) + (d1 < 0.0 ? 1 : 0) + (d2 < 0.0 ? 1 : 0);
      if (!(fabs(d0) > 0.0 && fabs(d1) > 0.0 && fabs(d2) > 0.0) || ric_nneg > (cfg.wrap ? 2 : 0)) {
        ok = false;
        break;
      }
      if (lane < 9) {
        const int c = lane;
        // rows 9-11 of column c (= entries c of rows 9-11: G is symmetric up to round-off)
        const double r0 = G[9 * kGS + c], r1 = G[10 * kGS + c], r2 = G[11 * kGS + c];
        {
          const double f1 = r1 - l10 * r0, f2 = r2 - l20 * r0 - l21 * f1;
          const double z2 = f2 * id2;
          const double z1 = f1 * id1 - l21 * z2;
          const double z0 = r0 * id0 - l10 * z1 - l20 * z2;
          st[kOffK + c] = -z0;
          st[kOffK + 9 + c] = -z1;
          st[kOffK + 18 + c] = -z2;
#pragma unroll
          for (int a = 0; a < 9; ++a)
            pr[a] = G[a * kGS + c] - G[9 * kGS + a] * z0 - G[10 * kGS + a] * z1 - G[11 * kGS + a] * z2;
        }
#pragma unroll
        for (int j = 0; j < kRhs; ++j) {
          if (j < nrhs) {
            const double g0 = G[9 * kGS + 12 + j], g1 = G[10 * kGS + 12 + j], g2 = G[11 * kGS + 12 + j];
            const double f1 = g1 - l10 * g0, f2 = g2 - l20 * g0 - l21 * f1;
            const double z2 = f2 * id2;
            const double z1 = f1 * id1 - l21 * z2;
            const double z0 = g0 * id0 - l10 * z1 - l20 * z2;
            if (c == j) {
              st[kOffKf + j * 3] = -z0;
              st[kOffKf + j * 3 + 1] = -z1;
              st[kOffKf + j * 3 + 2] = -z2;
            }
            pj[j] = G[c * kGS + 12 + j] - r0 * z0 - r1 * z1 - r2 * z2;
          }
        }
      }
      // no barrier here: round 1 of the next stage reads registers and writes T, whose last
      // readers (round 2) are behind the barrier above
    }
    if (!ok) return false;
    syncwarp();
    // ---- forward sweep: lane groups 0-10 / 11-21 / 22-31 each roll out ONE right-hand side;
    //      the group leaders park (z_{k+1}, w_k) in stage k's (now dead) gain slot -----------
    const int grp = (nrhs == 1) ? 0 : (lane < 11 ? 0 : (lane < 22 ? 1 : 2));
    const bool leader = (nrhs == 1) ? (lane == 0) : (lane == 0 || lane == 11 || lane == 22);
    const double bsel = (grp == 0) ? 1.0 : 0.0;
    double z[9];
    for (int a = 0; a < 9; ++a) z[a] = 0.0;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
      double* st = ws.st(k);
      double A[kLinEuler], Ax[12];
      for (int c = 0; c < kLinEuler; ++c) A[c] = st[kOffA + c];
      if (IG)
        for (int c = 0; c < 12; ++c) Ax[c] = st[kOffAx + c];
      double w[3];
      for (int i = 0; i < 3; ++i) {
        double acc = st[kOffKf + grp * 3 + i];
        for (int a = 0; a < 9; ++a) acc += st[kOffK + i * 9 + a] * z[a];
        w[i] = acc;
      }
      double zn[9];
      applyMx<IG>(z, w, A, Ax, Ts, zn);
      for (int a = 0; a < 6; ++a) zn[a] += bsel * st[kOffB + a];
      for (int a = 0; a < 9; ++a) z[a] = zn[a];
      syncwarp();                       // every lane has read K_k before it is overwritten
      if (leader) {
        double* dst = st + kOffK + grp * 9;
        for (int a = 0; a < 7; ++a) dst[a] = zn[a];
        dst[7] = w[0];
        dst[8] = w[1];
      }
    }
    syncwarp();
    return true;      // stage k's gain slot now holds (dz_{k+1} (7), dw_k (2)) of the 3 right-hand sides
  }

  // -------------------------------------------------------------------------
  // Full KKT solve: Riccati + Woodbury combination + costate recursion.
  // Outputs dx, dsh, du (primal step) and newlam / newy0 = the NEW equality
  // multipliers (y + dy) of the linearised problem.
  // -------------------------------------------------------------------------
  double newlam[6], newy0[7];

  MPCR_D bool kkt_solve(double dw) {
    if (!riccati(dw)) return false;
    // roll-out results of this lane's stage, read where they are used (they were solver members
    // once, i.e. a round trip through local memory)
    double mz[kRhs][7];   // (dx_{k+1} (6), ds_{k+1}) for the 3 right-hand sides
    double mw[kRhs][2];   // (dthr_k, dst_k)
    {
      const int nrhs = cfg.wrap ? kRhs : 1;
      const double* src = ws.st(act ? lane : 0) + kOffK;
#pragma unroll
      for (int j = 0; j < kRhs; ++j) {
        const bool use = act && j < nrhs;
#pragma unroll
        for (int a = 0; a < 7; ++a) mz[j][a] = use ? src[j * 9 + a] : 0.0;
        mw[j][0] = use ? src[j * 9 + 7] : 0.0;
        mw[j][1] = use ? src[j * 9 + 8] : 0.0;
      }
    }
    // Woodbury: x = x0 + sum_k D_k t_k X_k with (I - M D) t = t0, M_ik = a_i' X_k
    double c1 = 0.0, c2 = 0.0;
    if (cfg.wrap) {
      double t0[2], Mm[2][2];
      for (int i = 0; i < 2; ++i) {
        t0[i] = shfl(mw[0][i], 0) - shfl(mw[0][i], N - 1);
        for (int kk = 0; kk < 2; ++kk)
          Mm[i][kk] = shfl(mw[1 + kk][i], 0) - shfl(mw[1 + kk][i], N - 1);
      }
      const double a00 = 1.0 - Mm[0][0] * Dw[0], a01 = -Mm[0][1] * Dw[1];
      const double a10 = -Mm[1][0] * Dw[0], a11 = 1.0 - Mm[1][1] * Dw[1];
      const double det = a00 * a11 - a01 * a10;
      // Inertia of the full matrix: with S = -Dw^-1 - A' K0^-1 A = -C Dw^-1 (C = I - M Dw, the matrix
      // solved here) In(K) = In(K0) + In(S) - In(-Dw^-1), so the reduced Hessian of K is positive
      // definite iff S has exactly ric_nneg positive eigenvalues: sign(det S) = sign(det C),
      // trace S = -(C00 / Dw0 + C11 / Dw1).
      if (ric_nneg == 1 && !(det < 0.0)) return false;
      if (ric_nneg == 2 && !(det > 0.0 && a00 * Dw[1] + a11 * Dw[0] < 0.0)) return false;
      const double t1 = (a11 * t0[0] - a01 * t0[1]) / det;
      const double t2 = (-a10 * t0[0] + a00 * t0[1]) / det;
      c1 = Dw[0] * t1;
      c2 = Dw[1] * t2;
    }
    for (int a = 0; a < 6; ++a) dx[a] = mz[0][a] + c1 * mz[1][a] + c2 * mz[2][a];
    dsh = mz[0][6] + c1 * mz[1][6] + c2 * mz[2][6];
    for (int i = 0; i < 2; ++i) du[i] = mw[0][i] + c1 * mw[1][i] + c2 * mw[2][i];
    // costates: pi_j = local_j + A_j' pi_{j+1}, local_j = Q_j dz_j + S_j' dw_j + q_j
    double dzp[7];
    for (int a = 0; a < 6; ++a) {
      double t = shfl(dx[a], prev_lane());
      dzp[a] = (lane == 0) ? 0.0 : t;
    }
    {
      double t = shfl(dsh, prev_lane());
      dzp[6] = (lane == 0) ? 0.0 : t;
    }
    double* loc = ws.T();   // scratch [(N+1)][7]: T (162) + G (192) are contiguous
    if (act) {
      const double* st = ws.st(lane);
      const double v12[12] = {dzp[0], dzp[1], dzp[2], dzp[3], dzp[4], dzp[5], dzp[6], 0, 0,
                              du[0], du[1], 0};
      for (int a = 0; a < 7; ++a) {
        double acc = st[kOffGs + a];
        for (int c = 0; c < 11; ++c) {
          if (c == 7 || c == 8) continue;
          const int slot = slot_of(a, c);
          if (slot == kZeroSlot) continue;
          acc += stage_entry(st, slot, dw) * v12[c];
        }
        loc[lane * 7 + a] = acc;
      }
      if (lane == N - 1) {
        const double* sN = ws.st(N);
        const double vN[7] = {dx[0], dx[1], dx[2], dx[3], dx[4], dx[5], dsh};
        for (int a = 0; a < 7; ++a) {
          double acc = sN[kOffGs + a];
          for (int c = 0; c < 7; ++c) {
            const int slot = slot_of(a, c);
            if (slot == kZeroSlot) continue;
            acc += stage_entry(sN, slot, dw) * vN[c];
          }
          loc[N * 7 + a] = acc;
        }
      }
    }
    syncwarp();
    double pi[7];
    for (int a = 0; a < 7; ++a) pi[a] = loc[N * 7 + a];
    for (int a = 0; a < 6; ++a) newlam[a] = 0.0;
    if (lane == N - 1)
      for (int a = 0; a < 6; ++a) newlam[a] = -pi[a];
#pragma unroll 1
    for (int j = N - 1; j >= 0; --j) {
      const double* st = ws.st(j);
      const double *aX = st + kOffA, *aY = st + kOffA + 3, *d = st + kOffA + 6;
      double q[7];
      q[0] = pi[0];
      q[1] = pi[1];
      q[2] = aX[0] * pi[0] + aY[0] * pi[1] + pi[2];
      q[3] = aX[1] * pi[0] + aY[1] * pi[1] + d[0] * pi[3] + d[3] * pi[4] + d[6] * pi[5];
      q[4] = aX[2] * pi[0] + aY[2] * pi[1] + d[1] * pi[3] + d[4] * pi[4] + d[7] * pi[5];
      q[5] = cfg.Ts * pi[2] + d[2] * pi[3] + d[5] * pi[4] + d[8] * pi[5];
      q[6] = pi[6];
      if (IG) {
        const double* ax = st + kOffAx;
        q[3] += ax[6] * pi[2];
        q[4] += ax[7] * pi[2];
        q[5] += ax[0] * pi[0] + ax[3] * pi[1] + (ax[10] - cfg.Ts) * pi[2];
      }
      for (int a = 0; a < 7; ++a) pi[a] = loc[j * 7 + a] + q[a];
      if (j == lane + 1)
        for (int a = 0; a < 6; ++a) newlam[a] = -pi[a];
    }
    for (int a = 0; a < 7; ++a) newy0[a] = -pi[a];
    syncwarp();
    return true;
  }

  // -------------------------------------------------------------------------
  // Optimality error E_mu (Waechter & Biegler eq. (5)) and its ingredients
  // -------------------------------------------------------------------------
  // One pass gives E_mu for EVERY mu: only the complementarity term depends on mu and
  // max_i |p_i - mu| = max(|p_max - mu|, |p_min - mu|) over the products p_i = slack_i * v_i.
  struct Errors {
    double dual_inf, pr_inf, prod_max, prod_min, is_d, is_c, viol;      // is_d = 1 / s_d, is_c = 1 / s_c
    MPCR_D double compl_inf(double mu_t) const {
      return fmax(fabs(prod_max - mu_t), fabs(prod_min - mu_t));
    }
    MPCR_D double E(double mu_t) const {
      return fmax(dual_inf * is_d, fmax(pr_inf, compl_inf(mu_t) * is_c));
    }
  };
  int n_eq_ineq, n_bounds;   // uniform counts: rows with multipliers, bound multipliers

  MPCR_D Errors compute_errors() const {
    // t = A' lam: contributions of this stage's dynamics rows to (x_k, u_k)
    double A[kLinEuler], Ax[12], lv[9], out[12];
    lin_arrays(A, Ax);
    for (int k = 0; k < 6; ++k) lv[k] = act ? lam[k] : 0.0;
    lv[6] = lv[7] = lv[8] = 0.0;
    mulMx<IG>(lv, A, Ax, cfg.Ts, out);
    const bool last = (lane == N - 1);
    double dinf = 0.0;
    double txn[6];
    for (int a = 0; a < 6; ++a) {
      double t = shfl(out[a], next_lane());
      txn[a] = last ? 0.0 : t;
    }
    const double ydl = on(kLane) ? R.yd[kLane] : 0.0;
    const double gx[6] = {gcost[0] + ydl * gC[0], gcost[1] + ydl * gC[1], in_resto() ? R.pg[2] : 0.0, gcost[3],
                          gcost[4], in_resto() ? R.pg[5] : 0.0};
    const double yprog_n_raw = shfl(R.yd[kProg], next_lane());
    const double yprog_n = last ? 0.0 : yprog_n_raw;
    const double ydthr_n = shfl(on(kDThr) ? R.yd[kDThr] : 0.0, next_lane());
    const double ydst_n = shfl(on(kDSt) ? R.yd[kDSt] : 0.0, next_lane());
    const double sig_n_raw = shfl(sigc_prev, next_lane());
    const double sig_n = last ? 0.0 : sig_n_raw;
    double prinf = 0.0, pmax = -1.0e300, pmin = 1.0e300, sum_y = 0.0, sum_v = 0.0;
    double vio = 0.0;      // violation of the ORIGINAL constraints (unscaled, unrelaxed bounds)
    if (act) {
      for (int a = 0; a < 6; ++a) dinf = fmax(dinf, fabs(gx[a] + lam[a] - txn[a]));
      dinf = fmax(dinf, fabs(gcost[2] + ydl * gC[2] + R.yd[kProg] - yprog_n));
      double ru0 = -out[9] + R.yd[kThrUb] + R.yd[kThrLb] + (on(kDThr) ? R.yd[kDThr] : 0.0) - ydthr_n +
                   (on(kDThr0) ? R.yd[kDThr0] : 0.0) + (in_resto() ? R.pg[7] : 0.0);
      double ru1 = -out[10] + R.yd[kStUb] + R.yd[kStLb] + (on(kDSt) ? R.yd[kDSt] : 0.0) - ydst_n +
                   (on(kDSt0) ? R.yd[kDSt0] : 0.0) + (in_resto() ? R.pg[8] : 0.0) +
                   2.0 * prm().beta_delta * df * (sigc_prev - sig_n);
      dinf = fmax(dinf, fmax(fabs(ru0), fabs(ru1)));
      if (lane == 0) {
        for (int a = 0; a < 6; ++a) dinf = fmax(dinf, fabs(y0[a] - out[a]));
        dinf = fmax(dinf, fabs(y0[6] - R.yd[kProg]));
        for (int a = 0; a < 7; ++a) sum_y += fabs(y0[a]);
      }
      for (int a = 0; a < 6; ++a) {
        prinf = fmax(prinf, fabs(dcs[a] * cres[a]));
        vio = fmax(vio, fabs(cres[a]));
        sum_y += fabs(lam[a] * rcp_pivot(dcs[a]));
      }
    }
    RowRegs nx{};
    row_fetch(0, nx);
#pragma unroll 1
    for (int r = 0; r < kNumRows; ++r) {
      const RowRegs c = nx;
      row_fetch(r + 1, nx);
      if (!on(r)) continue;
      double rs = -c.yd;
      if (row_hasL(r)) {
        rs -= c.vL;
        const double pr_ = (c.s - lo(r)) * c.vL;
        pmax = fmax(pmax, pr_);
        pmin = fmin(pmin, pr_);
        sum_v += fabs(c.vL);
      }
      if (row_hasU(r)) {
        rs += c.vU;
        const double pr_ = (hi(r) - c.s) * c.vU;
        pmax = fmax(pmax, pr_);
        pmin = fmin(pmin, pr_);
        sum_v += fabs(c.vU);
      }
      dinf = fmax(dinf, fabs(rs));
      if (!in_resto()) prinf = fmax(prinf, fabs(c.rd));   // restoration: the rows sit in the objective
      sum_y += fabs(c.yd);
      double bl, bh;
      row_bounds(r, bl, bh);
      double val = c.rd + c.s;
      if (r == kLane) val /= ddl;
      if (row_hasL(r)) vio = fmax(vio, bl - val);
      if (row_hasU(r)) vio = fmax(vio, val - bh);
    }
    Errors e;
    e.dual_inf = warp_max(dinf);
    e.pr_inf = warp_max(prinf);
    e.viol = warp_max(vio);
    e.prod_max = warp_max(pmax);
    e.prod_min = warp_min(pmin);
    sum_y = warp_sum(sum_y);
    sum_v = warp_sum(sum_v);
    const double smax = o.s_max;
    // s_d = max(s_max, mean multiplier) / s_max (Waechter & Biegler eq. (6)), kept as reciprocals
    // (exactly 1 while the mean stays below s_max)
    const double qd = fmax(smax, (sum_y + sum_v) * rcp_pivot((double)(n_eq_ineq + n_bounds)));
    const double qc = fmax(smax, sum_v * rcp_pivot((double)n_bounds));
    e.is_d = (qd == smax) ? 1.0 : smax * rcp_pivot(qd);
    e.is_c = (qc == smax) ? 1.0 : smax * rcp_pivot(qc);
    return e;
  }

  // filter --------------------------------------------------------------------
  MPCR_D bool filter_ok(double th, double ph) const { return filter_ok_in(th, ph, filt_th, filt_ph, filt_n); }
  MPCR_D bool filter_ok_in(double th, double ph, double fth, double fph, int fn) const {
    int n = fn < 32 ? fn : 32;
    bool ok = (lane >= n) || (th <= fth) || (ph <= fph);
    return warp_all(ok);
  }
  MPCR_D void filter_add(double th, double ph) {
    if (lane == (filt_n & 31)) { filt_th = th; filt_ph = ph; }
    filt_n++;
  }
  MPCR_D static bool cmp_le(double a, double b, double ref) {
    return a - b <= 10.0 * 2.220446049250313e-16 * fabs(ref);
  }

  // slack steps ds_r = J_r dx + rdv_r (into dsr) and the fraction-to-the-boundary step; the same
  // pass accumulates this lane's share of grad_s(phi)' ds and of theta = |c|_1 + |d - s|_1
  MPCR_D double slack_steps(const double* rdv, double dwu, double& gs_ds, double& th_loc) {
    double dup[2], dshp;
    dup[0] = shfl(du[0], prev_lane());
    dup[1] = shfl(du[1], prev_lane());
    double t = shfl(dsh, prev_lane());
    dshp = (lane == 0) ? 0.0 : t;
    double a = 1.0, g = 0.0, th = 0.0;
    if (act)
      for (int k = 0; k < 6; ++k) th += fabs(dcs[k] * cres[k]);
    RowRegs nx{};
    row_fetch(0, nx);
    double nrdv = (rdv && on(0)) ? rdv[0] : 0.0;
#pragma unroll 1
    for (int r = 0; r < kNumRows; ++r) {
      const RowRegs c = nx;
      const double crdv = nrdv;
      row_fetch(r + 1, nx);
      if (rdv && r + 1 < kNumRows && on(r + 1)) nrdv = rdv[r + 1];
      double d = 0.0;
      if (on(r)) {
        const double jdx = row_lin(r, du, dup, dsh, dshp, dx);
        d = jdx + (rdv ? crdv : 0.0);
        if (in_resto()) {
          // slack step of a row that sits in the objective, and its share of grad(f_R)' d
          double sg, gb;
          sigma_gsb(r, c, sg, gb);
          d = (o.resto_weight * d - gb) / (o.resto_weight + sg + dwu);
          g += o.resto_weight * c.rd * (jdx - d);
        } else {
          th += fabs(c.rd);
        }
        if (row_hasL(r)) {
          const double sl = c.s - lo(r);
          if (d < 0.0) a = fmin(a, -tau * sl * rcp_pivot(d));
          g += (-mu * rcp_pivot(sl) + (row_hasU(r) ? 0.0 : o.kappa_d * mu)) * d;
        }
        if (row_hasU(r)) {
          const double su = hi(r) - c.s;
          if (d > 0.0) a = fmin(a, tau * su * rcp_pivot(d));
          g += (mu * rcp_pivot(su) - (row_hasL(r) ? 0.0 : o.kappa_d * mu)) * d;
        }
      }
      R.dsr[r] = d;
    }
    gs_ds = g;
    th_loc = th;
    return warp_min(a);
  }

  // -------------------------------------------------------------------------
  MPCR_D void solve(const SolveOut& out) {
    const double inf = 1.0e300;
    // ---- rows ---------------------------------------------------------------
    rowmask = 0;
    if (act) {
      rowmask = 0x4F;                                  // rows 0-3, 6
      if (lane > 0 || cfg.wrap) rowmask |= (1u << kDThr) | (1u << kDSt);
      if (lane == 0 && pbv().has_prev_thr) rowmask |= 1u << kDThr0;
      if (lane == 0 && pbv().has_prev_st) rowmask |= 1u << kDSt0;
      if (cfg.lane_con) rowmask |= 1u << kLane;
    }
    // ---- initial guess (control/MPC.py:109-131) ----------------------------------
    {
      double u0 = pbv().u_prev[0], u1 = pbv().u_prev[1];
      if (pbv().u_warm && act) {
        int src = (lane + 1 < N) ? lane + 1 : N - 1;   // shift by one, repeat the last
        u0 = pbv().u_warm[src];
        u1 = pbv().u_warm[N + src];
      }
      // lanes >= N compute along with dummy inputs; non-zero ones, because an exactly zero
      // numerator sends the model's fp64 divisions down their slow path for the whole warp
      u[0] = act ? u0 : 0.25;
      u[1] = act ? u1 : 0.125;
      double xr[6];
      for (int k = 0; k < 6; ++k) { xr[k] = pbv().x0[k]; x[k] = pbv().x0[k]; }
#pragma unroll 1
      for (int k = 0; k < N; ++k) {
        double uk[2] = {shfl(u[0], k), shfl(u[1], k)};
        double fn[6];
        if (IG) rk4_model(kDynamic, xr, uk, cfg.Ts, nullptr, fn, nullptr, nullptr);
        else f_dynamic(xr, uk, cfg.Ts, fn);            // always f_vehicle (MPC.py:129)
        for (int a = 0; a < 6; ++a) xr[a] = fn[a];
        if (k == lane)
          for (int a = 0; a < 6; ++a) x[a] = fn[a];
      }
      sh = act ? pbv().s0 + (lane + 1) * cfg.Ts * prm().v_max : pbv().s0;
    }
    for (int k = 0; k < 6; ++k) { lam[k] = 0.0; dcs[k] = 1.0; cres[k] = 0.0; dx[k] = 0.0; }
    for (int k = 0; k < 7; ++k) y0[k] = 0.0;
    dsh = 0.0;
    du[0] = du[1] = 0.0;
#pragma unroll 1
    for (int r = 0; r < kNumRows; ++r) {
      R.s[r] = 0.0; R.yd[r] = 0.0; R.rd[r] = 0.0; R.dsr[r] = 0.0;
      R.vL[r] = (on(r) && row_hasL(r)) ? o.bound_mult_init_val : 0.0;
      R.vU[r] = (on(r) && row_hasU(r)) ? o.bound_mult_init_val : 0.0;
    }
    df = 1.0;
    ddl = 1.0;
    mu = o.mu_init;
    tau = fmax(o.tau_min, 1.0 - mu);
    filt_n = 0;
    filt_th = 0.0;
    filt_ph = 0.0;
    fval = 0.0;
    barB = 0.0;
    resto_ = false;
    wanted_resto = false;
    {
      int nr = 0, nb = 0;
#pragma unroll 1
      for (int r = 0; r < kNumRows; ++r)
        if (on(r)) { nr += 1; nb += (row_hasL(r) ? 1 : 0) + (row_hasU(r) ? 1 : 0); }
      n_eq_ineq = (int)(warp_sum((double)nr) + 0.5) + 6 * N + 7;
      n_bounds = (int)(warp_sum((double)nb) + 0.5);
    }
    // provisional (unscaled, unrelaxed) bounds so that lo()/hi() are defined early
    if (lane < kNumRows) { ws.misc()[lane] = -inf; ws.misc()[kNumRows + lane] = inf; }
    syncwarp();

    // ---- phase-loop state ------------------------------------------------------------
    int phase = kPhInit;
    int status = kMaxIterExceeded;
    int it = 0, acc_count = 0, n_ls = 0, soc_count = 0, ls_fail_mode = 0;
    bool cur_finite = true, soc_done = false, f_type = false;
    double theta_max = 0.0, theta_min = 0.0, mu_floor = 0.0;
    double dw = 0.0, dw_last = 0.0, last_alpha = 0.0, last_dw = 0.0;
    double alpha = 0.0, a_soc = 0.0, amin = 0.0, gpd = 0.0, theta = 0.0, phi = 0.0, th_old = 0.0;
    double c_soc[6], d_soc[kNumRows];
    // plain (non-SOC) step saved while a second-order correction is being tried
    double sv_dx[6], sv_dsh = 0.0, sv_du[2] = {0.0, 0.0}, sv_dsr[kNumRows], sv_lam[6], sv_y0[7];
#pragma unroll 1
    for (int r = 0; r < kNumRows; ++r) { d_soc[r] = 0.0; sv_dsr[r] = 0.0; }
    for (int a = 0; a < 6; ++a) { c_soc[a] = 0.0; sv_dx[a] = 0.0; sv_lam[a] = 0.0; }
    for (int a = 0; a < 7; ++a) sv_y0[a] = 0.0;
    Point tp;
    double tp_rd[kNumRows];
    // restoration bookkeeping (the outer problem's scalars are parked in R.osave)
    bool want_resto = false, reinit = false;
    int resto_it = 0;
    // polish (mpcr_config.polish_tol > 0): after the first termination the iteration continues to the
    // tighter tolerance; status / iters of the first termination are what is reported
    double tol_cur = o.tol;
    bool polishing = false;
    int first_status = 0, first_iters = 0;

    while (true) {
      if (RESTO && want_resto) {
        // ---- enter the restoration phase at the current iterate (x_R, s_R) ---------------------
        want_resto = false;
        R.osave[11] = theta;
        filter_add((1.0 - o.gamma_theta) * theta, phi - o.gamma_phi * theta);   // the start point enters the filter
        R.osave[0] = mu; R.osave[1] = tau; R.osave[2] = df; R.osave[3] = theta_max; R.osave[4] = theta_min;
        R.osave[5] = mu_floor; R.osave[6] = filt_th; R.osave[7] = filt_ph; R.osave[8] = (double)filt_n;
        R.osave[9] = (double)acc_count;
        double vmax = 0.0;
        for (int a = 0; a < 6; ++a) {
          R.xr[a] = x[a];
          if (act) vmax = fmax(vmax, fabs(dcs[a] * cres[a]));
          lam[a] = 0.0;
          dx[a] = 0.0;
        }
        R.xr[6] = sh; R.xr[7] = u[0]; R.xr[8] = u[1];
        dsh = 0.0;
        du[0] = du[1] = 0.0;
#pragma unroll 1
        for (int r = 0; r < kNumRows; ++r) {
          R.dsr[r] = 0.0;
          if (!on(r)) continue;
          const double rd0 = R.rd[r];
          vmax = fmax(vmax, fabs(rd0));
          const double dval = rd0 + R.s[r];
          double v = R.s[r];
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
          R.s[r] = v;
          R.rd[r] = dval - v;
          R.yd[r] = o.resto_weight * (dval - v);
          R.vL[r] = row_hasL(r) ? o.bound_mult_init_val : 0.0;
          R.vU[r] = row_hasU(r) ? o.bound_mult_init_val : 0.0;
        }
        for (int a = 0; a < 7; ++a) y0[a] = 0.0;
        mu = fmax(mu, warp_max(vmax));
        tau = fmax(o.tau_min, 1.0 - mu);
        R.osave[10] = sqrt(mu);
        df = 1.0;
        mu_floor = fmax(o.mu_min, fmin(o.tol, o.compl_inf_tol) / (o.barrier_tol_factor + 1.0));
        filt_n = 0;
        acc_count = 0;
        resto_ = true;
        resto_it = 0;
        alpha = 0.0;
        phase = kPhInit2;          // evaluate the start point with the restoration objective
      }
      // ================= A. derivatives, convergence tests, barrier update ===============
      if (phase == kPhInit || phase == kPhNewton) {
        eval_derivs(phase == kPhInit ? 1 : 0);
        if (phase == kPhInit && reinit) {
          // multiplier re-initialisation after restoration: scaling, bounds and mu stay as they are
        } else if (phase == kPhInit) {
          // ---- gradient-based scaling at the starting point ----------------------------
          double sig_n_raw = shfl(sigc_prev, next_lane());
          double sig_n = (lane == N - 1) ? 0.0 : sig_n_raw;
          double g = 0.0;
          if (act) {
            for (int k = 0; k < 5; ++k) g = fmax(g, fabs(gcost[k]));
            g = fmax(g, fabs(2.0 * prm().beta_delta * (sigc_prev - sig_n)));
          }
          g = warp_max(g);
          if (g > o.nlp_scaling_max_gradient)
            df = fmax(o.nlp_scaling_max_gradient / g, o.nlp_scaling_min_value);
          double rm[6];
          rm[0] = fmax(1.0, fmax(fabs(lin.aX[0]), fmax(fabs(lin.aX[1]), fabs(lin.aX[2]))));
          rm[1] = fmax(1.0, fmax(fabs(lin.aY[0]), fmax(fabs(lin.aY[1]), fabs(lin.aY[2]))));
          rm[2] = fmax(1.0, fabs(cfg.Ts));
          if (IG) {
            for (int a = 0; a < 3; ++a) { rm[0] = fmax(rm[0], fabs(linx[a])); rm[1] = fmax(rm[1], fabs(linx[3 + a])); }
            rm[2] = fmax(1.0, fabs(linx[10]));
            for (int a = 0; a < 4; ++a) rm[2] = fmax(rm[2], fabs(linx[6 + a]));
          }
          for (int i = 0; i < 3; ++i) {
            double m = 1.0;
            for (int k = 0; k < 3; ++k) m = fmax(m, fabs(lin.d[i][k]));
            for (int k = 0; k < 2; ++k) m = fmax(m, fabs(lin.bu[i][k]));
            rm[3 + i] = m;
          }
          for (int k = 0; k < 6; ++k)
            dcs[k] = (rm[k] > o.nlp_scaling_max_gradient)
                         ? fmax(o.nlp_scaling_max_gradient / rm[k], o.nlp_scaling_min_value) : 1.0;
          double gl = fmax(fabs(gC[0]), fmax(fabs(gC[1]), fabs(gC[2])));
          if (cfg.lane_con && gl > o.nlp_scaling_max_gradient)
            ddl = fmax(o.nlp_scaling_max_gradient / gl, o.nlp_scaling_min_value);
          // the lane-row scale is made uniform over the warp (its bounds live in shared memory)
          ddl = warp_min(act ? ddl : 1.0);
          for (int k = 0; k < 5; ++k) gcost[k] *= df;    // eval_derivs ran with df = ddl = 1
          for (int k = 0; k < 3; ++k) gC[k] *= ddl;
          mu_floor = fmax(o.mu_min, fmin(o.tol, o.compl_inf_tol * df) / (o.barrier_tol_factor + 1.0));
          // ---- relaxed bounds ------------------------------------------------------------
          if (lane < kNumRows) {
            double l, h;
            row_bounds(lane, l, h);
            if (lane == kLane) { l *= ddl; h *= ddl; }
            if (row_hasL(lane)) l -= fmin(o.constr_viol_tol, o.bound_relax_factor * fmax(1.0, fabs(l)));
            else l = -inf;
            if (row_hasU(lane)) h += fmin(o.constr_viol_tol, o.bound_relax_factor * fmax(1.0, fabs(h)));
            else h = inf;
            ws.misc()[lane] = l;
            ws.misc()[kNumRows + lane] = h;
          }
          syncwarp();
        } else {
          if (!cur_finite) { status = kInvalidNumber; break; }
          bool done = false;
          {
            const Errors e = compute_errors();
            const double E0 = e.E(0.0), cmp0 = e.compl_inf(0.0);
            const double viol = e.viol;
            if (out.trace && lane == 0) {
              double* tr = out.trace + (size_t)it * kTraceCols;
              tr[0] = fval / df; tr[1] = E0; tr[2] = e.pr_inf; tr[3] = e.dual_inf;
              tr[4] = cmp0; tr[5] = mu; tr[6] = last_alpha; tr[7] = last_dw;
            }
            bool again;
            do {
              again = false;
              const bool can_polish = !polishing && !in_resto() && o.polish_tol > 0.0 && o.polish_tol < tol_cur;
              if (!finite_d(E0)) {
                status = kInvalidNumber;
                done = true;
              } else if (E0 <= tol_cur && e.dual_inf / df <= o.dual_inf_tol && (in_resto() || viol <= o.constr_viol_tol) &&
                         cmp0 / df <= o.compl_inf_tol) {
                // restoration: a stationary point of the infeasibility measure that the filter rejects
                status = in_resto() ? kRestorationNeeded : kSolveSucceeded;
                done = true;
                if (can_polish) again = true;
              } else {
                if (!polishing && !in_resto() && E0 <= o.acceptable_tol && e.dual_inf / df <= o.acceptable_dual_inf_tol &&
                    viol <= o.acceptable_constr_viol_tol && cmp0 / df <= o.acceptable_compl_inf_tol) {
                  acc_count++;
                  if (acc_count >= o.acceptable_iter) {
                    status = kSolvedToAcceptableLevel;
                    done = true;
                    if (can_polish) again = true;
                  }
                } else {
                  acc_count = 0;
                }
                // IPOPT does not start the restoration phase at an almost feasible point: a failed line
                // search there restores the stored acceptable point (here: only when that is this
                // iterate) or ends "Restoration Failed" (IpBacktrackingLineSearch.cpp).
                ls_fail_mode = (!in_resto() && e.pr_inf <= 1e-2 * o.tol) ? (acc_count > 0 ? 1 : 2) : 0;
                if (!done && it >= o.max_iter) { status = kMaxIterExceeded; done = true; }
              }
              if (again) {
                // polish: keep iterating to the tighter tolerance; this termination is what is reported
                polishing = true;
                first_status = status;
                first_iters = it;
                tol_cur = o.polish_tol;
                mu_floor = fmax(o.mu_min, fmin(tol_cur, o.compl_inf_tol * df) / (o.barrier_tol_factor + 1.0));
                done = false;
              }
            } while (again);
            // monotone barrier update (Waechter & Biegler eq. (7)); may fire repeatedly
            while (!done && e.E(mu) <= o.barrier_tol_factor * mu && mu > mu_floor) {
              mu = fmax(mu_floor, fmin(o.mu_linear_decrease_factor * mu,
                                       powd(mu, o.mu_superlinear_decrease_power)));
              tau = fmax(o.tau_min, 1.0 - mu);
              filt_n = 0;
            }
          }
          if (done) break;
          dw = 0.0;
        }
      }
      // ================= B. linear solve (single call site) ===================================
      if (phase == kPhInit || phase == kPhNewton || phase == kPhSoc) {
        const int mode = (phase == kPhInit) ? 1 : 0;
        const double* rdv = (phase == kPhInit) ? nullptr : ((phase == kPhSoc) ? d_soc : R.rd);
        const double* cv = (phase == kPhInit) ? nullptr : ((phase == kPhSoc) ? c_soc : cres);
        double dwu = (phase == kPhInit) ? 1.0 : dw;
        bool first_try = true, ok = false, failed = false;
#pragma unroll 1
        while (true) {
          write_row_terms(mode, dwu, rdv, cv);
          syncwarp();
          ok = kkt_solve(dwu);
          if (ok || phase != kPhNewton) break;
          // inertia correction (IPOPT's PDPerturbationHandler schedule)
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
        double gs_ds, th_loc;
        const double a_new = slack_steps(rdv, dwu, gs_ds, th_loc);     // dsr = J_d dx + rdv ; FTB step
        if (phase == kPhInit) {
          // ---- least-squares multipliers: y = -(QP multipliers), yd = (vU - vL) - J_d dx ----
          double ymax = 0.0;
          if (ok && act)
            for (int a = 0; a < 6; ++a) { lam[a] = -newlam[a]; ymax = fmax(ymax, fabs(lam[a] / dcs[a])); }
#pragma unroll 1
          for (int r = 0; r < kNumRows; ++r) {
            if (!on(r) || !ok) continue;
            R.yd[r] = ((row_hasU(r) ? R.vU[r] : 0.0) - (row_hasL(r) ? R.vL[r] : 0.0)) - R.dsr[r];
            ymax = fmax(ymax, fabs(R.yd[r]));
          }
          for (int a = 0; a < 7; ++a) { y0[a] = ok ? -newy0[a] : 0.0; ymax = fmax(ymax, fabs(y0[a])); }
          ymax = warp_max(ymax);
          if (!ok || !(ymax <= o.constr_mult_init_max)) {
            for (int a = 0; a < 6; ++a) lam[a] = 0.0;
            for (int a = 0; a < 7; ++a) y0[a] = 0.0;
#pragma unroll 1
            for (int r = 0; r < kNumRows; ++r) R.yd[r] = 0.0;
          }
          for (int a = 0; a < 6; ++a) dx[a] = 0.0;     // the init phases evaluate the point itself
          dsh = 0.0;
          du[0] = du[1] = 0.0;
#pragma unroll 1
          for (int r = 0; r < kNumRows; ++r) R.dsr[r] = 0.0;
          alpha = 0.0;
        } else if (phase == kPhNewton) {
          if (dw > 0.0) dw_last = dw;
          last_dw = dw;
          alpha = a_new;
          // grad(phi)' d
          double g = gs_ds;
          {
            const double dup1 = shfl(du[1], prev_lane());
            if (act && in_resto()) {
              const double d9[9] = {dx[0], dx[1], dx[2], dx[3], dx[4], dx[5], dsh, du[0], du[1]};
#pragma unroll 1
              for (int a = 0; a < 9; ++a) g += R.pg[a] * d9[a];
            } else if (act) {
              g += gcost[0] * dx[0] + gcost[1] * dx[1] + gcost[2] * dsh + gcost[3] * dx[3] +
                   gcost[4] * dx[4];
              if (lane >= 1) g += 2.0 * prm().beta_delta * df * sigc_prev * (du[1] - dup1);
            }
          }
          gpd = warp_sum(g);
          theta = warp_sum(th_loc);
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
            // abandon the correction: restore the plain step and keep backtracking
            for (int a = 0; a < 6; ++a) { dx[a] = sv_dx[a]; newlam[a] = sv_lam[a]; }
            for (int a = 0; a < 7; ++a) newy0[a] = sv_y0[a];
            dsh = sv_dsh;
            du[0] = sv_du[0];
            du[1] = sv_du[1];
#pragma unroll 1
            for (int r = 0; r < kNumRows; ++r) R.dsr[r] = sv_dsr[r];
            alpha *= o.alpha_red_factor;
            phase = kPhBacktrack;
            if (!(alpha >= amin && n_ls < 60)) {
              if (RESTO && o.restoration && !resto_ && ls_fail_mode == 0) { want_resto = true; continue; }
              if (!RESTO && o.restoration && ls_fail_mode == 0) wanted_resto = true;
              status = ls_fail_mode == 1 ? kSolvedToAcceptableLevel : kRestorationNeeded;
              break;
            }
          } else {
            a_soc = a_new;
          }
        }
      }
      // ================= C. trial point (single call site) =======================================
      {
        const bool check = (phase == kPhRestoCheck);
        const double at = check ? 0.0 : ((phase == kPhSoc) ? a_soc : alpha);
        for (int a = 0; a < 6; ++a) tp.x[a] = x[a] + at * dx[a];
        tp.sh = sh + at * dsh;
        tp.u[0] = u[0] + at * du[0];
        tp.u[1] = u[1] + at * du[1];
        eval_point(tp, tp_rd, at, in_resto() && !check, check ? R.osave[2] : df);
      }
      // ================= D. decisions ===========================================================
      if (RESTO && phase == kPhRestoCheck) {
        // ---- leave the restoration phase?  The point must be acceptable to the ORIGINAL filter and
        //      have reduced the original infeasibility by kappa_resto ------------------------------
        bool leave = tp.finite && tp.theta <= o.kappa_resto * R.osave[11] && tp.theta <= R.osave[3];
        if (leave)
          leave = filter_ok_in(tp.theta, tp.f + R.osave[0] * tp.barB, R.osave[6], R.osave[7], (int)R.osave[8]);
        if (leave) {
          resto_ = false;
          mu = R.osave[0]; tau = R.osave[1]; df = R.osave[2]; theta_max = R.osave[3]; theta_min = R.osave[4];
          mu_floor = R.osave[5]; filt_th = R.osave[6]; filt_ph = R.osave[7]; filt_n = (int)R.osave[8];
          acc_count = (int)R.osave[9];
#pragma unroll 1
          for (int r = 0; r < kNumRows; ++r) {
            R.vL[r] = (on(r) && row_hasL(r)) ? o.bound_mult_init_val : 0.0;
            R.vU[r] = (on(r) && row_hasU(r)) ? o.bound_mult_init_val : 0.0;
          }
          reinit = true;           // least-squares multipliers at the restored point, then resume
          phase = kPhInit;
          continue;
        }
        if (resto_it >= o.resto_max_iter) { status = kRestorationNeeded; break; }
        phase = kPhNewton;
        continue;
      }
      if (phase == kPhInit && reinit) {
        phase = kPhInit2;          // the slacks of the restored point are kept
        continue;
      }
      if (phase == kPhInit) {
        // slacks pushed into their (relaxed) bounds; tp.rd holds the row values (s was 0)
#pragma unroll 1
        for (int r = 0; r < kNumRows; ++r) {
          if (!on(r)) { R.s[r] = 0.0; continue; }
          double v = tp_rd[r];
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
          R.s[r] = v;
        }
        phase = kPhInit2;
        continue;
      }
      if (phase == kPhInit2) {
        for (int k = 0; k < 6; ++k) cres[k] = tp.c[k];
#pragma unroll 1
        for (int r = 0; r < kNumRows; ++r) R.rd[r] = tp_rd[r];
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
            // correction failed: back to the plain step, shorten it
            for (int a = 0; a < 6; ++a) { dx[a] = sv_dx[a]; newlam[a] = sv_lam[a]; }
            for (int a = 0; a < 7; ++a) newy0[a] = sv_y0[a];
            dsh = sv_dsh;
            du[0] = sv_du[0];
            du[1] = sv_du[1];
#pragma unroll 1
            for (int r = 0; r < kNumRows; ++r) R.dsr[r] = sv_dsr[r];
            alpha *= o.alpha_red_factor;
            phase = kPhBacktrack;
            if (!(alpha >= amin && n_ls < 60)) {
              if (RESTO && o.restoration && !resto_ && ls_fail_mode == 0) { want_resto = true; continue; }
              if (!RESTO && o.restoration && ls_fail_mode == 0) wanted_resto = true;
              status = ls_fail_mode == 1 ? kSolvedToAcceptableLevel : kRestorationNeeded;
              break;
            }
          } else {
            th_old = tp.theta;
            for (int a = 0; a < 6; ++a) c_soc[a] = a_soc * c_soc[a] + tp.c[a];
#pragma unroll 1
            for (int r = 0; r < kNumRows; ++r) d_soc[r] = a_soc * d_soc[r] + tp_rd[r];
          }
          continue;
        }
        // plain step rejected: second-order correction on the first trial, else backtrack
        if (!soc_done && n_ls == 1 && tp.finite && tp.theta >= theta && o.max_soc > 0) {
          soc_done = true;
          soc_count = 0;
          th_old = theta;
          for (int a = 0; a < 6; ++a) {
            c_soc[a] = alpha * cres[a] + tp.c[a];
            sv_dx[a] = dx[a];
            sv_lam[a] = newlam[a];
          }
          for (int a = 0; a < 7; ++a) sv_y0[a] = newy0[a];
          sv_dsh = dsh;
          sv_du[0] = du[0];
          sv_du[1] = du[1];
#pragma unroll 1
          for (int r = 0; r < kNumRows; ++r) {
            d_soc[r] = alpha * R.rd[r] + tp_rd[r];
            sv_dsr[r] = R.dsr[r];
          }
          phase = kPhSoc;
          continue;
        }
        alpha *= o.alpha_red_factor;
        phase = kPhBacktrack;
        if (!(alpha >= amin && n_ls < 60)) {
              if (RESTO && o.restoration && !resto_ && ls_fail_mode == 0) { want_resto = true; continue; }
              if (!RESTO && o.restoration && ls_fail_mode == 0) wanted_resto = true;
              status = ls_fail_mode == 1 ? kSolvedToAcceptableLevel : kRestorationNeeded;
              break;
            }
        continue;
      }
      // ================= E. accept the trial point ===============================================
      if (!f_type) filter_add((1.0 - o.gamma_theta) * theta, phi - o.gamma_phi * theta);
      const double astep = (phase == kPhSoc) ? a_soc : alpha;
      // dual steps: dyd = D ds + gsb - yd ; dv from the linearised complementarity
      double az = 1.0;
      RowRegs nx{};
      row_fetch(0, nx);
#pragma unroll 1
      for (int r = 0; r < kNumRows; ++r) {
        const RowRegs c = nx;
        row_fetch(r + 1, nx);
        if (!on(r)) continue;
        double sg, gb, isl, isu;
        sigma_gsb(r, c, sg, gb, isl, isu);
        const double dyd = (sg + dw) * c.dsr + gb - c.yd;
        R.yd[r] = c.yd + astep * dyd;
        if (row_hasL(r)) {
          const double a = (mu - c.vL * c.dsr) * isl - c.vL;      // dvL = S^-1 (mu - vL ds) - vL
          if (a < 0.0) az = fmin(az, -tau * c.vL * rcp_pivot(a));
        }
        if (row_hasU(r)) {
          const double b = (mu + c.vU * c.dsr) * isu - c.vU;
          if (b < 0.0) az = fmin(az, -tau * c.vU * rcp_pivot(b));
        }
      }
      az = warp_min(az);
      for (int a = 0; a < 6; ++a) {
        x[a] = tp.x[a];
        cres[a] = tp.c[a];
        if (act) lam[a] += astep * (newlam[a] - lam[a]);
      }
      for (int a = 0; a < 7; ++a) y0[a] += astep * (newy0[a] - y0[a]);
      sh = tp.sh;
      u[0] = tp.u[0];
      u[1] = tp.u[1];
      fval = tp.f;
      barB = tp.barB;
      cur_finite = tp.finite;
      row_fetch(0, nx);
      const double iks = 1.0 / o.kappa_sigma;
      double ntrd = on(0) ? tp_rd[0] : 0.0;
#pragma unroll 1
      for (int r = 0; r < kNumRows; ++r) {
        const RowRegs c = nx;
        const double ctrd = ntrd;
        row_fetch(r + 1, nx);
        if (r + 1 < kNumRows && on(r + 1)) ntrd = tp_rd[r + 1];
        if (!on(r)) continue;
        const double sold = c.s;
        const double snew = sold + astep * c.dsr;
        if (row_hasL(r)) {
          const double islo = rcp_pivot(sold - lo(r)), ms = mu * rcp_pivot(snew - lo(r));
          const double dv = (mu - c.vL * c.dsr) * islo - c.vL;
          const double v = c.vL + az * dv;
          R.vL[r] = fmax(fmin(v, o.kappa_sigma * ms), ms * iks);
        }
        if (row_hasU(r)) {
          const double isuo = rcp_pivot(hi(r) - sold), ms = mu * rcp_pivot(hi(r) - snew);
          const double dv = (mu + c.vU * c.dsr) * isuo - c.vU;
          const double v = c.vU + az * dv;
          R.vU[r] = fmax(fmin(v, o.kappa_sigma * ms), ms * iks);
        }
        R.s[r] = snew;
        R.rd[r] = ctrd;
        if (in_resto()) R.yd[r] = o.resto_weight * ctrd;
      }
      last_alpha = astep;
      it++;
      if (in_resto()) resto_it++;
      phase = in_resto() ? kPhRestoCheck : kPhNewton;
    }

    // ---- outputs ---------------------------------------------------------------------------
    {
      double xp[6], shp, up[2];
      gather_prev(x, sh, u, xp, shp, up);
      Poly4 px, py;
      poly4_eval(pbv().cx, shp, 1, px);
      poly4_eval(pbv().cy, shp, 1, py);
      const double dX = xp[0] - px.p, dY = xp[1] - py.p;
      if (act) {
        out.U[lane] = u[0];
        out.U[N + lane] = u[1];
        for (int a = 0; a < 6; ++a) out.X[a * (N + 1) + lane + 1] = x[a];
        out.S[lane + 1] = sh;
        out.eC[lane] = py.d1 * dX - px.d1 * dY;      // e_hat_C(S_hat[i], States[:, i]), i < N
        out.eL[lane] = -px.d1 * dX - py.d1 * dY;
      }
      if (lane == 0) {
        for (int a = 0; a < 6; ++a) out.X[a * (N + 1)] = pbv().x0[a];
        out.S[0] = pbv().s0;
        out.obj[0] = fval / df;
        out.status[0] = polishing ? first_status : status;   // polish: the first termination is what is reported
        out.iters[0] = polishing ? first_iters : it;
      }
      if (out.lam_g) {
        const int bs = 7 + (cfg.lane_con ? 1 : 0);
        const int base2 = 7 + N * bs;
        const double idf = 1.0 / df;
        if (lane == 0) {
          out.lam_g[0] = y0[6] * idf;
          for (int a = 0; a < 6; ++a) out.lam_g[1 + a] = y0[a] * idf;
        }
        if (act) {
          double* l1 = out.lam_g + 7 + lane * bs;     // rows of stage i = lane + 1
          for (int a = 0; a < 6; ++a) l1[a] = lam[a] * idf;
          l1[6] = R.yd[kProg] * idf;
          if (cfg.lane_con) l1[7] = R.yd[kLane] * ddl * idf;
          int off = cfg.wrap ? 6 * lane : (lane == 0 ? 0 : 4 + 6 * (lane - 1));
          double* l2 = out.lam_g + base2 + off;
          l2[0] = R.yd[kThrUb] * idf;
          l2[1] = R.yd[kThrLb] * idf;
          l2[2] = R.yd[kStUb] * idf;
          l2[3] = R.yd[kStLb] * idf;
          if (on(kDThr)) { l2[4] = R.yd[kDThr] * idf; l2[5] = R.yd[kDSt] * idf; }
          if (lane == 0) {
            double* l3 = out.lam_g + base2 + (cfg.wrap ? 6 * N : 6 * N - 2);
            int q = 0;
            if (on(kDThr0)) l3[q++] = R.yd[kDThr0] * idf;
            if (on(kDSt0)) l3[q++] = R.yd[kDSt0] * idf;
          }
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------
// One problem of a batch: load inputs, run the solver, store the outputs.
// Called by every lane of a warp with the same arguments.
// ---------------------------------------------------------------------------
MPCR_HD int lam_g_stride(const Config& c) {
  return 7 + c.N * (7 + (c.lane_con ? 1 : 0)) + 6 * c.N + 2;
}

// Returns true when the plain build gave up where the restoration phase would have started.
template <int MODEL, bool RESTO = false, int IG = 0>
MPCR_D bool solve_problem(const Config& cfg, const BatchIn& in, const BatchOut& out, int b,
                          int trace_rows, double* wsbase) {
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
    // a row outside the accepted range poisons this problem only: the solver ends it with
    // Invalid_Number_Detected (include/mpcr.h)
    bool ok = true;
    for (int k = 0; k < 5; ++k) ok = ok && (q[k] == q[k]) && fabs(q[k]) < 1.0e300;
    if (!(ok && q[0] >= 0.0 && q[2] >= 0.0 && q[4] >= 0.0 && q[3] >= 0.5 && q[3] < 64.0)) {
      prm.alpha_c = nan("");
      prm.n_exp = 2;
    }
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
    for (int k = lane_id(); k < ls; k += 32) so.lam_g[k] = 0.0;
    syncwarp();
  }
  RowState rows;
  WarpSolver<MODEL, RESTO, IG> solver(cfg, rows, wsbase);
  syncwarp();                       // the previous problem of this warp is done with the workspace
  if (lane_id() == 0) {
    *reinterpret_cast<Params*>(solver.ws.prm_mem()) = prm;
    *reinterpret_cast<ProblemView*>(solver.ws.pb_mem()) = pv;
  }
  for (int k = lane_id(); k <= N; k += 32) solver.ws.st(k)[kOffZero] = 0.0;   // the always-zero slot of every stage record
  syncwarp();
  solver.solve(so);
  return solver.wanted_resto;
}

// The restoration build as an out-of-line call (own register allocation; arguments by value so that
// no address of a kernel parameter is taken on the hot path): problems that need the phase are
// re-solved from the start with it.
template <int MODEL, int IG>
MPCR_NOINLINE void solve_problem_with_restoration(const Config* cfg, BatchIn in, BatchOut out, int b,
                                                  int trace_rows, double* wsbase) {
  solve_problem<MODEL, true, IG>(*cfg, in, out, b, trace_rows, wsbase);
}

// ---------------------------------------------------------------------------
// Kernel body: the warp drains the batch's work queue - fetch a problem, solve it, fetch the next
// (converged problems free their warp immediately).
// ---------------------------------------------------------------------------
#if defined(MPCR_EMU)
MPCR_D int atomic_next(int* counter) { return (*counter)++; }
#else
MPCR_D int atomic_next(int* counter) { return atomicAdd(counter, 1); }
#endif

struct WorkQueue {
  int* counter;        // next position of the dispatch order
  const int* order;    // dispatch order or null (= index order)
  int B;
};

// RMODE: 0 = no restoration phase; 1 = two-stage (mpcr_config.restoration = 1): every problem runs
// the plain solver, the few that need the phase are solved again by the build that carries it;
// 2 = every problem runs the restoration build (mpcr_config.restoration = 2: workloads in which many
// problems need the phase, e.g. the lane rows - re-solving 5 % of a batch costs more than the 6 %
// the phase's code costs everybody).
template <int MODEL, int RMODE, int IG>
MPCR_D void solve_batch(const Config& cfg, const BatchIn& in, const BatchOut& out, int trace_rows,
                        const WorkQueue& q, double* wsbase) {
  while (true) {
    int b = 0;
    if (lane_id() == 0) {
      b = atomic_next(q.counter);
      if (q.order && b < q.B) b = q.order[b];      // dispatch order (schedule kernels); results do not depend on it
    }
    b = shfl_i(b, 0);
    if (b >= q.B) break;
    if (RMODE == 2) {
      solve_problem<MODEL, true, IG>(cfg, in, out, b, trace_rows, wsbase);
      syncwarp();
    } else {
      // two-stage: the hot path is the plain solver (no restoration code, registers or instruction-
      // cache footprint); a problem whose line search gives up where IPOPT would restore is solved
      // again, from the start, by the build that carries the phase - the result that build gives on
      // its own, since it follows the same path up to there.
      const bool again = solve_problem<MODEL, false, IG>(cfg, in, out, b, trace_rows, wsbase);
      syncwarp();
      if (RMODE == 1 && again) {
        solve_problem_with_restoration<MODEL, IG>(&cfg, in, out, b, trace_rows, wsbase);
        syncwarp();
      }
    }
  }
}

template <int RMODE, int IG>
MPCR_D void solve_batch_model(const Config& cfg, const BatchIn& in, const BatchOut& out, int trace_rows,
                              const WorkQueue& q, double* wsbase) {
  if (cfg.model == kDynamic) solve_batch<kDynamic, RMODE, IG>(cfg, in, out, trace_rows, q, wsbase);
  else if (cfg.model == kKinematic) solve_batch<kKinematic, RMODE, IG>(cfg, in, out, trace_rows, q, wsbase);
  else solve_batch<kBlended, RMODE, IG>(cfg, in, out, trace_rows, q, wsbase);
}

// host builds (tests/emu): runtime dispatch over every instantiation
MPCR_D void solve_batch_any(const Config& cfg, const BatchIn& in, const BatchOut& out, int trace_rows,
                            const WorkQueue& q, double* wsbase) {
  const int rs = cfg.o.restoration;
  const bool ig = cfg.integrator != 0;
  if (rs == 2) { if (ig) solve_batch_model<2, 1>(cfg, in, out, trace_rows, q, wsbase); else solve_batch_model<2, 0>(cfg, in, out, trace_rows, q, wsbase); }
  else if (rs == 1) { if (ig) solve_batch_model<1, 1>(cfg, in, out, trace_rows, q, wsbase); else solve_batch_model<1, 0>(cfg, in, out, trace_rows, q, wsbase); }
  else { if (ig) solve_batch_model<0, 1>(cfg, in, out, trace_rows, q, wsbase); else solve_batch_model<0, 0>(cfg, in, out, trace_rows, q, wsbase); }
}

}  // namespace mpcr
