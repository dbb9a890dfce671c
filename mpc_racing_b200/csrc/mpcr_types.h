// Plain-old-data types shared by the host API, the CUDA kernels and the
// SIMT emulation harness used by the CPU tests.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MPCR_HD __host__ __device__ __forceinline__
#define MPCR_D __device__ __forceinline__
#else
#define MPCR_HD inline
#define MPCR_D inline
#endif

namespace mpcr {

constexpr int kMaxN = 32;          // one warp lane per stage transition
constexpr int kNumRows = 10;       // slack rows owned by one lane (see ipm.cuh)

enum Model : int { kDynamic = 0, kKinematic = 1, kBlended = 2 };

// control/ControllerParameters.py:4-32 + control/MPC.py:59 (min_s_delta)
struct Params {
  double lambda_s, alpha_L;
  double min_steer, max_steer, min_throttle;
  double max_steer_delta, min_steer_delta, max_throttle_delta, min_throttle_delta;
  double q_v_max, v_max, min_s_delta;
  double alpha_c, d_max, q_v_y, beta_delta;
  int n_exp;
};

// IPOPT options: the ones control/MPC.py:152-161 sets + IPOPT 3.14 defaults.
struct IpmOptions {
  int max_iter;
  double tol, acceptable_tol;
  int acceptable_iter;
  double dual_inf_tol, constr_viol_tol, compl_inf_tol;
  double acceptable_dual_inf_tol, acceptable_constr_viol_tol, acceptable_compl_inf_tol;
  double nlp_scaling_max_gradient, nlp_scaling_min_value;
  double bound_relax_factor, bound_push, bound_frac;
  double mu_init, mu_linear_decrease_factor, mu_superlinear_decrease_power;
  double barrier_tol_factor, mu_min, tau_min, kappa_d, kappa_sigma;
  double constr_mult_init_max, bound_mult_init_val, s_max;
  double theta_max_fact, theta_min_fact, gamma_theta, gamma_phi;
  double delta, s_theta, s_phi, eta_phi, alpha_min_frac, alpha_red_factor;
  int max_soc;
  double kappa_soc;
  double delta_w_init, delta_w_min, delta_w_max;
  double kappa_w_minus, kappa_w_plus, kappa_w_plus_first;
  // feasibility restoration (quadratic-penalty variant): 0 = off (status -2 instead)
  int restoration;
  int resto_max_iter;
  double resto_weight, kappa_resto;
  double polish_tol;     // > 0: keep iterating to this tolerance after the first termination
};

struct Config {
  int N;            // horizon (<= kMaxN)
  double Ts;
  int model;        // Model
  int wrap;         // keep the i=0 wrap-around rate rows of control/MPC.py:142-143
  int lane_con;     // enable the lane row commented out at control/MPC.py:135
  int integrator;   // 0 = explicit Euler (reference), 1 = RK4 (extension)
  Params p;
  IpmOptions o;
};

// return statuses: >= 0 is what CasADi treats as success
enum Status : int {
  kSolveSucceeded = 0,
  kSolvedToAcceptableLevel = 1,
  kMaxIterExceeded = -1,
  kRestorationNeeded = -2,
  kErrorInStepComputation = -3,
  kInvalidNumber = -4,
};

// Batched problem data, all row-major [B, ...] fp64, device pointers.
struct BatchIn {
  const double* x0;       // [B,6]
  const double* u_prev;   // [B,2]  NaN = None
  const double* s0;       // [B]
  const double* cx;       // [B,5]  highest order first
  const double* cy;       // [B,5]
  const double* max_err;  // [B] or null
  const double* params;   // [B,5] (alpha_c, d_max, q_v_y, n, beta_delta) or null
  const double* u_warm;   // [B,2,N] or null
};

struct BatchOut {
  double* U;        // [B,2,N]
  double* X;        // [B,6,N+1]
  double* S;        // [B,N+1]
  double* eC;       // [B,N]
  double* eL;       // [B,N]
  double* lam_g;    // [B,13N+7(+2)(+N)] or null
  double* obj;      // [B]
  int32_t* status;  // [B]
  int32_t* iters;   // [B]
  double* trace;    // optional debug trace [B, max_iter+1, kTraceCols] or null
};

constexpr int kTraceCols = 8;   // obj, E0, pr_inf, dual_inf, compl, mu, alpha, delta_w

}  // namespace mpcr
