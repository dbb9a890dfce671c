/* mpcr.h - C ABI of the B200-native batched NMPC solver.
 *
 * Drop-in boundary for the hot path of AlexGisi/mpc-racing (reference paths are
 * relative to that repository):
 *
 *   mpcr_solve_batch   replaces  control/MPC.py:10-181  (MPC.__init__: build the
 *                      Opti NLP, solve it with IPOPT) and control/MPC.py:183-184
 *                      (MPC.solution()), for B problems at once.  Argument
 *                      meaning follows the constructor: state0 -> x0 (+ u_prev =
 *                      state0.throttle/steer, NaN = None), s0, centerline_{x,y}_
 *                      poly_coeffs -> cx/cy (5 coefficients, highest order first,
 *                      absolute progress), max_error, runtime_params -> params,
 *                      last_controls -> u_warm, Ts / N -> mpcr_config.
 *   mpcr_config        mirrors control/ControllerParameters.py:4-32 and the IPOPT
 *                      options of control/MPC.py:152-161.
 *   mpcr_track_*       replace splines/ParameterizedLine.py:19-105 (Gx..ddGy,
 *                      x/y_as_coeffs, projection) and
 *                      splines/ParameterizedCenterline.py:61-91 (lookup_error,
 *                      error_sign) for batches of queries.
 *
 * All arrays are fp64, row-major [B, ...], caller-owned.  No exceptions cross the
 * boundary: every function returns 0 on success or a negative error code and
 * mpcr_last_error() gives the text.  Per-problem solver outcomes are reported
 * only through status[] (>= 0 is what CasADi's Opti treats as success).
 * There is no CPU fallback: without a CUDA device mpcr_create fails.
 */
#ifndef MPCR_H_
#define MPCR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCR_VERSION 105

/* models (control/MPC.py:186 f_vehicle = dynamic, :231 f_vehicle_kinematic,
 * models/BlendedBicycleModel.py:23-46 blend of the two) */
#define MPCR_MODEL_DYNAMIC 0
#define MPCR_MODEL_KINEMATIC 1
#define MPCR_MODEL_BLENDED 2

/* solve-kernel layout: one problem per warp (lane k = stage k).  AUTO = WARP; the field is kept for
 * further layouts (round 2 measured "two problems per warp" for horizons up to 16: the two half-
 * warps run as divergent paths of one warp and gain nothing). */
#define MPCR_KERNEL_AUTO 0
#define MPCR_KERNEL_WARP 1

/* integrators of the NLP's dynamics rows (also used by mpcr_model_eval) */
#define MPCR_INTEGRATOR_EULER 0
#define MPCR_INTEGRATOR_RK4 1

/* dispatch order of the problems of a batch (mpcr_config.schedule) */
#define MPCR_SCHEDULE_AUTO 0       /* curvature-first when the batch exceeds one wave, else index order */
#define MPCR_SCHEDULE_INDEX 1      /* problems in index order */
#define MPCR_SCHEDULE_CURVATURE 2  /* always expected-longest-first */

/* per-problem status codes */
#define MPCR_SOLVE_SUCCEEDED 0
#define MPCR_SOLVED_TO_ACCEPTABLE_LEVEL 1
#define MPCR_MAXITER_EXCEEDED (-1)
#define MPCR_RESTORATION_FAILED (-2)
#define MPCR_ERROR_IN_STEP_COMPUTATION (-3)
#define MPCR_INVALID_NUMBER (-4)

/* function return codes */
#define MPCR_OK 0
#define MPCR_ERR_INVALID_ARG (-1)
#define MPCR_ERR_CUDA (-2)
#define MPCR_ERR_NO_DEVICE (-3)

typedef struct mpcr_config {
  /* problem shape */
  int32_t N;                 /* horizon, 2..32 (FixedControllerParameters.N = 30): one lane per stage */
  int32_t model;             /* MPCR_MODEL_* */
  int32_t wrap_constraint;   /* keep the i=0 wrap rows of control/MPC.py:142-143 (1 = parity) */
  int32_t lane_constraint;   /* enable the lane row commented out at control/MPC.py:135 */
  int32_t device;            /* CUDA device ordinal */
  int32_t pointers_on_device;/* 1: all array arguments are device pointers; 0: host pointers */
  int32_t max_iter;          /* FixedControllerParameters.max_iter = 500 */
  int32_t acceptable_iter;   /* IPOPT default 15 */
  double Ts;                 /* FixedControllerParameters.Ts = 0.05 */
  /* control/ControllerParameters.py:4-23 */
  double lambda_s, alpha_L;
  double min_steer, max_steer, min_throttle;
  double max_steer_delta, min_steer_delta, max_throttle_delta, min_throttle_delta;
  double q_v_max, v_max;
  double min_s_delta;        /* control/MPC.py:59 */
  /* control/ControllerParameters.py:26-32 */
  double alpha_c, d_max, q_v_y, n_exp, beta_delta;
  /* IPOPT options of control/MPC.py:152-161 + the termination defaults */
  double tol, acceptable_tol;
  double dual_inf_tol, constr_viol_tol, compl_inf_tol;
  double mu_init;
  int32_t solve_kernel;      /* MPCR_KERNEL_* (0 = auto) */
  int32_t integrator;        /* MPCR_INTEGRATOR_EULER (reference, default) or MPCR_INTEGRATOR_RK4 (extension:
                              * the dynamics rows become x_i = RK4(x_{i-1}, u_{i-1}), same kernel) */
  int32_t restoration;       /* when the filter line search fails, run the feasibility-restoration phase (IPOPT
                              * would; quadratic-penalty variant) instead of returning
                              * MPCR_RESTORATION_FAILED.  0 = off (default).  1 = two-stage: every problem runs
                              * the plain solver and the few that need the phase are solved again, from the
                              * start, by the build that carries it (no cost for the others).  2 = every
                              * problem runs the build with the phase (~6 % slower per iteration; cheaper than
                              * 1 once more than a few percent of a batch need the phase, e.g. lane rows).
                              * 1 and 2 return the same solutions up to round-off. */
  int32_t schedule;          /* MPCR_SCHEDULE_*: order in which the persistent warps pick problems off the batch.
                              * Results do not depend on it; kernel time does (stragglers picked last leave
                              * the GPU idle).  0 = auto: expected-longest-first (reference-path curvature)
                              * for batches larger than one wave of resident warps */
  double polish_tol;         /* > 0 (and < tol): after a problem terminates successfully the iteration continues
                              * until the optimality error is below polish_tol (or the line search gives up);
                              * status[] / iters[] still report the first termination, U/X/S/obj/lam_g the
                              * polished point.  Test / analysis aid: two solvers that stop at tol = 1e-4 agree
                              * only to ~1e-3 in flat directions, polished points agree to ~1e-8.  0 = off */
} mpcr_config;

/* A handle owns one dispatch counter, one set of schedule buffers and one pair of timing events:
 * calls on the same handle are serialised on the device (a launch on another stream than the
 * previous one waits for that one), and a handle must not be used from two host threads at once.
 * Use one handle per stream / thread for concurrent solves. */
typedef struct mpcr_handle mpcr_handle;

int mpcr_version(void);
void mpcr_config_default(mpcr_config* cfg);
int mpcr_create(const mpcr_config* cfg, mpcr_handle** out);
int mpcr_destroy(mpcr_handle* h);
const char* mpcr_last_error(mpcr_handle* h);

/* length of one lam_g row in the output (13N+9, +N with the lane row); rows of
 * problems without u_prev are zero-padded at the end */
int mpcr_lam_g_stride(mpcr_handle* h);

int mpcr_solve_batch(mpcr_handle* h, int B,
                     const double* x0      /* [B,6] */,
                     const double* u_prev  /* [B,2], NaN = None */,
                     const double* s0      /* [B] */,
                     const double* cx      /* [B,5] */,
                     const double* cy      /* [B,5] */,
                     const double* max_err /* [B] or NULL */,
                     const double* params  /* [B,5] alpha_c,d_max,q_v_y,n,beta_delta or NULL; a row that is not finite,
                                              has n < 1 (n is rounded to the nearest integer) or negative weights makes
                                              that problem end with MPCR_INVALID_NUMBER, the others are unaffected */,
                     const double* u_warm  /* [B,2,N] or NULL */,
                     double* U             /* [B,2,N] */,
                     double* X             /* [B,6,N+1] */,
                     double* S             /* [B,N+1] */,
                     double* eC            /* [B,N] */,
                     double* eL            /* [B,N] */,
                     double* lam_g         /* [B,stride] or NULL */,
                     double* obj           /* [B] */,
                     int32_t* status       /* [B] */,
                     int32_t* iters        /* [B] */,
                     void* cuda_stream     /* cudaStream_t or NULL */);

/* Same call, but with HOST pointers it returns as soon as the copies and the kernels are enqueued on
 * `cuda_stream`: the inputs must stay valid and the outputs are undefined until the caller has
 * synchronised that stream (use page-locked host memory, or the copies are not asynchronous).  With
 * device pointers it is identical to mpcr_solve_batch (always asynchronous).  Two handles on two
 * streams used alternately form a pipeline: the copies of one batch and the tail of its launch (the
 * last long problems of the persistent kernel) overlap the head of the next batch
 * (mpc_racing_b200/pipeline.py).  The reference has no counterpart (control/MPC.py:152-181 solves
 * one problem per call, blocking). */
int mpcr_solve_batch_async(mpcr_handle* h, int B, const double* x0, const double* u_prev,
                           const double* s0, const double* cx, const double* cy,
                           const double* max_err, const double* params, const double* u_warm,
                           double* U, double* X, double* S, double* eC, double* eL,
                           double* lam_g, double* obj, int32_t* status, int32_t* iters,
                           void* cuda_stream);

/* Debug variant: also writes the per-iteration trace [B, max_iter+1, 8]
 * (obj, E0, pr_inf, dual_inf, compl, mu, alpha, delta_w). */
int mpcr_solve_batch_trace(mpcr_handle* h, int B, const double* x0, const double* u_prev,
                           const double* s0, const double* cx, const double* cy,
                           const double* max_err, const double* params, const double* u_warm,
                           double* U, double* X, double* S, double* eC, double* eL,
                           double* lam_g, double* obj, int32_t* status, int32_t* iters,
                           double* trace, void* cuda_stream);

/* Expected-cost keys for the dispatch order of the NEXT solves of this handle (any schedule that
 * reorders): keys [B] int32 on the device, e.g. the iteration counts the same problems took in the
 * previous control step of a closed loop; larger = dispatched earlier.  NULL = back to the built-in
 * key (curvature of the reference path).  The array is read when a solve is launched; it must stay
 * valid until then.  Only the order of the work changes, never a result. */
int mpcr_set_dispatch_keys(mpcr_handle* h, const int32_t* keys);

/* Device time of the last solve in milliseconds: the solve kernel and, when the batch was
 * reordered (mpcr_config.schedule), the two schedule kernels before it (CUDA events on the
 * launch stream); blocks until that launch has finished. */
int mpcr_last_kernel_ms(mpcr_handle* h, float* ms);
/* Launch geometry of the last solve launch (or of the warp kernel before any launch):
 * blocks, warps per block, dynamic shared bytes. */
int mpcr_launch_info(mpcr_handle* h, int* blocks, int* warps_per_block, int* smem_bytes);
/* Number of this library's kernels the last solve call launched (1 solve kernel, + 2 schedule
 * kernels when the batch was reordered, mpcr_config.schedule). */
int mpcr_last_launch_count(mpcr_handle* h, int* kernels);

/* ---- track geometry (replaces splines/ParameterizedLine.py:19-105 and
 * splines/ParameterizedCenterline.py:61-91 for batches of queries) -----------------------
 * A track is a piecewise-cubic table: n pieces with break points brk[n+1] (brk[0] = 0,
 * brk[n] = length) and local power-basis coefficients cx/cy[n,4] (lowest order first) of the
 * reference's not-a-knot interpolating spline, plus the 0.5 m max-error table
 * (lanes/<track>_max_error.csv: columns right, left; may be NULL / n_err = 0).
 * Table pointers are HOST pointers (copied once); query arrays follow
 * mpcr_config.pointers_on_device. */
typedef struct mpcr_track mpcr_track;
int mpcr_track_create(mpcr_handle* h, int n, const double* brk, const double* cx, const double* cy,
                      int n_err, const double* err_right, const double* err_left,
                      mpcr_track** out);
int mpcr_track_destroy(mpcr_track* t);
/* Gx, Gy, dGx, dGy, ddGx, ddGy at s[B] (any real, wrapped like `s % length`) -> out[B,6] */
int mpcr_track_eval(mpcr_track* t, int B, const double* s, double* out, void* cuda_stream);
/* x_as_coeffs / y_as_coeffs(s, lookahead, deg=4) -> cx[B,5], cy[B,5] (highest order first) */
int mpcr_track_coeffs(mpcr_track* t, int B, const double* s, double lookahead, double* cx,
                      double* cy, void* cuda_stream);
/* projection(X, Y, bounds=(lo, hi)); NaN bounds or a window wider than 5 m = global search.
 * Outputs s*[B], dist[B] and error_sign[B] (+1 right of the line, as error_sign(X, Y, s*)). */
int mpcr_track_project(mpcr_track* t, int B, const double* X, const double* Y, const double* lo,
                       const double* hi, double* s_out, double* dist_out, int32_t* sign_out,
                       void* cuda_stream);
/* lookup_error(s, lookahead) -> out[B] */
int mpcr_track_lookup_error(mpcr_track* t, int B, const double* s, double lookahead, double* out,
                            void* cuda_stream);

/* ---- plant step (closed-loop rollouts): x_out[b] = <Model>BicycleModel(x[b]).step(u[b,0],
 * u[b,1], dt) of models/KinematicBicycleModel.py:11-48, DynamicBicycleModel.py:16-78,
 * BlendedBicycleModel.py:18-58 (model = MPCR_MODEL_*).  x, x_out [B,6]; u [B,2]. */
int mpcr_plant_step(mpcr_handle* h, int model, int B, const double* x, const double* u, double dt,
                    double* x_out, void* cuda_stream);

/* ---- closed-loop control step around the solve (replaces, for R vehicles at once, what the
 * reference's agent does per simulator tick, agent.py:225-296): two kernels on device-resident
 * state, so that a receding-horizon step is  prepare -> mpcr_solve_batch -> apply  with no host
 * work in between (capturable in a CUDA graph).  All pointers are DEVICE pointers owned by the
 * caller; the handle must have been created with pointers_on_device = 1.
 *   prepare: [state estimate: yaw rate by finite difference of the yaw with the +-2 pi unwrap,
 *            velocities world -> body, agent.py:240-254]; progress = projection(X, Y,
 *            bounds = progress_bound()) (agent.py:80-92, :271); cx, cy = x/y_as_coeffs(progress -
 *            lookback, lookahead, 4) (agent.py:156-165); lap bookkeeping (travelled distance, lap and
 *            checkpoint times, largest centerline distance).
 *   apply:   u = first input of the plan, or (0.5, 0) before start_control_at (agent.py:278-283);
 *            x = <plant>.step(u, Ts) (models/*.py); failed solves counted; vehicles that finished stay put. */
typedef struct mpcr_rollout_bufs {
  int32_t R;                 /* vehicles */
  int32_t N;                 /* horizon of the plan U [R,2,N] */
  int32_t n_seg;             /* checkpoint segments (0 = none) */
  int32_t plant_model;       /* MPCR_MODEL_* of the simulator */
  int32_t estimate_state;    /* 1: the MPC sees the agent's state estimate instead of the plant state */
  int32_t start_control_at;  /* control steps with the fixed input (0.5, 0) before the MPC takes over */
  double Ts, window, lookback, lookahead;   /* agent.py: 0.05, 2.0, 5.0, 45.0 */
  double target;             /* distance after which a vehicle is finished (laps * length) */
  double* x;                 /* [R,6] plant state (X, Y, yaw, v_x, v_y, r), updated by apply */
  double* x_mpc;             /* [R,6] state0 handed to the solve, written by prepare */
  double* progress;          /* [R] */
  double* yaw_prev;          /* [R] yaw of the previous tick (state estimate) */
  double* u_prev;            /* [R,2] input applied at the previous tick */
  double* cx;                /* [R,5] */
  double* cy;                /* [R,5] */
  double* dist;              /* [R] distance to the centerline */
  int32_t* sign;             /* [R] error_sign */
  double* travelled;         /* [R] */
  double* lap_time;          /* [R] NaN until finished */
  double* max_err;           /* [R] */
  const double* seg_bounds;  /* [n_seg + 1] cumulative distances of the checkpoints, seg_bounds[0] = 0 */
  double* seg_t;             /* [R,n_seg] time at which a checkpoint was passed, NaN before */
  int32_t* steps;            /* [R] control steps taken while not finished */
  int32_t* fails;            /* [R] failed solves while not finished */
  int32_t* done;             /* [R] */
  int32_t* counters;         /* [4]: tick index k, vehicles finished, converged solves of unfinished vehicles, spare */
} mpcr_rollout_bufs;
int mpcr_rollout_prepare(mpcr_track* t, const mpcr_rollout_bufs* b, void* cuda_stream);
/* U [R,2,N] and status [R] of the solve that followed prepare (ignored before start_control_at) */
int mpcr_rollout_apply(mpcr_handle* h, const mpcr_rollout_bufs* b, const double* U, const int32_t* status,
                       void* cuda_stream);

/* ---- vehicle-model step with analytic derivatives (replaces CasADi's AD of
 * control/MPC.py:186-260 f_vehicle / f_vehicle_kinematic and of the blend of
 * models/BlendedBicycleModel.py:23-46): f[B,6] = x+, jac[B,6,8] = d x+/d(x,u) (or NULL),
 * hess[B,8,8] = sum_k lam[b,k] d2 x+_k/d(x,u)^2 (or NULL; needs lam[B,6]).
 * integrator: MPCR_INTEGRATOR_EULER = the explicit-Euler maps the reference's NLP uses,
 * MPCR_INTEGRATOR_RK4 = classical RK4 of the same vector field (extension, north-star item 2). */
int mpcr_model_eval(mpcr_handle* h, int model, int integrator, int B, const double* x, const double* u,
                    double Ts, const double* lam, double* f, double* jac, double* hess,
                    void* cuda_stream);

/* Measured fp64 FMA peak of the device (DFMA-chain microbenchmark, TFLOP/s):
 * the roofline denominator for the solve kernel. */
int mpcr_fp64_peak_tflops(mpcr_handle* h, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* MPCR_H_ */
