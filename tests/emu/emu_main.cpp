// TEST INFRASTRUCTURE ONLY: compiles the warp-level solver source for the host
// with the SIMT emulator and exposes a batch entry point for ctypes.
#define MPCR_EMU 1
#include "simt_emu.h"

#include <vector>

#include "../../mpc_racing_b200/csrc/mpcr_config.h"
#include "../../mpc_racing_b200/csrc/mpcr_ipm.cuh"
namespace {
struct Job {
  const mpcr::Config* cfg;
  const mpcr::BatchIn* in;
  const mpcr::BatchOut* out;
  mpcr::WorkQueue q;
  int trace_rows;
  double* ws;
};
// every fiber (= lane) runs the kernel body; the warp drains the whole batch through the work queue
void lane_body(void* p) {
  Job* j = (Job*)p;
  mpcr::solve_batch_any(*j->cfg, *j->in, *j->out, j->trace_rows, j->q, j->ws);
}
}  // namespace

extern "C" int emu_lam_g_stride(const mpcr_config* c) {
  mpcr::Config cfg;
  mpcr::config_convert(*c, &cfg);
  return mpcr::lam_g_stride(cfg);
}

extern "C" void emu_config_default(mpcr_config* c) { mpcr::config_default(c); }

extern "C" int emu_solve_batch(const mpcr_config* c, int B, const double* x0, const double* u_prev,
                               const double* s0, const double* cx, const double* cy,
                               const double* max_err, const double* params, const double* u_warm,
                               double* U, double* X, double* S, double* eC, double* eL,
                               double* lam_g, double* obj, int32_t* status, int32_t* iters,
                               double* trace) {
  const char* why = nullptr;
  if (!mpcr::config_valid(*c, &why)) return -1;
  mpcr::Config cfg;
  mpcr::config_convert(*c, &cfg);
  mpcr::BatchIn in{x0, u_prev, s0, cx, cy, max_err, params, u_warm};
  mpcr::BatchOut out{U, X, S, eC, eL, lam_g, obj, status, iters, trace};
  std::vector<double> ws(mpcr::ws_doubles(cfg.N, cfg.integrator));
  int counter = 0;
  Job j{&cfg, &in, &out, mpcr::WorkQueue{&counter, nullptr, B}, cfg.o.max_iter + 1, ws.data()};
  emu::run_warp(lane_body, &j);
  return 0;
}

// stage-level debug entry: model value / Jacobian / Hessian for parity tests
extern "C" void emu_model_lin(int model, const double* x, const double* u, double Ts,
                              const double* lam, double* f, double* lin21, double* hm21) {
  mpcr::StageLin L;
  mpcr::lin_model(model, x, u, Ts, lam, f, L, hm21);
  for (int k = 0; k < 3; ++k) { lin21[k] = L.aX[k]; lin21[3 + k] = L.aY[k]; }
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 3; ++k) lin21[6 + 3 * i + k] = L.d[i][k];
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 2; ++k) lin21[15 + 2 * i + k] = L.bu[i][k];
}

// Euler / RK4 model evaluation with dense Jacobian (6x8) and lam-weighted Hessian (8x8)
#include "../../mpc_racing_b200/csrc/mpcr_rk4.cuh"
extern "C" void emu_model_eval(int model, int integrator, int B, const double* x, const double* u,
                               double Ts, const double* lam, double* f, double* jac, double* hess) {
  for (int b = 0; b < B; ++b)
    mpcr::model_eval_dense(model, integrator, x + 6 * b, u + 2 * b, Ts, lam ? lam + 6 * b : nullptr,
                           f + 6 * b, jac ? jac + 48 * b : nullptr, hess ? hess + 64 * b : nullptr);
}

// ---- track geometry: the same host/device functions the CUDA track kernels call ------------
#include "../../mpc_racing_b200/csrc/mpcr_track.cuh"

static mpcr::TrackView emu_view(int n, const double* brk, const double* cx, const double* cy,
                                int n_err, const double* er, const double* el) {
  mpcr::TrackView t;
  t.n = n; t.length = brk[n]; t.brk = brk; t.cx = cx; t.cy = cy;
  t.n_err = n_err; t.err_right = er; t.err_left = el;
  return t;
}

extern "C" void emu_track_eval(int n, const double* brk, const double* cx, const double* cy, int B,
                               const double* s, double* out) {
  mpcr::TrackView t = emu_view(n, brk, cx, cy, 0, nullptr, nullptr);
  for (int b = 0; b < B; ++b) mpcr::track_eval(t, s[b], 2, out + 6 * b);
}

extern "C" void emu_track_coeffs(int n, const double* brk, const double* cx, const double* cy, int B,
                                 const double* s, double lookahead, double* ocx, double* ocy) {
  mpcr::TrackView t = emu_view(n, brk, cx, cy, 0, nullptr, nullptr);
  for (int b = 0; b < B; ++b) mpcr::track_fit4(t, s[b], lookahead, ocx + 5 * b, ocy + 5 * b);
}

extern "C" void emu_track_project(int n, const double* brk, const double* cx, const double* cy, int B,
                                  const double* X, const double* Y, const double* lo, const double* hi,
                                  double* s_out, double* d_out, int32_t* sign_out) {
  mpcr::TrackView t = emu_view(n, brk, cx, cy, 0, nullptr, nullptr);
  for (int b = 0; b < B; ++b) {
    mpcr::track_project(t, X[b], Y[b], lo[b], hi[b], s_out + b, d_out + b);
    sign_out[b] = mpcr::track_error_sign(t, X[b], Y[b], s_out[b]);
  }
}

extern "C" void emu_track_lookup_error(int n, const double* brk, const double* cx, const double* cy,
                                       int n_err, const double* er, const double* el, int B,
                                       const double* s, double lookahead, double* out) {
  mpcr::TrackView t = emu_view(n, brk, cx, cy, n_err, er, el);
  for (int b = 0; b < B; ++b) out[b] = mpcr::track_lookup_error(t, s[b], lookahead);
}

// ---- plant models ------------------------------------------------------------------------------
#include "../../mpc_racing_b200/csrc/mpcr_plant.cuh"
extern "C" void emu_plant_step(int model, int B, const double* x, const double* u, const double* dt,
                               double* out) {
  for (int b = 0; b < B; ++b) mpcr::plant_step(model, x + 6 * b, u + 2 * b, dt[b], out + 6 * b);
}
