// libmpcr.so: CUDA kernels + the C ABI of include/mpcr.h.
//
// Build (see mpc_racing_b200/build.py):
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
//
// Solve kernel (mpcr_solve_variant.cu, one translation unit per template variant): persistent,
// one warp (= one block of 32 threads) per problem, problems handed out through an atomic counter
// so that converged problems free their warp immediately.  No tensor cores (9x9 / 3x3 fp64
// blocks), no CPU path.
#include <cuda_runtime.h>

#include <stdlib.h>

#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "mpcr_config.h"
#include "mpcr_solve_kernel.cuh"
#include "mpcr_track.cuh"

namespace mpcr {

template <> SolveKernelFn solve_kernel_table<0, 0>(int);
template <> SolveKernelFn solve_kernel_table<0, 1>(int);
template <> SolveKernelFn solve_kernel_table<1, 0>(int);
template <> SolveKernelFn solve_kernel_table<1, 1>(int);
template <> SolveKernelFn solve_kernel_table<2, 0>(int);
template <> SolveKernelFn solve_kernel_table<2, 1>(int);

SolveKernelFn solve_kernel_variant(int model, int resto, int ig) {
  if (resto == 2) return ig ? solve_kernel_table<2, 1>(model) : solve_kernel_table<2, 0>(model);
  if (resto == 1) return ig ? solve_kernel_table<1, 1>(model) : solve_kernel_table<1, 0>(model);
  return ig ? solve_kernel_table<0, 1>(model) : solve_kernel_table<0, 0>(model);
}

// ---------------------------------------------------------------------------------------
// Dispatch schedule.  The persistent warps pick problems off a queue; a problem that needs 150
// iterations and is picked last keeps one warp busy long after the others ran dry (measured on
// the bench batch: index order 110 ms, longest-first with the iteration counts known 93 ms).
// Iteration counts are not known in advance, but they follow the curvature of the reference path
// the horizon covers (rank correlation 0.58 on the bench batch; tight corners = active rate and
// input bounds = more iterations).  So: key = mean |curvature| of the centerline polynomial over
// [s0, s0 + max(5 m, 1.2 v_x N Ts)], 256 log-spaced bins, tightest first (counting sort: histogram,
// scan, scatter; the order inside a bin is whatever the atomics give).  Only the ORDER of the
// work changes - every problem is solved exactly as before.
// ---------------------------------------------------------------------------------------
constexpr int kSchedBins = 256;

__global__ void __launch_bounds__(256)
schedule_bin_kernel(const BatchIn in, int B, int N, double Ts, int* __restrict__ hist,
                    unsigned char* __restrict__ bin, const int32_t* __restrict__ keys) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (keys) {
    // caller-supplied expected cost (closed loop: the iteration count of the same vehicle's previous
    // solve): largest first, one bin per unit up to 255
    int kq = keys[b];
    if (kq < 0) kq = 0;
    if (kq > kSchedBins - 1) kq = kSchedBins - 1;
    const int qk = (kSchedBins - 1) - kq;
    bin[b] = (unsigned char)qk;
    atomicAdd(&hist[qk], 1);
    return;
  }
  const double* cx = in.cx + (size_t)5 * b;
  const double* cy = in.cy + (size_t)5 * b;
  const double s0 = in.s0[b];
  const double vx = in.x0[(size_t)6 * b + 3];
  const double L = fmax(5.0, 1.2 * vx * N * Ts);
  double acc = 0.0;
  for (int j = 0; j < 16; ++j) {
    const double s = s0 + L * (double)j * (1.0 / 15.0);
    Poly4 px, py;
    poly4_eval(cx, s, 2, px);
    poly4_eval(cy, s, 2, py);
    const double n2 = px.d1 * px.d1 + py.d1 * py.d1;
    acc += fabs(px.d1 * py.d2 - py.d1 * px.d2) / (n2 * sqrt(n2));
  }
  const double kbar = acc * (1.0 / 16.0);
  // bins over [1e-4, 1] 1/m in log2; anything not comparable (NaN inputs) goes last
  double t = (log2(fmax(kbar, 1e-300)) + 13.287712379549449) * (1.0 / 13.287712379549449);
  if (!(t >= 0.0)) t = 0.0;
  if (t > 1.0) t = 1.0;
  int q = (kSchedBins - 1) - (int)(t * (kSchedBins - 1));
  bin[b] = (unsigned char)q;
  atomicAdd(&hist[q], 1);
}

__global__ void __launch_bounds__(256)
schedule_scatter_kernel(int B, const int* __restrict__ hist, int* __restrict__ cursor,
                        const unsigned char* __restrict__ bin, int* __restrict__ order) {
  __shared__ int base[kSchedBins];
  {   // exclusive scan of the 256-entry histogram, redundantly per block (Hillis-Steele)
    const int t = threadIdx.x;
    int v = hist[t];
    base[t] = v;
    __syncthreads();
    for (int d = 1; d < kSchedBins; d <<= 1) {
      int add = (t >= d) ? base[t - d] : 0;
      __syncthreads();
      base[t] += add;
      __syncthreads();
    }
    int incl = base[t];
    __syncthreads();
    base[t] = incl - v;
    __syncthreads();
  }
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int q = bin[b];
  order[base[q] + atomicAdd(&cursor[q], 1)] = b;
}

// DFMA-chain microbenchmark: the measured fp64 roofline denominator
// (MEASURED_PEAKS.json has no fp64 entry).  8 independent chains per thread.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace mpcr

using namespace mpcr;

struct mpcr_handle {
  mpcr_config pub;
  Config cfg;
  std::string err;
  int device = 0;
  int num_sms = 0;
  int* d_counter = nullptr;
  struct Layout {                        // the solve kernel of this handle (model, restoration, integrator)
    SolveKernelFn kernel = nullptr;
    int smem_bytes = 0;
    int blocks_per_sm = 0;
  } warp;
  cudaEvent_t ev_done = nullptr;         // end of the last launch: orders launches from different streams
  bool launched = false;
  int last_blocks = 0, last_smem = 0;    // geometry of the last launch
  int schedule = MPCR_SCHEDULE_AUTO;     // config / MPCR_SCHEDULE env
  const int32_t* dispatch_keys = nullptr; // mpcr_set_dispatch_keys: device array [B] or null (= curvature key)
  int* d_order = nullptr;                // dispatch order [order_cap] | histogram + cursors [2 * 256]
  unsigned char* d_bin = nullptr;
  int order_cap = 0;
  int last_launches = 0;                 // kernels of this library launched by the last solve call
  int carve_env = -1, carve_now = 0;     // L1/shared carve-out: env override, value currently set
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;
  // staging buffers for the host-pointer path
  int cap = 0;
  double* d_buf = nullptr;     // all fp64 inputs + outputs
  int32_t* d_ibuf = nullptr;   // status + iters
  size_t buf_doubles = 0;
};

#define CK(call)                                                      \
  do {                                                                \
    cudaError_t e_ = (call);                                          \
    if (e_ != cudaSuccess) {                                          \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);    \
      return MPCR_ERR_CUDA;                                           \
    }                                                                 \
  } while (0)

extern "C" int mpcr_version(void) { return MPCR_VERSION; }

extern "C" void mpcr_config_default(mpcr_config* cfg) { config_default(cfg); }

extern "C" const char* mpcr_last_error(mpcr_handle* h) { return h ? h->err.c_str() : "null handle"; }

extern "C" int mpcr_create(const mpcr_config* cfg, mpcr_handle** out) {
  if (!cfg || !out) return MPCR_ERR_INVALID_ARG;
  *out = nullptr;
  const char* why = nullptr;
  if (!config_valid(*cfg, &why)) return MPCR_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return MPCR_ERR_NO_DEVICE;
  if (cfg->device < 0 || cfg->device >= ndev) return MPCR_ERR_INVALID_ARG;
  mpcr_handle* h = new (std::nothrow) mpcr_handle();
  if (!h) return MPCR_ERR_INVALID_ARG;
  *out = h;     // returned even on CUDA failure so that the caller can read the error text
  h->pub = *cfg;
  config_convert(*cfg, &h->cfg);
  h->device = cfg->device;
  CK(cudaSetDevice(h->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, h->device));
  h->num_sms = prop.multiProcessorCount;
  h->schedule = cfg->schedule;
  {
    const char* e = getenv("MPCR_SCHEDULE");
    if (e && !strcmp(e, "index")) h->schedule = MPCR_SCHEDULE_INDEX;
    if (e && !strcmp(e, "curvature")) h->schedule = MPCR_SCHEDULE_CURVATURE;
  }
  h->carve_env = -1;
  {
    const char* e = getenv("MPCR_CARVEOUT");      // percent of the unified L1/shared array
    if (e && atoi(e) > 0 && atoi(e) <= 100) h->carve_env = atoi(e);
  }
  const int resto = h->cfg.o.restoration;       // 0 off, 1 two-stage, 2 every problem in the restoration build
  const int ig = h->cfg.integrator;
  {
    mpcr_handle::Layout& L = h->warp;
    L.kernel = solve_kernel_variant(h->cfg.model, resto, ig);
    L.smem_bytes = ws_doubles(h->cfg.N, ig) * (int)sizeof(double);
    CK(cudaFuncSetAttribute(L.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.smem_bytes));
    h->carve_now = h->carve_env > 0 ? h->carve_env : (int)cudaSharedmemCarveoutMaxShared;
    CK(cudaFuncSetAttribute(L.kernel, cudaFuncAttributePreferredSharedMemoryCarveout, h->carve_now));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&L.blocks_per_sm, L.kernel, 32, L.smem_bytes));
    if (L.blocks_per_sm < 1) {
      h->err = "solve kernel does not fit on this device";
      return MPCR_ERR_CUDA;
    }
    const char* e = getenv("MPCR_BLOCKS_PER_SM");
    if (e && atoi(e) > 0 && atoi(e) < L.blocks_per_sm) L.blocks_per_sm = atoi(e);
  }
  CK(cudaMalloc(&h->d_counter, sizeof(int)));
  CK(cudaEventCreate(&h->ev0));
  CK(cudaEventCreate(&h->ev1));
  CK(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
  return MPCR_OK;
}

extern "C" int mpcr_destroy(mpcr_handle* h) {
  if (!h) return MPCR_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  if (h->d_counter) cudaFree(h->d_counter);
  if (h->d_order) cudaFree(h->d_order);
  if (h->d_bin) cudaFree(h->d_bin);
  if (h->d_buf) cudaFree(h->d_buf);
  if (h->d_ibuf) cudaFree(h->d_ibuf);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_done) cudaEventDestroy(h->ev_done);
  delete h;
  return MPCR_OK;
}

extern "C" int mpcr_set_dispatch_keys(mpcr_handle* h, const int32_t* keys) {
  if (!h) return MPCR_ERR_INVALID_ARG;
  if (keys && !h->pub.pointers_on_device) {
    h->err = "mpcr_set_dispatch_keys: device pointer, handle with pointers_on_device = 1";
    return MPCR_ERR_INVALID_ARG;
  }
  h->dispatch_keys = keys;
  return MPCR_OK;
}

extern "C" int mpcr_lam_g_stride(mpcr_handle* h) {
  if (!h) return MPCR_ERR_INVALID_ARG;
  return lam_g_stride(h->cfg);
}

extern "C" int mpcr_launch_info(mpcr_handle* h, int* blocks, int* warps_per_block, int* smem) {
  if (!h) return MPCR_ERR_INVALID_ARG;
  if (blocks) *blocks = h->last_blocks ? h->last_blocks : h->num_sms * h->warp.blocks_per_sm;
  if (warps_per_block) *warps_per_block = 1;
  if (smem) *smem = h->last_blocks ? h->last_smem : h->warp.smem_bytes;
  return MPCR_OK;
}

extern "C" int mpcr_last_launch_count(mpcr_handle* h, int* kernels) {
  if (!h || !kernels) return MPCR_ERR_INVALID_ARG;
  *kernels = h->last_launches;
  return MPCR_OK;
}

extern "C" int mpcr_last_kernel_ms(mpcr_handle* h, float* ms) {
  if (!h || !ms) return MPCR_ERR_INVALID_ARG;
  if (!h->timed) { h->err = "no solve launched yet"; return MPCR_ERR_INVALID_ARG; }
  CK(cudaSetDevice(h->device));
  CK(cudaEventSynchronize(h->ev1));
  CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
  return MPCR_OK;
}

static int launch_solve(mpcr_handle* h, int B, const BatchIn& in, const BatchOut& out,
                        int trace_rows, cudaStream_t st) {
  // The handle's dispatch counter, order buffers and timing events are shared by all of its
  // launches: a launch on another stream than the previous one waits for that one first.
  // Inside a stream capture (CUDA graph of a closed-loop step) the launches are ordered by the graph
  // and the timing events are left out.
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  CK(cudaStreamIsCapturing(st, &cap));
  const bool capturing = cap != cudaStreamCaptureStatusNone;
  if (h->launched && !capturing) CK(cudaStreamWaitEvent(st, h->ev_done, 0));
  const mpcr_handle::Layout& L = h->warp;
  int grid = h->num_sms * L.blocks_per_sm;
  if (grid > B) grid = B;
  h->last_blocks = grid;
  h->last_smem = L.smem_bytes;
  if (h->carve_env < 0) {
    // Small batches leave shared memory unused: give it to L1, where the per-lane local-memory
    // arrays live (measured: B = 1 3.99 -> 3.32 ms, B = 256 17.4 -> 14.4 ms).  Full grids keep the
    // maximum shared carve-out.
    const int resident = (grid + h->num_sms - 1) / h->num_sms;
    int carve = (int)cudaSharedmemCarveoutMaxShared;
    if (resident < L.blocks_per_sm) {
      carve = (int)((100.0 * resident * (L.smem_bytes + 1024)) / (228.0 * 1024.0)) + 1;
      if (carve > 100) carve = 100;
    }
    // set on every launch: the attribute belongs to the kernel function, not to this handle
    // (another handle may have changed it)
    CK(cudaFuncSetAttribute(L.kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    h->carve_now = carve;
  }
  const bool sched = h->schedule == MPCR_SCHEDULE_CURVATURE ||
                     (h->schedule == MPCR_SCHEDULE_AUTO && B > grid);
  if (sched && B > h->order_cap) {
    if (h->d_order) cudaFree(h->d_order);
    if (h->d_bin) cudaFree(h->d_bin);
    h->d_order = nullptr;
    h->d_bin = nullptr;
    h->order_cap = 0;
    CK(cudaMalloc(&h->d_order, ((size_t)B + 2 * kSchedBins) * sizeof(int)));
    CK(cudaMalloc(&h->d_bin, (size_t)B));
    h->order_cap = B;
  }
  CK(cudaMemsetAsync(h->d_counter, 0, sizeof(int), st));
  if (!capturing) CK(cudaEventRecord(h->ev0, st));          // the schedule kernels are inside the timed region
  const int* order = nullptr;
  h->last_launches = 1;
  if (sched) {
    int* hist = h->d_order + h->order_cap;
    CK(cudaMemsetAsync(hist, 0, 2 * kSchedBins * sizeof(int), st));
    const int nb = (B + 255) / 256;
    schedule_bin_kernel<<<nb, 256, 0, st>>>(in, B, h->cfg.N, h->cfg.Ts, hist, h->d_bin, h->dispatch_keys);
    schedule_scatter_kernel<<<nb, 256, 0, st>>>(B, hist, hist + kSchedBins, h->d_bin, h->d_order);
    CK(cudaGetLastError());
    order = h->d_order;
    h->last_launches = 3;
  }
  L.kernel<<<grid, 32, L.smem_bytes, st>>>(h->cfg, in, out, B, trace_rows, h->d_counter, order);
  CK(cudaGetLastError());
  if (!capturing) {
    CK(cudaEventRecord(h->ev1, st));
    CK(cudaEventRecord(h->ev_done, st));
    h->launched = true;
    h->timed = true;
  }
  return MPCR_OK;
}

static int solve_impl(mpcr_handle* h, int B, const double* x0, const double* u_prev,
                      const double* s0, const double* cx, const double* cy, const double* max_err,
                      const double* params, const double* u_warm, double* U, double* X, double* S,
                      double* eC, double* eL, double* lam_g, double* obj, int32_t* status,
                      int32_t* iters, double* trace, void* stream, bool sync = true) {
  if (!h) return MPCR_ERR_INVALID_ARG;
  if (B < 0 || !x0 || !u_prev || !s0 || !cx || !cy || !U || !X || !S || !eC || !eL || !obj ||
      !status || !iters) {
    h->err = "null required argument";
    return MPCR_ERR_INVALID_ARG;
  }
  if (h->cfg.lane_con && !max_err) {
    h->err = "lane_constraint needs max_err";
    return MPCR_ERR_INVALID_ARG;
  }
  if (B == 0) return MPCR_OK;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int N = h->cfg.N;
  const int ls = lam_g_stride(h->cfg);
  const int trace_rows = h->cfg.o.max_iter + 1;
  if (h->pub.pointers_on_device) {
    BatchIn in{x0, u_prev, s0, cx, cy, max_err, params, u_warm};
    BatchOut out{U, X, S, eC, eL, lam_g, obj, status, iters, trace};
    return launch_solve(h, B, in, out, trace_rows, st);
  }
  // ---- host pointers: stage through device buffers, copies on the same stream ----------
  const size_t nin = (size_t)B * (6 + 2 + 1 + 5 + 5 + 1 + 5 + 2 * N);
  const size_t nout = (size_t)B * (2 * N + 6 * (N + 1) + (N + 1) + N + N + ls + 1);
  const size_t ntr = trace ? (size_t)B * trace_rows * kTraceCols : 0;
  const size_t need = nin + nout + ntr;
  if (need > h->buf_doubles || B > h->cap) {
    if (h->d_buf) cudaFree(h->d_buf);
    if (h->d_ibuf) cudaFree(h->d_ibuf);
    h->d_buf = nullptr;
    h->d_ibuf = nullptr;
    h->buf_doubles = 0;
    h->cap = 0;
    CK(cudaMalloc(&h->d_buf, need * sizeof(double)));
    CK(cudaMalloc(&h->d_ibuf, (size_t)B * 2 * sizeof(int32_t)));
    h->buf_doubles = need;
    h->cap = B;
  }
  double* p = h->d_buf;
  auto take = [&](size_t n) { double* r = p; p += n; return r; };
  double* d_x0 = take((size_t)B * 6);
  double* d_up = take((size_t)B * 2);
  double* d_s0 = take(B);
  double* d_cx = take((size_t)B * 5);
  double* d_cy = take((size_t)B * 5);
  double* d_me = take(B);
  double* d_pr = take((size_t)B * 5);
  double* d_uw = take((size_t)B * 2 * N);
  double* d_U = take((size_t)B * 2 * N);
  double* d_X = take((size_t)B * 6 * (N + 1));
  double* d_S = take((size_t)B * (N + 1));
  double* d_eC = take((size_t)B * N);
  double* d_eL = take((size_t)B * N);
  double* d_lam = take((size_t)B * ls);
  double* d_obj = take(B);
  double* d_tr = trace ? take(ntr) : nullptr;
  int32_t* d_status = h->d_ibuf;
  int32_t* d_iters = h->d_ibuf + B;
  auto h2d = [&](double* d, const double* s, size_t n) {
    return cudaMemcpyAsync(d, s, n * sizeof(double), cudaMemcpyHostToDevice, st);
  };
  CK(h2d(d_x0, x0, (size_t)B * 6));
  CK(h2d(d_up, u_prev, (size_t)B * 2));
  CK(h2d(d_s0, s0, B));
  CK(h2d(d_cx, cx, (size_t)B * 5));
  CK(h2d(d_cy, cy, (size_t)B * 5));
  if (max_err) CK(h2d(d_me, max_err, B));
  if (params) CK(h2d(d_pr, params, (size_t)B * 5));
  if (u_warm) CK(h2d(d_uw, u_warm, (size_t)B * 2 * N));
  if (trace) CK(cudaMemsetAsync(d_tr, 0, ntr * sizeof(double), st));   // rows past the last iteration stay 0
  BatchIn in{d_x0, d_up, d_s0, d_cx, d_cy, max_err ? d_me : nullptr, params ? d_pr : nullptr,
             u_warm ? d_uw : nullptr};
  BatchOut out{d_U, d_X, d_S, d_eC, d_eL, lam_g ? d_lam : nullptr, d_obj, d_status, d_iters, d_tr};
  int rc = launch_solve(h, B, in, out, trace_rows, st);
  if (rc != MPCR_OK) return rc;
  auto d2h = [&](double* dst, const double* src, size_t n) {
    return cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, st);
  };
  CK(d2h(U, d_U, (size_t)B * 2 * N));
  CK(d2h(X, d_X, (size_t)B * 6 * (N + 1)));
  CK(d2h(S, d_S, (size_t)B * (N + 1)));
  CK(d2h(eC, d_eC, (size_t)B * N));
  CK(d2h(eL, d_eL, (size_t)B * N));
  if (lam_g) CK(d2h(lam_g, d_lam, (size_t)B * ls));
  CK(d2h(obj, d_obj, B));
  if (trace) CK(d2h(trace, d_tr, ntr));
  CK(cudaMemcpyAsync(status, d_status, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(iters, d_iters, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  {
    // the staging buffers are in use until the copies above are done: the next solve on this handle
    // (any stream) is ordered behind them
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CK(cudaStreamIsCapturing(st, &cap));
    if (cap == cudaStreamCaptureStatusNone) CK(cudaEventRecord(h->ev_done, st));
  }
  if (sync) CK(cudaStreamSynchronize(st));
  return MPCR_OK;
}

extern "C" int mpcr_solve_batch(mpcr_handle* h, int B, const double* x0, const double* u_prev,
                                const double* s0, const double* cx, const double* cy,
                                const double* max_err, const double* params, const double* u_warm,
                                double* U, double* X, double* S, double* eC, double* eL,
                                double* lam_g, double* obj, int32_t* status, int32_t* iters,
                                void* cuda_stream) {
  return solve_impl(h, B, x0, u_prev, s0, cx, cy, max_err, params, u_warm, U, X, S, eC, eL, lam_g,
                    obj, status, iters, nullptr, cuda_stream);
}

extern "C" int mpcr_solve_batch_async(mpcr_handle* h, int B, const double* x0, const double* u_prev,
                                      const double* s0, const double* cx, const double* cy,
                                      const double* max_err, const double* params, const double* u_warm,
                                      double* U, double* X, double* S, double* eC, double* eL,
                                      double* lam_g, double* obj, int32_t* status, int32_t* iters,
                                      void* cuda_stream) {
  return solve_impl(h, B, x0, u_prev, s0, cx, cy, max_err, params, u_warm, U, X, S, eC, eL, lam_g,
                    obj, status, iters, nullptr, cuda_stream, false);
}

extern "C" int mpcr_solve_batch_trace(mpcr_handle* h, int B, const double* x0,
                                      const double* u_prev, const double* s0, const double* cx,
                                      const double* cy, const double* max_err, const double* params,
                                      const double* u_warm, double* U, double* X, double* S,
                                      double* eC, double* eL, double* lam_g, double* obj,
                                      int32_t* status, int32_t* iters, double* trace,
                                      void* cuda_stream) {
  return solve_impl(h, B, x0, u_prev, s0, cx, cy, max_err, params, u_warm, U, X, S, eC, eL, lam_g,
                    obj, status, iters, trace, cuda_stream);
}

extern "C" int mpcr_fp64_peak_tflops(mpcr_handle* h, double* tflops) {
  if (!h || !tflops) return MPCR_ERR_INVALID_ARG;
  CK(cudaSetDevice(h->device));
  const int threads = 256, blocks = h->num_sms * 8, iters = 1 << 16;
  double* d = nullptr;
  CK(cudaMalloc(&d, (size_t)threads * blocks * sizeof(double)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0, 0));
    fp64_peak_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-7);
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  const double flops = 2.0 * 8.0 * (double)iters * threads * blocks;
  *tflops = flops / (best * 1e-3) / 1e12;
  return MPCR_OK;
}

// =====================================================================================
// Track kernels: one thread per query; the coefficient table is staged into shared memory
// with ONE TMA bulk copy (cp.async.bulk, completion on an mbarrier) when it fits, else the
// kernel reads the table from global memory (lane-boundary splines have up to 8k pieces).
// =====================================================================================
namespace mpcr {

struct TrackDev {
  int n, n_err;
  double length;
  const double* table;    // packed: brk[n+1] | cx[4n] | cy[4n] | err_right[n_err] | err_left[n_err]
  size_t bytes;           // padded to a multiple of 16
};

__device__ __forceinline__ TrackView make_view(const TrackDev& d, const double* base) {
  TrackView v;
  v.n = d.n;
  v.length = d.length;
  v.brk = base;
  v.cx = base + (d.n + 1);
  v.cy = v.cx + 4 * d.n;
  v.n_err = d.n_err;
  v.err_right = v.cy + 4 * d.n;
  v.err_left = v.err_right + d.n_err;
  return v;
}

// Stage `bytes` (multiple of 16, 16-byte aligned) from global to shared with a TMA bulk copy.
__device__ __forceinline__ void tma_stage(void* smem_dst, const void* gsrc, unsigned bytes,
                                          unsigned long long* bar) {
  const unsigned bar_addr = (unsigned)__cvta_generic_to_shared(bar);
  const unsigned dst_addr = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(dst_addr), "l"(gsrc), "r"(bytes), "r"(bar_addr)
        : "memory");
  }
  unsigned done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar_addr)
        : "memory");
  }
}

enum TrackOp : int { kOpEval = 0, kOpCoeffs = 1, kOpProject = 2, kOpLookup = 3 };

struct TrackArgs {
  int B;
  const double* a0;   // s | s | X | s
  const double* a1;   // - | - | Y | -
  const double* a2;   // - | - | lo | -
  const double* a3;   // - | - | hi | -
  double scalar;      // - | lookahead | - | lookahead
  double* o0;         // out[B,6] | cx[B,5] | s_out | out
  double* o1;         // -        | cy[B,5] | dist  | -
  int32_t* o2;        // -        | -       | sign  | -
};

template <int OP, bool STAGED>
__global__ void __launch_bounds__(128) track_kernel(const TrackDev dev, const TrackArgs a) {
  extern __shared__ __align__(16) unsigned char tsm[];
  __shared__ __align__(8) unsigned long long bar;
  const double* base = dev.table;
  if (STAGED) {
    tma_stage(tsm, dev.table, (unsigned)dev.bytes, &bar);
    base = reinterpret_cast<const double*>(tsm);
  }
  const TrackView t = make_view(dev, base);
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < a.B; b += gridDim.x * blockDim.x) {
    if (OP == kOpEval) {
      double o[6];
      track_eval(t, a.a0[b], 2, o);
      for (int k = 0; k < 6; ++k) a.o0[(size_t)b * 6 + k] = o[k];
    } else if (OP == kOpCoeffs) {
      double cx[5], cy[5];
      track_fit4(t, a.a0[b], a.scalar, cx, cy);
      for (int k = 0; k < 5; ++k) { a.o0[(size_t)b * 5 + k] = cx[k]; a.o1[(size_t)b * 5 + k] = cy[k]; }
    } else if (OP == kOpProject) {
      double s, d;
      track_project(t, a.a0[b], a.a1[b], a.a2[b], a.a3[b], &s, &d);
      a.o0[b] = s;
      a.o1[b] = d;
      if (a.o2) a.o2[b] = track_error_sign(t, a.a0[b], a.a1[b], s);
    } else {
      a.o0[b] = track_lookup_error(t, a.a0[b], a.scalar);
    }
  }
}

}  // namespace mpcr

struct mpcr_track {
  mpcr_handle* h;
  TrackDev dev;
  double* d_table = nullptr;
  bool staged = false;
  // staging for host-pointer queries
  double* d_q = nullptr;
  size_t q_doubles = 0;
};

static const size_t kTrackSmemLimit = 200 * 1024;

extern "C" int mpcr_track_create(mpcr_handle* h, int n, const double* brk, const double* cx,
                                 const double* cy, int n_err, const double* err_right,
                                 const double* err_left, mpcr_track** out) {
  if (!h || !out || n < 1 || !brk || !cx || !cy || n_err < 0 || (n_err > 0 && (!err_right || !err_left))) {
    if (h) h->err = "mpcr_track_create: invalid argument";
    return MPCR_ERR_INVALID_ARG;
  }
  CK(cudaSetDevice(h->device));
  mpcr_track* t = new (std::nothrow) mpcr_track();
  if (!t) return MPCR_ERR_INVALID_ARG;
  t->h = h;
  size_t nd = (size_t)(n + 1) + 8 * (size_t)n + 2 * (size_t)n_err;
  nd = (nd + 1) & ~(size_t)1;                 // 16-byte multiple
  std::vector<double> host(nd, 0.0);
  double* p = host.data();
  memcpy(p, brk, (n + 1) * sizeof(double)); p += n + 1;
  memcpy(p, cx, 4 * (size_t)n * sizeof(double)); p += 4 * (size_t)n;
  memcpy(p, cy, 4 * (size_t)n * sizeof(double)); p += 4 * (size_t)n;
  if (n_err) {
    memcpy(p, err_right, n_err * sizeof(double)); p += n_err;
    memcpy(p, err_left, n_err * sizeof(double));
  }
  cudaError_t e = cudaMalloc(&t->d_table, nd * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(t->d_table, host.data(), nd * sizeof(double), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    h->err = std::string("mpcr_track_create: ") + cudaGetErrorString(e);
    if (t->d_table) cudaFree(t->d_table);
    delete t;
    return MPCR_ERR_CUDA;
  }
  t->dev.n = n;
  t->dev.n_err = n_err;
  t->dev.length = brk[n];
  t->dev.table = t->d_table;
  t->dev.bytes = nd * sizeof(double);
  t->staged = t->dev.bytes <= kTrackSmemLimit;
  *out = t;
  return MPCR_OK;
}

extern "C" int mpcr_track_destroy(mpcr_track* t) {
  if (!t) return MPCR_ERR_INVALID_ARG;
  cudaSetDevice(t->h->device);
  if (t->d_table) cudaFree(t->d_table);
  if (t->d_q) cudaFree(t->d_q);
  delete t;
  return MPCR_OK;
}

template <int OP>
static int track_launch(mpcr_track* t, const TrackArgs& a, cudaStream_t st) {
  mpcr_handle* h = t->h;
  const int threads = 128;
  int blocks = (a.B + threads - 1) / threads;
  const int cap = h->num_sms * 8;
  if (blocks > cap) blocks = cap;
  if (t->staged) {
    CK(cudaFuncSetAttribute(track_kernel<OP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)t->dev.bytes));
    track_kernel<OP, true><<<blocks, threads, t->dev.bytes, st>>>(t->dev, a);
  } else {
    track_kernel<OP, false><<<blocks, threads, 0, st>>>(t->dev, a);
  }
  CK(cudaGetLastError());
  return MPCR_OK;
}

// generic driver: nin input arrays of B doubles, outputs of given widths
static int track_run(mpcr_track* t, int op, int B, const double* const* in, int nin, double scalar,
                     double* const* outp, const int* out_w, int nout, int32_t* sign_out, void* stream) {
  if (!t) return MPCR_ERR_INVALID_ARG;
  mpcr_handle* h = t->h;
  if (B < 0) { h->err = "negative batch"; return MPCR_ERR_INVALID_ARG; }
  if (B == 0) return MPCR_OK;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const double* din[4] = {nullptr, nullptr, nullptr, nullptr};
  double* dout[2] = {nullptr, nullptr};
  int32_t* dsign = nullptr;
  const bool on_dev = h->pub.pointers_on_device != 0;
  if (on_dev) {
    for (int i = 0; i < nin; ++i) din[i] = in[i];
    for (int i = 0; i < nout; ++i) dout[i] = outp[i];
    dsign = sign_out;
  } else {
    size_t need = (size_t)B * nin;
    for (int i = 0; i < nout; ++i) need += (size_t)B * out_w[i];
    need += (size_t)B;     // sign (int32 in a double slot)
    if (need > t->q_doubles) {
      if (t->d_q) cudaFree(t->d_q);
      t->d_q = nullptr;
      t->q_doubles = 0;
      CK(cudaMalloc(&t->d_q, need * sizeof(double)));
      t->q_doubles = need;
    }
    double* p = t->d_q;
    for (int i = 0; i < nin; ++i) {
      CK(cudaMemcpyAsync(p, in[i], (size_t)B * sizeof(double), cudaMemcpyHostToDevice, st));
      din[i] = p;
      p += B;
    }
    for (int i = 0; i < nout; ++i) { dout[i] = p; p += (size_t)B * out_w[i]; }
    dsign = sign_out ? reinterpret_cast<int32_t*>(p) : nullptr;
  }
  TrackArgs a{B, din[0], din[1], din[2], din[3], scalar, dout[0], dout[1], dsign};
  int rc;
  if (op == kOpEval) rc = track_launch<kOpEval>(t, a, st);
  else if (op == kOpCoeffs) rc = track_launch<kOpCoeffs>(t, a, st);
  else if (op == kOpProject) rc = track_launch<kOpProject>(t, a, st);
  else rc = track_launch<kOpLookup>(t, a, st);
  if (rc != MPCR_OK) return rc;
  if (!on_dev) {
    for (int i = 0; i < nout; ++i)
      CK(cudaMemcpyAsync(outp[i], dout[i], (size_t)B * out_w[i] * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (sign_out)
      CK(cudaMemcpyAsync(sign_out, dsign, (size_t)B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  return MPCR_OK;
}

extern "C" int mpcr_track_eval(mpcr_track* t, int B, const double* s, double* out, void* stream) {
  const double* in[1] = {s};
  double* o[1] = {out};
  const int w[1] = {6};
  return track_run(t, kOpEval, B, in, 1, 0.0, o, w, 1, nullptr, stream);
}

extern "C" int mpcr_track_coeffs(mpcr_track* t, int B, const double* s, double lookahead, double* cx,
                                 double* cy, void* stream) {
  const double* in[1] = {s};
  double* o[2] = {cx, cy};
  const int w[2] = {5, 5};
  return track_run(t, kOpCoeffs, B, in, 1, lookahead, o, w, 2, nullptr, stream);
}

extern "C" int mpcr_track_project(mpcr_track* t, int B, const double* X, const double* Y,
                                  const double* lo, const double* hi, double* s_out, double* dist_out,
                                  int32_t* sign_out, void* stream) {
  const double* in[4] = {X, Y, lo, hi};
  double* o[2] = {s_out, dist_out};
  const int w[2] = {1, 1};
  return track_run(t, kOpProject, B, in, 4, 0.0, o, w, 2, sign_out, stream);
}

extern "C" int mpcr_track_lookup_error(mpcr_track* t, int B, const double* s, double lookahead,
                                       double* out, void* stream) {
  if (t && t->dev.n_err == 0) { t->h->err = "track has no max-error table"; return MPCR_ERR_INVALID_ARG; }
  const double* in[1] = {s};
  double* o[1] = {out};
  const int w[1] = {1};
  return track_run(t, kOpLookup, B, in, 1, lookahead, o, w, 1, nullptr, stream);
}

// =====================================================================================
// Plant step (closed-loop rollouts): x_out[b] = model.step(u[b], dt) of models/*.py
// =====================================================================================
#include "mpcr_plant.cuh"

namespace mpcr {
__global__ void __launch_bounds__(128) plant_kernel(int model, int B, const double* __restrict__ x,
                                                    const double* __restrict__ u, double dt,
                                                    double* __restrict__ out) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    double xi[6], ui[2], o[6];
    for (int k = 0; k < 6; ++k) xi[k] = x[(size_t)b * 6 + k];
    ui[0] = u[(size_t)b * 2];
    ui[1] = u[(size_t)b * 2 + 1];
    plant_step(model, xi, ui, dt, o);
    for (int k = 0; k < 6; ++k) out[(size_t)b * 6 + k] = o[k];
  }
}
}  // namespace mpcr

extern "C" int mpcr_plant_step(mpcr_handle* h, int model, int B, const double* x, const double* u,
                               double dt, double* x_out, void* stream) {
  if (!h || !x || !u || !x_out || B < 0 || model < 0 || model > 2) {
    if (h) h->err = "mpcr_plant_step: invalid argument";
    return MPCR_ERR_INVALID_ARG;
  }
  if (B == 0) return MPCR_OK;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = 128;
  int blocks = (B + threads - 1) / threads;
  if (blocks > h->num_sms * 16) blocks = h->num_sms * 16;
  if (h->pub.pointers_on_device) {
    plant_kernel<<<blocks, threads, 0, st>>>(model, B, x, u, dt, x_out);
    CK(cudaGetLastError());
    return MPCR_OK;
  }
  double* d = nullptr;
  CK(cudaMalloc(&d, (size_t)B * 14 * sizeof(double)));
  cudaError_t e = cudaMemcpyAsync(d, x, (size_t)B * 6 * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d + (size_t)B * 6, u, (size_t)B * 2 * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    plant_kernel<<<blocks, threads, 0, st>>>(model, B, d, d + (size_t)B * 6, dt, d + (size_t)B * 8);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(x_out, d + (size_t)B * 8, (size_t)B * 6 * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) { h->err = std::string("mpcr_plant_step: ") + cudaGetErrorString(e); return MPCR_ERR_CUDA; }
  return MPCR_OK;
}

// =====================================================================================
// Closed-loop control step around the solve: prepare (estimate, project, refit, bookkeeping) and
// apply (first input of the plan -> plant step) on device-resident state, one thread per vehicle.
// =====================================================================================
namespace mpcr {

template <bool STAGED>
__global__ void __launch_bounds__(128) rollout_prepare_kernel(const TrackDev dev, const mpcr_rollout_bufs b) {
  extern __shared__ __align__(16) unsigned char tsm[];
  __shared__ __align__(8) unsigned long long bar;
  const double* base = dev.table;
  if (STAGED) {
    tma_stage(tsm, dev.table, (unsigned)dev.bytes, &bar);
    base = reinterpret_cast<const double*>(tsm);
  }
  const TrackView t = make_view(dev, base);
  const int k = b.counters[0];
  const double L = t.length;
  int newly_done = 0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < b.R; r += gridDim.x * blockDim.x) {
    double x[6];
    for (int a = 0; a < 6; ++a) x[a] = b.x[(size_t)r * 6 + a];
    // ---- what the MPC sees (agent.py:240-254) ------------------------------------------------
    double xe[6] = {x[0], x[1], x[2], x[3], x[4], x[5]};
    if (b.estimate_state) {
      const double yaw = x[2];
      // world velocity of the plant, rotated back into the vehicle frame by the agent
      const double wx = x[3] * cos(yaw) - x[4] * sin(yaw), wy = x[3] * sin(yaw) + x[4] * cos(yaw);
      xe[3] = wx * cos(-yaw) - wy * sin(-yaw);
      xe[4] = wx * sin(-yaw) + wy * cos(-yaw);
      if (k > 0) {
        const double old = b.yaw_prev[r];
        double d;
        if (yaw - old > 3.0) d = yaw - (old + 3.141592653589793 * 2);
        else if (yaw - old < -3.0) d = yaw - (old - 3.141592653589793 * 2);
        else d = yaw - old;
        xe[5] = d / b.Ts;
      }
      b.yaw_prev[r] = yaw;
    }
    for (int a = 0; a < 6; ++a) b.x_mpc[(size_t)r * 6 + a] = xe[a];
    // ---- progress = projection(X, Y, bounds = progress_bound()) (agent.py:80-92, :271) -------
    const double p_old = b.progress[r];
    const double lo = py_mod(p_old - b.window, L), hi = py_mod(p_old + b.window, L);
    double s_new, d;
    track_project(t, x[0], x[1], fmin(lo, hi), fmax(lo, hi), &s_new, &d);   // swapped at the wrap -> global search
    b.progress[r] = s_new;
    b.dist[r] = d;
    b.sign[r] = track_error_sign(t, x[0], x[1], s_new);
    // ---- reference path of the horizon (agent.py:156-165) ------------------------------------
    double cx[5], cy[5];
    track_fit4(t, s_new - b.lookback, b.lookahead, cx, cy);
    for (int a = 0; a < 5; ++a) { b.cx[(size_t)r * 5 + a] = cx[a]; b.cy[(size_t)r * 5 + a] = cy[a]; }
    // ---- lap bookkeeping -----------------------------------------------------------------------
    if (!b.done[r]) {
      double tr = b.travelled[r];
      if (k > 0) tr += py_mod(s_new - p_old + 0.5 * L, L) - 0.5 * L;
      b.travelled[r] = tr;
      b.max_err[r] = fmax(b.max_err[r], d);
      for (int j = 0; j < b.n_seg; ++j) {
        double* st = b.seg_t + (size_t)r * b.n_seg + j;
        if (*st != *st && tr >= b.seg_bounds[j + 1] - 1e-12) *st = k * b.Ts;
      }
      if (tr >= b.target) {
        b.lap_time[r] = k * b.Ts;
        b.done[r] = 1;
        newly_done++;
      }
    }
  }
  if (newly_done) atomicAdd(&b.counters[1], newly_done);
}

__global__ void __launch_bounds__(128) rollout_apply_kernel(const mpcr_rollout_bufs b, const double* __restrict__ U,
                                                            const int32_t* __restrict__ status) {
  const int k = b.counters[0];
  const bool mpc = k >= b.start_control_at;
  int conv = 0;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < b.R; r += gridDim.x * blockDim.x) {
    double u[2] = {0.5, 0.0};                                  // agent.py:283
    if (mpc) { u[0] = U[(size_t)r * 2 * b.N]; u[1] = U[(size_t)r * 2 * b.N + b.N]; }
    b.u_prev[(size_t)r * 2] = u[0];
    b.u_prev[(size_t)r * 2 + 1] = u[1];
    if (b.done[r]) continue;
    double x[6], o[6];
    for (int a = 0; a < 6; ++a) x[a] = b.x[(size_t)r * 6 + a];
    plant_step(b.plant_model, x, u, b.Ts, o);
    for (int a = 0; a < 6; ++a) b.x[(size_t)r * 6 + a] = o[a];
    b.steps[r] += 1;
    if (mpc) {
      if (status[r] < 0) b.fails[r] += 1;
      else conv++;
    }
  }
  if (conv) atomicAdd(&b.counters[2], conv);
  __syncthreads();
  // the tick counter moves when the last block is through (every block read k at its start; the
  // next kernel on the stream starts after this one has finished)
  if (threadIdx.x == 0) {
    const int arrived = atomicAdd(&b.counters[3], 1) + 1;
    if (arrived == (int)gridDim.x) {
      b.counters[3] = 0;
      b.counters[0] = k + 1;
    }
  }
}
}  // namespace mpcr

static bool rollout_bufs_ok(const mpcr_rollout_bufs* b) {
  return b && b->R >= 0 && b->N >= 1 && b->n_seg >= 0 && b->plant_model >= 0 && b->plant_model <= 2 &&
         b->Ts > 0.0 && b->x && b->x_mpc && b->progress && b->yaw_prev && b->u_prev && b->cx && b->cy &&
         b->dist && b->sign && b->travelled && b->lap_time && b->max_err && b->steps && b->fails && b->done &&
         b->counters && (b->n_seg == 0 || (b->seg_bounds && b->seg_t));
}

extern "C" int mpcr_rollout_prepare(mpcr_track* t, const mpcr_rollout_bufs* b, void* stream) {
  if (!t) return MPCR_ERR_INVALID_ARG;
  mpcr_handle* h = t->h;
  if (!rollout_bufs_ok(b) || !h->pub.pointers_on_device) {
    h->err = "mpcr_rollout_prepare: invalid buffers (device pointers, handle with pointers_on_device = 1)";
    return MPCR_ERR_INVALID_ARG;
  }
  if (b->R == 0) return MPCR_OK;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = 128;
  int blocks = (b->R + threads - 1) / threads;
  if (blocks > h->num_sms * 8) blocks = h->num_sms * 8;
  if (t->staged) {
    CK(cudaFuncSetAttribute(rollout_prepare_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                            (int)t->dev.bytes));
    rollout_prepare_kernel<true><<<blocks, threads, t->dev.bytes, st>>>(t->dev, *b);
  } else {
    rollout_prepare_kernel<false><<<blocks, threads, 0, st>>>(t->dev, *b);
  }
  CK(cudaGetLastError());
  return MPCR_OK;
}

extern "C" int mpcr_rollout_apply(mpcr_handle* h, const mpcr_rollout_bufs* b, const double* U,
                                  const int32_t* status, void* stream) {
  if (!h) return MPCR_ERR_INVALID_ARG;
  if (!rollout_bufs_ok(b) || !U || !status || !h->pub.pointers_on_device) {
    h->err = "mpcr_rollout_apply: invalid buffers (device pointers, handle with pointers_on_device = 1)";
    return MPCR_ERR_INVALID_ARG;
  }
  if (b->R == 0) return MPCR_OK;
  CK(cudaSetDevice(h->device));
  const int threads = 128;
  int blocks = (b->R + threads - 1) / threads;
  if (blocks > h->num_sms * 16) blocks = h->num_sms * 16;
  rollout_apply_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(*b, U, status);
  CK(cudaGetLastError());
  return MPCR_OK;
}

// =====================================================================================
// Model step with analytic Jacobian / Hessian, Euler or RK4 (mpcr_rk4.cuh): one thread per query
// =====================================================================================
#include "mpcr_rk4.cuh"

namespace mpcr {
__global__ void __launch_bounds__(64) model_eval_kernel(int model, int integrator, int B,
                                                        const double* __restrict__ x,
                                                        const double* __restrict__ u, double Ts,
                                                        const double* __restrict__ lam,
                                                        double* __restrict__ f, double* __restrict__ jac,
                                                        double* __restrict__ hess) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    double xi[6], ui[2], li[6], fo[6], J[48], H[64];
    for (int k = 0; k < 6; ++k) { xi[k] = x[(size_t)b * 6 + k]; li[k] = lam ? lam[(size_t)b * 6 + k] : 0.0; }
    ui[0] = u[(size_t)b * 2];
    ui[1] = u[(size_t)b * 2 + 1];
    model_eval_dense(model, integrator, xi, ui, Ts, lam ? li : nullptr, fo, jac ? J : nullptr,
                     (hess && lam) ? H : nullptr);
    for (int k = 0; k < 6; ++k) f[(size_t)b * 6 + k] = fo[k];
    if (jac)
      for (int k = 0; k < 48; ++k) jac[(size_t)b * 48 + k] = J[k];
    if (hess)
      for (int k = 0; k < 64; ++k) hess[(size_t)b * 64 + k] = lam ? H[k] : 0.0;
  }
}
}  // namespace mpcr

extern "C" int mpcr_model_eval(mpcr_handle* h, int model, int integrator, int B, const double* x,
                               const double* u, double Ts, const double* lam, double* f, double* jac,
                               double* hess, void* stream) {
  if (!h || !x || !u || !f || B < 0 || model < 0 || model > 2 || integrator < 0 || integrator > 1 ||
      !(Ts > 0.0) || (hess && !lam)) {
    if (h) h->err = "mpcr_model_eval: invalid argument";
    return MPCR_ERR_INVALID_ARG;
  }
  if (B == 0) return MPCR_OK;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = 64;
  int blocks = (B + threads - 1) / threads;
  if (blocks > h->num_sms * 16) blocks = h->num_sms * 16;
  if (h->pub.pointers_on_device) {
    model_eval_kernel<<<blocks, threads, 0, st>>>(model, integrator, B, x, u, Ts, lam, f, jac, hess);
    CK(cudaGetLastError());
    return MPCR_OK;
  }
  double* d = nullptr;
  const size_t nin = (size_t)B * 14, nout = (size_t)B * (6 + 48 + 64);
  CK(cudaMalloc(&d, (nin + nout) * sizeof(double)));
  double *dx = d, *du = d + (size_t)B * 6, *dl = du + (size_t)B * 2, *df = d + nin, *dj = df + (size_t)B * 6,
         *dh = dj + (size_t)B * 48;
  cudaError_t e = cudaMemcpyAsync(dx, x, (size_t)B * 6 * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(du, u, (size_t)B * 2 * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && lam) e = cudaMemcpyAsync(dl, lam, (size_t)B * 6 * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    model_eval_kernel<<<blocks, threads, 0, st>>>(model, integrator, B, dx, du, Ts, lam ? dl : nullptr, df,
                                                  jac ? dj : nullptr, hess ? dh : nullptr);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(f, df, (size_t)B * 6 * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && jac) e = cudaMemcpyAsync(jac, dj, (size_t)B * 48 * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && hess) e = cudaMemcpyAsync(hess, dh, (size_t)B * 64 * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) { h->err = std::string("mpcr_model_eval: ") + cudaGetErrorString(e); return MPCR_ERR_CUDA; }
  return MPCR_OK;
}
