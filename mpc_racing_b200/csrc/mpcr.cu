// libmpcr.so: CUDA kernels + the C ABI of include/mpcr.h.
//
// Build (see __graft_entry__.build):
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
//
// Solve kernel: persistent, one warp (= one block of 32 threads) per problem,
// problems handed out through an atomic counter so that converged problems free
// their warp immediately.  No tensor cores (9x9 / 3x3 fp64 blocks), no CPU path.
#include <cuda_runtime.h>

#include <stdlib.h>

#include <new>
#include <string>

#include "mpcr_config.h"
#include "mpcr_ipm.cuh"

namespace mpcr {

// STAGE_GLOBAL: the per-stage Riccati storage lives in a per-block global slab (L1/L2
// resident) instead of shared memory; shared memory then only holds the ~4 KB scratch.
template <int MODEL, bool STAGE_GLOBAL>
__global__ void __launch_bounds__(32, 1)
solve_kernel(const __grid_constant__ Config cfg, const BatchIn in, const BatchOut out, int B,
             int trace_rows, int* __restrict__ counter, double* __restrict__ slabs) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31;
  double* stage = STAGE_GLOBAL ? slabs + (size_t)blockIdx.x * stage_doubles(cfg.N) : nullptr;
  while (true) {
    int b = 0;
    if (lane == 0) b = atomicAdd(counter, 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= B) break;
    solve_problem<MODEL>(cfg, in, out, b, trace_rows, smem, stage);
    __syncwarp();
  }
}

typedef void (*SolveKernelFn)(const Config, const BatchIn, const BatchOut, int, int, int*, double*);

static SolveKernelFn pick_kernel(int model, bool stage_global) {
  if (stage_global) {
    if (model == kDynamic) return solve_kernel<kDynamic, true>;
    if (model == kKinematic) return solve_kernel<kKinematic, true>;
    return solve_kernel<kBlended, true>;
  }
  if (model == kDynamic) return solve_kernel<kDynamic, false>;
  if (model == kKinematic) return solve_kernel<kKinematic, false>;
  return solve_kernel<kBlended, false>;
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
  int blocks_per_sm = 0;
  int smem_bytes = 0;
  int* d_counter = nullptr;
  double* d_slabs = nullptr;
  bool stage_global = false;
  SolveKernelFn kernel = nullptr;
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
  {
    const char* e = getenv("MPCR_STAGE_GLOBAL");
    h->stage_global = e ? (atoi(e) != 0) : false;
  }
  h->kernel = pick_kernel(h->cfg.model, h->stage_global);
  h->smem_bytes = (h->stage_global ? kScratch : ws_doubles(h->cfg.N)) * (int)sizeof(double);
  CK(cudaFuncSetAttribute(h->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_bytes));
  if (!h->stage_global)
    CK(cudaFuncSetAttribute(h->kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                            cudaSharedmemCarveoutMaxShared));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&h->blocks_per_sm, h->kernel, 32,
                                                   h->smem_bytes));
  if (h->blocks_per_sm < 1) {
    h->err = "solve kernel does not fit on this device";
    return MPCR_ERR_CUDA;
  }
  {
    const char* e = getenv("MPCR_BLOCKS_PER_SM");
    if (e && atoi(e) > 0 && atoi(e) < h->blocks_per_sm) h->blocks_per_sm = atoi(e);
  }
  if (h->stage_global)
    CK(cudaMalloc(&h->d_slabs, (size_t)h->num_sms * h->blocks_per_sm * stage_doubles(h->cfg.N) *
                                   sizeof(double)));
  CK(cudaMalloc(&h->d_counter, sizeof(int)));
  CK(cudaEventCreate(&h->ev0));
  CK(cudaEventCreate(&h->ev1));
  return MPCR_OK;
}

extern "C" int mpcr_destroy(mpcr_handle* h) {
  if (!h) return MPCR_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  if (h->d_counter) cudaFree(h->d_counter);
  if (h->d_slabs) cudaFree(h->d_slabs);
  if (h->d_buf) cudaFree(h->d_buf);
  if (h->d_ibuf) cudaFree(h->d_ibuf);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  delete h;
  return MPCR_OK;
}

extern "C" int mpcr_lam_g_stride(mpcr_handle* h) {
  if (!h) return MPCR_ERR_INVALID_ARG;
  return lam_g_stride(h->cfg);
}

extern "C" int mpcr_launch_info(mpcr_handle* h, int* blocks, int* warps_per_block, int* smem) {
  if (!h) return MPCR_ERR_INVALID_ARG;
  if (blocks) *blocks = h->num_sms * h->blocks_per_sm;
  if (warps_per_block) *warps_per_block = 1;
  if (smem) *smem = h->smem_bytes;
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
  int grid = h->num_sms * h->blocks_per_sm;
  if (grid > B) grid = B;
  CK(cudaMemsetAsync(h->d_counter, 0, sizeof(int), st));
  CK(cudaEventRecord(h->ev0, st));
  h->kernel<<<grid, 32, h->smem_bytes, st>>>(h->cfg, in, out, B, trace_rows, h->d_counter, h->d_slabs);
  CK(cudaGetLastError());
  CK(cudaEventRecord(h->ev1, st));
  h->timed = true;
  return MPCR_OK;
}

static int solve_impl(mpcr_handle* h, int B, const double* x0, const double* u_prev,
                      const double* s0, const double* cx, const double* cy, const double* max_err,
                      const double* params, const double* u_warm, double* U, double* X, double* S,
                      double* eC, double* eL, double* lam_g, double* obj, int32_t* status,
                      int32_t* iters, double* trace, void* stream) {
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
  CK(cudaStreamSynchronize(st));
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
