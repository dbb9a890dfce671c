// The solve kernel and its template variants.  Every (RESTO, IG) combination is compiled in
// its own translation unit (mpcr_solve_variant.cu with -DMPCR_V_RESTO / _IG, see build.py) so
// that the variants build in parallel; mpcr.cu picks one per handle.
#pragma once
#include "mpcr_ipm.cuh"

namespace mpcr {

typedef void (*SolveKernelFn)(const Config, const BatchIn, const BatchOut, int, int, int*, const int*);

// kernel for (model, restoration mode 0 / 1 / 2, integrator); defined in mpcr.cu over the per-variant tables
SolveKernelFn solve_kernel_variant(int model, int resto, int ig);

#if defined(__CUDACC__)
// Persistent: block = one warp that pulls problems off the batch's work queue until it runs dry.
template <int MODEL, int RMODE, int IG>
__global__ void __launch_bounds__(32, 1)
solve_kernel(const __grid_constant__ Config cfg, const BatchIn in, const BatchOut out, int B,
             int trace_rows, int* __restrict__ counter, const int* __restrict__ order) {
  extern __shared__ __align__(16) double smem[];
  WorkQueue q{counter, order, B};
  solve_batch<MODEL, RMODE, IG>(cfg, in, out, trace_rows, q, smem);
}

template <int RMODE, int IG>
SolveKernelFn solve_kernel_table(int model);      // one explicit specialisation per translation unit
#endif

}  // namespace mpcr
