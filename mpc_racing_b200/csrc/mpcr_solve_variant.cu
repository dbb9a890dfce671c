// One template variant of the solve kernel (all three vehicle models).  Compiled once per
// (MPCR_V_RESTO, MPCR_V_IG) by build.py.
#include <cuda_runtime.h>

#include "mpcr_solve_kernel.cuh"

#if !defined(MPCR_V_RESTO) || !defined(MPCR_V_IG)
#error "build with -DMPCR_V_RESTO=0|1|2 -DMPCR_V_IG=0|1"
#endif

namespace mpcr {

template <>
SolveKernelFn solve_kernel_table<MPCR_V_RESTO, MPCR_V_IG>(int model) {
  constexpr int R = MPCR_V_RESTO;
  constexpr int I = MPCR_V_IG;
  if (model == kDynamic) return solve_kernel<kDynamic, R, I>;
#if !defined(MPCR_ONLY_DYNAMIC)      // development builds: one model, a third of the compile time
  if (model == kKinematic) return solve_kernel<kKinematic, R, I>;
  if (model == kBlended) return solve_kernel<kBlended, R, I>;
#endif
  return nullptr;
}

}  // namespace mpcr
