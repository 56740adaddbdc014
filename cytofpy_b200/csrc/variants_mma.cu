// Instantiations of the tensor-core kernel for J, K <= 32 (csrc/elbo_kernel_mma.cuh).
#include "elbo_kernel_mma.cuh"

namespace cytof {

#define CYTOF_MMA_ROW(a, b) \
  { elbo_fwdbwd_mma_kernel<a, b, false, false, true>, elbo_fwdbwd_mma_kernel<a, b, true, false, true>, \
    elbo_fwdbwd_mma_kernel<a, b, false, true, true>, elbo_fwdbwd_mma_kernel<a, b, true, true, true>, \
    elbo_fwdbwd_mma_kernel<a, b, false, false, false>, elbo_fwdbwd_mma_kernel<a, b, true, false, false>, \
    elbo_fwdbwd_mma_kernel<a, b, false, true, false>, elbo_fwdbwd_mma_kernel<a, b, true, true, false> }

KernelFn mma_kernel_fn(int lidx, int fi) {
#ifdef CYTOF_TUNE_ONLY55     // tuning builds (tools/kvariants.py): only the L = 5/5 kernels are compiled; the other rows must not be used
  static const KernelFn table[4][8] = { CYTOF_MMA_ROW(5, 5), CYTOF_MMA_ROW(5, 5), CYTOF_MMA_ROW(5, 5), CYTOF_MMA_ROW(5, 5) };
#else
  static const KernelFn table[4][8] = { CYTOF_MMA_ROW(3, 3), CYTOF_MMA_ROW(5, 3), CYTOF_MMA_ROW(5, 5), CYTOF_MMA_ROW(8, 8) };
#endif
  return table[lidx][fi];
}

}  // namespace cytof
