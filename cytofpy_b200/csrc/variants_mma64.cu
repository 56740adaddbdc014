// Instantiations of the tensor-core kernel for J <= 64, K <= 64 beyond the 8-cell-tile kernel's J, K <= 32
// (csrc/elbo_kernel_mma64.cuh).
#include "elbo_kernel_mma64.cuh"

namespace cytof {

#define CYTOF_MMA64_ROW(a, b, miss) \
  { elbo_fwdbwd_mma64_kernel<a, b, false, false, 1, miss>, elbo_fwdbwd_mma64_kernel<a, b, true, false, 1, miss>, \
    elbo_fwdbwd_mma64_kernel<a, b, false, true, 1, miss>, elbo_fwdbwd_mma64_kernel<a, b, true, true, 1, miss>, \
    elbo_fwdbwd_mma64_kernel<a, b, false, false, 2, miss>, elbo_fwdbwd_mma64_kernel<a, b, true, false, 2, miss>, \
    elbo_fwdbwd_mma64_kernel<a, b, false, true, 2, miss>, elbo_fwdbwd_mma64_kernel<a, b, true, true, 2, miss> }

// WIDE layout (48 < J <= 64): every cell takes a second full pass over markers 32..63 (NTP = 4); fi = noisy + 2 * has_idx
#define CYTOF_MMA64_WIDE_ROW(a, b, miss) \
  { elbo_fwdbwd_mma64_kernel<a, b, false, false, 4, miss>, elbo_fwdbwd_mma64_kernel<a, b, true, false, 4, miss>, \
    elbo_fwdbwd_mma64_kernel<a, b, false, true, 4, miss>, elbo_fwdbwd_mma64_kernel<a, b, true, true, 4, miss>, \
    elbo_fwdbwd_mma64_kernel<a, b, false, false, 4, miss>, elbo_fwdbwd_mma64_kernel<a, b, true, false, 4, miss>, \
    elbo_fwdbwd_mma64_kernel<a, b, false, true, 4, miss>, elbo_fwdbwd_mma64_kernel<a, b, true, true, 4, miss> }

KernelFn mma64_kernel_fn(int lidx, int missing, int fi) {
  static const KernelFn table[3][2][8] = { { CYTOF_MMA64_ROW(5, 5, false), CYTOF_MMA64_ROW(5, 5, true) },
                                           { CYTOF_MMA64_ROW(8, 8, false), CYTOF_MMA64_ROW(8, 8, true) },
                                           { CYTOF_MMA64_WIDE_ROW(5, 5, false), CYTOF_MMA64_WIDE_ROW(5, 5, true) } };
  return table[lidx][missing][fi];
}

}  // namespace cytof
