// Instantiation table of the CTA-wide tcgen05 kernel (elbo_kernel_umma.cuh), in its own translation unit so that the
// library builds in parallel.  lidx: 0 = (3,3), 1 = (5,3), 2 = (5,5) mixture sizes; fi = noisy + 2 * has_idx + 4 * no_missing.
#include "elbo_kernel_umma.cuh"

namespace cytof {

#define CYTOF_UMMA_ROW(a, b) { elbo_fwdbwd_umma_kernel<a, b, false, false, true>, elbo_fwdbwd_umma_kernel<a, b, true, false, true>, \
                               elbo_fwdbwd_umma_kernel<a, b, false, true, true>, elbo_fwdbwd_umma_kernel<a, b, true, true, true>, \
                               elbo_fwdbwd_umma_kernel<a, b, false, false, false>, elbo_fwdbwd_umma_kernel<a, b, true, false, false>, \
                               elbo_fwdbwd_umma_kernel<a, b, false, true, false>, elbo_fwdbwd_umma_kernel<a, b, true, true, false> }

KernelFn umma_kernel_fn(int lidx, int fi) {
  static const KernelFn table[3][8] = { CYTOF_UMMA_ROW(3, 3), CYTOF_UMMA_ROW(5, 3), CYTOF_UMMA_ROW(5, 5) };
  return table[lidx][fi];
}
int umma_smem_bytes() { return kUSmemBytes; }
int umma_threads() { return kUThreads; }

}  // namespace cytof
