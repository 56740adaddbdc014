// extern "C" entry points of libcytof_b200 (see include/cytof_b200.h) + the small
// step-loop kernels (fused Adam with NaN scrub, without-replacement minibatch sampler).
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace cytof {
int validate_desc(const cytof_desc* d);
size_t workspace_bytes(const cytof_desc* d);
int launch_fwdbwd(const cytof_desc* d, const float* y, const int32_t* idx, const float* eps,
                  const float* globals, const uint64_t* zmask, float* grad_block, double* ll_lqy,
                  double* packed, void* workspace, size_t workspace_bytes_, cudaStream_t stream,
                  cudaError_t* err, const cytof_peer_desc* peer = nullptr);
size_t peer_mailbox_bytes(const cytof_desc* d, int world);
int launch_labels(const cytof_desc* d, const float* y, const float* globals, const uint64_t* zmask,
                  int32_t* lam_out, float* probs_out, cudaStream_t stream, cudaError_t* err);
struct GlobalParams;
int global_sizes(const cytof_global_desc* g, int64_t out[4]);
int launch_global_fwd(const cytof_global_desc* g, const double* theta, const double* eps_in, double* stash,
                      float* globals, uint64_t* zmask, double* scalars, cudaStream_t s, cudaError_t* err);
int launch_global_bwd(const cytof_global_desc* g, double* theta, double* stash, const double* packed,
                      const double* scalars_in, double* adam_m, double* adam_v, double* grad_out,
                      double* out_scalars, int64_t* step_dev, int32_t* flag, int do_update, cudaStream_t s,
                      cudaError_t* err);
int launch_emit(const cytof_desc* d, const float* y, const int32_t* idx, const float* eps,
                const float* globals, float* y_out, float* eps_out, cudaStream_t stream, cudaError_t* err);

// ---- Adam (torch.optim.Adam defaults as used by reference fit.py:83) + NaN scrub fit.py:24-30
__global__ void adam_step_kernel(double* __restrict__ param, const double* __restrict__ grad,
                                 double* __restrict__ m, double* __restrict__ v, long long n,
                                 long long n_scrub, double grad_scale, double lr, double beta1,
                                 double beta2, double eps, long long step_host,
                                 long long* __restrict__ step_dev, int* __restrict__ flag) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // device-resident step counter (CUDA-graph replay): every thread reads the pre-increment
  // value; the bump is done by a trailing 1-thread kernel (adam_bump_kernel).
  const long long step = step_dev ? (*step_dev + 1) : step_host;
  if (t >= n) return;
  double g = grad[t] * grad_scale;
  if (t < n_scrub && g != g) {
    g = 0.0;
    if (flag) *flag = 1;
  }
  const double mt = m[t] + (1.0 - beta1) * (g - m[t]);
  const double vt = beta2 * v[t] + (1.0 - beta2) * g * g;
  m[t] = mt;
  v[t] = vt;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2s = sqrt(1.0 - pow(beta2, (double)step));
  const double denom = sqrt(vt) / bc2s + eps;
  param[t] -= (lr / bc1) * (mt / denom);
}
__global__ void adam_bump_kernel(long long* step_dev) { *step_dev += 1; }

// ---- sampler kernels: the keyed bijection itself is in common.cuh (the per-cell kernels evaluate it too)
__global__ void sample_minibatch_kernel(long long N, long long m, unsigned long long key, int hbits,
                                        int32_t* __restrict__ out, unsigned long long seed,
                                        const long long* __restrict__ step_dev, int sample) {
  pdl_enter();
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  if (step_dev) key = sampler_key(seed, (unsigned long long)(*step_dev + 1), sample);
  const unsigned long long x = feistel_perm(key, hbits, (unsigned long long)N, (unsigned long long)t);
  out[t] = (int32_t)x;
}
// all I samples in ONE launch (blockIdx.y = sample): the CUDA-graph step needs a single sampler node
struct MultiSample {
  long long N[CYTOF_MAX_SAMPLES], m[CYTOF_MAX_SAMPLES], off[CYTOF_MAX_SAMPLES];
  int hbits[CYTOF_MAX_SAMPLES];
};
__global__ void sample_minibatch_multi_kernel(const MultiSample ms, unsigned long long seed,
                                              const long long* __restrict__ step_dev, unsigned long long step_host,
                                              int32_t* __restrict__ out) {
  pdl_enter();
  const int i = blockIdx.y;
  const long long N = ms.N[i], m = ms.m[i];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  int32_t* o = out + ms.off[i];
  if (m >= N) { o[t] = (int32_t)t; return; }
  const unsigned long long key = sampler_key(seed, step_dev ? (unsigned long long)(*step_dev + 1) : step_host, i);
  const int hbits = ms.hbits[i];
  const unsigned long long x = feistel_perm(key, hbits, (unsigned long long)N, (unsigned long long)t);
  o[t] = (int32_t)x;
}
__global__ void iota_kernel(long long N, int32_t* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < N) out[t] = (int32_t)t;
}

static thread_local char g_cuda_err[256] = "";
static void set_cuda_err(cudaError_t e) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
}
}  // namespace cytof

using namespace cytof;

extern "C" {

int cytof_abi_version(void) { return CYTOF_ABI_VERSION; }

const char* cytof_strerror(int code) {
  switch (code) {
    case CYTOF_OK: return "ok";
    case CYTOF_EINVAL: return "invalid descriptor or null pointer";
    case CYTOF_ESHAPE: return "shape outside compiled limits (I<=32, J<=64, K<=64, L<=8)";
    case CYTOF_EWORKSPACE: return "workspace too small (see cytof_workspace_bytes)";
    case CYTOF_ECUDA: return "CUDA error (see cytof_last_cuda_error)";
    case CYTOF_EALIGN: return "pointer not sufficiently aligned";
    default: return "unknown error";
  }
}
const char* cytof_last_cuda_error(void) { return g_cuda_err; }

int cytof_globals_layout(const cytof_desc* d, int64_t off[11]) {
  if (!d || !off) return CYTOF_EINVAL;
  const GlobalsLayout g = globals_layout(d->I, d->J, d->K, d->L0, d->L1);
  const int v[11] = {g.mu0, g.mu1, g.sig, g.le0, g.le1, g.lw, g.eps, g.b, g.vm, g.vls, g.total};
  for (int i = 0; i < 11; ++i) off[i] = v[i];
  return CYTOF_OK;
}
int cytof_grad_block_layout(const cytof_desc* d, int64_t off[12]) {
  if (!d || !off) return CYTOF_EINVAL;
  const BlockLayout g = block_layout(d->I, d->J, d->K, d->L0, d->L1);
  const int v[12] = {g.dmu0, g.dmu1, g.dsig, g.dle0, g.dle1, g.dlw, g.dZ, g.deps, g.dvm, g.dvls, g.dlq, g.total};
  for (int i = 0; i < 12; ++i) off[i] = v[i];
  return CYTOF_OK;
}

size_t cytof_workspace_bytes(const cytof_desc* d) {
  if (validate_desc(d) != CYTOF_OK) return 0;
  return workspace_bytes(d);
}

int cytof_elbo_fwdbwd_f32(const cytof_desc* d, const float* y, const int32_t* idx, const float* eps,
                          const float* globals, const uint64_t* zmask, float* grad_block,
                          double* ll_lqy, void* workspace, size_t workspace_bytes_, void* stream) {
  int rc = validate_desc(d);
  if (rc != CYTOF_OK) return rc;
  if (!globals || !zmask || !grad_block || !ll_lqy || !workspace) return CYTOF_EINVAL;
  if (!y && d->cell_begin[d->I] > 0) return CYTOF_EINVAL;
  if ((reinterpret_cast<uintptr_t>(ll_lqy) & 7) || (reinterpret_cast<uintptr_t>(zmask) & 7) ||
      (reinterpret_cast<uintptr_t>(y) & 3) || (reinterpret_cast<uintptr_t>(globals) & 3))
    return CYTOF_EALIGN;
  cudaError_t err = cudaSuccess;
  rc = launch_fwdbwd(d, y, idx, eps, globals, zmask, grad_block, ll_lqy, nullptr, workspace,
                     workspace_bytes_, (cudaStream_t)stream, &err);
  if (rc == CYTOF_ECUDA) set_cuda_err(err);
  return rc;
}

int cytof_elbo_fwdbwd_packed(const cytof_desc* d, const float* y, const int32_t* idx, const float* eps,
                             const float* globals, const uint64_t* zmask, double* packed,
                             void* workspace, size_t workspace_bytes_, void* stream) {
  int rc = validate_desc(d);
  if (rc != CYTOF_OK) return rc;
  if (!globals || !zmask || !packed || !workspace) return CYTOF_EINVAL;
  if (!y && d->cell_begin[d->I] > 0) return CYTOF_EINVAL;
  if ((reinterpret_cast<uintptr_t>(packed) & 7) || (reinterpret_cast<uintptr_t>(zmask) & 7)) return CYTOF_EALIGN;
  cudaError_t err = cudaSuccess;
  rc = launch_fwdbwd(d, y, idx, eps, globals, zmask, nullptr, nullptr, packed, workspace, workspace_bytes_,
                     (cudaStream_t)stream, &err);
  if (rc == CYTOF_ECUDA) set_cuda_err(err);
  return rc;
}

size_t cytof_peer_mailbox_bytes(const cytof_desc* d, int32_t world) {
  if (validate_desc(d) != CYTOF_OK || world < 1 || world > CYTOF_MAX_PEERS) return 0;
  return peer_mailbox_bytes(d, world);
}
int cytof_peer_alloc(size_t bytes, void** ptr) {
  if (!ptr || bytes == 0) return CYTOF_EINVAL;
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { set_cuda_err(e); return CYTOF_ECUDA; }
  return CYTOF_OK;
}
int cytof_peer_free(void* ptr) {
  const cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) { set_cuda_err(e); return CYTOF_ECUDA; }
  return CYTOF_OK;
}
int cytof_ipc_get_handle(void* ptr, unsigned char handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!ptr || !handle) return CYTOF_EINVAL;
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) { set_cuda_err(e); return CYTOF_ECUDA; }
  memcpy(handle, &h, 64);
  return CYTOF_OK;
}
int cytof_ipc_open_handle(const unsigned char handle[64], void** ptr) {
  if (!ptr || !handle) return CYTOF_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  const cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { set_cuda_err(e); return CYTOF_ECUDA; }
  return CYTOF_OK;
}
int cytof_ipc_close(void* ptr) {
  const cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) { set_cuda_err(e); return CYTOF_ECUDA; }
  return CYTOF_OK;
}
int cytof_peer_debug(const void* mailbox, uint64_t out[4]) {
  if (!mailbox || !out) return CYTOF_EINVAL;
  const cudaError_t e = cudaMemcpy(out, static_cast<const unsigned char*>(mailbox) + 192, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { set_cuda_err(e); return CYTOF_ECUDA; }
  return CYTOF_OK;
}
int cytof_elbo_fwdbwd_allreduce(const cytof_desc* d, const float* y, const int32_t* idx, const float* eps,
                                const float* globals, const uint64_t* zmask, double* packed,
                                void* workspace, size_t workspace_bytes_, const cytof_peer_desc* peer,
                                void* stream) {
  int rc = validate_desc(d);
  if (rc != CYTOF_OK) return rc;
  if (!globals || !zmask || !packed || !workspace || !peer) return CYTOF_EINVAL;
  if (!y && d->cell_begin[d->I] > 0) return CYTOF_EINVAL;
  if (peer->world < 1 || peer->world > CYTOF_MAX_PEERS || peer->rank < 0 || peer->rank >= peer->world ||
      peer->seq == 0)
    return CYTOF_EINVAL;
  for (int q = 0; q < peer->world; ++q)
    if (peer->world > 1 && !peer->mailbox[q]) return CYTOF_EINVAL;
  if ((reinterpret_cast<uintptr_t>(packed) & 7) || (reinterpret_cast<uintptr_t>(zmask) & 7)) return CYTOF_EALIGN;
  cudaError_t err = cudaSuccess;
  rc = launch_fwdbwd(d, y, idx, eps, globals, zmask, nullptr, nullptr, packed, workspace, workspace_bytes_,
                     (cudaStream_t)stream, &err, peer);
  if (rc == CYTOF_ECUDA) set_cuda_err(err);
  return rc;
}

int cytof_global_sizes(const cytof_global_desc* g, int64_t out[4]) {
  if (!g || !out) return CYTOF_EINVAL;
  return global_sizes(g, out);
}

int cytof_global_fwd_f64(const cytof_global_desc* g, const double* theta, const double* eps_in,
                         double* stash, float* globals, uint64_t* zmask, double* scalars, void* stream) {
  if (!g || !theta || !stash || !globals || !zmask || !scalars) return CYTOF_EINVAL;
  if (g->noise_mode == CYTOF_NOISE_EXTERNAL && !eps_in) return CYTOF_EINVAL;
  cudaError_t err = cudaSuccess;
  const int rc = launch_global_fwd(g, theta, eps_in, stash, globals, zmask, scalars, (cudaStream_t)stream, &err);
  if (rc == CYTOF_ECUDA) set_cuda_err(err);
  return rc;
}

int cytof_global_bwd_adam_f64(const cytof_global_desc* g, double* theta, double* stash,
                              const double* packed, const double* scalars_in, double* adam_m,
                              double* adam_v, double* grad_out, double* out_scalars,
                              int64_t* step_dev, int32_t* flag, int32_t do_update, void* stream) {
  if (!g || !theta || !stash || !packed || !scalars_in || !out_scalars) return CYTOF_EINVAL;
  if (do_update && (!adam_m || !adam_v)) return CYTOF_EINVAL;
  cudaError_t err = cudaSuccess;
  const int rc = launch_global_bwd(g, theta, stash, packed, scalars_in, adam_m, adam_v, grad_out, out_scalars,
                                   step_dev, flag, do_update, (cudaStream_t)stream, &err);
  if (rc == CYTOF_ECUDA) set_cuda_err(err);
  return rc;
}

int cytof_impute_emit_f32(const cytof_desc* d, const float* y, const int32_t* idx, const float* eps,
                          const float* globals, float* y_out, void* stream) {
  int rc = validate_desc(d);
  if (rc != CYTOF_OK) return rc;
  if (!y || !globals || !y_out) return CYTOF_EINVAL;
  cudaError_t err = cudaSuccess;
  rc = launch_emit(d, y, idx, eps, globals, y_out, nullptr, (cudaStream_t)stream, &err);
  if (rc == CYTOF_ECUDA) set_cuda_err(err);
  return rc;
}

int cytof_label_sample_f32(const cytof_desc* d, const float* y_imp, const float* globals,
                           const uint64_t* zmask, int32_t* lam_out, float* probs_out, void* stream) {
  int rc = validate_desc(d);
  if (rc != CYTOF_OK) return rc;
  if (!y_imp || !globals || !zmask || !lam_out) return CYTOF_EINVAL;
  cudaError_t err = cudaSuccess;
  rc = launch_labels(d, y_imp, globals, zmask, lam_out, probs_out, (cudaStream_t)stream, &err);
  if (rc == CYTOF_ECUDA) set_cuda_err(err);
  return rc;
}

int cytof_noise_emit_f32(const cytof_desc* d, const int32_t* idx, float* eps_out, void* stream) {
  int rc = validate_desc(d);
  if (rc != CYTOF_OK) return rc;
  if (!eps_out) return CYTOF_EINVAL;
  cytof_desc dd = *d;
  dd.noise_mode = CYTOF_NOISE_PHILOX;
  cudaError_t err = cudaSuccess;
  rc = launch_emit(&dd, nullptr, idx, nullptr, nullptr, nullptr, eps_out, (cudaStream_t)stream, &err);
  if (rc == CYTOF_ECUDA) set_cuda_err(err);
  return rc;
}

int cytof_adam_step_f64(double* param, const double* grad, double* exp_avg, double* exp_avg_sq,
                        int64_t n, int64_t n_scrub, double grad_scale, double lr, double beta1,
                        double beta2, double eps, int64_t step_count, int64_t* step_dev,
                        int32_t* flag, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || n < 0) return CYTOF_EINVAL;
  if (n == 0) return CYTOF_OK;
  cudaStream_t s = (cudaStream_t)stream;
  adam_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
      param, grad, exp_avg, exp_avg_sq, n, n_scrub, grad_scale, lr, beta1, beta2, eps,
      (long long)step_count, reinterpret_cast<long long*>(step_dev), flag);
  if (step_dev) adam_bump_kernel<<<1, 1, 0, s>>>(reinterpret_cast<long long*>(step_dev));
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_cuda_err(err); return CYTOF_ECUDA; }
  return CYTOF_OK;
}

static int sample_minibatch_impl(int64_t N, int64_t m, uint64_t seed, uint64_t step, const int64_t* step_dev,
                                 int32_t sample, int32_t* idx_out, void* stream) {
  if (!idx_out || N < 0 || m < 0 || N > 0x7fffffffLL) return CYTOF_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  if (N == 0 || m == 0) return CYTOF_OK;
  if (m >= N) {
    iota_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(N, idx_out);
  } else {
    int bits = 1;
    while ((1LL << bits) < N) ++bits;
    int hbits = (bits + 1) / 2;
    if (hbits < 1) hbits = 1;
    const unsigned long long z = sampler_key(seed, step, sample);
    sample_minibatch_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(
        N, m, z, hbits, idx_out, seed, reinterpret_cast<const long long*>(step_dev), sample);
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_cuda_err(err); return CYTOF_ECUDA; }
  return CYTOF_OK;
}

int cytof_sample_minibatch(int64_t N, int64_t m, uint64_t seed, uint64_t step, int32_t sample,
                           int32_t* idx_out, void* stream) {
  return sample_minibatch_impl(N, m, seed, step, nullptr, sample, idx_out, stream);
}
int cytof_sample_minibatch_dev(int64_t N, int64_t m, uint64_t seed, const int64_t* step_dev,
                               int32_t sample, int32_t* idx_out, void* stream) {
  if (!step_dev) return CYTOF_EINVAL;
  return sample_minibatch_impl(N, m, seed, 0, step_dev, sample, idx_out, stream);
}
int cytof_sample_minibatch_multi(int32_t I, const int64_t* N, const int64_t* m, uint64_t seed, uint64_t step,
                                 const int64_t* step_dev, int32_t* idx_out, void* stream) {
  if (!idx_out || !N || !m || I < 1 || I > CYTOF_MAX_SAMPLES) return CYTOF_EINVAL;
  MultiSample ms;
  long long off = 0, mmax = 0;
  for (int i = 0; i < I; ++i) {
    if (N[i] < 0 || m[i] < 0 || N[i] > 0x7fffffffLL) return CYTOF_EINVAL;
    ms.N[i] = N[i];
    ms.m[i] = m[i] < N[i] ? m[i] : N[i];
    ms.off[i] = off;
    off += ms.m[i];
    int bits = 1;
    while ((1LL << bits) < N[i]) ++bits;
    ms.hbits[i] = (bits + 1) / 2 < 1 ? 1 : (bits + 1) / 2;
    if (ms.m[i] > mmax) mmax = ms.m[i];
  }
  if (mmax == 0) return CYTOF_OK;
  const dim3 grid((unsigned)((mmax + 255) / 256), (unsigned)I);
  cudaError_t err = launch_pdl(sample_minibatch_multi_kernel, grid, dim3(256), 0, (cudaStream_t)stream, ms, seed,
                               reinterpret_cast<const long long*>(step_dev), step, idx_out);
  if (err != cudaSuccess) { set_cuda_err(err); return CYTOF_ECUDA; }
  return CYTOF_OK;
}

}  // extern "C"
