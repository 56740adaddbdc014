// Device helpers shared by the sm_100a kernels of cytof_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/cytof_b200.h"

namespace cytof {

#ifndef CYTOF_THREADS
#define CYTOF_THREADS 256
#endif
constexpr int kThreads = CYTOF_THREADS;   // warps per CTA = kThreads / 32
constexpr int kWarps = kThreads / 32;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kHalfLog2Pi = 0.9189385332046727f;  // 0.5*log(2*pi)
constexpr float kNegBig = -1.0e30f;                   // stands in for log(0) without inf-inf NaNs

__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// sm_100a warp reductions that do not go through the shuffle network:
// CREDUX.MAX.F32 (one instruction, result in a uniform register) ...
__device__ __forceinline__ float warp_max_redux(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
// ... and an exact-order-independent sum of values in [0, 64) via 2^-24 fixed point + REDUX.SUM
// (abs. error <= 32 * 2^-25; used for softmax denominators whose largest term is 1).
__device__ __forceinline__ float warp_sum_unit(float v) {
  unsigned q = __float2uint_rn(v * 16777216.0f), r;
  asm volatile("redux.sync.add.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(q));
  return (float)r * (1.0f / 16777216.0f);
}

// streaming 4-byte load that does not pollute L1 (each y row is read exactly once)
__device__ __forceinline__ float ld_stream(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based, no state ----
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

// Standard-normal draws for (position, marker j) of step `step`; position = row_global[i] + index of the cell in
// the sample's batch -- the global row for full batches (independent of how the cells are split over warps, CTAs or
// GPUs), the slot of the minibatch for gathered ones (the reference draws a dense (B_i, J) normal per step too,
// vae.py:33; until round 2 gathered batches were keyed by the row, which cost one Philox block per cell).
// One Philox4x32-10 block is keyed by (position / 4, marker) and yields FOUR normals (two Box-Muller pairs), one
// for each of the 4 consecutive positions of the group: a warp that walks a tile pays one block per 4 cells.
// Third counter word of the per-cell draws: marker | sample << 8, so that row r of different samples (every sample's
// rows start at row_global[i], often 0) gets independent draws, as the reference's dense normal_ does (vae.py:33).
__device__ __forceinline__ uint32_t philox_ctr(int j, int sample) { return (uint32_t)j | ((uint32_t)sample << 8); }
__device__ __forceinline__ float philox_unit(uint32_t w) { return ((float)(w >> 8) + 0.5f) * (1.0f / 16777216.0f); }
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t step, uint64_t grp, uint32_t j, float out[4]) {
  uint32_t c[4] = {(uint32_t)grp, (uint32_t)(grp >> 32), j, step};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float r0 = sqrtf(-2.0f * kLn2 * fast_lg2(philox_unit(c[0])));
  const float r1 = sqrtf(-2.0f * kLn2 * fast_lg2(philox_unit(c[2])));
  const float a0 = 6.283185307179586f * philox_unit(c[1]), a1 = 6.283185307179586f * philox_unit(c[3]);
  out[0] = r0 * __cosf(a0);
  out[1] = r0 * __sinf(a0);
  out[2] = r1 * __cosf(a1);
  out[3] = r1 * __sinf(a1);
}
// the draw of one (row, marker): the same values as philox_normal4, only the needed one evaluated
__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t step, uint64_t grow, uint32_t j) {
  const uint64_t grp = grow >> 2;
  uint32_t c[4] = {(uint32_t)grp, (uint32_t)(grp >> 32), j, step};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const bool second = (grow & 2ull) != 0ull;
  const float r = sqrtf(-2.0f * kLn2 * fast_lg2(philox_unit(second ? c[2] : c[0])));
  const float a = 6.283185307179586f * philox_unit(second ? c[3] : c[1]);
  return r * ((grow & 1ull) ? __sinf(a) : __cosf(a));
}

// Offsets (in floats) of the sections of `globals` (include/cytof_b200.h).
struct GlobalsLayout {
  int mu0, mu1, sig, le0, le1, lw, eps, b, vm, vls, total;
};
__host__ __device__ inline GlobalsLayout globals_layout(int I, int J, int K, int L0, int L1) {
  GlobalsLayout g;
  int o = 0;
  g.mu0 = o; o += L0;
  g.mu1 = o; o += L1;
  g.sig = o; o += I;
  g.le0 = o; o += I * J * L0;
  g.le1 = o; o += I * J * L1;
  g.lw = o; o += I * K;
  g.eps = o; o += I;
  g.b = o; o += I * 3;
  g.vm = o; o += I * J;
  g.vls = o; o += I * J;
  g.total = o;
  return g;
}

// Offsets of the sections of `grad_block`.
struct BlockLayout {
  int dmu0, dmu1, dsig, dle0, dle1, dlw, dZ, deps, dvm, dvls, dlq, total;
};
__host__ __device__ inline BlockLayout block_layout(int I, int J, int K, int L0, int L1) {
  BlockLayout g;
  int o = 0;
  g.dmu0 = o; o += L0;
  g.dmu1 = o; o += L1;
  g.dsig = o; o += I;
  g.dle0 = o; o += I * J * L0;
  g.dle1 = o; o += I * J * L1;
  g.dlw = o; o += I * K;
  g.dZ = o; o += J * K;
  g.deps = o; o += I;
  g.dvm = o; o += I * J;
  g.dvls = o; o += I * J;
  g.dlq = o; o += I * J;
  g.total = o;
  return g;
}

// Per-CTA partial-sum record written by the fused kernel and consumed by the finalise
// kernel.  One CTA only ever sees cells of ONE sample, so per-sample sections need no
// sample index here.
struct PartialLayout {
  int sw0, sw1, swd0, swd1, sc, gW, dZ, dvm, dvls, nm, dbl, total;
  // sc: [0]=sum w d^2  [1]=sum fac*rho  [2]=sum fac*(1-rho)   dbl: 2 doubles (ll, lqy) as 4 floats
};
__host__ __device__ inline PartialLayout partial_layout(int J, int K, int L0, int L1) {
  PartialLayout g;
  int o = 0;
  g.sw0 = o; o += J * L0;
  g.sw1 = o; o += J * L1;
  g.swd0 = o; o += L0;
  g.swd1 = o; o += L1;
  g.sc = o; o += 4;
  g.gW = o; o += K;
  g.dZ = o; o += J * K;
  g.dvm = o; o += J;
  g.dvls = o; o += J;
  g.nm = o; o += J;
  o = (o + 3) & ~3;
  g.dbl = o; o += 4;
  g.total = (o + 3) & ~3;
  return g;
}

struct ElboParams {
  const float* __restrict__ y;
  const int32_t* __restrict__ idx;
  const float* __restrict__ eps;
  const float* __restrict__ globals;
  const unsigned long long* __restrict__ zmask;
  float* partials;
  float* y_out;  // impute-emit mode only
  int I, J, K, L0, L1;
  int noise_mode, model_noisy;
  float noisy_sd, scale;
  unsigned long long seed;
  uint32_t step;
  const long long* step_dev;   // non-null: the step is *step_dev + 1 (CUDA-graph replay)
  int nblocks;
  long long cell_begin[CYTOF_MAX_SAMPLES + 1];
  long long row_base[CYTOF_MAX_SAMPLES];
  long long row_global[CYTOF_MAX_SAMPLES];
  float fac[CYTOF_MAX_SAMPLES];
  int blk_begin[CYTOF_MAX_SAMPLES + 1];
  // CYTOF_FLAG_SAMPLE_IN_KERNEL: the kernel draws the minibatch rows itself (RowSampler below); idx is null
  int sampler_on;
  unsigned long long sampler_seed, sampler_step;     // the step is *step_dev + 1 when step_dev is set
  int sampler_N[CYTOF_MAX_SAMPLES];
  unsigned char sampler_hbits[CYTOF_MAX_SAMPLES];
};

// The heavy template instantiations live in their own translation units (variants_mma.cu,
// variants_mma64.cu) so that the library builds in parallel; elbo_kernel.cu asks for the kernels.
using KernelFn = void (*)(const ElboParams);
// lidx: 0 = (3,3), 1 = (5,3), 2 = (5,5), 3 = (8,8) mixture sizes; fi = noisy + 2 * has_idx + 4 * no_missing
KernelFn mma_kernel_fn(int lidx, int fi);
// lidx: 0 = (5,5), 1 = (8,8), 2 = (5,5) in the WIDE layout (48 < J <= 64; fi = noisy + 2 * has_idx);
// missing: imputation path compiled in; fi = noisy + 2 * has_idx + 4 * (tail passes - 1)
KernelFn mma64_kernel_fn(int lidx, int missing, int fi);
// CTA-wide tcgen05 kernel (elbo_kernel_umma.cuh): lidx 0 = (3,3), 1 = (5,3), 2 = (5,5); fi as for mma_kernel_fn
KernelFn umma_kernel_fn(int lidx, int fi);
int umma_smem_bytes();
int umma_threads();

__device__ __forceinline__ uint32_t elbo_step(const ElboParams& p) {
  return p.step_dev ? (uint32_t)(*p.step_dev + 1) : p.step;
}

// ---- without-replacement minibatch: keyed bijection of [0, N) (balanced Feistel network + cycle walking); position t
// of the minibatch is row perm(t).  Replaces np.random.choice(N_i, mb, replace=False) of reference fit.py:96-102.
// Used by the sampler kernels (abi.cu) and, with CYTOF_FLAG_SAMPLE_IN_KERNEL, evaluated by the per-cell kernels
// themselves, so that a minibatch step needs neither a sampler launch nor an index buffer.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__host__ __device__ inline unsigned long long sampler_key(unsigned long long seed, unsigned long long step, int sample) {
  // splitmix64 of (seed, step, sample) -> Feistel key
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (step + 1) + 0xD1B54A32D192ED03ull * (unsigned long long)(sample + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return z;
}
__host__ __device__ inline int sampler_half_bits(long long N) {
  int bits = 1;
  while ((1LL << bits) < N) ++bits;
  return (bits + 1) / 2 < 1 ? 1 : (bits + 1) / 2;
}
__host__ __device__ __forceinline__ unsigned long long feistel_perm(unsigned long long key, int hbits, unsigned long long N,
                                                                    unsigned long long t) {
  const uint32_t hmask = (1u << hbits) - 1u;
  const uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
  unsigned long long x = t;
  do {
    uint32_t L = (uint32_t)(x >> hbits) & hmask, R = (uint32_t)x & hmask;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const uint32_t f = mix32(R ^ (k0 + 0x9E3779B9u * (uint32_t)r) ^ (k1 << (r & 7))) & hmask;
      const uint32_t nl = R;
      R = L ^ f;
      L = nl;
    }
    x = ((unsigned long long)L << hbits) | R;
  } while (x >= N);
  return x;
}
// rows of sample i's minibatch inside a per-cell kernel: pos = position in the sample's batch on this rank
struct RowSampler {
  unsigned long long key, N;
  int hbits;
  bool identity;      // the batch is the whole sample
  __device__ __forceinline__ int row(long long pos) const {
    return identity ? (int)pos : (int)feistel_perm(key, hbits, N, (unsigned long long)pos);
  }
};
__device__ __forceinline__ RowSampler make_row_sampler(const ElboParams& p, int i) {
  RowSampler s;
  s.N = (unsigned long long)p.sampler_N[i];
  s.hbits = p.sampler_hbits[i];
  s.identity = (unsigned long long)(p.cell_begin[i + 1] - p.cell_begin[i]) >= s.N;
  s.key = sampler_key(p.sampler_seed, p.step_dev ? (unsigned long long)(*p.step_dev + 1) : p.sampler_step, i);
  return s;
}

// ---- programmatic dependent launch (sm_90+): every kernel of the step is launched with the
// programmatic-stream-serialisation attribute, so its launch and scheduling overlap the drain of the
// kernel before it; each kernel starts with pdl_enter(), which (1) lets ITS successor start launching
// and (2) blocks until the predecessor grid has completed and its writes are visible.  Because every
// kernel in the chain waits before touching memory, completion is transitive.  MEASURED (B200, whole step as
// one CUDA graph, L2 flushed between steps): cfg2 68.7 us with it vs 57.5 us without, cfg4 0.325 vs 0.304 ms
// -- graph kernel->kernel edges are already cheap and the early-resident dependents get in the way -- so it
// is OFF unless CYTOF_PDL=1 is set in the environment (plain stream order; pdl_enter() is then a no-op).
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
inline bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("CYTOF_PDL"); return e && e[0] == '1'; }();
  return on;
}
template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr;
  memset(&attr, 0, sizeof(attr));
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
}

}  // namespace cytof
