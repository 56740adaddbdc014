// CTA-wide, warp-specialised Blackwell kernel for J <= 32, K <= 32 (BASELINE cfg1-cfg4): the tcgen05 / TMEM / bulk-copy
// form of elbo_kernel_mma.cuh.  OPT-IN (CYTOF_UMMA=1): it passes the same parity suite (tests/test_gpu_umma_variant.py)
// and measures 1.8x SLOWER than the per-warp kernel (0.486 vs 0.267 ms per 1 M cells at cfg4) -- tcgen05.ld moves
// 32 B/clk/SM, SS-mode MMAs stream their operands from shared memory at 128 B/clk, and the forward -> backward round
// trip through the issuers and the cell warps is ~5,000 cycles against a TMEM stash of four quarters
// (profiles/README.md, round 2).
//
// Why it was built: ncu on the per-warp kernel (profiles/README.md, r1e) showed four schedulers at 55 % issue with
// two 254-register warps each; the per-warp mma.sync contractions (7 HMMA + LDSM + fragment splits + ~18 LDS/STS
// per cell, a ~1,000-cycle dependent chain per tile) sit on the same schedulers as the MUFU / FMA work.  Here
//   * the three J x K contractions of Model.py:246-249 (+ autograd twins) are tcgen05.mma.kind::tf32 with M = 128,
//     N = 32, issued by ONE thread each, operands in shared memory, accumulators in tensor memory:
//         f   [n, k]  = A1[n, :] . Z[:, k] + N0[n, :] . (Z[:, k] - 1)            (N0 = -A0; rows = cells)
//         c1^T[j, n]  = Z[j, :] . g[n, :]                (rows = 4 x 32 replicated markers, columns = 32 cells)
//         dZ  [j, k] += sum_n (A1 + N0)[n, j] g[n, k]    (rows = 4 operand kinds x 32 markers, contraction over cells)
//     all operands are K-major tiles with 128-byte rows and the 128-byte swizzle (the only MN-major form
//     kind::tf32 accepts is another swizzle, so the cell-contraction reads TRANSPOSED copies [marker][cell],
//     [feature][cell] instead; tools/tcgen05_probe2.cu); values are split x = hi + lo (hi rounded to TF32, lo the
//     exact remainder) so the fp32-accumulated products carry ~22 mantissa bits;
//   * warp roles: 8 MIXTURE warps (lane = marker j: mixture log-densities forward, responsibilities backward; the
//     (j, l) statistics live in the lane's registers), 4 CELL warps (thread = cell, one TMEM quadrant each: logsumexp
//     over K, noisy-cell class, g), 1 PRODUCER warp (y rows by cp.async.bulk + mbarrier), 2 single-thread MMA issuers;
//   * the 10 responsibilities p_zl of a (cell, marker) are parked in TENSOR MEMORY between the forward and the
//     backward of the cell (tcgen05.st / tcgen05.ld, 320 of the 512 columns): a warp holds no per-cell state across
//     the tensor-core round trip, so the mixture warps need ~1/2 of the registers of the old kernel and never wait
//     on a dependent MMA chain;
//   * work unit = a QUARTER (32 cells = one TMEM quadrant); a mixture warp runs forward(Q) then backward(Q - 3):
//     three quarters of other work cover the round trip (A1, N0) -> f -> g -> c1.
// Same inputs, same per-CTA partial record and therefore the same finalise kernel as elbo_kernel_mma.cuh.
#pragma once
#include "elbo_kernel_mma.cuh"

namespace cytof {

constexpr int kUThreads = 512;
constexpr int kUMixWarps = 8;
constexpr int kULag = 3;                               // quarters between the forward and the backward of a cell
// shared memory (bytes from the 1024-aligned base); a "block" = 32 rows x 128 B = 4 KB
constexpr int kUOffY = 0;                              // [8 quarter slots][32 cells][128 B] y rows (plain)
constexpr int kUOffAR = 8 * 4096;                      // [4 kinds][2 slots] blocks, row = cell: A1_hi, A1_lo, N0_hi, N0_lo
constexpr int kUOffGR = kUOffAR + 8 * 4096;            // [2 kinds][2 slots] blocks, row = cell: g_hi, g_lo
constexpr int kUOffAT = kUOffGR + 4 * 4096;            // [4 slots][4 kinds] blocks, row = marker, 32 cells per row
constexpr int kUOffGT = kUOffAT + 16 * 4096;           // [2 slots][2 kinds] blocks, row = feature, 32 cells per row
constexpr int kUOffWZ = kUOffGT + 4 * 4096;            // [32 k][32 j] Z^T
constexpr int kUOffWZm1 = kUOffWZ + 4096;              // [32 k][32 j] Z^T - 1
constexpr int kUOffZrep = kUOffWZm1 + 4096;            // [128 m][32 k] Z[m % 32][k]
constexpr int kUOffScal = kUOffZrep + 16384;           // fr[128] | fo[128] | lw[32] floats
constexpr int kUOffBar = kUOffScal + 2048;             // mbarriers, abort flag, TMEM base
constexpr int kUSmemBytes = kUOffBar + 512 + 1024;     // + slack for the manual 1024-byte alignment
// mbarrier indices
// (a waiter must see EVERY phase of a barrier it waits on -- a parity wait cannot tell "phase u - 1 still pending" from
// "phase u complete" -- hence one f barrier per cell warp although the f accumulators are only double-buffered)
constexpr int kBYFull = 0, kBYFree = 8, kBD = 16, kBF = 20, kBFFree = 24, kBG = 26, kBC1 = 30, kBC1Free = 33, kBDZ = 36;
constexpr int kBCount = 40;
// tensor-memory columns
constexpr uint32_t kTmDZ = 0, kTmF = 32, kTmC1 = 96, kTmStash = 192;

__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(a) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// bounded wait on phase parity `par`; a time-out (or another role's time-out) raises the CTA's abort flag and every
// role leaves its loop: the launch ends with NaN in this CTA's record instead of hanging the GPU
__device__ __forceinline__ bool uwait(uint32_t bar, uint32_t par, volatile int* abort_flag) {
  for (int it = 0; it < (1 << 17); ++it) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(bar), "r"(par) : "memory");
    if (done) return true;
    if ((it & 31) == 31 && *abort_flag) return false;
  }
#ifdef CYTOF_UMMA_DEBUG
  if (!*abort_flag && (threadIdx.x & 31) == 0)
    printf("umma timeout: block %d warp %d barrier %d parity %u\n", (int)blockIdx.x, (int)(threadIdx.x >> 5),
           (int)((bar & 1023u) >> 3), par);     // the barrier table starts 1024-aligned + kUOffBar (a multiple of 1024)
#endif
  *abort_flag = 1;
  return false;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// shared-memory operand descriptors, 128-byte swizzle (layout type 2), rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t udesc_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::tf32, fp32 accumulate, both operands K-major, M = 128, N = 32
constexpr uint32_t kUIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
__device__ __forceinline__ void umma_issue(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(kUIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_to(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

// ---- tensor memory <-> registers (32 lanes x N consecutive columns; lane = TMEM lane of the warp's quadrant)
template <int N> struct TmemRegs;
__device__ __forceinline__ void tmem_st2(uint32_t a, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" :: "r"(a), "r"(r[0]), "r"(r[1]) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t a, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t a, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t a, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t a, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t a, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t a, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
      "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(a) : "memory");
}
// N words to / from N consecutive columns (N = 6, 8, 10: the stash sizes of the (3,3), (5,3), (5,5) kernels)
template <int N> __device__ __forceinline__ void tmem_stN(uint32_t a, const uint32_t* r) {
  static_assert(N == 6 || N == 8 || N == 10, "stash width");
  if (N == 6) { tmem_st4(a, r); tmem_st2(a + 4, r + 4); }
  if (N == 8) tmem_st8(a, r);
  if (N == 10) { tmem_st8(a, r); tmem_st2(a + 8, r + 8); }
}
template <int N> __device__ __forceinline__ void tmem_ldN(uint32_t a, uint32_t* r) {
  if (N == 6) { tmem_ld4(a, r); tmem_ld2(a + 4, r + 4); }
  if (N == 8) tmem_ld8(a, r);
  if (N == 10) { tmem_ld8(a, r); tmem_ld2(a + 8, r + 8); }
}
// tcgen05.wait::ld that the compiler cannot move the consumers of r[0..N) across
template <int N> __device__ __forceinline__ void tmem_wait_ld(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int q = 0; q < N; ++q) asm volatile("" : "+r"(r[q]));
}

__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ float4 lds_f128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" :: "r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

#ifdef CYTOF_UMMA_DEBUG
// per-warp event trace of block 0 (lane 0): (event code, quarter, clock), printed at the end of the kernel
#define UTRACE(code, qv) do { if (blockIdx.x == 0 && lane == 0 && ntr < 24) { tr_c[ntr] = (code) * 1000 + (qv); tr_t[ntr] = clock64(); ++ntr; } } while (0)
#define UTRACE_DECL int ntr = 0; int tr_c[24]; long long tr_t[24];
#define UTRACE_DUMP do { if (blockIdx.x == 0 && lane == 0) for (int z = 0; z < ntr; ++z) printf("trace warp %2d ev %5d t %lld\n", warp, tr_c[z], tr_t[z] - t_start); } while (0)
#else
#define UTRACE(code, qv) do { } while (0)
#define UTRACE_DECL
#define UTRACE_DUMP do { } while (0)
#endif

template <int REGS> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(REGS)); }
template <int REGS> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(REGS)); }

template <int LM0, int LM1, bool NOISY, bool HAS_IDX, bool MISSING>
__global__ void __launch_bounds__(kUThreads, 1)
elbo_fwdbwd_umma_kernel(const ElboParams p) {
  pdl_enter();
  constexpr int NC = LM0 + LM1, NPT = (NC + 1) / 2, NST = 2 * NPT;     // components, packed pairs, stash words per (cell, marker)
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* sgen = smem_raw + (sbase - smem_u32(smem_raw));          // generic pointer to the aligned base
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  volatile int* abort_flag = reinterpret_cast<volatile int*>(sgen + kUOffBar + kBCount * 8);
  volatile uint32_t* tmem_base_s = reinterpret_cast<volatile uint32_t*>(sgen + kUOffBar + kBCount * 8 + 8);
  auto bar = [&](int b) -> uint32_t { return sbase + kUOffBar + 8u * (uint32_t)b; };

  const int I = p.I, J = p.J, K = p.K, L0 = p.L0, L1 = p.L1;
  int i = 0;
  while (i + 1 < I && (int)blockIdx.x >= p.blk_begin[i + 1]) ++i;
  const int nb = p.blk_begin[i + 1] - p.blk_begin[i];
  const int lb = (int)blockIdx.x - p.blk_begin[i];
  const long long cb = p.cell_begin[i];
  const long long Bi = p.cell_begin[i + 1] - cb;
  const long long row_base = p.row_base[i];
  const unsigned long long grow_base = (unsigned long long)p.row_global[i];
  // this CTA's cells [c_lo, c_hi) of the sample; full batch: cut points at global rows that are multiples of 8
  auto cut = [&](const long long q) -> long long {
    long long x = Bi * q / nb;
    if (q > 0 && q < nb) x -= (long long)((grow_base + (unsigned long long)x) & 7ull);
    return x < 0 ? 0 : x;
  };
  const long long c_lo = cut(lb), c_hi = cut(lb + 1);
  const int ncell = (int)(c_hi - c_lo);
  const int NQ = (ncell + 31) >> 5;
  const GlobalsLayout gl = globals_layout(I, J, K, L0, L1);
  const float* __restrict__ G = p.globals;
  const float fac = p.fac[i];

  // ---------------- one-time setup: constant operand tiles, barriers, tensor memory ----------------
  {
    const unsigned kmask = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
    for (int e = threadIdx.x; e < 32 * 32; e += kUThreads) {        // weight tiles of f: row = feature k, 32 markers
      const int k = e >> 5, j = e & 31;
      const bool live = j < J && k < K;
      const float z = (live && ((((unsigned)p.zmask[j] & kmask) >> k) & 1u)) ? 1.f : 0.f;
      const uint32_t off = (uint32_t)(k * 128 + (((j >> 2) ^ (k & 7)) << 4) + (j & 3) * 4);
      sts_f32(sbase + kUOffWZ + off, z);
      sts_f32(sbase + kUOffWZm1 + off, live ? z - 1.f : 0.f);
    }
    for (int e = threadIdx.x; e < 128 * 32; e += kUThreads) {       // A operand of c1^T: row m = marker m % 32, 32 features
      const int m = e >> 5, k = e & 31, j = m & 31;
      const float z = (j < J && k < K && ((((unsigned)p.zmask[j] & kmask) >> k) & 1u)) ? 1.f : 0.f;
      sts_f32(sbase + kUOffZrep + (uint32_t)(m * 128 + (((k >> 2) ^ (m & 7)) << 4) + (k & 3) * 4), z);
    }
    if (threadIdx.x < 32)
      sts_f32(sbase + kUOffScal + 1024 + 4 * threadIdx.x, (int)threadIdx.x < K ? G[gl.lw + i * K + threadIdx.x] * kLog2e : kNegBig);
    if (threadIdx.x < 256) sts_f32(sbase + kUOffScal + 4 * threadIdx.x, 0.f);
    if (threadIdx.x == 0) {
      for (int b = 0; b < 8; ++b) { mbar_init(bar(kBYFull + b), 1); mbar_init(bar(kBYFree + b), kUMixWarps); }
      for (int b = 0; b < 4; ++b) { mbar_init(bar(kBD + b), kUMixWarps); mbar_init(bar(kBG + b), 1); mbar_init(bar(kBDZ + b), 1); mbar_init(bar(kBF + b), 1); }
      for (int b = 0; b < 2; ++b) mbar_init(bar(kBFFree + b), 1);
      for (int b = 0; b < 3; ++b) { mbar_init(bar(kBC1 + b), 1); mbar_init(bar(kBC1Free + b), kUMixWarps); }
      *abort_flag = 0;
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(sbase + kUOffBar + kBCount * 8 + 8) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  const uint32_t tmem_base = *tmem_base_s;
#ifdef CYTOF_UMMA_DEBUG
  const long long t_start = clock64();
#endif
  UTRACE_DECL
  const PartialLayout pl = partial_layout(J, K, L0, L1);
  float* sred = reinterpret_cast<float*>(sgen);          // 12 per-warp records alias the tiles after the main loop
  float* mine = sred + (warp < 12 ? warp : 0) * pl.total;
  bool good = true;

  if (warp < kUMixWarps) {
    // =========================== MIXTURE warps: lane = marker j ===========================
    reg_inc<168>();
    const int w = warp;
    const bool ok = lane < J;
    const float okf = ok ? 1.f : 0.f;
    const int jj = ok ? lane : 0;
    const float sig = G[gl.sig + i];
    const float inv_sig2 = 1.0f / (sig * sig);
    const float nh2 = -0.5f * inv_sig2 * kLog2e;
    const float cbase = (-logf(sig) - kHalfLog2Pi) * kLog2e;
    const float b0 = G[gl.b + 3 * i], b1 = G[gl.b + 3 * i + 1], b2 = G[gl.b + 3 * i + 2];
    // mixture constants exactly as in elbo_kernel_mma.cuh: component x of the flat list (mixture 0 first),
    // v_x = nh y'^2 + [be_x y' + al_x] with y' = y - (mid-range of the mixture's means)
    float2 mup[NPT], be[NPT], al[NPT];
    float mc0, mc1;
    {
      const float m0a = G[gl.mu0], m0b = G[gl.mu0 + L0 - 1], m1a = G[gl.mu1], m1b = G[gl.mu1 + L1 - 1];
      mc0 = 0.5f * (m0a + m0b);
      mc1 = 0.5f * (m1a + m1b);
      float m[2 * NPT], b[2 * NPT], a[2 * NPT];
#pragma unroll
      for (int x = 0; x < 2 * NPT; ++x) {
        const bool z1 = x >= LM0;
        const int l = z1 ? x - LM0 : x, Lz = z1 ? L1 : L0;
        const bool live = l < Lz && x < LM0 + LM1;
        const int ll = live ? l : 0;
        m[x] = live ? G[(z1 ? gl.mu1 : gl.mu0) + ll] - (z1 ? mc1 : mc0) : 0.f;
        b[x] = live ? -2.f * nh2 * m[x] : 0.f;
        a[x] = (ok && live) ? fmaf(nh2 * m[x], m[x], fmaf(G[(z1 ? gl.le1 : gl.le0) + (i * J + jj) * Lz + ll], kLog2e, cbase))
                            : (l == 0 ? 0.f : kNegBig);
      }
#pragma unroll
      for (int q = 0; q < NPT; ++q) {
        mup[q] = mkp(m[2 * q], m[2 * q + 1]);
        be[q] = mkp(b[2 * q], b[2 * q + 1]);
        al[q] = mkp(a[2 * q], a[2 * q + 1]);
      }
    }
    const float vm = G[gl.vm + i * J + jj], vls = G[gl.vls + i * J + jj], sd = expf(vls), inv_sd = 1.0f / sd;
    const float inv_s2n = NOISY ? 1.0f / (p.noisy_sd * p.noisy_sd) : 0.f;

    float2 sw[NPT], swy[NPT];
    float sq = 0.f, dvm = 0.f, dvls = 0.f, nm = 0.f, lqy_l = 0.f, lmiss_l = 0.f;
#pragma unroll
    for (int q = 0; q < NPT; ++q) sw[q] = swy[q] = bcp(0.f);

    const uint32_t tm_lane = (uint32_t)(32 * (w & 3)) << 16;
    const uint32_t stash0 = tmem_base + tm_lane + kTmStash + 160u * (uint32_t)(w >> 2);
    const uint32_t c1col0 = tmem_base + tm_lane + kTmC1 + 4u * (uint32_t)w;
    // cell rows 4 w + c of a quarter block: byte offset of this lane's word inside the swizzled row
    uint32_t loff[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) loff[c] = (uint32_t)((4 * w + c) * 128 + ((((lane >> 2) ^ ((4 * w + c) & 7)) << 4) | ((lane & 3) << 2)));
    // transposed blocks: row = marker (lane), this warp's 4 cells are the 16-byte chunk w of the row
    const uint32_t toff = (uint32_t)(lane * 128 + ((w ^ (lane & 7)) << 4));
    const uint32_t yoff = (uint32_t)(w * 512 + 4 * lane);
    unsigned missring = 0u, anyring = 0u;      // bit 4 (Q & 3) + c: this lane's entry / any entry of the cell is missing
    const int32_t* __restrict__ idx = p.idx + (HAS_IDX ? (cb + c_lo) : 0);
    const int first_row = (int)c_lo;

    for (int it = 0; it < NQ + kULag && good; ++it) {
      if (it < NQ) {
        // ---------------- forward of quarter Q (Model.py:223-233): A1, N0 operands + responsibilities ----------------
        const int Q = it, qq = Q & 3;
        good = uwait(bar(kBYFull + (Q & 7)), (uint32_t)(Q >> 3) & 1u, abort_flag) && good;
        if (Q >= 2) good = uwait(bar(kBF + ((Q - 2) & 3)), (uint32_t)((Q - 2) >> 2) & 1u, abort_flag) && good;   // f MMA of Q - 2 read its rows
        if (Q >= 4) good = uwait(bar(kBDZ + qq), (uint32_t)((Q >> 2) - 1) & 1u, abort_flag) && good;      // dZ MMA of Q - 4 read its block
        if (!good) break;
        UTRACE(1, Q);
        const uint32_t ybase = sbase + kUOffY + (uint32_t)(Q & 7) * 4096u + yoff;
        float xs[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          xs[c] = lds_f32(ybase + 128u * c);
          if (!ok) xs[c] = 0.f;             // J < 32: the pad words of a row are never written
        }
        if (MISSING) {
          // imputation (vae.py:17-33): NaN -> vm + sd e, written back into the y rows (the cell warps and the
          // backward read the imputed value); only quarters that hold a NaN pay for the draws
          unsigned mm = 0u;
#pragma unroll
          for (int c = 0; c < 4; ++c) mm |= (xs[c] != xs[c]) ? (1u << c) : 0u;
          const unsigned am = __reduce_or_sync(FULL, mm);
          missring = (missring & ~(0xfu << (4 * qq))) | (mm << (4 * qq));
          anyring = (anyring & ~(0xfu << (4 * qq))) | (am << (4 * qq));
          if (am) {
            float e[4] = {0.f, 0.f, 0.f, 0.f};
            const int cell0 = Q * 32 + 4 * w;
            const uint32_t ctr_j = philox_ctr(lane, i);
            if (p.noise_mode == CYTOF_NOISE_EXTERNAL) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                if (p.eps && ((mm >> c) & 1u)) e[c] = __ldg(p.eps + (cb + c_lo + cell0 + c) * J + lane);
            } else {
              // the draws are keyed by the POSITION in the batch (= the row for full batches): the 4 cells lie in one
              // Philox block of 4 positions (aligned) or in two
              const unsigned long long grow0 = grow_base + (unsigned long long)(first_row + cell0);
              const unsigned a = (unsigned)grow0 & 3u;
              float n8[8];
              philox_normal4(p.seed, elbo_step(p), grow0 >> 2, ctr_j, n8);
              if (a) philox_normal4(p.seed, elbo_step(p), (grow0 >> 2) + 1ull, ctr_j, n8 + 4);
              else n8[4] = n8[5] = n8[6] = n8[7] = 0.f;
#pragma unroll
              for (int c = 0; c < 4; ++c) e[c] = a == 0u ? n8[c] : (a == 1u ? n8[c + 1] : (a == 2u ? n8[c + 2] : n8[c + 3]));
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if ((mm >> c) & 1u) {
                xs[c] = fmaf(sd, e[c], vm);
                sts_f32(ybase + 128u * c, xs[c]);
              }
          }
        }
        uint32_t h1[4], l1[4], h0[4], l0[4];
        const uint32_t arow = sbase + kUOffAR + (uint32_t)(Q & 1) * 4096u;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float yv = xs[c];
          const float y0 = yv - mc0, y1 = yv - mc1;
          const float2 Y00 = bcp(y0), Y11 = bcp(y1), Y01 = mkp(y0, y1);
          float2 v[NPT];
#pragma unroll
          for (int q = 0; q < NPT; ++q) {
            const bool lo1 = 2 * q >= LM0, hi1 = 2 * q + 1 >= LM0;
            v[q] = fma2(be[q], lo1 ? Y11 : (hi1 ? Y01 : Y00), al[q]);
          }
          float M0 = kNegBig, M1 = kNegBig;
#pragma unroll
          for (int q = 0; q < NPT; ++q) constexpr_pair_max<LM0, NC>(q, v[q], M0, M1);
          const float2 M00 = bcp(M0), M11 = bcp(M1), M01 = mkp(M0, M1);
          float2 ee[NPT];
          float2 s0 = bcp(0.f), s1 = bcp(0.f);
          float s0x = 0.f, s1x = 0.f;
#pragma unroll
          for (int q = 0; q < NPT; ++q) {
            const bool lo1 = 2 * q >= LM0, hi1 = 2 * q + 1 >= LM0, hi_live = 2 * q + 1 < NC;
            const float2 tt = sub2(v[q], lo1 ? M11 : (hi1 ? M01 : M00));
            ee[q] = mkp(fast_ex2(lo2(tt)), hi_live ? fast_ex2(hi2(tt)) : 0.f);
            if (lo1) s1 = add2(s1, ee[q]);
            else if (!hi1) s0 = add2(s0, ee[q]);
            else { s0x = lo2(ee[q]); s1x = hi2(ee[q]); }
          }
          const float S0 = (lo2(s0) + hi2(s0)) + s0x, S1 = (lo2(s1) + hi2(s1)) + s1x;
          const float A0 = fmaf(nh2 * y0, y0, M0) + fast_lg2(S0), A1 = fmaf(nh2 * y1, y1, M1) + fast_lg2(S1);
          const float rr = fast_rcp(S0 * S1);          // one reciprocal for both mixtures (S in [1, L])
          const float r0 = rr * S1, r1 = rr * S0;
          const float2 R00 = bcp(r0), R11 = bcp(r1), R01 = mkp(r0, r1);
          uint32_t st[NST];
#pragma unroll
          for (int q = 0; q < NPT; ++q) {
            const bool lo1 = 2 * q >= LM0, hi1 = 2 * q + 1 >= LM0;
            const float2 pp = mul2(ee[q], lo1 ? R11 : (hi1 ? R01 : R00));
            st[2 * q] = __float_as_uint(lo2(pp));
            st[2 * q + 1] = __float_as_uint(hi2(pp));
          }
          tmem_stN<NST>(stash0 + (uint32_t)((4 * qq + c) * NST), st);
          split_tf32(okf * A1, h1[c], l1[c]);
          split_tf32(-okf * A0, h0[c], l0[c]);
          const uint32_t o = arow + loff[c];           // row of the cell in the four row-major operand blocks
          sts_u32(o, h1[c]);
          sts_u32(o + 8192u, l1[c]);
          sts_u32(o + 16384u, h0[c]);
          sts_u32(o + 24576u, l0[c]);
        }
        {   // the same values transposed (row = marker, 4 consecutive cells = one 16-byte chunk): operand of dZ
          const uint32_t o = sbase + kUOffAT + (uint32_t)qq * 16384u + toff;
          sts_u128(o, h1[0], h1[1], h1[2], h1[3]);
          sts_u128(o + 4096u, l1[0], l1[1], l1[2], l1[3]);
          sts_u128(o + 8192u, h0[0], h0[1], h0[2], h0[3]);
          sts_u128(o + 12288u, l0[0], l0[1], l0[2], l0[3]);
        }
        tc_wait_st();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kBD + qq));
        UTRACE(2, Q);
      }
      if (it >= kULag) {
        // ---------------- backward of quarter Qb: w_zl = c_z p_zl and the (j, l) statistics ----------------
        const int Qb = it - kULag, qq = Qb & 3, cs = Qb % 3;
        good = uwait(bar(kBC1 + cs), (uint32_t)(Qb / 3) & 1u, abort_flag) && good;
        good = uwait(bar(kBG + qq), (uint32_t)(Qb >> 2) & 1u, abort_flag) && good;     // (complete already: orders the scalar rows)
        if (!good) break;
        UTRACE(3, Qb);
        tc_fence_after();
        uint32_t c1r[4];
        tmem_ld4(c1col0 + 32u * (uint32_t)cs, c1r);
        uint32_t sa[NST], sb[NST];
        tmem_ldN<NST>(stash0 + (uint32_t)((4 * qq) * NST), sa);
        tmem_wait_ld<4>(c1r);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kBC1Free + cs));
        const uint32_t ybase = sbase + kUOffY + (uint32_t)(Qb & 7) * 4096u + yoff;
        const float4 fr4 = lds_f128(sbase + kUOffScal + (uint32_t)(4 * (32 * qq + 4 * w)));
        float4 fo4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MISSING && NOISY) fo4 = lds_f128(sbase + kUOffScal + 512u + (uint32_t)(4 * (32 * qq + 4 * w)));
        const float frs[4] = {fr4.x, fr4.y, fr4.z, fr4.w}, fos[4] = {fo4.x, fo4.y, fo4.z, fo4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t* cur = (c & 1) ? sb : sa;
          uint32_t* nxt = (c & 1) ? sa : sb;
          tmem_wait_ld<NST>(cur);
          if (c < 3) tmem_ldN<NST>(stash0 + (uint32_t)((4 * qq + c + 1) * NST), nxt);
          float x = lds_f32(ybase + 128u * c);
          if (!ok) x = 0.f;
          const float c1 = __uint_as_float(c1r[c]);      // = 0 for markers >= J (zero rows of Z)
          const float c0 = okf * frs[c] - c1;
          const float y0 = x - mc0, y1 = x - mc1;
          const float2 Y00 = bcp(y0), Y11 = bcp(y1), Y01 = mkp(y0, y1);
          const float2 C00 = bcp(c0), C11 = bcp(c1), C01 = mkp(c0, c1);
          sq = fmaf(c0 * y0, y0, fmaf(c1 * y1, y1, sq));
          float2 wm = bcp(0.f);
#pragma unroll
          for (int q = 0; q < NPT; ++q) {
            const bool lo1 = 2 * q >= LM0, hi1 = 2 * q + 1 >= LM0;
            const float2 pp = mkp(__uint_as_float(cur[2 * q]), __uint_as_float(cur[2 * q + 1]));
            const float2 wv = mul2(lo1 ? C11 : (hi1 ? C01 : C00), pp);
            acc_add2(sw[q], wv);
            acc_fma2(swy[q], wv, lo1 ? Y11 : (hi1 ? Y01 : Y00));
            if (MISSING) acc_fma2(wm, wv, mup[q]);
          }
          if (MISSING && ((anyring >> (4 * qq + c)) & 1u)) {
            // logistic missingness (Model.py:269-274) and the d/dy chain into the imputation parameters; lanes
            // without a missing entry add 0.  sum_l w_l d_l = c0 y0' + c1 y1' - sum_l w_l mu'_l
            const bool miss = (missring >> (4 * qq + c)) & 1u;
            const float mf = miss ? 1.f : 0.f;
            const float e = miss ? (x - vm) * inv_sd : 0.f;
            const float swd = fmaf(c0, y0, c1 * y1) - (lo2(wm) + hi2(wm));
            const float uu = fmaf(fmaf(b2, x, b1), x, b0);
            const float tm = fast_ex2(-fabsf(uu) * kLog2e);
            const float rq = fast_rcp(1.f + tm);
            const float ls = -(fmaxf(-uu, 0.f) + kLn2 * fast_lg2(1.f + tm));
            const float sneg = (uu > 0.f ? tm : 1.f) * rq;   // 1 - sigmoid(u)
            float dy = -swd * inv_sig2 + fac * sneg * fmaf(2.f * b2, x, b1);
            if (NOISY) dy = fmaf(-fos[c] * inv_s2n, x, dy);
            dy = miss ? dy : 0.f;
            lmiss_l = fmaf(ls, mf, lmiss_l);
            dvm += dy;
            dvls = fmaf(dy * sd, e, dvls);
            nm += mf;
            lqy_l = fmaf(fmaf(-0.5f * e, e, -vls - kHalfLog2Pi), mf, lqy_l);
          }
        }
        // this warp is done with the y rows of the quarter: the slot goes back to the producer
        if (MISSING) fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(kBYFree + (Qb & 7)));
        UTRACE(4, Qb);
      }
    }
    // ---- this warp's share of the record (sections owned by the mixture warps)
    float swdf[2 * NPT];
    float swdd_l = sq;
#pragma unroll
    for (int x = 0; x < 2 * NPT; ++x) {
      const float m = (x & 1) ? hi2(mup[x / 2]) : lo2(mup[x / 2]);
      const float a = (x & 1) ? hi2(sw[x / 2]) : lo2(sw[x / 2]);
      const float b = (x & 1) ? hi2(swy[x / 2]) : lo2(swy[x / 2]);
      const float dsum = fmaf(-m, a, b);
      swdd_l -= m * (b + dsum);
      swdf[x] = warp_sum(dsum);
    }
    const float swdd_w = warp_sum(swdd_l);
    lqy_l = warp_sum(lqy_l);
    lmiss_l = warp_sum(lmiss_l);
    tc_fence_before();
    __syncthreads();            // (A) every role is done with the tiles: the records may overwrite them
    for (int t = lane; t < pl.total; t += 32) mine[t] = 0.f;
    __syncwarp();
    if (ok) {
#pragma unroll
      for (int x = 0; x < NC; ++x) {
        const float val = (x & 1) ? hi2(sw[x / 2]) : lo2(sw[x / 2]);
        if (x < LM0) { if (x < L0) mine[pl.sw0 + x * J + lane] = val; }
        else if (x - LM0 < L1) mine[pl.sw1 + (x - LM0) * J + lane] = val;
      }
      mine[pl.dvm + lane] = dvm;
      mine[pl.dvls + lane] = dvls;
      mine[pl.nm + lane] = nm;
    }
    if (lane == 0) {
#pragma unroll
      for (int x = 0; x < NC; ++x) {
        if (x < LM0) { if (x < L0) mine[pl.swd0 + x] = swdf[x]; }
        else if (x - LM0 < L1) mine[pl.swd1 + (x - LM0)] = swdf[x];
      }
      mine[pl.sc + 0] = swdd_w;
      double* md = reinterpret_cast<double*>(mine + pl.dbl);
      md[0] = (double)fac * (double)lmiss_l;
      md[1] = (double)fac * (double)p.scale * (double)lqy_l;
    }
  } else if (warp < 12) {
    // =========================== CELL warps: thread = cell (TMEM lane), one quadrant each ===========================
    const int q = warp - 8;
    float gW[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) gW[k] = 0.f;
    float sfr = 0.f, sfo = 0.f;
    double ll2 = 0.0;
    float l1me2 = 0.f, zconst2 = 0.f, ninv2s2 = 0.f;
    if (NOISY) {
      const float e_i = G[gl.eps + i];
      l1me2 = log1pf(-e_i) * kLog2e;
      zconst2 = ((float)J * (-kHalfLog2Pi - logf(p.noisy_sd)) + logf(e_i)) * kLog2e;
      ninv2s2 = -0.5f / (p.noisy_sd * p.noisy_sd) * kLog2e;
    }
    const uint32_t tm_lane = (uint32_t)(32 * q) << 16;
    const uint32_t myrow = (uint32_t)(32 * q + lane);
    const int jchunks = (J + 3) >> 2;
    uint32_t gtoff[8];            // transposed g blocks: element (feature k, cell = lane) inside row k, by k % 8
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) gtoff[kk] = (uint32_t)((((lane >> 2) ^ kk) << 4) | ((lane & 3) << 2));
    for (int Q = q; Q < NQ && good; Q += 4) {
      const int fb = Q & 1;
      good = uwait(bar(kBF + q), (uint32_t)(Q >> 2) & 1u, abort_flag) && good;
      good = uwait(bar(kBD + q), (uint32_t)(Q >> 2) & 1u, abort_flag) && good;      // (complete already: orders the y rows)
      if (!good) break;
      UTRACE(5, Q);
      tc_fence_after();
      uint32_t fr[32];
      tmem_ld32(tmem_base + tm_lane + kTmF + 32u * (uint32_t)fb, fr);
      tmem_wait_ld<32>(fr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBFFree + fb));
      UTRACE(6, Q);
      // sum_j y_j^2 of this thread's cell from its y row (chunk order rotated by the row: conflict-free)
      float y2s = 0.f;
      if (NOISY) {
        const uint32_t yrow = sbase + kUOffY + (uint32_t)(Q & 7) * 4096u + (uint32_t)lane * 128u;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int ch = (c + lane) & 7;
          if (ch < jchunks) {
            const float4 v = lds_f128(yrow + 16u * (uint32_t)ch);
            y2s = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, y2s))));
          }
        }
      }
      // f_k + log W_k, logsumexp over K (Model.py:252-255)
      float f[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 lw = lds_f128(sbase + kUOffScal + 1024u + 16u * c);
        f[4 * c] = __uint_as_float(fr[4 * c]) + lw.x;
        f[4 * c + 1] = __uint_as_float(fr[4 * c + 1]) + lw.y;
        f[4 * c + 2] = __uint_as_float(fr[4 * c + 2]) + lw.z;
        f[4 * c + 3] = __uint_as_float(fr[4 * c + 3]) + lw.w;
      }
      float mx = max3(f[0], f[1], f[2]);
#pragma unroll
      for (int k = 3; k + 1 < 32; k += 2) mx = max3(mx, f[k], f[k + 1]);
      mx = fmaxf(mx, f[31]);
      float ssum = 0.f;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        f[k] = fast_ex2(f[k] - mx);
        ssum += f[k];
      }
      const bool valid = Q * 32 + lane < ncell;
      const float facc = valid ? fac : 0.f;
      const float lse2 = mx + fast_lg2(ssum);
      const float rinv = fast_rcp(ssum);
      float rho = 1.f, omr = 0.f, L2 = lse2;
      if (NOISY) {      // Model.py:259-264
        const float q2 = lse2 + l1me2;
        const float z2 = fmaf(ninv2s2, y2s, zconst2);
        const float dd = q2 - z2;
        const float tq = fast_ex2(-fabsf(dd));
        const float r1 = fast_rcp(1.f + tq);
        L2 = fmaxf(q2, z2) + fast_lg2(1.f + tq);
        rho = dd >= 0.f ? r1 : tq * r1;
        omr = dd >= 0.f ? tq * r1 : r1;
      }
      if (valid) ll2 += (double)L2;
      const float frv = facc * rho, fov = facc * omr, sc = frv * rinv;
      sfr += frv;
      sfo += fov;
      // the g blocks of slot Q & 1 were last read by the MMAs of quarter Q - 2 (c1^T, then dZ: in order)
      if (Q >= 2) good = uwait(bar(kBDZ + ((Q - 2) & 3)), (uint32_t)((Q - 2) >> 2) & 1u, abort_flag) && good;
      if (!good) break;
      // g_k = fac rho softmax_k(f): row of the row-major g blocks (hi / lo), 16-byte chunks swizzled by the row,
      // and column `lane` of the transposed blocks (row = feature)
      const uint32_t grow = sbase + kUOffGR + (uint32_t)fb * 4096u + (uint32_t)lane * 128u;
      const uint32_t gt = sbase + kUOffGT + (uint32_t)fb * 8192u;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int k = 4 * c + u;
          const float gk = sc * f[k];
          gW[k] += gk;
          split_tf32(gk, h[u], l[u]);
          sts_u32(gt + (uint32_t)(k * 128) + gtoff[k & 7], h[u]);
          sts_u32(gt + 4096u + (uint32_t)(k * 128) + gtoff[k & 7], l[u]);
        }
        const uint32_t o = grow + (uint32_t)((c ^ (lane & 7)) << 4);
        sts_u128(o, h[0], h[1], h[2], h[3]);
        sts_u128(o + 8192u, l[0], l[1], l[2], l[3]);
      }
      sts_f32(sbase + kUOffScal + 4u * myrow, frv);
      sts_f32(sbase + kUOffScal + 512u + 4u * myrow, fov);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(kBG + q));
      UTRACE(7, Q);
    }
    tc_fence_before();
    __syncthreads();            // (A)
    tc_fence_after();
    for (int t = lane; t < pl.total; t += 32) mine[t] = 0.f;
    __syncwarp();
    // sum_n g_k over this warp's cells
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float s = warp_sum(gW[k]);
      if (lane == 0 && k < K) mine[pl.gW + k] = s;
    }
    sfr = warp_sum(sfr);
    sfo = warp_sum(sfo);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ll2 += __shfl_xor_sync(FULL, ll2, o);
    if (lane == 0) {
      mine[pl.sc + 1] = sfr;
      mine[pl.sc + 2] = sfo;
      reinterpret_cast<double*>(mine + pl.dbl)[0] = (double)fac * ll2 * 0.6931471805599453;
    }
    // dZ: TMEM lane 32 t + j holds sum_n kind_t[n][j] g[n][k] (t = A1_hi, A1_lo, N0_hi, N0_lo): this warp owns kind q
    if (NQ > 0 && !*abort_flag) {
      uint32_t dz[32];
      tmem_ld32(tmem_base + tm_lane + kTmDZ, dz);
      tmem_wait_ld<32>(dz);
      if (lane < J) {
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (k < K) mine[pl.dZ + k * J + lane] = __uint_as_float(dz[k]);
      }
    }
  } else {
    reg_dec<40>();
    if (warp == 12) {
      // =========================== PRODUCER: y rows by bulk copy, one quarter (32 cells) at a time ===========================
      const int32_t* __restrict__ idx = p.idx + (HAS_IDX ? (cb + c_lo) : 0);
      const float* __restrict__ ysamp = p.y + (size_t)row_base * J;
      const bool whole = !HAS_IDX && J == 32;          // a quarter is one contiguous 4 KB block of the matrix
      RowSampler rsm;                                  // CYTOF_FLAG_SAMPLE_IN_KERNEL
      if (HAS_IDX && p.sampler_on) rsm = make_row_sampler(p, i);
      const uint32_t rowbytes = 4u * (uint32_t)J;
      for (int Q = 0; Q < NQ && good; ++Q) {
        const int slot = Q & 7;
        if (Q >= 8) good = uwait(bar(kBYFree + slot), (uint32_t)((Q >> 3) - 1) & 1u, abort_flag) && good;
        if (!good) break;
        const int nrows = min(32, ncell - 32 * Q);
        const uint32_t dst = sbase + kUOffY + (uint32_t)slot * 4096u;
        // rows past the end of the range are zero-filled (generic stores to bytes the copy never touches)
        for (int r = nrows + (lane >> 3); r < 32; r += 4) sts_u128(dst + 128u * r + 16u * (lane & 7), 0u, 0u, 0u, 0u);
        __syncwarp();
        if (lane == 0) mbar_expect_tx(bar(kBYFull + slot), (uint32_t)nrows * rowbytes);
        __syncwarp();
        UTRACE(12, Q);
        if (whole) {
          if (lane == 0) bulk_g2s(dst, ysamp + (size_t)((int)c_lo + 32 * Q) * 32, (uint32_t)nrows * 128u, bar(kBYFull + slot));
        } else if (lane < nrows) {
          const int row = !HAS_IDX ? (int)c_lo + 32 * Q + lane
                                   : (p.sampler_on ? rsm.row(c_lo + 32 * Q + lane) : __ldg(idx + 32 * Q + lane));
          bulk_g2s(dst + 128u * lane, ysamp + (size_t)row * J, rowbytes, bar(kBYFull + slot));
        }
      }
    } else if (warp == 13) {
      // =========================== MMA issuer 1: f = A1 . Z + N0 . (Z - 1) of the quarter ===========================
      // M = 128 always reads 128 rows: the start address is moved back by (Q & 3) blocks so that the quarter's 32 rows
      // are rows 32 (Q & 3) .. of the instruction, i.e. land in the TMEM quadrant of cell warp Q & 3; the other 96
      // rows are whatever neighbours the block (finite or not: rows do not mix)
      if (lane == 0) {
        for (int Q = 0; Q < NQ && good; ++Q) {
          good = uwait(bar(kBD + (Q & 3)), (uint32_t)(Q >> 2) & 1u, abort_flag) && good;
          if (Q >= 2) good = uwait(bar(kBFFree + (Q & 1)), (uint32_t)((Q >> 1) - 1) & 1u, abort_flag) && good;
          if (!good) break;
          UTRACE(8, Q);
          tc_fence_after();
          const uint32_t d = tmem_base + kTmF + 32u * (uint32_t)(Q & 1);
          const uint32_t a0 = sbase + kUOffAR + (uint32_t)(Q & 1) * 4096u - (uint32_t)(Q & 3) * 4096u;
#pragma unroll
          for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_issue(d, udesc_sw128(a0 + 8192u * t + 32u * ks), udesc_sw128(sbase + (t < 2 ? kUOffWZ : kUOffWZm1) + 32u * ks),
                         (t | ks) ? 1u : 0u);
          umma_commit_to(bar(kBF + (Q & 3)));
          UTRACE(9, Q);
        }
      }
    } else if (warp == 14) {
      // ============ MMA issuer 2: c1^T = Zrep . g^T of the quarter, then dZ += (A1 | N0)^T . g over its 32 cells ============
      if (lane == 0) {
        for (int Q = 0; Q < NQ && good; ++Q) {
          const int qq = Q & 3, cs = Q % 3, gs = Q & 1;
          good = uwait(bar(kBG + qq), (uint32_t)(Q >> 2) & 1u, abort_flag) && good;
          if (Q >= 3) good = uwait(bar(kBC1Free + cs), (uint32_t)(Q / 3 - 1) & 1u, abort_flag) && good;
          if (!good) break;
          UTRACE(10, Q);
          tc_fence_after();
          const uint32_t dc = tmem_base + kTmC1 + 32u * (uint32_t)cs;
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_issue(dc, udesc_sw128(sbase + kUOffZrep + 32u * ks), udesc_sw128(sbase + kUOffGR + 8192u * t + 4096u * gs + 32u * ks),
                         (t | ks) ? 1u : 0u);
          umma_commit_to(bar(kBC1 + cs));
#pragma unroll
          for (int t = 0; t < 2; ++t)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_issue(tmem_base + kTmDZ, udesc_sw128(sbase + kUOffAT + 16384u * qq + 32u * ks),
                         udesc_sw128(sbase + kUOffGT + 8192u * gs + 4096u * t + 32u * ks), (Q | t | ks) ? 1u : 0u);
          umma_commit_to(bar(kBDZ + qq));
          UTRACE(11, Q);
        }
        // the last commit covers every MMA of this thread: the accumulators are final
        if (NQ > 0 && good) good = uwait(bar(kBDZ + ((NQ - 1) & 3)), (uint32_t)((NQ - 1) >> 2) & 1u, abort_flag) && good;
      }
    }
    tc_fence_before();
    __syncthreads();            // (A)
  }

  // ---------------- CTA reduction of the 12 records (fixed order) ----------------
  __syncthreads();              // (B) records complete
  UTRACE(13, 0);
  UTRACE_DUMP;
  const bool bad = *abort_flag != 0;
  float* out = p.partials + (size_t)blockIdx.x * pl.total;
  for (int t = threadIdx.x; t < pl.total; t += kUThreads) {
    if (t >= pl.dbl && t < pl.dbl + 4) continue;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 12; ++w) a += sred[w * pl.total + t];
    out[t] = bad ? __int_as_float(0x7fc00000) : a;
  }
  if (threadIdx.x < 2) {
    double a = 0.0;
    for (int w = 0; w < 12; ++w) a += reinterpret_cast<const double*>(sred + w * pl.total + pl.dbl)[threadIdx.x];
    reinterpret_cast<double*>(out + pl.dbl)[threadIdx.x] = bad ? __longlong_as_double(0x7ff8000000000000ll) : a;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
}

}  // namespace cytof
