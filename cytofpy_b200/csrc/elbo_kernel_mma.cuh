// Tensor-core variant of the fused ELBO kernel for J <= 32 and K <= 32 (BASELINE cfg1-cfg4).
//
// ncu on the packed-fp32 kernel (profiles/README.md, r1c) showed the path is NOT HBM-bound (DRAM
// 3 %): it is bound by issue slots / the FMA pipe, a third of which went into the three J x K
// contractions, and by latency (one cell in flight per warp).  BASELINE.json's north star asks
// for tensor cores in exactly that case.  This kernel
//   * processes EIGHT cells per warp iteration: the mixture phases keep lane = marker j (so the
//     (j,l)-indexed statistics stay in the lane's registers) and are unrolled over the 8 cells,
//     which gives the FMA / MUFU pipes 8 independent instruction streams per warp;
//   * moves the three contractions of Model.py:246-249 and their backward to warp-level tensor
//     core MMAs (mma.sync m16n8k8, TF32 operands, FP32 accumulate) with the 8 cells as the N (or
//     K) dimension:
//         f^T [k, n] = Z^T[k, j] . D^T[j, n] + log W_k       (A = Z^T constant, B = D tile)
//         c1^T[j, n] = Z  [j, k] . g^T[k, n]                 (A = Z   constant, B = g tile)
//         dZ  [j, k] += D^T[j, n] . g [n, k]                 (contraction over the 8 cells)
//     Z is exactly 0/1 in TF32.  D and g are split x = hi + lo (hi = x rounded to TF32, lo the
//     exact fp32 remainder), and the products hi.Z + lo.Z (hi.hi + hi.lo + lo.hi for dZ) are
//     accumulated in fp32, so the result carries ~22 mantissa bits: the 1e-4 / 1e-5 parity bars
//     hold (single-pass TF32 would not);
//   * streams y with cp.async (LDGSTS) into a per-warp double buffer one tile ahead;
//   * issues the 24 HMMA of the dZ contraction inside the long basic block of the mixture backward / next forward
//     (full-batch kernels without the imputation path), where the tensor pipe is otherwise idle: the contraction stage
//     is a latency chain (ldmatrix -> HMMA chain -> 3 shuffle rounds -> ex2 -> STS -> ldmatrix -> HMMA chain) that only
//     the scheduler's other warp can cover, so every tensor-pipe cycle taken out of it counts (round 2, r2n: -3.8 %).
// Per-cell scalars (logsumexp over K, noisy class) are evaluated in the accumulator layout of the
// first MMA (thread (g,t) owns cells 2t, 2t+1 and features g, g+8, g+16, g+24).
#pragma once
#include <cstdio>
#include "common.cuh"

namespace cytof {

constexpr int kTile = 8;                 // cells per warp iteration (= N of the MMAs)
constexpr int kTS = 36;                  // row stride of the per-warp tiles: conflict-free ldmatrix
constexpr int kTileFloats = kTile * kTS;
// per-warp shared memory: D, A0/c1, y^2, g tiles, 1/S0 and 1/S1, per-cell scalars, y double buffer
constexpr int kMmaWarpFloats = 4 * kTileFloats + 2 * kTile * 32 + 32 + 2 * kTile * 32 + 12 * 32;   // sW: 12 rows (Philox staging)
constexpr int kMmaZFloats = 2 * 8 * 32 * 4;     // A-fragment tables of Z^T and Z
// A CTA's warps form kMmaSplit GROUPS; a group is what the host's grid plan calls a block: it owns a cell range of ONE
// sample and writes one partial record.  Two groups of 4 warps per CTA halve the granularity of the split over the
// samples (cfg4: 296 groups = 37 per sample exactly, where 148 CTAs split 19 / 18 and the slowest CTA carries 2.7 %
// more cells than the mean).  MEASURED SLOWER (profiles/README.md, r2n: cfg4 0.2753 vs 0.2693 ms, 125 k cells 54.3 vs
// 50.2 us, 6,000 cells 21.5 vs 19.5 us -- twice the records for the finalise kernel outweigh the balance), so the
// default is one group = the whole CTA.
#ifndef CYTOF_MMA_SPLIT
#define CYTOF_MMA_SPLIT 1
#endif
constexpr int kMmaSplit = CYTOF_MMA_SPLIT;
constexpr int kGroupWarps = kWarps / kMmaSplit;
static_assert(kGroupWarps * kMmaSplit == kWarps && kGroupWarps >= 1, "groups must tile the CTA");


// Packed fp32 through the CUDA 12.9 sm_100 intrinsics (FFMA2 / FADD2 / FMUL2 on float2): unlike
// inline-PTX f32x2 pairs the compiler sees the two halves, so no MOVs are spent on packing and
// unpacking register pairs.
__device__ __forceinline__ float2 mkp(float lo, float hi) { return make_float2(lo, hi); }
__device__ __forceinline__ float2 bcp(float a) { return make_float2(a, a); }
__device__ __forceinline__ float lo2(float2 a) { return a.x; }
__device__ __forceinline__ float hi2(float2 a) { return a.y; }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ void acc_add2(float2& acc, float2 x) { acc = __fadd2_rn(acc, x); }
__device__ __forceinline__ void acc_fma2(float2& acc, float2 a, float2 b) { acc = __ffma2_rn(a, b, acc); }
__device__ __forceinline__ float2 shfl_xor2(float2 a, int o) {
  return make_float2(__shfl_xor_sync(0xffffffffu, a.x, o), __shfl_xor_sync(0xffffffffu, a.y, o));
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint4 a, const uint32_t b0, const uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}
// x = hi + lo with hi = x rounded to TF32 (round half away) and lo the exact remainder, truncated
// to TF32 (the truncation error is <= 2^-21 |x| and follows the sign of lo, i.e. it is unbiased)
__device__ __forceinline__ void split_tf32(const float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) & 0xffffe000u;
}
// split for operands the tensor core reads from shared memory (tcgen05 kind::tf32 uses the upper 19
// bits of each fp32 word): hi = x truncated to TF32, lo = the exact remainder, stored as is
__device__ __forceinline__ void split_trunc(const float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  lo = x - hi;
}
// four 8x4 fp32 blocks (ldmatrix on 8x8 b16 tiles): thread (g,t) receives element [g][4m + t] of
// block m; every lane passes the address of row (lane & 7) of block (lane >> 3)
__device__ __forceinline__ void ldsm_x4(const float* row_ptr, uint32_t (&r)[4]) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(row_ptr);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
// ---- tcgen05 / TMEM helpers (dZ accumulation; layouts validated with tools/tcgen05_probe.cu) ----
// K-major, no-swizzle canonical operand tile of a kind::tf32 MMA with K = 8: element (row m, k) lives at
//   (m / 8) * 256 + (k / 4) * 128 + (m % 8) * 16 + (k % 4) * 4  bytes   (SBO = 256, LBO = 128)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc_kmajor(const void* tile) {
  uint64_t d = (uint64_t)((smem_u32(tile) >> 4) & 0x3fff);
  d |= (uint64_t)(128 >> 4) << 16;      // leading byte offset: between the two k halves
  d |= (uint64_t)(256 >> 4) << 32;      // stride byte offset: between groups of 8 rows
  d |= (uint64_t)1 << 46;               // descriptor version (Blackwell); no swizzle
  return d;
}
// c = F32, a = b = TF32, both K-major, N = 64, M = 64
constexpr uint32_t kUmmaIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((64u >> 4) << 24);
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               :: "r"(tmem_d), "l"(da), "l"(db), "r"(kUmmaIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned long long* mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(mbar)) : "memory");
}
// bounded wait for the phase with parity `par` of an mbarrier (never hang the GPU: gives up after ~1 s)
__device__ __forceinline__ bool mbar_wait(unsigned long long* mbar, uint32_t par) {
  uint32_t done = 0;
  for (int it = 0; it < (1 << 18) && !done; ++it)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(done) : "r"(smem_u32(mbar)), "r"(par) : "memory");
  return done != 0;
}

__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// running maxima of the two mixtures over packed pair q of the flat component list
template <int LM0, int NC>
__device__ __forceinline__ void constexpr_pair_max(const int q, const float2 v, float& M0, float& M1) {
  const bool lo1 = 2 * q >= LM0, hi1 = 2 * q + 1 >= LM0, hi_live = 2 * q + 1 < NC;
  if (!hi_live) { if (lo1) M1 = fmaxf(M1, lo2(v)); else M0 = fmaxf(M0, lo2(v)); }
  else if (lo1) M1 = max3(M1, lo2(v), hi2(v));
  else if (!hi1) M0 = max3(M0, lo2(v), hi2(v));
  else { M0 = fmaxf(M0, lo2(v)); M1 = fmaxf(M1, hi2(v)); }
}

#ifndef CYTOF_DZ_FLOAT
#define CYTOF_DZ_FLOAT 1   // 1: the dZ MMAs are issued inside the backward / next-forward block (full-batch kernels without the imputation path); 2: in all kernels; 0: at the end of the contraction stage
#endif
#ifndef CYTOF_A2_SKEW
#define CYTOF_A2_SKEW 2    // the log / reciprocal half of the forward runs this many cells behind the ex2 half
#endif
// maxima of mixture 0 (components [0, LM0)) and mixture 1 ([LM0, NC)) of the packed list v, max3 over runs of the flat list
template <int LM0, int NC, int NPT>
__device__ __forceinline__ void flat_max(const float2 (&v)[NPT], float& M0, float& M1) {
  auto at = [&](const int x) -> float { return (x & 1) ? hi2(v[x >> 1]) : lo2(v[x >> 1]); };
#pragma unroll
  for (int z = 0; z < 2; ++z) {
    const int a = z ? LM0 : 0, b = z ? NC : LM0;
    float m = at(a);
#pragma unroll
    for (int x = a + 1; x < b; x += 2) m = (x + 1 < b) ? max3(m, at(x), at(x + 1)) : fmaxf(m, at(x));
    (z ? M1 : M0) = m;
  }
}


// MISSING = false is the caller's promise (cytof_desc.flags & CYTOF_FLAG_NO_MISSING) that y holds
// no NaN: the imputation pre-pass and the d/dy chain are compiled out.
#ifndef CYTOF_EARLY_IDX
#define CYTOF_EARLY_IDX 1   // gathered minibatches with the imputation path: early request measured -2 us at 6,000 cells, neutral at 800 k
#endif
template <int LM0, int LM1, bool NOISY, bool HAS_IDX, bool MISSING>
__global__ void __launch_bounds__(kThreads, 1)
elbo_fwdbwd_mma_kernel(const ElboParams p) {
  pdl_enter();
#ifdef CYTOF_STAMPS
  unsigned long long stamp_t[6];
  auto stamp_now = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
  stamp_t[0] = stamp_now();
#endif
  constexpr int NC = LM0 + LM1, NPT = (NC + 1) / 2;          // components, packed pairs
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = kMmaSplit == 1 ? 0 : warp / kGroupWarps, wg = warp - grp * kGroupWarps;     // record group of the warp, warp within it
  const int vb = (int)blockIdx.x * kMmaSplit + grp;                       // the group's block of the grid plan
  const bool live = kMmaSplit == 1 || vb < p.nblocks;                     // an odd plan leaves the last group idle
  const int g = lane >> 2, t4 = lane & 3;                   // MMA fragment coordinates
  uint4* sZT = reinterpret_cast<uint4*>(smem);              // [mt][c][lane] A fragments of Z^T (k x j)
  uint4* sZ = sZT + 8 * 32;                                 // [mt][c][lane] A fragments of Z   (j x k)
  float* wbase = smem + kMmaZFloats + warp * kMmaWarpFloats;     // kMmaZFloats includes the CTA header
  float* sD = wbase;                       // [8][kTS] D_j = A1_j - A0_j
  float* sA = sD + kTileFloats;            // [8][kTS] A0_j, later c1_j
  float* sY = sA + kTileFloats;            // [8][kTS] y_j^2 (noisy class)
  float* sG = sY + kTileFloats;            // [8][kTS] g_k
  float* sR0 = sG + kTileFloats;           // [8][32] 1/S0
  float* sR1 = sR0 + kTile * 32;           // [8][32] 1/S1
  float* sS = sR1 + kTile * 32;            // [0..7] fac*rho, [8..15] fac*(1-rho), [16..23] fac*rho/sum_k e_k per cell
  float* sYb = sS + 32;                    // [2][8][32] y rows, double buffered
  float* sW = sYb + 2 * kTile * 32;        // [8][32] sum_l w d (imputation gradient); [12][32] normals during the pre-pass
  float* sred = smem;                      // final reduction aliases everything (after the loop)

  const int I = p.I, J = p.J, K = p.K, L0 = p.L0, L1 = p.L1;
  int i = 0;
  const int vbc = live ? vb : p.nblocks - 1;
  while (i + 1 < I && vbc >= p.blk_begin[i + 1]) ++i;
  const int nb = p.blk_begin[i + 1] - p.blk_begin[i];
  const int lb = vbc - p.blk_begin[i];
  const long long cb = p.cell_begin[i];
  const long long Bi = p.cell_begin[i + 1] - cb;
  const long long row_base = p.row_base[i];
  const unsigned long long grow_base = (unsigned long long)p.row_global[i];
  // this CTA's cells [c_lo, c_hi) of the sample.  Full batch: the cut points are moved down to a global row
  // that is a multiple of 8, so that the 8-cell tiles coincide with two 4-row Philox blocks (pre-pass)
  auto cut = [&](const long long q) -> long long {
    long long x = Bi * q / nb;
    if (q > 0 && q < nb) x -= (long long)((grow_base + (unsigned long long)x) & 7ull);
    return x < 0 ? 0 : x;
  };
  const long long c_lo = cut(lb), c_hi = cut(lb + 1);

  const float* __restrict__ G = p.globals;
  const bool ok = lane < J;
  // ---------------- L2 prefetch of what the prologue reads: this sample's slices of the globals and the Z masks
  // (one request per 128-byte line, all in flight before anything waits), then the first y tile of every warp.
  // Measured (B200, kernel + finalise, L2 flushed): 52.2 -> 50.2 us at 125 k cells, 265.3 -> 262.7 us at 1 M.  The
  // full-batch variants WITH the imputation path keep the request next to its wait: hoisting it there changed the
  // register allocation of the main loop (20 -> 40 bytes of spills) and cost 3.6 % at 30 % NaN.
  constexpr bool EARLY = !MISSING || (CYTOF_EARLY_IDX && HAS_IDX);
  // (the kernels with the imputation path and the gathered ones spill when the MMAs move: r2n / r2u in profiles/README.md)
  constexpr bool DZ_FLOAT = CYTOF_DZ_FLOAT == 2 || (CYTOF_DZ_FLOAT == 1 && !MISSING && !HAS_IDX);
  constexpr int SKEW = CYTOF_A2_SKEW;
  if (EARLY) {
    const GlobalsLayout glp = globals_layout(I, J, K, L0, L1);
    for (int piece = wg; piece < 8; piece += kGroupWarps) {     // the group's warps share the 8 slices of its sample
      const float* base = G;
      int nfl = L0 + L1 + I;
      switch (piece) {
        case 1: base = G + glp.le0 + i * J * L0; nfl = J * L0; break;
        case 2: base = G + glp.le1 + i * J * L1; nfl = J * L1; break;
        case 3: base = G + glp.lw + i * K; nfl = K; break;
        case 4: base = G + glp.eps; nfl = 4 * I; break;
        case 5: base = G + glp.vm + i * J; nfl = J; break;
        case 6: base = G + glp.vls + i * J; nfl = J; break;
        case 7: base = reinterpret_cast<const float*>(p.zmask); nfl = 2 * J; break;
        default: break;
      }
      const uintptr_t a0 = reinterpret_cast<uintptr_t>(base) & ~(uintptr_t)127, a1 = reinterpret_cast<uintptr_t>(base + nfl);
      for (uintptr_t q = a0 + 128u * lane; q < a1; q += 128u * 32u) asm volatile("prefetch.global.L2 [%0];" :: "l"(q));
    }
  }
  // ---------------- stream the cells of this CTA, 8 per warp iteration ----------------
  const int32_t* __restrict__ idx = p.idx + (HAS_IDX ? (cb + c_lo) : 0);
  const float* __restrict__ ysamp = p.y + (size_t)row_base * J;
  const int ncell = live ? (int)(c_hi - c_lo) : 0;
  const int first_row = (int)c_lo;            // full batch: row = c_lo + cell
  constexpr int STRIDE = kGroupWarps * kTile;
  const bool vec16 = J == 32 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0;   // 128-byte rows

  // y rows of the tile starting at cell tn -> sYb[b]; rows past the end of the range and lanes
  // >= J are zero-filled, so the arithmetic below needs no validity selects
  // gathered batches: lane l fetches the row of cell (l & 7) of the tile -- from idx, or (CYTOF_FLAG_SAMPLE_IN_KERNEL)
  // by evaluating the sampler's keyed bijection at the cell's position -- and the lanes exchange them by shuffle
  RowSampler rsm;
  if (HAS_IDX && p.sampler_on) rsm = make_row_sampler(p, i);
  auto tile_rows = [&](const int tn) -> int {
    const int u8 = lane & 7;
    if (!HAS_IDX || tn + u8 >= ncell) return 0;
    return p.sampler_on ? rsm.row(c_lo + tn + u8) : __ldg(idx + tn + u8);
  };
  auto issue_tile = [&](const int tn, const int b) {
    float* dst = sYb + b * (kTile * 32);
    const int row8 = tile_rows(tn);
    if (vec16) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h, u = c >> 3, part = c & 7;
        const int rowg = HAS_IDX ? __shfl_sync(FULL, row8, u) : 0;
        if (tn + u < ncell) {
          const int row = HAS_IDX ? rowg : first_row + tn + u;
          cp_async16(dst + u * 32 + 4 * part, ysamp + (size_t)row * 32 + 4 * part);
        } else {
          *reinterpret_cast<float4*>(dst + u * 32 + 4 * part) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < kTile; ++u) {
        const int rowg = HAS_IDX ? __shfl_sync(FULL, row8, u) : 0;
        if (ok && tn + u < ncell) {
          const int row = HAS_IDX ? rowg : first_row + tn + u;
          cp_async4(dst + u * 32 + lane, ysamp + (size_t)row * J + lane);
        } else {
          dst[u * 32 + lane] = 0.f;
        }
      }
    }
    cp_async_commit();
  };

  if (EARLY && wg * kTile < ncell) issue_tile(wg * kTile, 0);

  // ---------------- Z fragment tables (once per CTA) ----------------
  {
    const unsigned kmask = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
    for (int e = threadIdx.x; e < 2 * 8 * 32; e += kThreads) {
      const int which = e >> 8, mt = (e >> 7) & 1, c = (e >> 5) & 3, ln = e & 31;
      const int gg = ln >> 2, tt = ln & 3;
      uint32_t v[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int row = 16 * mt + gg + ((r & 1) ? 8 : 0), col = 8 * c + tt + ((r & 2) ? 4 : 0);
        const int j = which ? row : col, k = which ? col : row;   // which = 0: Z^T[k][j], 1: Z[j][k]
        const unsigned zr = j < J ? ((unsigned)p.zmask[j] & kmask) : 0u;
        v[r] = ((zr >> k) & 1u) ? 0x3f800000u : 0u;
      }
      (which ? sZ : sZT)[(mt * 4 + c) * 32 + ln] = make_uint4(v[0], v[1], v[2], v[3]);
    }
  }


  // ---------------- per-sample constants -> registers ----------------
  const GlobalsLayout gl = globals_layout(I, J, K, L0, L1);
  const float sig = G[gl.sig + i];
  const float inv_sig2 = 1.0f / (sig * sig);
  const float nh2 = -0.5f * inv_sig2 * kLog2e;
  const float cbase = (-logf(sig) - kHalfLog2Pi) * kLog2e;
  const float fac = p.fac[i];
  const float b0 = G[gl.b + 3 * i], b1 = G[gl.b + 3 * i + 1], b2 = G[gl.b + 3 * i + 2];
  const float okf = ok ? 1.f : 0.f;
  const int jj = ok ? lane : 0;

  // The LM0 + LM1 mixture components as ONE list (mixture 0 first), processed as packed pairs.
  // With y' = y - m_z (m_z = mid-range of the means of mixture z, keeps |y'| small where the
  // mixture has mass) the log2-density of component l is
  //     v_l = cz_l + nh (y' - mu'_l)^2 = nh y'^2 + [ be_l y' + al_l ],   mu'_l = mu_l - m_z,
  //     be_l = -2 nh mu'_l (lane-uniform),  al_l = cz_l + nh mu'_l^2 (per marker),
  // and nh y'^2 is common to the components of a mixture: it cancels in the softmax over l and
  // in D = A1 - A0 up to the two centres, so the forward costs ONE FFMA2 per component pair.
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
  const float vm = G[gl.vm + i * J + jj], vls = G[gl.vls + i * J + jj], sd = expf(vls);
  // log W (log2 domain) of the four feature rows this thread owns in the accumulator layout
  float lwk[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int k = g + 8 * r;
    lwk[r] = k < K ? G[gl.lw + i * K + k] * kLog2e : kNegBig;
  }
  float l1me2 = 0.f, zconst2 = 0.f, ninv2s2 = 0.f, inv_s2n = 0.f;
  if (NOISY) {
    const float e_i = G[gl.eps + i];
    l1me2 = log1pf(-e_i) * kLog2e;
    zconst2 = ((float)J * (-kHalfLog2Pi - logf(p.noisy_sd)) + logf(e_i)) * kLog2e;
    inv_s2n = 1.0f / (p.noisy_sd * p.noisy_sd);
    ninv2s2 = -0.5f * inv_s2n * kLog2e;
  }

  // ---------------- accumulators ----------------
  // sum_n w, sum_n w y' per component (sum w d and sum w d^2 follow in the epilogue) and
  // sum_n (c0 y0'^2 + c1 y1'^2)
  float2 sw[NPT], swy[NPT];
  float sq = 0.f;
  float dZ[2][4][4];                     // C fragments of dZ (permuted rows / columns, see epilogue)
  float gW[4] = {0.f, 0.f, 0.f, 0.f};    // sum_n g_k for k = g, g+8, g+16, g+24 (this thread's 2 cells)
  float dvm = 0.f, dvls = 0.f, nm = 0.f, sfr = 0.f, sfo = 0.f, lqy_l = 0.f, lmiss_l = 0.f;
  double ll2 = 0.0;
#pragma unroll
  for (int q = 0; q < NPT; ++q) sw[q] = swy[q] = bcp(0.f);
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) dZ[a][b][c] = 0.f;

  // imputation of missing entries (vae.py:17-33): the only branchy part, kept apart so that the
  // mixture arithmetic is one long basic block.  The N(0,1) draw replaces the NaN in the y buffer.
  auto prepass = [&](const int tn, const int b, unsigned& mmask, unsigned& amask) {
    float* src = sYb + b * (kTile * 32);
    mmask = 0u;
#pragma unroll
    for (int u = 0; u < kTile; ++u) {
      const float yr = src[u * 32 + lane];
      mmask |= (yr != yr) ? (1u << u) : 0u;
    }
    amask = __reduce_or_sync(FULL, mmask);       // cells of the tile with at least one missing marker
    if (amask && p.noise_mode != CYTOF_NOISE_EXTERNAL) {
      // The draws are keyed by the POSITION of the cell in the sample's batch (= its row for full batches; for a
      // gathered minibatch the reference draws a dense (B_i, J) normal per step as well, vae.py:33), so the tile lies
      // in 2 (aligned) or 3 Philox blocks of 4 positions.  Their normals go to a [12][32] staging area
      // (row = position - 4 grp0) and the NaNs are replaced by plain copies.
      const unsigned long long grow0 = grow_base + (unsigned long long)(first_row + tn);
      const unsigned a = (unsigned)grow0 & 3u;
      const unsigned long long grp0 = grow0 >> 2;
      const int ng = (int)((a + kTile + 3u) >> 2);
#pragma unroll 1      // a real loop: ONE copy of the Philox rounds in the instruction stream
      for (int gi = 0; gi < ng; ++gi) {
        const int u0 = 4 * gi - (int)a;          // first cell of the block (may be negative)
        const unsigned gm = (u0 >= 0 ? (0xfu << u0) : (0xfu >> -u0)) & 0xffu;
        if (!(amask & gm)) continue;             // warp-uniform: no cell of this block misses anything
        float nz[4];
        philox_normal4(p.seed, elbo_step(p), grp0 + (unsigned long long)gi, philox_ctr(lane, i), nz);
#pragma unroll
        for (int w = 0; w < 4; ++w) sW[(4 * gi + w) * 32 + lane] = nz[w];
      }
      const float* nzr = sW + a * 32 + lane;     // every lane reads back its own column only
#pragma unroll
      for (int u = 0; u < kTile; ++u)
        if ((mmask >> u) & 1u) src[u * 32 + lane] = nzr[u * 32];
    } else if (amask) {      // external noise: eps[cell of the batch][marker]
#pragma unroll
      for (int u = 0; u < kTile; ++u)
        if ((mmask >> u) & 1u) src[u * 32 + lane] = p.eps ? __ldg(p.eps + (cb + c_lo + tn + u) * J + lane) : 0.f;
    }
  };

  float2 ee[kTile][NPT];      // exp(a_l - max) of the tile in flight, lane = marker j
  const float2 nh = bcp(nh2);

  // mixture log-densities of cell u, forward (Model.py:223-233), log2 domain, component pairs.
  // Two halves: a1 issues the 10 ex2 of the cell; a2 (called CYTOF_A2_SKEW = 2 cells later in the main loop, so
  // that the MUFU pipe already works on the next cells' exponentials while the sums of this one
  // settle; r2q: skew 0 / 1 / 2 / 3 = 0.2776 / 0.2549 / 0.2510 / 0.2571 ms per 1 M cells) takes the logs / reciprocal
  // and writes the tiles.  The component maxima run over the FLAT list (flat_max: 4 instead of 6 FMNMX3 per cell at
  // L = 5 / 5, on the chain in front of the ex2: r2p, -3.8 %).
  struct AState { float S0, S1, B0, B1, yy; };
  auto phase_a1 = [&](const int u, const float* src, const unsigned mmask) -> AState {
    const float x = src[u * 32 + lane];
    const float yv = (MISSING && ((mmask >> u) & 1u)) ? fmaf(sd, x, vm) : x;
    const float y0 = yv - mc0, y1 = yv - mc1;
    const float2 Y00 = bcp(y0), Y11 = bcp(y1), Y01 = mkp(y0, y1);
    float2 v[NPT];
#pragma unroll
    for (int q = 0; q < NPT; ++q) {
      const bool lo1 = 2 * q >= LM0, hi1 = 2 * q + 1 >= LM0;
      v[q] = fma2(be[q], lo1 ? Y11 : (hi1 ? Y01 : Y00), al[q]);
    }
    float M0 = kNegBig, M1 = kNegBig;
    flat_max<LM0, NC, NPT>(v, M0, M1);
    const float2 M00 = bcp(M0), M11 = bcp(M1), M01 = mkp(M0, M1);
    float2 s0 = bcp(0.f), s1 = bcp(0.f);
    float s0x = 0.f, s1x = 0.f;
#pragma unroll
    for (int q = 0; q < NPT; ++q) {
      const bool lo1 = 2 * q >= LM0, hi1 = 2 * q + 1 >= LM0, hi_live = 2 * q + 1 < NC;
      const float2 tt = sub2(v[q], lo1 ? M11 : (hi1 ? M01 : M00));
      ee[u][q] = mkp(fast_ex2(lo2(tt)), hi_live ? fast_ex2(hi2(tt)) : 0.f);
      if (lo1) s1 = add2(s1, ee[u][q]);
      else if (!hi1) s0 = add2(s0, ee[u][q]);
      else { s0x = lo2(ee[u][q]); s1x = hi2(ee[u][q]); }
    }
    AState st;
    st.S0 = (lo2(s0) + hi2(s0)) + s0x;
    st.S1 = (lo2(s1) + hi2(s1)) + s1x;
    st.B0 = fmaf(nh2 * y0, y0, M0);      // A_z = B_z + lg2(S_z)
    st.B1 = fmaf(nh2 * y1, y1, M1);
    st.yy = yv * yv;
    return st;
  };
  auto phase_a2 = [&](const int u, const AState& st) {
    const float A0 = st.B0 + fast_lg2(st.S0), A1 = st.B1 + fast_lg2(st.S1);
    const float rr = fast_rcp(st.S0 * st.S1);      // one reciprocal for both mixtures (S in [1, L])
    reinterpret_cast<float2*>(sR0)[u * 32 + lane] = make_float2(rr * st.S1, rr * st.S0);   // 1/S0, 1/S1
    sD[u * kTS + lane] = A1 - A0;           // markers >= J: Z has zero rows there, the value is never used
    sA[u * kTS + lane] = okf * A0;
    if (NOISY) sY[u * kTS + lane] = st.yy;
  };

  // mixtures of cell u, backward: w_zl = c_z p_zl; statistics sum w and sum w y' per component,
  // sum c_z y_z'^2 per cell (sum_l w_zl = c_z) -- three FMA-pipe ops per component pair
  auto phase_b = [&](const int u, const float* src, const unsigned mmask, const unsigned amask) {
    const float c1 = sA[u * kTS + lane];    // = 0 for markers >= J (zero rows of Z)
    const float c0 = okf * sS[u] - c1;
    const float x = src[u * 32 + lane];
    const float yv = (MISSING && ((mmask >> u) & 1u)) ? fmaf(sd, x, vm) : x;
    const float y0 = yv - mc0, y1 = yv - mc1;
    const float2 Y00 = bcp(y0), Y11 = bcp(y1), Y01 = mkp(y0, y1);
    const float2 rs01 = reinterpret_cast<const float2*>(sR0)[u * 32 + lane];
    const float r0 = c0 * rs01.x, r1 = c1 * rs01.y;
    const float2 R00 = bcp(r0), R11 = bcp(r1), R01 = mkp(r0, r1);
    sq = fmaf(c0 * y0, y0, fmaf(c1 * y1, y1, sq));
    float2 wm = bcp(0.f);
#pragma unroll
    for (int q = 0; q < NPT; ++q) {
      const bool lo1 = 2 * q >= LM0, hi1 = 2 * q + 1 >= LM0;
      const float2 w = mul2(lo1 ? R11 : (hi1 ? R01 : R00), ee[u][q]);
      acc_add2(sw[q], w);
      acc_fma2(swy[q], w, lo1 ? Y11 : (hi1 ? Y01 : Y00));
      if (MISSING) acc_fma2(wm, w, mup[q]);
    }
    // sum_l w_l d_l = c0 y0' + c1 y1' - sum_l w_l mu'_l: only the imputation gradient needs it
    if (MISSING && amask) sW[u * 32 + lane] = fmaf(c0, y0, c1 * y1) - (lo2(wm) + hi2(wm));
  };

  // contraction 3: dZ[j, k] += D^T . g over the 8 cells of the tile (tensor cores).  Nothing in the tile's backward
  // depends on it.  CYTOF_DZ_FLOAT = 1 issues it INSIDE the basic block of the mixture backward / next forward, so that
  // its 24 HMMA (252 tensor-pipe cycles per tile) overlap this warp's own FMA / MUFU work instead of stalling the warp.
  struct DzOps { uint4 ah[2], al2[2]; float gl4[4], gh4[4]; };
  auto dz_load = [&](DzOps& o) {       // operands out of this tile's D / e tiles (cross-lane reads: before the tiles are rewritten)
    const float4 d_lo = *reinterpret_cast<const float4*>(sD + t4 * kTS + 4 * g);        // cell t4,   markers 4g..4g+3
    const float4 d_hi = *reinterpret_cast<const float4*>(sD + (t4 + 4) * kTS + 4 * g);  // cell t4+4
    const float4 e_lo = *reinterpret_cast<const float4*>(sG + t4 * kTS + 4 * g);        // cell t4,   features 4g..4g+3
    const float4 e_hi = *reinterpret_cast<const float4*>(sG + (t4 + 4) * kTS + 4 * g);
    const float sca = sS[16 + t4], scb = sS[16 + t4 + 4];
    const float dl[4] = {d_lo.x, d_lo.y, d_lo.z, d_lo.w}, dh[4] = {d_hi.x, d_hi.y, d_hi.z, d_hi.w};
    o.gl4[0] = sca * e_lo.x; o.gl4[1] = sca * e_lo.y; o.gl4[2] = sca * e_lo.z; o.gl4[3] = sca * e_lo.w;
    o.gh4[0] = scb * e_hi.x; o.gh4[1] = scb * e_hi.y; o.gh4[2] = scb * e_hi.z; o.gh4[3] = scb * e_hi.w;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {   // rows g / g+8 of m-tile mt = markers 4g+2mt / 4g+2mt+1
      split_tf32(dl[2 * mt], o.ah[mt].x, o.al2[mt].x);
      split_tf32(dl[2 * mt + 1], o.ah[mt].y, o.al2[mt].y);
      split_tf32(dh[2 * mt], o.ah[mt].z, o.al2[mt].z);
      split_tf32(dh[2 * mt + 1], o.ah[mt].w, o.al2[mt].w);
    }
  };
  auto dz_mma = [&](const DzOps& o) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {   // column g of n-tile nt = feature 4g+nt
      uint32_t b0h, b0l, b1h, b1l;
      split_tf32(o.gl4[nt], b0h, b0l);
      split_tf32(o.gh4[nt], b1h, b1l);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        mma_tf32(dZ[mt][nt], o.ah[mt], b0h, b1h);
        mma_tf32(dZ[mt][nt], o.ah[mt], b0l, b1l);
        mma_tf32(dZ[mt][nt], o.al2[mt], b0h, b1h);
      }
    }
  };

#ifdef CYTOF_PHASE_CLOCKS
  long long clkS = 0, clkBA = 0;
  int ntiles = 0;
#endif
  int t0 = wg * kTile;
  int cur = 0;
  unsigned missmask = 0u, anymask = 0u;
#ifdef CYTOF_STAMPS
  stamp_t[1] = stamp_now();
#endif
  if (t0 < ncell) {       // prologue: phase A of this warp's first tile (MISSING = false: its y rows were requested at the top)
    if (!EARLY) issue_tile(t0, 0);
    cp_async_wait<0>();
    __syncwarp();
    if (MISSING) prepass(t0, 0, missmask, anymask);
    {
      AState prev = phase_a1(0, sYb, missmask);
#pragma unroll
      for (int u = 1; u < kTile; ++u) {
        const AState now = phase_a1(u, sYb, missmask);
        phase_a2(u - 1, prev);
        prev = now;
      }
      phase_a2(kTile - 1, prev);
    }
    if (t0 + STRIDE < ncell) issue_tile(t0 + STRIDE, 1);
  }
  __syncthreads();   // Z tables visible

#ifdef CYTOF_STAMPS
  stamp_t[2] = stamp_now();
#endif
  for (; t0 < ncell; t0 += STRIDE, cur ^= 1) {
    const int nvalid = min(kTile, ncell - t0);
    const bool has_next = t0 + STRIDE < ncell;
    __syncwarp();      // logically redundant (the iteration ends with one), but without it ptxas schedules the stage 6 % slower (profiles/README.md, r2s)
#ifdef CYTOF_PHASE_CLOCKS
    const long long ck0 = clock64();
#endif

    // ================= contraction 1: f^T[k, n] = Z^T . D^T + log W (tensor cores) =================
    uint32_t bh[8], bl[8];      // B fragments of the D tile: block m = markers 4m..4m+3
    {
      uint32_t r0[4], r1[4];
      ldsm_x4(sD + (lane & 7) * kTS + 4 * (lane >> 3), r0);
      ldsm_x4(sD + (lane & 7) * kTS + 16 + 4 * (lane >> 3), r1);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        split_tf32(__uint_as_float(r0[m]), bh[m], bl[m]);
        split_tf32(__uint_as_float(r1[m]), bh[4 + m], bl[4 + m]);
      }
    }
    // this thread owns cells 2*t4, 2*t4+1 and features g, g+8, g+16, g+24 in the accumulator
    // layout; reductions over k (and over the markers below) run over the 8 lanes with equal t4
    const int ca = 2 * t4;
    // sum_j A0_j and sum_j y_j^2 of the two cells: independent of the MMAs, issued first so that
    // their shuffle rounds overlap the accumulator chains
    float2 ra, ry = bcp(0.f);
    {
      const float4 a4 = *reinterpret_cast<const float4*>(sA + ca * kTS + 4 * g);
      const float4 b4 = *reinterpret_cast<const float4*>(sA + (ca + 1) * kTS + 4 * g);
      ra = mkp((a4.x + a4.y) + (a4.z + a4.w), (b4.x + b4.y) + (b4.z + b4.w));
      if (NOISY) {
        const float4 c4 = *reinterpret_cast<const float4*>(sY + ca * kTS + 4 * g);
        const float4 d4 = *reinterpret_cast<const float4*>(sY + (ca + 1) * kTS + 4 * g);
        ry = mkp((c4.x + c4.y) + (c4.z + c4.w), (d4.x + d4.y) + (d4.z + d4.w));
      }
    }
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      ra = add2(ra, shfl_xor2(ra, o));
      if (NOISY) ry = add2(ry, shfl_xor2(ry, o));
    }
    float fT[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      fT[mt][0] = fT[mt][1] = lwk[2 * mt];
      fT[mt][2] = fT[mt][3] = lwk[2 * mt + 1];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint4 a = sZT[(mt * 4 + c) * 32 + lane];
        mma_tf32(fT[mt], a, bh[2 * c], bh[2 * c + 1]);
        mma_tf32(fT[mt], a, bl[2 * c], bl[2 * c + 1]);
      }
    }
    float mxa = fmaxf(fmaxf(fT[0][0], fT[0][2]), fmaxf(fT[1][0], fT[1][2]));
    float mxb = fmaxf(fmaxf(fT[0][1], fT[0][3]), fmaxf(fT[1][1], fT[1][3]));
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      mxa = fmaxf(mxa, __shfl_xor_sync(FULL, mxa, o));
      mxb = fmaxf(mxb, __shfl_xor_sync(FULL, mxb, o));
    }
    // un-normalised softmax numerators e_k = 2^(f_k - max): the second contraction runs on them
    // (c1 = fac rho / sum_k e_k * Z e), so it does not wait for the logsumexp / noisy-class chain
    float ex[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      ex[mt][0] = fast_ex2(fT[mt][0] - mxa);
      ex[mt][1] = fast_ex2(fT[mt][1] - mxb);
      ex[mt][2] = fast_ex2(fT[mt][2] - mxa);
      ex[mt][3] = fast_ex2(fT[mt][3] - mxb);
      sG[ca * kTS + 16 * mt + g] = ex[mt][0];
      sG[(ca + 1) * kTS + 16 * mt + g] = ex[mt][1];
      sG[ca * kTS + 16 * mt + g + 8] = ex[mt][2];
      sG[(ca + 1) * kTS + 16 * mt + g + 8] = ex[mt][3];
    }
    __syncwarp();

    // ================= contraction 2 on the numerators: Z . e^T (tensor cores) =================
    {
      uint32_t r0[4], r1[4];
      ldsm_x4(sG + (lane & 7) * kTS + 4 * (lane >> 3), r0);
      ldsm_x4(sG + (lane & 7) * kTS + 16 + 4 * (lane >> 3), r1);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        split_tf32(__uint_as_float(r0[m]), bh[m], bl[m]);
        split_tf32(__uint_as_float(r1[m]), bh[4 + m], bl[4 + m]);
      }
    }
    float cT[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) cT[mt][0] = cT[mt][1] = cT[mt][2] = cT[mt][3] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint4 a = sZ[(mt * 4 + c) * 32 + lane];
        mma_tf32(cT[mt], a, bh[2 * c], bh[2 * c + 1]);
        mma_tf32(cT[mt], a, bl[2 * c], bl[2 * c + 1]);
      }
    }
    // ---- per-cell scalars (overlap the MMAs above): logsumexp over K, noisy class (Model.py:255-264)
    float2 rs = mkp((ex[0][0] + ex[0][2]) + (ex[1][0] + ex[1][2]), (ex[0][1] + ex[0][3]) + (ex[1][1] + ex[1][3]));
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) rs = add2(rs, shfl_xor2(rs, o));
    float sc2[2], fr2[2], om2[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float mx = h ? mxb : mxa;
      const float ssum = h ? hi2(rs) : lo2(rs);
      const float a0s = h ? hi2(ra) : lo2(ra);
      const bool valid = ca + h < nvalid;
      const float facc = valid ? fac : 0.f;
      const float lse2 = (mx + a0s) + fast_lg2(ssum);
      const float rinv = fast_rcp(ssum);
      float rho = 1.f, omr = 0.f, L2 = lse2;
      if (NOISY) {
        const float y2s = h ? hi2(ry) : lo2(ry);
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
      const float fr = facc * rho;
      fr2[h] = fr;
      om2[h] = facc * omr;
      sc2[h] = fr * rinv;                   // g_k = sc e_k = fac rho softmax_k(f)
      sfr += fr;
      sfo += om2[h];
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      gW[2 * mt] += fmaf(sc2[0], ex[mt][0], sc2[1] * ex[mt][1]);
      gW[2 * mt + 1] += fmaf(sc2[0], ex[mt][2], sc2[1] * ex[mt][3]);
    }
    if (g == 0) {
      sS[ca] = fr2[0];
      sS[ca + 1] = fr2[1];
      sS[8 + ca] = om2[0];
      sS[8 + ca + 1] = om2[1];
      sS[16 + ca] = sc2[0];
      sS[16 + ca + 1] = sc2[1];
    }
    __syncwarp();      // the A0 sums of every lane are taken: c1 may overwrite the tile
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      sA[ca * kTS + 16 * mt + g] = sc2[0] * cT[mt][0];
      sA[(ca + 1) * kTS + 16 * mt + g] = sc2[1] * cT[mt][1];
      sA[ca * kTS + 16 * mt + g + 8] = sc2[0] * cT[mt][2];
      sA[(ca + 1) * kTS + 16 * mt + g + 8] = sc2[1] * cT[mt][3];
    }

    // ================= contraction 3: dZ += D^T . g (tensor cores) =================
    DzOps dzo;
    dz_load(dzo);
    if (!DZ_FLOAT) dz_mma(dzo);     // issued last in the stage: nothing in this tile's backward depends on it
    __syncwarp();


#ifdef CYTOF_PHASE_CLOCKS
    const long long ck1 = clock64();
    clkS += ck1 - ck0;
#endif
    // ===== phase B of this tile, software-pipelined with phase A of the warp's NEXT tile: the
    // backward is FMA-pipe work, the forward is MUFU work, and cell u of the next tile reuses the
    // registers cell u of this tile has just released -- one basic block, both pipes busy
    const float* ycur = sYb + cur * (kTile * 32);
    unsigned miss_n = 0u, any_n = 0u;
    if (has_next) {
      const float* ynxt = sYb + (cur ^ 1) * (kTile * 32);
      cp_async_wait<0>();
      __syncwarp();
      if (MISSING) prepass(t0 + STRIDE, cur ^ 1, miss_n, any_n);
      if (DZ_FLOAT) dz_mma(dzo);
      {
        // cell u: backward of this tile, first forward half of the next tile, second forward half CYTOF_A2_SKEW cells late
        AState st[kTile];
#pragma unroll
        for (int u = 0; u < kTile + SKEW; ++u) {
          if (u < kTile) {
            phase_b(u, ycur, missmask, anymask);
            st[u] = phase_a1(u, ynxt, miss_n);
          }
          if (u >= SKEW) phase_a2(u - SKEW, st[u - SKEW]);
        }
      }
    } else {
      if (DZ_FLOAT) dz_mma(dzo);
#pragma unroll
      for (int u = 0; u < kTile; ++u) phase_b(u, ycur, missmask, anymask);
    }
    // logistic missingness (Model.py:269-274) and d/dy chain into the imputation parameters
    if (MISSING && anymask) {
#pragma unroll
      for (int u = 0; u < kTile; ++u) {     // branch-free: lanes without a missing entry add 0
        const bool miss = (missmask >> u) & 1u;
        const float mf = (miss && u < nvalid) ? 1.f : 0.f;
        const float x = ycur[u * 32 + lane];
        const float e = miss ? x : 0.f;
        const float yv = miss ? fmaf(sd, e, vm) : x;
        const float facc = u < nvalid ? fac : 0.f;
        const float uu = fmaf(fmaf(b2, yv, b1), yv, b0);
        const float tm = fast_ex2(-fabsf(uu) * kLog2e);
        const float rr = fast_rcp(1.f + tm);
        const float ls = -(fmaxf(-uu, 0.f) + kLn2 * fast_lg2(1.f + tm));
        const float sneg = (uu > 0.f ? tm : 1.f) * rr;   // 1 - sigmoid(u)
        float dy = -sW[u * 32 + lane] * inv_sig2 + facc * sneg * fmaf(2.f * b2, yv, b1);
        if (NOISY) dy = fmaf(-sS[8 + u] * inv_s2n, yv, dy);
        dy = miss ? dy : 0.f;
        lmiss_l = fmaf(ls, mf, lmiss_l);
        dvm += dy;
        dvls = fmaf(dy * sd, e, dvls);
        nm += mf;
        lqy_l = fmaf(fmaf(-0.5f * e, e, -vls - kHalfLog2Pi), mf, lqy_l);
      }
    }
    __syncwarp();   // every lane is done with this tile's y buffer and scalars
    if (t0 + 2 * STRIDE < ncell) issue_tile(t0 + 2 * STRIDE, cur);
    missmask = miss_n;
    anymask = any_n;
#ifdef CYTOF_PHASE_CLOCKS
    clkBA += clock64() - ck1;
    ++ntiles;
#endif
  }
#ifdef CYTOF_PHASE_CLOCKS
  if (blockIdx.x == 3 && lane == 0)
    printf("warp %d tiles %d  stage S %lld cyc/tile  B+A %lld cyc/tile\n", warp, ntiles, clkS / max(ntiles, 1), clkBA / max(ntiles, 1));
#endif

#ifdef CYTOF_STAMPS
  stamp_t[3] = stamp_now();
#endif
  // ---------------- CTA reduction (fixed order) and partial record ----------------
  const PartialLayout pl = partial_layout(J, K, L0, L1);
  // sum w d = sum w y' - mu' sum w;  sum w d^2 = sum c y'^2 - sum_l mu'_l (2 sum w y' - mu'_l sum w)
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
  // accumulator-layout sums: features over the 4 lanes of a quad, per-cell scalars over t4 at g == 0
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    gW[r] += __shfl_xor_sync(FULL, gW[r], 1);
    gW[r] += __shfl_xor_sync(FULL, gW[r], 2);
  }
  sfr += __shfl_xor_sync(FULL, sfr, 1);
  sfr += __shfl_xor_sync(FULL, sfr, 2);
  sfo += __shfl_xor_sync(FULL, sfo, 1);
  sfo += __shfl_xor_sync(FULL, sfo, 2);
  ll2 += __shfl_xor_sync(FULL, ll2, 1);
  ll2 += __shfl_xor_sync(FULL, ll2, 2);

  float* mine = sred + warp * pl.total;
  __syncthreads();            // every warp is done with its tiles and with the Z tables
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
  // dZ fragments: (mt, nt, r) -> marker 4g + 2mt + (r >> 1), feature 8 t4 + 4 (r & 1) + nt
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int j = 4 * g + 2 * mt + (r >> 1), k = 8 * t4 + 4 * (r & 1) + nt;
        if (j < J && k < K) mine[pl.dZ + k * J + j] = dZ[mt][nt][r];
      }
  if (t4 == 0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) if (g + 8 * r < K) mine[pl.gW + g + 8 * r] = gW[r];
  }
  if (lane == 0) {
#pragma unroll
    for (int x = 0; x < NC; ++x) {
      if (x < LM0) { if (x < L0) mine[pl.swd0 + x] = swdf[x]; }
      else if (x - LM0 < L1) mine[pl.swd1 + (x - LM0)] = swdf[x];
    }
    mine[pl.sc + 0] = swdd_w;
    mine[pl.sc + 1] = sfr;
    mine[pl.sc + 2] = sfo;
    double* md = reinterpret_cast<double*>(mine + pl.dbl);
    md[0] = (double)fac * (ll2 * 0.6931471805599453 + (double)lmiss_l);
    md[1] = (double)fac * (double)p.scale * (double)lqy_l;
  }
  __syncthreads();
#ifdef CYTOF_STAMPS
  stamp_t[4] = stamp_now();
#endif
  // every group adds the records of its own warps (fixed order) and stores the record of its block
  float* out = p.partials + (size_t)vbc * pl.total;
  const float* gred = sred + grp * kGroupWarps * pl.total;
  const int gtid = (int)threadIdx.x - grp * (kGroupWarps * 32);
  if (live) {
    for (int t = gtid; t < pl.total; t += kGroupWarps * 32) {
      if (t >= pl.dbl && t < pl.dbl + 4) continue;
      float a = 0.f;
#pragma unroll
      for (int w = 0; w < kGroupWarps; ++w) a += gred[w * pl.total + t];
      out[t] = a;
    }
    if (gtid < 2) {
      double a = 0.0;
      for (int w = 0; w < kGroupWarps; ++w) a += reinterpret_cast<const double*>(gred + w * pl.total + pl.dbl)[gtid];
      reinterpret_cast<double*>(out + pl.dbl)[gtid] = a;
    }
  }
#ifdef CYTOF_STAMPS
  stamp_t[5] = stamp_now();
  if ((blockIdx.x == 0 || blockIdx.x == 77) && threadIdx.x == 0)
    printf("cta %d ncell %d: prologue %llu  first tile + phase A %llu  loop %llu  warp sums + record %llu  cta sum + store %llu ns\n",
           (int)blockIdx.x, ncell, stamp_t[1] - stamp_t[0], stamp_t[2] - stamp_t[1], stamp_t[3] - stamp_t[2],
           stamp_t[4] - stamp_t[3], stamp_t[5] - stamp_t[4]);
#endif
}

}  // namespace cytof
