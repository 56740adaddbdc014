// Tensor-core fused ELBO kernel for 32 < J <= 48 markers, K <= 64 features, L <= 8 components
// (BASELINE cfg5: J = 40, K = 50, L = 8/8), for J <= 32 markers with 32 < K <= 64 features (main lanes >= J are
// masked, the tail pass runs empty), and -- NTP = 4, the WIDE layout -- for 48 < J <= 64 with L <= 5 (every cell takes a
// second full pass over markers 32..63).  MISSING = false compiles the imputation path out.
//
// The first SIMT kernel needs 1,767 warp-instructions per cell at these shapes (bit-mask Z gating,
// two markers per lane, register spills; profiles/README.md).  This kernel keeps the structure of
// elbo_kernel_mma4.cuh (4 cells per warp tile, hi/lo halves of the TF32 split as MMA columns,
// expanded-quadratic forward, moment-based backward) and adds what the larger shapes need:
//   * markers 0..31 run with lane = marker, one warp pass per cell; the J - 32 TAIL markers of
//     several cells are packed into ONE warp pass (lane = (cell, tail marker): 4 cells x 8 markers or
//     2 cells x 16), so the 19 MUFU per (cell, marker) are issued with 100 % / 83 % of the lanes live
//     instead of 62 % with "two markers per lane";
//   * Z^T . D^T is 4 m-tiles x ceil(J/8) k-chunks, Z . e^T is 3 m-tiles x ceil(K/8) k-chunks;
//   * dZ (48 x 64) would be 96 accumulator registers per thread if every warp kept its own copy.
//     Instead the CTA works in ROUNDS of 8 tiles and dZ lives in TENSOR MEMORY: every warp pair
//     writes the hi / lo split of its D and g values into K-major operand tiles (8 cells = K), one
//     __syncthreads, and one thread issues 4 tcgen05.mma.kind::tf32 (M = N = 128 = hi | lo rows,
//     K = 8) into the ONE 128 x 128 fp32 accumulator of the CTA in TMEM (fixed order: reproducible);
//     tcgen05.commit -> mbarrier tells when the double-buffered operand tiles may be rewritten, and
//     the epilogue reads the accumulator with tcgen05.ld.  (-DCYTOF_X_TMEM=0: the same with mma.sync,
//     warp w accumulating n-tile w over the tiles of all warps -- 5 % slower.)
#pragma once
#include "elbo_kernel_mma.cuh"

namespace cytof {

#ifndef CYTOF_X_A2_SKEW
#define CYTOF_X_A2_SKEW 2   // the log / reciprocal half of a mixture unit runs this many units behind its ex2 half (r2t: cfg5 0.849 -> 0.845 ms, J = 64 0.856 -> 0.837 ms, J = 32 / K = 50 0.660 -> 0.663 ms per 1 M cells)
#endif
#ifndef CYTOF_X_SMEMCONST
#define CYTOF_X_SMEMCONST 1
#endif
#ifndef CYTOF_X_MUP_RELOAD
#define CYTOF_X_MUP_RELOAD 1
#endif
#ifndef CYTOF_X_ZPF
#define CYTOF_X_ZPF 6       // A fragments of Z / Z^T requested this many MMAs ahead (r2y: J=32, K=50, L=5/5 0.6625 -> 0.6485 ms per 1 M cells with 6, 0.6535 with 4; the L=8/8 instances compile to the same code either way)
#endif
#ifndef CYTOF_X_NOGUARD
#define CYTOF_X_NOGUARD 1   // 1 (measured faster): always run the full 4x6 / 3x8 MMA grids (padding is zero in the Z tables)
#endif
constexpr int kXThreads = 256;
constexpr int kXWarps = kXThreads / 32;
constexpr int kXT = 4;                    // cells per warp tile
constexpr int kXDS = 52;                  // row stride of the D / A0 / y^2 tiles (48 markers, conflict-free ldmatrix)
constexpr int kXGS = 68;                  // row stride of the e / g tile (64 features)
constexpr int kXJS = 48;                  // row stride of the y buffer
constexpr int kXMTK = 4, kXMTJ = 3;       // m-tiles over features (64) and over markers (48)
constexpr int kXNCJ = 6, kXNCK = 8;       // k-chunks of 8 over markers and over features
#ifndef CYTOF_X_TMEM
#define CYTOF_X_TMEM 1   // 1: dZ on tcgen05.mma + TMEM (one 128 x 128 accumulator per CTA); 0: CTA-cooperative mma.sync
#endif
// CTA-level region after the Z tables (CYTOF_X_TMEM): header (2 mbarriers, TMEM base) and the operand
// tiles of the 4 warp PAIRS, double-buffered: A = [D_hi (64 rows) ; D_lo (64 rows)] x 8 cells, B = [g_hi ; g_lo]
constexpr int kXUFloats = CYTOF_X_TMEM ? 16 + 4 * 2 * 2 * 1024 : 0;
// CTA-level constant tables (CYTOF_X_SMEMCONST, narrow layout): intercepts of the tail markers [8 pairs][32 lanes] float2
// and log W in accumulator order [8 g][8] -- 24 registers of a kernel that spills at L = 8/8
constexpr int kXCFloats = 8 * 32 * 2 + 64;
constexpr int kXZFloats = (kXMTK * kXNCJ + kXMTJ * kXNCK) * 32 * 4 + kXUFloats + kXCFloats;
// WIDE instances (48 < J <= 64, NTP = 4: every cell takes a second full pass over markers 32..63): 64-marker tiles
constexpr int kXDSW = 68, kXJSW = 64, kXMTJW = 4, kXNCJW = 8;
constexpr int kXZFloatsW = (kXMTK * kXNCJW + kXMTJW * kXNCK) * 32 * 4 + kXUFloats + kXCFloats;
// c = F32, a = b = TF32, both K-major, N = 128, M = 128
constexpr uint32_t kUmmaIdesc128 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
// per warp (floats): D hi/lo [2][8][52] | g hi/lo [2][8][68] | A0/c1 [4][52] | y^2 [4][52] |
// 1/S0,1/S1 [6 units][32] float2 | per-cell scalars [16] | y double buffer [2][4][48] | sum w d [6][32]
constexpr int kXWarpFloats = 2 * 8 * kXDS + 2 * 8 * kXGS + 2 * kXT * kXDS + 6 * 32 * 2 + 16 + 2 * kXT * kXJS + 6 * 32;
constexpr int kXWarpFloatsW = 2 * 8 * kXDSW + 2 * 8 * kXGS + 2 * kXT * kXDSW + 8 * 32 * 2 + 16 + 2 * kXT * kXJSW + 8 * 32;

template <int LM0, int NC, int NPT>
__device__ __forceinline__ void mix_forward(const float yv, const float mc0, const float mc1, const float nh2,
                                            const float2 (&be)[NPT], const float2 (&al)[NPT], float2 (&ee)[NPT],
                                            float& S0, float& S1, float& B0, float& B1) {
  const float y0 = yv - mc0, y1 = yv - mc1;
  const float2 Y00 = bcp(y0), Y11 = bcp(y1), Y01 = mkp(y0, y1);
  float2 v[NPT];
#pragma unroll
  for (int q = 0; q < NPT; ++q) {
    const bool lo1 = 2 * q >= LM0, hi1 = 2 * q + 1 >= LM0;
    v[q] = fma2(be[q], lo1 ? Y11 : (hi1 ? Y01 : Y00), al[q]);
  }
  float M0 = kNegBig, M1 = kNegBig;
  flat_max<LM0, NC, NPT>(v, M0, M1);      // 2 FMNMX3 per 5 components (3 over the packed pairs)
  const float2 M00 = bcp(M0), M11 = bcp(M1), M01 = mkp(M0, M1);
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
  S0 = (lo2(s0) + hi2(s0)) + s0x;
  S1 = (lo2(s1) + hi2(s1)) + s1x;
  B0 = fmaf(nh2 * y0, y0, M0);      // A_z = B_z + lg2(S_z)
  B1 = fmaf(nh2 * y1, y1, M1);
}

// returns sum_l w_l d_l = c0 y0' + c1 y1' - sum_l w_l mu'_l when WD (the imputation gradient needs it)
template <int LM0, int NPT, bool WD>
__device__ __forceinline__ float mix_backward(const float yv, const float mc0, const float mc1, const float c0,
                                              const float c1, const float2 rs01, const float2 (&ee)[NPT],
                                              const float2 (&mup)[NPT], float2 (&sw)[NPT], float2 (&swy)[NPT],
                                              float& sq) {
  const float y0 = yv - mc0, y1 = yv - mc1;
  const float2 Y00 = bcp(y0), Y11 = bcp(y1), Y01 = mkp(y0, y1);
  const float r0 = c0 * rs01.x, r1 = c1 * rs01.y;
  const float2 R00 = bcp(r0), R11 = bcp(r1), R01 = mkp(r0, r1);
  sq = fmaf(c0 * y0, y0, fmaf(c1 * y1, y1, sq));
  float2 wm = bcp(0.f);
#pragma unroll
  for (int q = 0; q < NPT; ++q) {
    const bool lo1 = 2 * q >= LM0, hi1 = 2 * q + 1 >= LM0;
    const float2 w = mul2(lo1 ? R11 : (hi1 ? R01 : R00), ee[q]);
    acc_add2(sw[q], w);
    acc_fma2(swy[q], w, lo1 ? Y11 : (hi1 ? Y01 : Y00));
    if (WD) acc_fma2(wm, w, mup[q]);
  }
  return WD ? fmaf(c0, y0, c1 * y1) - (lo2(wm) + hi2(wm)) : 0.f;
}

// NTP = tail passes per 4-cell tile: 1 (J - 32 <= 8: 4 cells x 8 tail markers per pass),
//                                    2 (J - 32 <= 16: 2 cells x 16 tail markers per pass) or
//                                    4 (J - 32 <= 32: 1 cell x 32 tail markers per pass; WIDE tile layout)
template <int LM0, int LM1, bool NOISY, bool HAS_IDX, int NTP, bool MISSING>
__global__ void __launch_bounds__(kXThreads, 1)
elbo_fwdbwd_mma64_kernel(const ElboParams p) {
  pdl_enter();
  constexpr int NC = LM0 + LM1, NPT = (NC + 1) / 2;
  // NTP == 4 is the WIDE layout (48 < J <= 64); these shadow the namespace-scope constants of the narrow layout
  constexpr bool WIDE = NTP == 4;
  constexpr int kXDS = WIDE ? kXDSW : cytof::kXDS, kXJS = WIDE ? kXJSW : cytof::kXJS;
  constexpr int kXMTJ = WIDE ? kXMTJW : cytof::kXMTJ, kXNCJ = WIDE ? kXNCJW : cytof::kXNCJ;
  constexpr int kXZFloats = WIDE ? kXZFloatsW : cytof::kXZFloats, kXWarpFloats = WIDE ? kXWarpFloatsW : cytof::kXWarpFloats;
  constexpr int NUX = WIDE ? 8 : 6;                          // rows of the per-unit tables sR / sW
  static_assert(!WIDE || CYTOF_X_TMEM, "the WIDE layout keeps dZ in tensor memory");
  constexpr int CPT = kXT / NTP, TP = 32 / CPT;            // cells and (padded) tail markers per tail pass
  constexpr int NU = kXT + NTP;                             // mixture units (warp passes) per tile
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;
  uint4* sZT = reinterpret_cast<uint4*>(smem);              // [mt < 4][c < 6][lane]  A fragments of Z^T (k x j)
  uint4* sZ = sZT + kXMTK * kXNCJ * 32;                     // [mt < 3][c < 8][lane]  A fragments of Z   (j x k)
  float* wbase0 = smem + kXZFloats;      // kXZFloats includes the CTA-level operand region
#if CYTOF_X_TMEM
  float* uhdr = smem + (kXMTK * kXNCJ + kXMTJ * kXNCK) * 32 * 4;
  unsigned long long* xbar = reinterpret_cast<unsigned long long*>(uhdr);          // [2] one mbarrier per buffer
  uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(uhdr + 4);
  unsigned char* sU = reinterpret_cast<unsigned char*>(uhdr + 16);                  // [pair][buf][A | B][4096 B]
  unsigned char* myU = sU + (warp >> 1) * (2 * 2 * 4096);                           // this warp pair's tiles
#endif
  // only the L > 5 instances spill; for the others the table loads cost more than the registers (J=32, K=50, L=5/5:
  // 0.650 -> 0.658 ms per 1 M cells with the tables, cfg5 shapes at L=8/8 0.836 -> 0.766)
  constexpr bool CSM = CYTOF_X_SMEMCONST && !WIDE && NPT > 5;
  float2* sAlt = reinterpret_cast<float2*>(smem + (kXMTK * kXNCJ + kXMTJ * kXNCK) * 32 * 4 + kXUFloats);   // [8][32]
  float* sLw = reinterpret_cast<float*>(sAlt + 8 * 32);                                                     // [8 g][8]
  float* wbase = wbase0 + warp * kXWarpFloats;
  float* sDb = wbase;                              // [2][8][kXDS] row 2c = hi(D) of cell c, 2c+1 = lo(D)
  float* sGb = sDb + 2 * 8 * kXDS;                 // [2][8][kXGS] e, then g = fac rho softmax: hi / lo rows
  float* sA = sGb + 2 * 8 * kXGS;                  // [4][kXDS] A0_j, later c1_j
  float* sY = sA + kXT * kXDS;                     // [4][kXDS] y_j^2
  float2* sR = reinterpret_cast<float2*>(sY + kXT * kXDS);   // [NU][32] (1/S0, 1/S1) per unit and lane
  float* sS = reinterpret_cast<float*>(sR + NUX * 32);        // [0..3] fac rho, [4..7] fac (1-rho), [8..11] sc
  float* sYb = sS + 16;                            // [2][4][kXJS] y rows
  float* sW = sYb + 2 * kXT * kXJS;                // [NU][32] sum_l w d per unit (imputation gradient)
  float* sred = smem;

  const int I = p.I, J = p.J, K = p.K, L0 = p.L0, L1 = p.L1;
  int i = 0;
  while (i + 1 < I && (int)blockIdx.x >= p.blk_begin[i + 1]) ++i;
  const int nb = p.blk_begin[i + 1] - p.blk_begin[i];
  const int lb = (int)blockIdx.x - p.blk_begin[i];
  const long long cb = p.cell_begin[i];
  const long long Bi = p.cell_begin[i + 1] - cb;
  // this CTA's cells [c_lo, c_hi) of the sample; the cut points are moved down to a global position that is a multiple
  // of 4, so that a 4-cell tile is exactly ONE Philox block of the imputation pre-pass
  auto cut = [&](const long long q) -> long long {
    long long x = Bi * q / nb;
    if (q > 0 && q < nb) x -= (long long)(((unsigned long long)p.row_global[i] + (unsigned long long)x) & 3ull);
    return x < 0 ? 0 : x;
  };
  const long long c_lo = cut(lb), c_hi = cut(lb + 1);
  const long long row_base = p.row_base[i];
  const int ncj = (J + 7) >> 3, nck = (K + 7) >> 3, mtk = (K + 15) >> 4, mtj = (J + 15) >> 4;

  // ---------------- Z fragment tables, zeroed tiles ----------------
  for (int e = threadIdx.x; e < (kXMTK * kXNCJ + kXMTJ * kXNCK) * 32; e += kXThreads) {
    const bool second = e >= kXMTK * kXNCJ * 32;
    const int e2 = second ? e - kXMTK * kXNCJ * 32 : e;
    const int ln = e2 & 31, mc = e2 >> 5;
    const int mt = second ? mc / kXNCK : mc / kXNCJ, c = second ? mc % kXNCK : mc % kXNCJ;
    const int gg = ln >> 2, tt = ln & 3;
    uint32_t v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = 16 * mt + gg + ((r & 1) ? 8 : 0), col = 8 * c + tt + ((r & 2) ? 4 : 0);
      const int j = second ? row : col, k = second ? col : row;   // first: Z^T[k][j], second: Z[j][k]
      const unsigned long long zr = (j < J && k < K) ? p.zmask[j] : 0ull;
      v[r] = ((zr >> k) & 1ull) ? 0x3f800000u : 0u;
    }
    (second ? sZ : sZT)[mc * 32 + ln] = make_uint4(v[0], v[1], v[2], v[3]);
  }
  for (int e = lane; e < 2 * 8 * kXDS + 2 * 8 * kXGS + 2 * kXT * kXDS; e += 32) wbase[e] = 0.f;   // D, g, A0, y^2 tiles
  for (int e = lane; e < 2 * kXT * kXJS; e += 32) sYb[e] = 0.f;       // y buffers: columns >= J are never written by the copies
#if CYTOF_X_TMEM
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(smem_u32(tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(xbar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(xbar + 1)) : "memory");
  }
  for (int e = threadIdx.x; e < 4 * 2 * 2 * 1024; e += kXThreads) reinterpret_cast<float*>(sU)[e] = 0.f;   // pad rows stay zero
#endif

  // ---------------- per-sample constants ----------------
  const GlobalsLayout gl = globals_layout(I, J, K, L0, L1);
  const float* __restrict__ G = p.globals;
  const float sig = G[gl.sig + i];
  const float inv_sig2 = 1.0f / (sig * sig);
  const float nh2 = -0.5f * inv_sig2 * kLog2e;
  const float cbase = (-logf(sig) - kHalfLog2Pi) * kLog2e;
  const float fac = p.fac[i];
  // main lanes: marker `lane`, live if lane < J (always when J > 32)
  const bool okm = lane < J;
  const float okmf = okm ? 1.f : 0.f;
  const int jm = okm ? lane : 0;
  // tail lane -> (cell within the pass, tail marker)
  const int tcell = lane / TP, jt = 32 + lane % TP;
  const bool tail_ok = jt < J;
  const float tokf = tail_ok ? 1.f : 0.f;

  float2 mup[NPT], be[NPT], alm[NPT], alt[NPT];      // alm: marker `lane`, alt: tail marker jt
  float mc0, mc1;
  {
    const float m0a = G[gl.mu0], m0b = G[gl.mu0 + L0 - 1], m1a = G[gl.mu1], m1b = G[gl.mu1 + L1 - 1];
    mc0 = 0.5f * (m0a + m0b);
    mc1 = 0.5f * (m1a + m1b);
    float m[2 * NPT], b[2 * NPT], am[2 * NPT], at[2 * NPT];
    const int jtt = tail_ok ? jt : 0;
#pragma unroll
    for (int x = 0; x < 2 * NPT; ++x) {
      const bool z1 = x >= LM0;
      const int l = z1 ? x - LM0 : x, Lz = z1 ? L1 : L0;
      const bool live = l < Lz && x < LM0 + LM1;
      const int ll = live ? l : 0;
      m[x] = live ? G[(z1 ? gl.mu1 : gl.mu0) + ll] - (z1 ? mc1 : mc0) : 0.f;
      b[x] = live ? -2.f * nh2 * m[x] : 0.f;
      const float base = fmaf(nh2 * m[x], m[x], cbase);
      const float* le = G + (z1 ? gl.le1 : gl.le0);
      am[x] = (okm && live) ? fmaf(le[(i * J + jm) * Lz + ll], kLog2e, base) : ((okm || l != 0) ? kNegBig : 0.f);
      at[x] = (tail_ok && live) ? fmaf(le[(i * J + jtt) * Lz + ll], kLog2e, base) : (l == 0 ? 0.f : kNegBig);
    }
#pragma unroll
    for (int q = 0; q < NPT; ++q) {
      mup[q] = mkp(m[2 * q], m[2 * q + 1]);
      be[q] = mkp(b[2 * q], b[2 * q + 1]);
      alm[q] = mkp(am[2 * q], am[2 * q + 1]);
      alt[q] = mkp(at[2 * q], at[2 * q + 1]);
    }
  }
  // log W (log2 domain) of the 8 feature rows of this thread in the accumulator layout: 16 mt + g (+8)
  float lwk[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int k = 16 * (r >> 1) + g + 8 * (r & 1);
    lwk[r] = k < K ? G[gl.lw + i * K + k] * kLog2e : kNegBig;
  }
  if (CSM && warp == 0) {     // every warp of the CTA holds the same values (same sample, same lane -> marker map)
#pragma unroll
    for (int q = 0; q < NPT; ++q) sAlt[q * 32 + lane] = alt[q];
    if (t4 == 0) {
#pragma unroll
      for (int r = 0; r < 8; ++r) sLw[g * 8 + r] = lwk[r];
    }
  }
  // imputation parameters of this lane's main marker and tail marker (vae.py:12-13), missingness b
  const int jtl = tail_ok ? jt : 0;
  const float vm_m = G[gl.vm + i * J + jm], vls_m = G[gl.vls + i * J + jm], sd_m = expf(vls_m);
  const float vm_t = G[gl.vm + i * J + jtl], vls_t = G[gl.vls + i * J + jtl], sd_t = expf(vls_t);
  const float b0 = G[gl.b + 3 * i], b1 = G[gl.b + 3 * i + 1], b2 = G[gl.b + 3 * i + 2];
  const unsigned long long grow_base = (unsigned long long)p.row_global[i];
  float l1me2 = 0.f, zconst2 = 0.f, ninv2s2 = 0.f, inv_s2n = 0.f;
  if (NOISY) {
    inv_s2n = 1.0f / (p.noisy_sd * p.noisy_sd);
    const float e_i = G[gl.eps + i];
    l1me2 = log1pf(-e_i) * kLog2e;
    zconst2 = ((float)J * (-kHalfLog2Pi - logf(p.noisy_sd)) + logf(e_i)) * kLog2e;
    ninv2s2 = -0.5f / (p.noisy_sd * p.noisy_sd) * kLog2e;
  }

  // ---------------- accumulators ----------------
  float2 swm[NPT], swym[NPT], swt[NPT], swyt[NPT];
  float sq = 0.f;                        // sum c_z y_z'^2 over all of this lane's units
#if !CYTOF_X_TMEM
  float dZ[kXMTJ][4];                    // n-tile `warp` of dZ, all 3 m-tiles (markers permuted, see epilogue)
#endif
  float gW[8];
  float sfr = 0.f, sfo = 0.f;
  float dvm_m = 0.f, dvls_m = 0.f, nm_m = 0.f, dvm_t = 0.f, dvls_t = 0.f, nm_t = 0.f, lqy_l = 0.f, lmiss_l = 0.f;
  double ll2 = 0.0;
#pragma unroll
  for (int q = 0; q < NPT; ++q) swm[q] = swym[q] = swt[q] = swyt[q] = bcp(0.f);
#if !CYTOF_X_TMEM
#pragma unroll
  for (int a = 0; a < kXMTJ; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) dZ[a][c] = 0.f;
#endif
#pragma unroll
  for (int r = 0; r < 8; ++r) gW[r] = 0.f;

  // ---------------- rounds ----------------
  const int32_t* __restrict__ idx = p.idx + (HAS_IDX ? (cb + c_lo) : 0);
  const float* __restrict__ ysamp = p.y + (size_t)row_base * J;
  const int ncell = (int)(c_hi - c_lo);
  const int first_row = (int)c_lo;
  const int nrounds = (ncell + kXWarps * kXT - 1) / (kXWarps * kXT);      // CTA-uniform
  const bool vec16 = (J & 3) == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0;   // rows are 16-byte aligned, gathered or not

  // y rows of this warp's tile of round r -> sYb[b] (row stride kXJS); rows past the range are zeros
  RowSampler rsm;      // CYTOF_FLAG_SAMPLE_IN_KERNEL: the rows come from the sampler's keyed bijection, not from idx
  if (HAS_IDX && p.sampler_on) rsm = make_row_sampler(p, i);
  auto issue_tile = [&](const int r, const int b) {
    const int tn = (r * kXWarps + warp) * kXT;
    float* dst = sYb + b * (kXT * kXJS);
    int row4 = 0;      // lane l: row of cell (l & 3) of the tile
    if (HAS_IDX && tn + (lane & 3) < ncell)
      row4 = p.sampler_on ? rsm.row(c_lo + tn + (lane & 3)) : __ldg(idx + tn + (lane & 3));
#pragma unroll
    for (int u = 0; u < kXT; ++u) {
      const int rowg = HAS_IDX ? __shfl_sync(FULL, row4, u) : 0;
      if (tn + u < ncell) {
        const int row = HAS_IDX ? rowg : first_row + tn + u;
        const float* src = ysamp + (size_t)row * J;
        if (vec16) {
          if (4 * lane < J) cp_async16(dst + u * kXJS + 4 * lane, src + 4 * lane);
        } else {
          if (lane < J) cp_async4(dst + u * kXJS + lane, src + lane);
          if (32 + lane < J) cp_async4(dst + u * kXJS + 32 + lane, src + 32 + lane);
        }
      } else {
        dst[u * kXJS + lane] = 0.f;
        if (lane < kXJS - 32) dst[u * kXJS + 32 + lane] = 0.f;
      }
    }
    cp_async_commit();
  };

  // imputation of missing entries (vae.py:17-33): NaNs of the tile in sYb[b] are replaced by the
  // N(0,1) draw; mm = cells whose MAIN marker `lane` is missing, tm = tail passes whose entry of this
  // lane is missing, am = units of the tile with any missing entry (warp-uniform)
  auto prepass = [&](const int r, const int b, unsigned& mm, unsigned& tm, unsigned& am) {
    const int tn = (r * kXWarps + warp) * kXT;
    float* src = sYb + b * (kXT * kXJS);
    mm = 0u;
    tm = 0u;
#pragma unroll
    for (int u = 0; u < kXT; ++u) {
      const float yr = src[u * kXJS + lane];
      mm |= (yr != yr) ? (1u << u) : 0u;
    }
#pragma unroll
    for (int ps = 0; ps < NTP; ++ps) {
      const float yr = tail_ok ? src[(ps * CPT + tcell) * kXJS + jt] : 0.f;
      tm |= (yr != yr) ? (1u << ps) : 0u;
    }
    am = __reduce_or_sync(FULL, mm | (tm << kXT));
    if (am) {
      unsigned long long grp_c = ~0ull;
      float nz[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int u = 0; u < kXT; ++u) {
        if (!((am >> u) & 1u)) continue;
        float e = 0.f;
        if (p.noise_mode == CYTOF_NOISE_EXTERNAL) {
          if (p.eps) e = __ldg(p.eps + (cb + c_lo + tn + u) * J + lane);
        } else {
          const unsigned long long grow = grow_base + (unsigned long long)(first_row + tn + u);   // position in the batch
          if ((grow >> 2) != grp_c) {
            grp_c = grow >> 2;
            philox_normal4(p.seed, elbo_step(p), grp_c, philox_ctr(lane, i), nz);
          }
          const unsigned w = (unsigned)grow & 3u;
          e = w == 0u ? nz[0] : (w == 1u ? nz[1] : (w == 2u ? nz[2] : nz[3]));
        }
        if ((mm >> u) & 1u) src[u * kXJS + lane] = e;
      }
#pragma unroll 1
      for (int ps = 0; ps < NTP; ++ps) {
        if (!((am >> (kXT + ps)) & 1u)) continue;
        const int cu = ps * CPT + tcell;
        if ((tm >> ps) & 1u) {      // only lanes whose tail entry is NaN (it exists, so the cell is in range)
          src[cu * kXJS + jt] = p.noise_mode == CYTOF_NOISE_EXTERNAL
                                    ? (p.eps ? __ldg(p.eps + (cb + c_lo + tn + cu) * J + jt) : 0.f)
                                    : philox_normal(p.seed, elbo_step(p), grow_base + (unsigned long long)(first_row + tn + cu), philox_ctr(jt, i));
        }
      }
    }
  };

  float2 ee[NU][NPT];

  // forward of unit `un` of the tile whose y rows are in `src`, in two halves: a1 issues the ex2 of
  // the unit, a2 (called one unit later, so the MUFU pipe already works on the next exponentials
  // while these sums settle) takes the logs / reciprocal and writes the tiles of buffer `b`
  struct AState { float S0, S1, B0, B1, yv; };
  auto phase_a1 = [&](const int un, const float* src, const unsigned mm, const unsigned tm, const bool in_loop = true) -> AState {
    AState st;
    if (un < kXT) {                               // main pass: cell un, marker lane
      const float x = src[un * kXJS + lane];
      st.yv = (MISSING && ((mm >> un) & 1u)) ? fmaf(sd_m, x, vm_m) : x;
      mix_forward<LM0, NC, NPT>(st.yv, mc0, mc1, nh2, be, alm, ee[un], st.S0, st.S1, st.B0, st.B1);
    } else {                                      // tail pass: cell cu, tail marker jt
      const int cu = (un - kXT) * CPT + tcell;
      const float x = tail_ok ? src[cu * kXJS + jt] : 0.f;
      st.yv = (MISSING && ((tm >> (un - kXT)) & 1u)) ? fmaf(sd_t, x, vm_t) : x;
      if (CSM && in_loop) {      // the tail intercepts come from the CTA table (one unit per tile uses them)
        float2 al_t[NPT];
#pragma unroll
        for (int q = 0; q < NPT; ++q) al_t[q] = sAlt[q * 32 + lane];
        mix_forward<LM0, NC, NPT>(st.yv, mc0, mc1, nh2, be, al_t, ee[un], st.S0, st.S1, st.B0, st.B1);
      } else {
        mix_forward<LM0, NC, NPT>(st.yv, mc0, mc1, nh2, be, alt, ee[un], st.S0, st.S1, st.B0, st.B1);
      }
    }
    return st;
  };
#if CYTOF_X_TMEM
  float udh[kXT], udl[kXT];    // hi / lo of D of the 4 cells of the tile: one 16-byte store each into the A operand tile
#endif
  auto phase_a2 = [&](const int un, const AState& st, const int b) {
    float* sD = sDb + b * (8 * kXDS);
    const float A0 = st.B0 + fast_lg2(st.S0), A1 = st.B1 + fast_lg2(st.S1);
    const float rr = fast_rcp(st.S0 * st.S1);
    sR[un * 32 + lane] = make_float2(rr * st.S1, rr * st.S0);
    uint32_t dh, dl;
    if (un < kXT) {
      split_tf32(okmf * (A1 - A0), dh, dl);
      sD[(2 * un) * kXDS + lane] = __uint_as_float(dh);
      sD[(2 * un + 1) * kXDS + lane] = __uint_as_float(dl);
      sA[un * kXDS + lane] = okmf * A0;
      if (NOISY) sY[un * kXDS + lane] = st.yv * st.yv;
#if CYTOF_X_TMEM
      udh[un] = __uint_as_float(dh);
      udl[un] = __uint_as_float(dl);
      if (un == kXT - 1) {     // A operand: row = marker (hi) / 64 + marker (lo), k = 4 (warp & 1) + cell
        unsigned char* ua = myU + b * (2 * 4096) + (lane >> 3) * 256 + (warp & 1) * 128 + (lane & 7) * 16;
        *reinterpret_cast<float4*>(ua) = make_float4(udh[0], udh[1], udh[2], udh[3]);
        *reinterpret_cast<float4*>(ua + 2048) = make_float4(udl[0], udl[1], udl[2], udl[3]);
      }
#endif
    } else {
      const int cu = (un - kXT) * CPT + tcell;
      split_tf32(tokf * (A1 - A0), dh, dl);
      sD[(2 * cu) * kXDS + jt] = __uint_as_float(dh);
      sD[(2 * cu + 1) * kXDS + jt] = __uint_as_float(dl);
      sA[cu * kXDS + jt] = tokf * A0;
      if (NOISY) sY[cu * kXDS + jt] = tokf * st.yv * st.yv;
#if CYTOF_X_TMEM
      {
        unsigned char* ua = myU + b * (2 * 4096) + (jt >> 3) * 256 + (warp & 1) * 128 + (jt & 7) * 16 + cu * 4;
        *reinterpret_cast<float*>(ua) = __uint_as_float(dh);
        *reinterpret_cast<float*>(ua + 2048) = __uint_as_float(dl);
      }
#endif
    }
  };
  auto phase_b = [&](const int un, const float* src, const unsigned mm, const unsigned tm, const unsigned am) {
    if (un < kXT) {
      const float c1 = sA[un * kXDS + lane];
      const float c0 = okmf * sS[un] - c1;
      const float x = src[un * kXJS + lane];
      const float yv = (MISSING && ((mm >> un) & 1u)) ? fmaf(sd_m, x, vm_m) : x;
      const float wd = mix_backward<LM0, NPT, MISSING>(yv, mc0, mc1, c0, c1, sR[un * 32 + lane], ee[un], mup, swm, swym, sq);
      if (MISSING && am) sW[un * 32 + lane] = wd;
    } else {
      const int cu = (un - kXT) * CPT + tcell;
      const float c1 = tail_ok ? sA[cu * kXDS + jt] : 0.f;
      const float c0 = tokf * sS[cu] - c1;
      const float x = tail_ok ? src[cu * kXJS + jt] : 0.f;
      const float yv = (MISSING && ((tm >> (un - kXT)) & 1u)) ? fmaf(sd_t, x, vm_t) : x;
      const float wd = mix_backward<LM0, NPT, MISSING>(yv, mc0, mc1, c0, c1, sR[un * 32 + lane], ee[un], mup, swt, swyt, sq);
      if (MISSING && am) sW[un * 32 + lane] = wd;
    }
  };
  // logistic missingness (Model.py:269-274) and the d/dy chain into the imputation parameters of
  // unit `un`, branch-free (lanes without a missing entry add 0)
  auto missing_unit = [&](const int un, const float* src, const unsigned mm, const unsigned tm, const int tn) {
    const bool main = un < kXT;
    const int cu = main ? un : (un - kXT) * CPT + tcell;
    const bool miss = main ? ((mm >> un) & 1u) : ((tm >> (un - kXT)) & 1u);
    const bool vu = tn + cu < ncell;
    const float mf = (miss && vu) ? 1.f : 0.f;
    const float x = main ? src[un * kXJS + lane] : (tail_ok ? src[cu * kXJS + jt] : 0.f);
    const float sd = main ? sd_m : sd_t, vm = main ? vm_m : vm_t, vls = main ? vls_m : vls_t;
    const float e = miss ? x : 0.f;
    const float yv = miss ? fmaf(sd, e, vm) : x;
    const float fu = vu ? fac : 0.f;
    const float uu = fmaf(fmaf(b2, yv, b1), yv, b0);
    const float tmm = fast_ex2(-fabsf(uu) * kLog2e);
    const float rr = fast_rcp(1.f + tmm);
    const float ls = -(fmaxf(-uu, 0.f) + kLn2 * fast_lg2(1.f + tmm));
    const float sneg = (uu > 0.f ? tmm : 1.f) * rr;   // 1 - sigmoid(u)
    float dy = -sW[un * 32 + lane] * inv_sig2 + fu * sneg * fmaf(2.f * b2, yv, b1);
    if (NOISY) dy = fmaf(-sS[4 + cu] * inv_s2n, yv, dy);
    dy = miss ? dy : 0.f;
    lmiss_l = fmaf(ls, mf, lmiss_l);
    lqy_l = fmaf(fmaf(-0.5f * e, e, -vls - kHalfLog2Pi), mf, lqy_l);
    if (main) { dvm_m += dy; dvls_m = fmaf(dy * sd, e, dvls_m); nm_m += mf; }
    else { dvm_t += dy; dvls_t = fmaf(dy * sd, e, dvls_t); nm_t += mf; }
  };

  unsigned mmask = 0u, tmask = 0u, amask = 0u;
  if (nrounds > 0) {
    issue_tile(0, 0);
    cp_async_wait<0>();
    __syncwarp();
    if (MISSING) prepass(0, 0, mmask, tmask, amask);
    {
      AState prev = phase_a1(0, sYb, mmask, tmask, false);
#pragma unroll
      for (int un = 1; un < NU; ++un) {
        const AState now = phase_a1(un, sYb, mmask, tmask, false);
        phase_a2(un - 1, prev, 0);
        prev = now;
      }
      phase_a2(NU - 1, prev, 0);
    }
    if (nrounds > 1) issue_tile(1, 1);
  }
  __syncthreads();   // Z tables visible
#if CYTOF_X_TMEM
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = *tmem_base_s;
  bool umma_ok = true;
#endif

  for (int r = 0; r < nrounds; ++r) {
    const int b = r & 1;
    const int t0 = (r * kXWarps + warp) * kXT;
    const bool valid = t0 + t4 < ncell;        // this thread's cell (t4) in the accumulator layout
    float* sD = sDb + b * (8 * kXDS);
    float* sG = sGb + b * (8 * kXGS);
    __syncwarp();

    // ===== contraction 1: f^T[k, (n, hi|lo)] = Z^T . [D_hi | D_lo]^T (+ log W on the hi column)
    float fT[kXMTK][4];
    {
      uint32_t rb[2 * kXNCJ];     // B fragments: block q = markers 4q..4q+3, column g = row g of the D tile
#pragma unroll
      for (int q4 = 0; q4 < kXNCJ / 2; ++q4) {
        uint32_t r4[4];
        ldsm_x4(sD + (lane & 7) * kXDS + 16 * q4 + 4 * (lane >> 3), r4);
#pragma unroll
        for (int m = 0; m < 4; ++m) rb[4 * q4 + m] = r4[m];
      }
      if (CSM) {
        const float4 l0 = *reinterpret_cast<const float4*>(sLw + g * 8), l1 = *reinterpret_cast<const float4*>(sLw + g * 8 + 4);
        fT[0][0] = l0.x; fT[0][2] = l0.y; fT[1][0] = l0.z; fT[1][2] = l0.w;
        fT[2][0] = l1.x; fT[2][2] = l1.y; fT[3][0] = l1.z; fT[3][2] = l1.w;
#pragma unroll
        for (int mt = 0; mt < kXMTK; ++mt) fT[mt][1] = fT[mt][3] = 0.f;
      } else {
#pragma unroll
        for (int mt = 0; mt < kXMTK; ++mt) {
          fT[mt][0] = lwk[2 * mt];
          fT[mt][2] = lwk[2 * mt + 1];
          fT[mt][1] = fT[mt][3] = 0.f;
        }
      }
#if CYTOF_X_ZPF > 0 && CYTOF_X_NOGUARD
      {   // the A fragments of Z^T are requested CYTOF_X_ZPF MMAs ahead of their use (each feeds ONE MMA here)
        constexpr int NT = kXNCJ * kXMTK, PF = CYTOF_X_ZPF;
        uint4 zq[PF];
#pragma unroll
        for (int t = 0; t < PF; ++t) zq[t] = sZT[((t % kXMTK) * kXNCJ + t / kXMTK) * 32 + lane];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const uint4 a = zq[t % PF];
          if (t + PF < NT) zq[t % PF] = sZT[(((t + PF) % kXMTK) * kXNCJ + (t + PF) / kXMTK) * 32 + lane];
          mma_tf32(fT[t % kXMTK], a, rb[2 * (t / kXMTK)], rb[2 * (t / kXMTK) + 1]);
        }
      }
#else
#pragma unroll
      for (int c = 0; c < kXNCJ; ++c) {
        if (CYTOF_X_NOGUARD || c < ncj) {
#pragma unroll
          for (int mt = 0; mt < kXMTK; ++mt)
            if (CYTOF_X_NOGUARD || mt < mtk) mma_tf32(fT[mt], sZT[(mt * kXNCJ + c) * 32 + lane], rb[2 * c], rb[2 * c + 1]);
        }
      }
#endif
    }
    float f[8];
#pragma unroll
    for (int mt = 0; mt < kXMTK; ++mt) {
      f[2 * mt] = fT[mt][0] + fT[mt][1];        // feature 16 mt + g
      f[2 * mt + 1] = fT[mt][2] + fT[mt][3];    // feature 16 mt + g + 8
    }
    float mx = fmaxf(fmaxf(fmaxf(f[0], f[1]), fmaxf(f[2], f[3])), fmaxf(fmaxf(f[4], f[5]), fmaxf(f[6], f[7])));
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    float ex[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      ex[q] = fast_ex2(f[q] - mx);
      uint32_t eh, el;
      split_tf32(ex[q], eh, el);
      const int k = 16 * (q >> 1) + g + 8 * (q & 1);
      sG[(2 * t4) * kXGS + k] = __uint_as_float(eh);
      sG[(2 * t4 + 1) * kXGS + k] = __uint_as_float(el);
    }
    __syncwarp();

    // ===== contraction 2 on the numerators: c1'^T[j, (n, hi|lo)] = Z . [e_hi | e_lo]^T
    float cT[kXMTJ][4];
    {
      uint32_t rb[16];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint32_t r4[4];
        ldsm_x4(sG + (lane & 7) * kXGS + 16 * q4 + 4 * (lane >> 3), r4);
#pragma unroll
        for (int m = 0; m < 4; ++m) rb[4 * q4 + m] = r4[m];
      }
#pragma unroll
      for (int mt = 0; mt < kXMTJ; ++mt) cT[mt][0] = cT[mt][1] = cT[mt][2] = cT[mt][3] = 0.f;
#if CYTOF_X_ZPF > 0 && CYTOF_X_NOGUARD
      {
        constexpr int NT = kXNCK * kXMTJ, PF = CYTOF_X_ZPF;
        uint4 zq[PF];
#pragma unroll
        for (int t = 0; t < PF; ++t) zq[t] = sZ[((t % kXMTJ) * kXNCK + t / kXMTJ) * 32 + lane];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const uint4 a = zq[t % PF];
          if (t + PF < NT) zq[t % PF] = sZ[(((t + PF) % kXMTJ) * kXNCK + (t + PF) / kXMTJ) * 32 + lane];
          mma_tf32(cT[t % kXMTJ], a, rb[2 * (t / kXMTJ)], rb[2 * (t / kXMTJ) + 1]);
        }
      }
#else
#pragma unroll
      for (int c = 0; c < kXNCK; ++c) {
        if (CYTOF_X_NOGUARD || c < nck) {
#pragma unroll
          for (int mt = 0; mt < kXMTJ; ++mt)
            if (CYTOF_X_NOGUARD || mt < mtj) mma_tf32(cT[mt], sZ[(mt * kXNCK + c) * 32 + lane], rb[2 * c], rb[2 * c + 1]);
        }
      }
#endif
    }
    // ===== per-cell scalars (cell t4): logsumexp over K, noisy class
    float rs = ((ex[0] + ex[1]) + (ex[2] + ex[3])) + ((ex[4] + ex[5]) + (ex[6] + ex[7]));
    float ra, ry = 0.f;
    {
      const float4 a4 = *reinterpret_cast<const float4*>(sA + t4 * kXDS + 4 * g);
      if (WIDE) {     // tail columns 32..63: four per g
        const float4 b4 = *reinterpret_cast<const float4*>(sA + t4 * kXDS + 32 + 4 * g);
        ra = ((a4.x + a4.y) + (a4.z + a4.w)) + ((b4.x + b4.y) + (b4.z + b4.w));
      } else {        // tail columns 32..47: two per g
        const float2 a2 = *reinterpret_cast<const float2*>(sA + t4 * kXDS + 32 + 2 * g);
        ra = ((a4.x + a4.y) + (a4.z + a4.w)) + (a2.x + a2.y);
      }
      if (NOISY) {
        const float4 c4 = *reinterpret_cast<const float4*>(sY + t4 * kXDS + 4 * g);
        if (WIDE) {
          const float4 d4 = *reinterpret_cast<const float4*>(sY + t4 * kXDS + 32 + 4 * g);
          ry = ((c4.x + c4.y) + (c4.z + c4.w)) + ((d4.x + d4.y) + (d4.z + d4.w));
        } else {
          const float2 c2 = *reinterpret_cast<const float2*>(sY + t4 * kXDS + 32 + 2 * g);
          ry = ((c4.x + c4.y) + (c4.z + c4.w)) + (c2.x + c2.y);
        }
      }
    }
    {
      float2 rr = mkp(rs, ra);
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        rr = add2(rr, shfl_xor2(rr, o));
        if (NOISY) ry += __shfl_xor_sync(FULL, ry, o);
      }
      rs = lo2(rr);
      ra = hi2(rr);
    }
    const float facc = valid ? fac : 0.f;
    const float lse2 = (mx + ra) + fast_lg2(rs);
    const float rinv = fast_rcp(rs);
    float rho = 1.f, omr = 0.f, L2 = lse2;
    if (NOISY) {
      const float q2 = lse2 + l1me2;
      const float z2 = fmaf(ninv2s2, ry, zconst2);
      const float dd = q2 - z2;
      const float tq = fast_ex2(-fabsf(dd));
      const float r1 = fast_rcp(1.f + tq);
      L2 = fmaxf(q2, z2) + fast_lg2(1.f + tq);
      rho = dd >= 0.f ? r1 : tq * r1;
      omr = dd >= 0.f ? tq * r1 : r1;
    }
    if (valid) ll2 += (double)L2;
    const float fr = facc * rho;
    const float sc = fr * rinv;               // g_k = sc e_k
    sfr += fr;
    sfo = fmaf(facc, omr, sfo);
    if (g == 0) {
      sS[t4] = fr;
      sS[4 + t4] = facc * omr;
    }
    __syncwarp();      // every lane has its e fragments (ldmatrix) and its A0 sums: the tiles may be rewritten
    // c1 = sc * (hi + lo columns) over the A0 tile; g = sc e, split, over the e tile (for dZ)
#pragma unroll
    for (int mt = 0; mt < kXMTJ; ++mt) {
      const int j0 = 16 * mt + g;
      if (j0 < kXJS) sA[t4 * kXDS + j0] = sc * (cT[mt][0] + cT[mt][1]);
      if (j0 + 8 < kXJS) sA[t4 * kXDS + j0 + 8] = sc * (cT[mt][2] + cT[mt][3]);
    }
#if CYTOF_X_TMEM
    // ===== contraction 3 on tcgen05 + TMEM: g = sc e, split, into the B operand tile of this warp pair
    // (row = feature (hi) / 64 + feature (lo), k = 4 (warp & 1) + cell t4); after the CTA barrier one
    // thread issues the 4 pair MMAs (TF32, M = N = 128, K = 8) of the round into the ONE accumulator of
    // the CTA -- a fixed issue order, so the sums are reproducible.  (Measured: per-pair accumulators
    // with pair barriers instead of the CTA barrier are 5 % slower.)
    {
      unsigned char* ub = myU + b * (2 * 4096) + 4096 + (warp & 1) * 128 + g * 16 + t4 * 4;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float gv = sc * ex[q];
        gW[q] += gv;
        uint32_t gh, gl2;
        split_tf32(gv, gh, gl2);
        *reinterpret_cast<float*>(ub + q * 256) = __uint_as_float(gh);            // feature 16 (q >> 1) + 8 (q & 1) + g
        *reinterpret_cast<float*>(ub + q * 256 + 2048) = __uint_as_float(gl2);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // the operand tiles of ALL warp pairs for this round are complete
    if (threadIdx.x == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int pr = 0; pr < 4; ++pr) {
        const unsigned char* t = sU + pr * (2 * 2 * 4096) + b * (2 * 4096);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                     :: "r"(tmem_acc), "l"(umma_desc_kmajor(t)), "l"(umma_desc_kmajor(t + 4096)), "r"(kUmmaIdesc128),
                        "r"((r > 0 || pr > 0) ? 1u : 0u) : "memory");
      }
      umma_commit(xbar + b);
    }
#else
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float gv = sc * ex[q];
      gW[q] += gv;
      uint32_t gh, gl2;
      split_tf32(gv, gh, gl2);
      const int k = 16 * (q >> 1) + g + 8 * (q & 1);
      sG[(2 * t4) * kXGS + k] = __uint_as_float(gh);
      sG[(2 * t4 + 1) * kXGS + k] = __uint_as_float(gl2);
    }
    __syncthreads();   // the D and g tiles of ALL warps for this round are complete

    // ===== contraction 3 (CTA-cooperative): this warp owns features 8 warp .. 8 warp + 7 of dZ and
    // walks the tiles of all 8 warps; contraction index = (cell, hi|lo of D); markers permuted so that
    // the 6 values a thread needs are contiguous: rows g / g+8 of m-tile mt = markers 6g+2mt / 6g+2mt+1
    if (8 * warp < K) {
#pragma unroll 2
      for (int s = 0; s < kXWarps; ++s) {
        const float* tD = wbase0 + s * kXWarpFloats + b * (8 * kXDS);
        const float* tG = wbase0 + s * kXWarpFloats + 2 * 8 * kXDS + b * (8 * kXGS);
        const uint32_t gh = __float_as_uint(tG[(2 * t4) * kXGS + 8 * warp + g]);
        const uint32_t gl2 = __float_as_uint(tG[(2 * t4 + 1) * kXGS + 8 * warp + g]);
#pragma unroll
        for (int mt = 0; mt < kXMTJ; ++mt) {
          const float2 dh = *reinterpret_cast<const float2*>(tD + (2 * t4) * kXDS + 6 * g + 2 * mt);
          const float2 dl = *reinterpret_cast<const float2*>(tD + (2 * t4 + 1) * kXDS + 6 * g + 2 * mt);
          const uint4 a = make_uint4(__float_as_uint(dh.x), __float_as_uint(dh.y), __float_as_uint(dl.x), __float_as_uint(dl.y));
          mma_tf32(dZ[mt], a, gh, gh);                                   // D_hi g_hi + D_lo g_hi
          mma_tf32(dZ[mt], make_uint4(a.x, a.y, 0u, 0u), gl2, 0u);       // D_hi g_lo
        }
      }
    }

#endif

    // ===== backward of this round's tile, interleaved unit by unit with the forward of the next
    const float* ycur = sYb + b * (kXT * kXJS);
    unsigned mm_n = 0u, tm_n = 0u, am_n = 0u;
#if CYTOF_X_TMEM
    // the next round's operands go to buffer b ^ 1: the MMAs of round r - 1 (same buffer) must be done
    if (r + 1 < nrounds && r >= 1) umma_ok = mbar_wait(xbar + (b ^ 1), (uint32_t)(((r + 1) >> 1) - 1) & 1u) && umma_ok;
#endif
    if (r + 1 < nrounds) {
      const float* ynxt = sYb + (b ^ 1) * (kXT * kXJS);
      cp_async_wait<0>();
      __syncwarp();
      if (MISSING) prepass(r + 1, b ^ 1, mm_n, tm_n, am_n);
      AState st[NU];      // the log / reciprocal half of unit un runs CYTOF_X_A2_SKEW units behind its ex2 half
#pragma unroll
      for (int un = 0; un < NU + CYTOF_X_A2_SKEW; ++un) {
        if (un < NU) {
          phase_b(un, ycur, mmask, tmask, amask);
          st[un] = phase_a1(un, ynxt, mm_n, tm_n);
        }
        if (un >= CYTOF_X_A2_SKEW) phase_a2(un - CYTOF_X_A2_SKEW, st[un - CYTOF_X_A2_SKEW], b ^ 1);
      }
    } else {
#pragma unroll
      for (int un = 0; un < NU; ++un) phase_b(un, ycur, mmask, tmask, amask);
    }
    if (MISSING && amask) {
#pragma unroll
      for (int un = 0; un < NU; ++un) missing_unit(un, ycur, mmask, tmask, t0);
    }
    mmask = mm_n;
    tmask = tm_n;
    amask = am_n;
    __syncwarp();
    if (r + 2 < nrounds) issue_tile(r + 2, b);
  }

  // ---------------- CTA reduction (fixed order) and partial record ----------------
  const PartialLayout pl = partial_layout(J, K, L0, L1);
  // sum w d = sum w y' - mu' sum w;  sum w d^2 = sum c y'^2 - sum_l mu'_l (2 sum w y' - mu'_l sum w)
  float swdf[2 * NPT];
  float swdd_l = sq;
#pragma unroll
  for (int x = 0; x < 2 * NPT; ++x) {
#if CYTOF_X_MUP_RELOAD
    // without the imputation path the centred means are only needed here: they are read again (the same expression as
    // in the prologue, through a load the compiler cannot merge with the first one) instead of holding 2 NPT
    // registers over the whole loop of a kernel that spills
    float m = (x & 1) ? hi2(mup[x / 2]) : lo2(mup[x / 2]);
    if (!MISSING) {
      const bool z1 = x >= LM0;
      const int l = z1 ? x - LM0 : x, Lz = z1 ? L1 : L0;
      m = 0.f;
      if (l < Lz && x < LM0 + LM1) {
        float v;
        asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(G + (z1 ? gl.mu1 : gl.mu0) + l));
        m = v - (z1 ? mc1 : mc0);
      }
    }
#else
    const float m = (x & 1) ? hi2(mup[x / 2]) : lo2(mup[x / 2]);
#endif
    const float a = ((x & 1) ? hi2(swm[x / 2]) : lo2(swm[x / 2]));
    const float bm = ((x & 1) ? hi2(swym[x / 2]) : lo2(swym[x / 2]));
    const float at = ((x & 1) ? hi2(swt[x / 2]) : lo2(swt[x / 2]));
    const float bt = ((x & 1) ? hi2(swyt[x / 2]) : lo2(swyt[x / 2]));
    const float dm = fmaf(-m, a, bm), dt = fmaf(-m, at, bt);
    swdd_l -= m * ((bm + dm) + (bt + dt));
    swdf[x] = warp_sum(dm + dt);
  }
  const float swdd_w = warp_sum(swdd_l);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    gW[r] += __shfl_xor_sync(FULL, gW[r], 1);
    gW[r] += __shfl_xor_sync(FULL, gW[r], 2);
  }
  sfr += __shfl_xor_sync(FULL, sfr, 1);
  sfr += __shfl_xor_sync(FULL, sfr, 2);
  sfo += __shfl_xor_sync(FULL, sfo, 1);
  sfo += __shfl_xor_sync(FULL, sfo, 2);
  ll2 += __shfl_xor_sync(FULL, ll2, 1);
  ll2 += __shfl_xor_sync(FULL, ll2, 2);
  lqy_l = warp_sum(lqy_l);
  lmiss_l = warp_sum(lmiss_l);
#pragma unroll
  for (int o = TP; o < 32; o <<= 1) {
    dvm_t += __shfl_xor_sync(FULL, dvm_t, o);
    dvls_t += __shfl_xor_sync(FULL, dvls_t, o);
    nm_t += __shfl_xor_sync(FULL, nm_t, o);
  }
  // the tail statistics of marker jt are spread over the CPT cell groups of the tail pass
  float swt_f[2 * NPT];
#pragma unroll
  for (int x = 0; x < 2 * NPT; ++x) {
    float v = (x & 1) ? hi2(swt[x / 2]) : lo2(swt[x / 2]);
#pragma unroll
    for (int o = TP; o < 32; o <<= 1) v += __shfl_xor_sync(FULL, v, o);
    swt_f[x] = v;
  }

  float* mine = sred + warp * pl.total;
#if CYTOF_X_TMEM
  // the MMAs of the last two rounds have completed (=> all of them)
  if (nrounds >= 1) umma_ok = mbar_wait(xbar + ((nrounds - 1) & 1), (uint32_t)((nrounds - 1) >> 1) & 1u) && umma_ok;
  if (nrounds >= 2) umma_ok = mbar_wait(xbar + (nrounds & 1), (uint32_t)((nrounds - 2) >> 1) & 1u) && umma_ok;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#endif
  __syncthreads();            // every warp is done with the tiles and the Z tables
#if CYTOF_X_TMEM
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#endif
  for (int t = lane; t < pl.total; t += 32) mine[t] = 0.f;
  __syncwarp();
#pragma unroll
  for (int x = 0; x < NC; ++x) {
    const float vm_ = (x & 1) ? hi2(swm[x / 2]) : lo2(swm[x / 2]);
    const bool z1 = x >= LM0;
    const int l = z1 ? x - LM0 : x;
    if (l < (z1 ? L1 : L0)) {
      const int base = (z1 ? pl.sw1 : pl.sw0) + l * J;
      if (okm) mine[base + lane] = vm_;
      if (lane < TP && tail_ok) mine[base + jt] = swt_f[x];
    }
  }
  if (MISSING && okm) {
    mine[pl.dvm + lane] = dvm_m;
    mine[pl.dvls + lane] = dvls_m;
    mine[pl.nm + lane] = nm_m;
  }
  if (MISSING) {
    if (lane < TP && tail_ok) {
      mine[pl.dvm + jt] = dvm_t;
      mine[pl.dvls + jt] = dvls_t;
      mine[pl.nm + jt] = nm_t;
    }
  }
#if CYTOF_X_TMEM
  // accumulator: lane = row (0..47 D_hi markers, 64..111 D_lo markers), column = feature (0..63 hi, 64..127 lo).
  // Warp w reads lanes 32 (w & 3) .. + 31 and the column half w >> 2; D_lo g_lo is dropped.
  {
    const int q = warp & 3, ch = warp >> 2;
    const int j = 32 * (q & 1) + lane;
    if (!(q >= 2 && ch == 1)) {
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t rr[16];
        const uint32_t taddr = tmem_acc + ((uint32_t)(32 * q) << 16) + (uint32_t)(64 * ch + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]),
              "=r"(rr[8]), "=r"(rr[9]), "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (j < J && nrounds > 0) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int k = c0 + c;
            if (k < K) mine[pl.dZ + k * J + j] += __uint_as_float(rr[c]);
          }
        }
      }
    }
  }
#else
  // dZ fragments of n-tile `warp`: (mt, r) -> marker 6g + 2mt + (r >> 1), feature 8 warp + 2 t4 + (r & 1)
#pragma unroll
  for (int mt = 0; mt < kXMTJ; ++mt)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int j = 6 * g + 2 * mt + (r >> 1), k = 8 * warp + 2 * t4 + (r & 1);
      if (j < J && k < K) mine[pl.dZ + k * J + j] = dZ[mt][r];
    }
#endif
  if (t4 == 0) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int k = 16 * (r >> 1) + g + 8 * (r & 1);
      if (k < K) mine[pl.gW + k] = gW[r];
    }
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
#if CYTOF_X_TMEM
    if (!umma_ok) md[0] = __longlong_as_double(0x7ff8000000000000ll);     // an MMA wait timed out: poison, never hang
#endif
  }
  __syncthreads();
  float* out = p.partials + (size_t)blockIdx.x * pl.total;
  for (int t = threadIdx.x; t < pl.total; t += kXThreads) {
    if (t >= pl.dbl && t < pl.dbl + 4) continue;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kXWarps; ++w) a += sred[w * pl.total + t];
    out[t] = a;
  }
  if (threadIdx.x < 2) {
    double a = 0.0;
    for (int w = 0; w < kXWarps; ++w) a += reinterpret_cast<const double*>(sred + w * pl.total + pl.dbl)[threadIdx.x];
    reinterpret_cast<double*>(out + pl.dbl)[threadIdx.x] = a;
  }
#if CYTOF_X_TMEM
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(tmem_acc) : "memory");
#endif
}

}  // namespace cytof
