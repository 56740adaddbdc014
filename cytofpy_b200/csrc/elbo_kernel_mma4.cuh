// Tensor-core fused ELBO kernel, FOUR cells per warp iteration, 12 warps per SM (J, K <= 32).
//
// Same arithmetic as elbo_kernel_mma.cuh (expanded-quadratic mixture forward, moment-based
// backward, TF32 hi/lo split MMAs for the three J x K contractions), re-tiled after measuring that
// kernel with clock64 (profiles/README.md, round-1 "phase clocks"): with 8 cells per warp the
// per-tile state (80 exponentials + 32 dZ accumulators per lane) costs 255 registers = 8 warps/SM,
// and both the contraction stage (a ~1200-cycle dependency chain) and the mixture stage ran at
// IPC 0.2-0.3 because only one other warp per scheduler was there to cover their latencies.
//   * A tile is 4 cells; the hi and lo halves of the TF32 split are two COLUMNS of the same MMA
//     (N = 8 = 4 cells x {hi, lo}), so one pass over the contraction index gives both partial
//     products (the accumulator chain is 4 MMAs deep instead of 8) and the tensor work per cell is
//     unchanged.  State per lane halves => <= 168 registers => 12 warps/SM, 3 per scheduler.
//   * D is split where it is produced (lane = marker, one value per lane and cell); the B
//     fragments then come straight out of ldmatrix with no arithmetic in the contraction stage.
//   * The second contraction runs on the UN-normalised softmax numerators e_k = 2^(f_k - max)
//     and its result is scaled per cell afterwards (c1 = fac rho / sum_k e_k * Z e), so it starts
//     right after the ex2 and overlaps the logsumexp / noisy-class scalar chain.
//   * dZ += D^T g is issued last; nothing in the tile's backward depends on it.
#pragma once
#include "elbo_kernel_mma.cuh"

namespace cytof {

#ifndef CYTOF_QTHREADS
#define CYTOF_QTHREADS 384
#endif
constexpr int kQThreads = CYTOF_QTHREADS;
constexpr int kQWarps = kQThreads / 32;
constexpr int kQT = 4;                   // cells per warp iteration
constexpr int kQTS = 36;                 // tile row stride (floats): conflict-free ldmatrix
// per-warp shared memory (floats): D hi/lo [8][36] | e hi/lo [8][36] | A0 / c1 [4][36] | y^2 [4][36] |
// 1/S0, 1/S1 [4][32] each | per-cell scalars [16] | y double buffer [2][4][32] | sum w d [4][32]
constexpr int kQWarpFloats = 2 * 8 * kQTS + 2 * kQT * kQTS + 2 * kQT * 32 + 16 + 2 * kQT * 32 + kQT * 32;

template <int LM0, int LM1, bool NOISY, bool HAS_IDX, bool MISSING>
__global__ void __launch_bounds__(kQThreads, 1)
elbo_fwdbwd_mma4_kernel(const ElboParams p) {
  pdl_enter();
  constexpr int NC = LM0 + LM1, NPT = (NC + 1) / 2;          // components, packed pairs
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t4 = lane & 3;                   // MMA fragment coordinates
  uint4* sZT = reinterpret_cast<uint4*>(smem);              // [mt][c][lane] A fragments of Z^T (k x j)
  uint4* sZ = sZT + 8 * 32;                                 // [mt][c][lane] A fragments of Z   (j x k)
  float* wbase = smem + kMmaZFloats + warp * kQWarpFloats;
  float* sD = wbase;                       // [8][kQTS] row 2c = hi(D) of cell c, row 2c+1 = lo(D)
  float* sG = sD + 8 * kQTS;               // [8][kQTS] row 2c = hi(e) of cell c, row 2c+1 = lo(e)
  float* sA = sG + 8 * kQTS;               // [4][kQTS] A0_j, later c1_j
  float* sY = sA + kQT * kQTS;             // [4][kQTS] y_j^2 (noisy class)
  float* sR0 = sY + kQT * kQTS;            // [4][32] 1/S0
  float* sR1 = sR0 + kQT * 32;             // [4][32] 1/S1
  float* sS = sR1 + kQT * 32;              // [0..3] fac*rho, [4..7] fac*(1-rho), [8..11] fac*rho/sum_k e_k
  float* sYb = sS + 16;                    // [2][4][32] y rows, double buffered
  float* sW = sYb + 2 * kQT * 32;          // [4][32] sum_l w d (imputation gradient)
  float* sred = smem;                      // final reduction aliases everything (after the loop)

  const int I = p.I, J = p.J, K = p.K, L0 = p.L0, L1 = p.L1;
  int i = 0;
  while (i + 1 < I && (int)blockIdx.x >= p.blk_begin[i + 1]) ++i;
  const int nb = p.blk_begin[i + 1] - p.blk_begin[i];
  const int lb = (int)blockIdx.x - p.blk_begin[i];
  const long long cb = p.cell_begin[i];
  const long long Bi = p.cell_begin[i + 1] - cb;
  const long long c_lo = Bi * lb / nb, c_hi = Bi * (lb + 1) / nb;
  const long long row_base = p.row_base[i];
  const unsigned long long grow_base = (unsigned long long)p.row_global[i];

  // ---------------- Z fragment tables (once per CTA) ----------------
  {
    const unsigned kmask = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
    for (int e = threadIdx.x; e < 2 * 8 * 32; e += kQThreads) {
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
  const float* __restrict__ G = p.globals;
  const float sig = G[gl.sig + i];
  const float inv_sig2 = 1.0f / (sig * sig);
  const float nh2 = -0.5f * inv_sig2 * kLog2e;
  const float cbase = (-logf(sig) - kHalfLog2Pi) * kLog2e;
  const float fac = p.fac[i];
  const float b0 = G[gl.b + 3 * i], b1 = G[gl.b + 3 * i + 1], b2 = G[gl.b + 3 * i + 2];
  const bool ok = lane < J;
  const float okf = ok ? 1.f : 0.f;
  const int jj = ok ? lane : 0;

  // mixture constants of the expanded quadratic (see elbo_kernel_mma.cuh): v_l = nh y'^2 + be_l y' + al_l
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
  float2 sw[NPT], swy[NPT];              // sum_n w, sum_n w y' per component
  float sq = 0.f;                        // sum_n (c0 y0'^2 + c1 y1'^2)
  float dZ[2][4][4];                     // C fragments of dZ (permuted rows / columns, see epilogue)
  float gW[4] = {0.f, 0.f, 0.f, 0.f};    // sum_n g_k for k = g, g+8, g+16, g+24 (this thread's cell)
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

  // ---------------- stream the cells of this CTA, 4 per warp iteration ----------------
  const int32_t* __restrict__ idx = p.idx + (HAS_IDX ? (cb + c_lo) : 0);
  const float* __restrict__ ysamp = p.y + (size_t)row_base * J;
  const int ncell = (int)(c_hi - c_lo);
  const int first_row = (int)c_lo;            // full batch: row = c_lo + cell
  constexpr int STRIDE = kQWarps * kQT;
  const bool vec16 = J == 32 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0;   // 128-byte rows

  // y rows of the tile starting at cell tn -> sYb[b]; rows past the end of the range and lanes
  // >= J are zero-filled, so the arithmetic below needs no validity selects
  auto issue_tile = [&](const int tn, const int b) {
    float* dst = sYb + b * (kQT * 32);
    if (vec16) {
      const int u = lane >> 3, part = lane & 7;
      if (tn + u < ncell) {
        const int row = HAS_IDX ? __ldg(idx + tn + u) : first_row + tn + u;
        cp_async16(dst + u * 32 + 4 * part, ysamp + (size_t)row * 32 + 4 * part);
      } else {
        *reinterpret_cast<float4*>(dst + u * 32 + 4 * part) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
#pragma unroll
      for (int u = 0; u < kQT; ++u) {
        if (ok && tn + u < ncell) {
          const int row = HAS_IDX ? __ldg(idx + tn + u) : first_row + tn + u;
          cp_async4(dst + u * 32 + lane, ysamp + (size_t)row * J + lane);
        } else {
          dst[u * 32 + lane] = 0.f;
        }
      }
    }
    cp_async_commit();
  };

  // imputation of missing entries (vae.py:17-33): the only branchy part, kept apart so that the
  // mixture arithmetic is one long basic block.  The N(0,1) draw replaces the NaN in the y buffer.
  auto prepass = [&](const int tn, const int b, unsigned& mmask, unsigned& amask) {
    float* src = sYb + b * (kQT * 32);
    mmask = 0u;
#pragma unroll
    for (int u = 0; u < kQT; ++u) {
      const float yr = src[u * 32 + lane];
      mmask |= (yr != yr) ? (1u << u) : 0u;
    }
    amask = __reduce_or_sync(FULL, mmask);       // cells of the tile with at least one missing marker
    if (amask) {
      unsigned long long grp_c = ~0ull;          // Philox block in the cache: 4 consecutive global rows
      float nz[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1      // a real loop: ONE copy of the Philox rounds in the instruction stream
      for (int u = 0; u < kQT; ++u) {
        if (!((amask >> u) & 1u)) continue;      // warp-uniform
        const int row = HAS_IDX ? __ldg(idx + tn + u) : first_row + tn + u;
        float e = 0.f;
        if (p.noise_mode == CYTOF_NOISE_EXTERNAL) {
          if (p.eps) e = __ldg(p.eps + (cb + c_lo + tn + u) * J + lane);
        } else {
          const unsigned long long grow = grow_base + (unsigned long long)row;
          if ((grow >> 2) != grp_c) {            // warp-uniform
            grp_c = grow >> 2;
            philox_normal4(p.seed, elbo_step(p), grp_c, philox_ctr(lane, i), nz);
          }
          const unsigned w = (unsigned)grow & 3u;
          e = w == 0u ? nz[0] : (w == 1u ? nz[1] : (w == 2u ? nz[2] : nz[3]));
        }
        if ((mmask >> u) & 1u) src[u * 32 + lane] = e;
      }
    }
  };

  float2 ee[kQT][NPT];      // exp(a_l - max) of the tile in flight, lane = marker j

  // mixture log-densities of cell u, forward (Model.py:223-233), log2 domain, component pairs
  auto phase_a = [&](const int u, const float* src, const unsigned mmask) {
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
#pragma unroll
    for (int q = 0; q < NPT; ++q) constexpr_pair_max<LM0, NC>(q, v[q], M0, M1);
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
    const float S0 = (lo2(s0) + hi2(s0)) + s0x, S1 = (lo2(s1) + hi2(s1)) + s1x;
    const float A0 = fmaf(nh2 * y0, y0, M0) + fast_lg2(S0), A1 = fmaf(nh2 * y1, y1, M1) + fast_lg2(S1);
    sR0[u * 32 + lane] = fast_rcp(S0);
    sR1[u * 32 + lane] = fast_rcp(S1);
    const float D = okf * (A1 - A0);
    uint32_t dh, dl;
    split_tf32(D, dh, dl);
    sD[(2 * u) * kQTS + lane] = __uint_as_float(dh);
    sD[(2 * u + 1) * kQTS + lane] = __uint_as_float(dl);
    sA[u * kQTS + lane] = okf * A0;
    if (NOISY) sY[u * kQTS + lane] = yv * yv;
  };

  // mixtures of cell u, backward: w_zl = c_z p_zl; statistics sum w and sum w y' per component,
  // sum c_z y_z'^2 per cell (sum_l w_zl = c_z) -- three FMA-pipe ops per component pair
  auto phase_b = [&](const int u, const float* src, const unsigned mmask, const unsigned amask) {
    const float c1 = okf * sA[u * kQTS + lane];
    const float c0 = okf * sS[u] - c1;
    const float x = src[u * 32 + lane];
    const float yv = (MISSING && ((mmask >> u) & 1u)) ? fmaf(sd, x, vm) : x;
    const float y0 = yv - mc0, y1 = yv - mc1;
    const float2 Y00 = bcp(y0), Y11 = bcp(y1), Y01 = mkp(y0, y1);
    const float r0 = c0 * sR0[u * 32 + lane], r1 = c1 * sR1[u * 32 + lane];
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

  int t0 = warp * kQT;
  int cur = 0;
  unsigned missmask = 0u, anymask = 0u;
  if (t0 < ncell) {       // prologue: phase A of this warp's first tile
    issue_tile(t0, 0);
    cp_async_wait<0>();
    __syncwarp();
    if (MISSING) prepass(t0, 0, missmask, anymask);
#pragma unroll
    for (int u = 0; u < kQT; ++u) phase_a(u, sYb, missmask);
    if (t0 + STRIDE < ncell) issue_tile(t0 + STRIDE, 1);
  }
  __syncthreads();   // Z tables visible

  for (; t0 < ncell; t0 += STRIDE, cur ^= 1) {
    const bool has_next = t0 + STRIDE < ncell;
    const bool valid = t0 + t4 < ncell;      // this thread's cell in the accumulator layout
    __syncwarp();

    // ===== contraction 1: f^T[k, (n, hi|lo)] = Z^T . [D_hi | D_lo]^T  (+ log W on the hi column)
    float fT[2][4];
    {
      uint32_t r0[4], r1[4];       // B fragments: block m = markers 4m..4m+3, column g = row g of sD
      ldsm_x4(sD + (lane & 7) * kQTS + 4 * (lane >> 3), r0);
      ldsm_x4(sD + (lane & 7) * kQTS + 16 + 4 * (lane >> 3), r1);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        fT[mt][0] = lwk[2 * mt];
        fT[mt][2] = lwk[2 * mt + 1];
        fT[mt][1] = fT[mt][3] = 0.f;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint4 a = sZT[(mt * 4 + c) * 32 + lane];
          if (c < 2) mma_tf32(fT[mt], a, r0[2 * c], r0[2 * c + 1]);
          else mma_tf32(fT[mt], a, r1[2 * c - 4], r1[2 * c - 3]);
        }
      }
    }
    // accumulator layout: column 2 t4 = hi part, 2 t4 + 1 = lo part of cell t4; rows g, g+8 of tile mt
    float f[4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      f[2 * mt] = fT[mt][0] + fT[mt][1];        // feature 16 mt + g
      f[2 * mt + 1] = fT[mt][2] + fT[mt][3];    // feature 16 mt + g + 8
    }
    float mx = fmaxf(fmaxf(f[0], f[1]), fmaxf(f[2], f[3]));
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o));
    float ex[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) ex[r] = fast_ex2(f[r] - mx);
    // un-normalised numerators, split, to the e tile (rows 2 t4, 2 t4 + 1); feature of ex[r] = g + 8 r
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      uint32_t eh, el;
      split_tf32(ex[r], eh, el);
      sG[(2 * t4) * kQTS + g + 8 * r] = __uint_as_float(eh);
      sG[(2 * t4 + 1) * kQTS + g + 8 * r] = __uint_as_float(el);
    }
    __syncwarp();

    // ===== contraction 2 on the numerators: c1'^T[j, (n, hi|lo)] = Z . [e_hi | e_lo]^T
    float cT[2][4];
    {
      uint32_t r0[4], r1[4];
      ldsm_x4(sG + (lane & 7) * kQTS + 4 * (lane >> 3), r0);
      ldsm_x4(sG + (lane & 7) * kQTS + 16 + 4 * (lane >> 3), r1);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) cT[mt][0] = cT[mt][1] = cT[mt][2] = cT[mt][3] = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint4 a = sZ[(mt * 4 + c) * 32 + lane];
          if (c < 2) mma_tf32(cT[mt], a, r0[2 * c], r0[2 * c + 1]);
          else mma_tf32(cT[mt], a, r1[2 * c - 4], r1[2 * c - 3]);
        }
      }
    }
    // ===== per-cell scalars (overlap the MMAs above): logsumexp over K, noisy class
    float rs = (ex[0] + ex[1]) + (ex[2] + ex[3]);
    float ra, ry = 0.f;
    {
      const float4 a4 = *reinterpret_cast<const float4*>(sA + t4 * kQTS + 4 * g);
      ra = (a4.x + a4.y) + (a4.z + a4.w);
      if (NOISY) {
        const float4 c4 = *reinterpret_cast<const float4*>(sY + t4 * kQTS + 4 * g);
        ry = (c4.x + c4.y) + (c4.z + c4.w);
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
#pragma unroll
    for (int r = 0; r < 4; ++r) gW[r] = fmaf(sc, ex[r], gW[r]);
    if (g == 0) {
      sS[t4] = fr;
      sS[4 + t4] = facc * omr;
      sS[8 + t4] = sc;
    }
    // c1 = sc * (hi + lo columns), to the A0 tile (its partial sums were taken above)
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      sA[t4 * kQTS + 16 * mt + g] = sc * (cT[mt][0] + cT[mt][1]);
      sA[t4 * kQTS + 16 * mt + g + 8] = sc * (cT[mt][2] + cT[mt][3]);
    }
    __syncwarp();

    // ===== contraction 3: dZ[j, k] += sum_n D[n, j] g[n, k], contraction index = (cell, hi|lo of D)
    {
      // A = [D_hi(cell t4) | D_lo(cell t4)]: rows g / g+8 of m-tile mt = markers 4g+2mt / 4g+2mt+1
      const float4 dh4 = *reinterpret_cast<const float4*>(sD + (2 * t4) * kQTS + 4 * g);
      const float4 dl4 = *reinterpret_cast<const float4*>(sD + (2 * t4 + 1) * kQTS + 4 * g);
      const uint4 a0 = make_uint4(__float_as_uint(dh4.x), __float_as_uint(dh4.y), __float_as_uint(dl4.x), __float_as_uint(dl4.y));
      const uint4 a1 = make_uint4(__float_as_uint(dh4.z), __float_as_uint(dh4.w), __float_as_uint(dl4.z), __float_as_uint(dl4.w));
      const uint4 a0h = make_uint4(a0.x, a0.y, 0u, 0u), a1h = make_uint4(a1.x, a1.y, 0u, 0u);   // D_hi only
      // B rows t4 and t4 + 4 (both = cell t4): g = sc e, split; column g of n-tile nt = feature 4g+nt
      const float4 eh4 = *reinterpret_cast<const float4*>(sG + (2 * t4) * kQTS + 4 * g);
      const float4 el4 = *reinterpret_cast<const float4*>(sG + (2 * t4 + 1) * kQTS + 4 * g);
      const float sct = sS[8 + t4];
      const float gv[4] = {sct * (eh4.x + el4.x), sct * (eh4.y + el4.y), sct * (eh4.z + el4.z), sct * (eh4.w + el4.w)};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t bh, bl;
        split_tf32(gv[nt], bh, bl);
        mma_tf32(dZ[0][nt], a0, bh, bh);      // D_hi g_hi + D_lo g_hi
        mma_tf32(dZ[1][nt], a1, bh, bh);
        mma_tf32(dZ[0][nt], a0h, bl, 0u);     // D_hi g_lo
        mma_tf32(dZ[1][nt], a1h, bl, 0u);
      }
    }
    __syncwarp();

    // ===== phase B of this tile, software-pipelined with phase A of the warp's NEXT tile
    const float* ycur = sYb + cur * (kQT * 32);
    unsigned miss_n = 0u, any_n = 0u;
    if (has_next) {
      const float* ynxt = sYb + (cur ^ 1) * (kQT * 32);
      cp_async_wait<0>();
      __syncwarp();
      if (MISSING) prepass(t0 + STRIDE, cur ^ 1, miss_n, any_n);
#pragma unroll
      for (int u = 0; u < kQT; ++u) {
        phase_b(u, ycur, missmask, anymask);
        phase_a(u, ynxt, miss_n);
      }
    } else {
#pragma unroll
      for (int u = 0; u < kQT; ++u) phase_b(u, ycur, missmask, anymask);
    }
    // logistic missingness (Model.py:269-274) and d/dy chain into the imputation parameters
    if (MISSING && anymask) {
#pragma unroll
      for (int u = 0; u < kQT; ++u) {     // branch-free: lanes without a missing entry add 0
        const bool miss = (missmask >> u) & 1u;
        const bool vu = t0 + u < ncell;
        const float mf = (miss && vu) ? 1.f : 0.f;
        const float x = ycur[u * 32 + lane];
        const float e = miss ? x : 0.f;
        const float yv = miss ? fmaf(sd, e, vm) : x;
        const float fu = vu ? fac : 0.f;
        const float uu = fmaf(fmaf(b2, yv, b1), yv, b0);
        const float tm = fast_ex2(-fabsf(uu) * kLog2e);
        const float rr = fast_rcp(1.f + tm);
        const float ls = -(fmaxf(-uu, 0.f) + kLn2 * fast_lg2(1.f + tm));
        const float sneg = (uu > 0.f ? tm : 1.f) * rr;   // 1 - sigmoid(u)
        float dy = -sW[u * 32 + lane] * inv_sig2 + fu * sneg * fmaf(2.f * b2, yv, b1);
        if (NOISY) dy = fmaf(-sS[4 + u] * inv_s2n, yv, dy);
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
  }

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
  // accumulator-layout sums: this thread's cell is t4, so features and per-cell scalars are summed
  // over the 4 lanes of a quad (the 8 quads g hold replicas of the per-cell scalars: take g == 0)
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
  float* out = p.partials + (size_t)blockIdx.x * pl.total;
  for (int t = threadIdx.x; t < pl.total; t += kQThreads) {
    if (t >= pl.dbl && t < pl.dbl + 4) continue;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kQWarps; ++w) a += sred[w * pl.total + t];
    out[t] = a;
  }
  if (threadIdx.x < 2) {
    double a = 0.0;
    for (int w = 0; w < kQWarps; ++w) a += reinterpret_cast<const double*>(sred + w * pl.total + pl.dbl)[threadIdx.x];
    reinterpret_cast<double*>(out + pl.dbl)[threadIdx.x] = a;
  }
}

}  // namespace cytof
