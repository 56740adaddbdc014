// sm_100a-tuned variant of the fused ELBO kernel for J <= 32 and K <= 32 (BASELINE cfg1-cfg4).
//
// Same mapping as elbo_fwdbwd_kernel (one warp per cell, lane = marker j, lane = feature k for
// the Z-gated sum) with two changes that ncu asked for (profiles/README.md):
//   * TWO cells are in flight per warp iteration: their instruction streams are independent, so
//     the MUFU / shuffle / shared-memory latencies of one cell are covered by the other (the
//     first kernel issued on only 52 % of cycles with 2 warps per scheduler);
//   * all mixture and contraction arithmetic is issued as Blackwell packed-fp32 instructions
//     (FADD2 / FMUL2 / FFMA2, PTX add/mul/fma.rn.f32x2): pairs of mixture components (l, l+1) in
//     the mixtures, pairs of markers / features in the three J x K contractions.  Same FP32 work,
//     about 40 % fewer issue slots.
#pragma once
#include "common.cuh"

namespace cytof {

struct f2 { unsigned long long v; };   // two packed fp32: lo = bits 31:0, hi = bits 63:32
__device__ __forceinline__ f2 mk2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo2(f2 a) { float l, h; asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(h) : "l"(a.v)); (void)h; return l; }
__device__ __forceinline__ float hi2(f2 a) { float l, h; asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(h) : "l"(a.v)); (void)l; return h; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ f2 bc2(float a) { return mk2(a, a); }
// in-place accumulators ("+l"): keeps ptxas from copying 64-bit register pairs every iteration
__device__ __forceinline__ void acc_add2(f2& acc, f2 x) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc.v) : "l"(x.v)); }
__device__ __forceinline__ void acc_fma2(f2& acc, f2 a, f2 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc.v) : "l"(a.v), "l"(b.v)); }
// schedulable (non-volatile) forms of the warp reductions: the two cells in flight must overlap
__device__ __forceinline__ float warp_max_redux_nv(float v) {
  float r;
  asm("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float warp_sum_unit_nv(float v) {
  unsigned q = __float2uint_rn(v * 16777216.0f), r;
  asm("redux.sync.add.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(q));
  return (float)r * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float ld_stream_nv(const float* p) {
  float r;
  asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ f2 shfl_xor2(f2 a, int o) {
  return mk2(__shfl_xor_sync(0xffffffffu, lo2(a), o), __shfl_xor_sync(0xffffffffu, hi2(a), o));
}

#ifndef CYTOF_CPW
#define CYTOF_CPW 1
#endif
#ifndef CYTOF_J32_MINBLOCKS
#define CYTOF_J32_MINBLOCKS 1
#endif
#ifndef CYTOF_CONST_SMEM
#define CYTOF_CONST_SMEM 0   // 1: mixture constants (mu, log eta) read from shared memory, not registers
#endif
constexpr int kCPW = CYTOF_CPW;   // cells in flight per warp
constexpr int kJ32ConstFloats = 2 * (2 * 4) + 64 * (2 * 4);   // smem for mu pairs + per-lane cz pairs (L <= 8)

template <int LM0, int LM1, bool NOISY, bool HAS_IDX>
__global__ void __launch_bounds__(kThreads, CYTOF_J32_MINBLOCKS)
elbo_fwdbwd_j32_kernel(const ElboParams p) {
  pdl_enter();
  constexpr int NP0 = (LM0 + 1) / 2, NP1 = (LM1 + 1) / 2;   // component pairs per mixture
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  f2* s_mu = reinterpret_cast<f2*>(smem);                 // [NP0 + NP1] broadcast
  f2* s_cz = reinterpret_cast<f2*>(smem + 16);            // [NP0 + NP1][32 lanes]
  float* sD = smem + kJ32ConstFloats + warp * (kCPW * 64);   // [cell][32] D rows
  float* sG = sD + kCPW * 32;                                // [cell][32] g rows
  float* sred = smem + kJ32ConstFloats + kWarps * (kCPW * 64);

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

  f2 mu0[NP0], mu1[NP1], cz0[NP0], cz1[NP1];
  {
    float m[2 * NP0], c[2 * NP0];
#pragma unroll
    for (int l = 0; l < 2 * NP0; ++l) {
      m[l] = l < L0 ? G[gl.mu0 + l] : 0.f;
      c[l] = (ok && l < L0) ? fmaf(G[gl.le0 + (i * J + jj) * L0 + (l < L0 ? l : 0)], kLog2e, cbase)
                            : (l == 0 ? 0.f : kNegBig);
    }
#pragma unroll
    for (int q = 0; q < NP0; ++q) { mu0[q] = mk2(m[2 * q], m[2 * q + 1]); cz0[q] = mk2(c[2 * q], c[2 * q + 1]); }
  }
  {
    float m[2 * NP1], c[2 * NP1];
#pragma unroll
    for (int l = 0; l < 2 * NP1; ++l) {
      m[l] = l < L1 ? G[gl.mu1 + l] : 0.f;
      c[l] = (ok && l < L1) ? fmaf(G[gl.le1 + (i * J + jj) * L1 + (l < L1 ? l : 0)], kLog2e, cbase)
                            : (l == 0 ? 0.f : kNegBig);
    }
#pragma unroll
    for (int q = 0; q < NP1; ++q) { mu1[q] = mk2(m[2 * q], m[2 * q + 1]); cz1[q] = mk2(c[2 * q], c[2 * q + 1]); }
  }
  if (CYTOF_CONST_SMEM) {
    if (warp == 0) {
#pragma unroll
      for (int q = 0; q < NP0; ++q) { s_cz[q * 32 + lane] = cz0[q]; if (lane == 0) s_mu[q] = mu0[q]; }
#pragma unroll
      for (int q = 0; q < NP1; ++q) { s_cz[(NP0 + q) * 32 + lane] = cz1[q]; if (lane == 0) s_mu[NP0 + q] = mu1[q]; }
    }
    __syncthreads();
  }
#if CYTOF_CONST_SMEM
#define MU0(q) s_mu[q]
#define MU1(q) s_mu[NP0 + (q)]
#define CZ0(q) s_cz[(q) * 32 + lane]
#define CZ1(q) s_cz[(NP0 + (q)) * 32 + lane]
#else
#define MU0(q) mu0[q]
#define MU1(q) mu1[q]
#define CZ0(q) cz0[q]
#define CZ1(q) cz1[q]
#endif
  const float vm = G[gl.vm + i * J + jj], vls = G[gl.vls + i * J + jj], sd = expf(vls);
  // Z as packed 0/1 floats: column `lane` over j (forward), row `lane` over k (backward)
  f2 zc[16], zr[16];
  {
    const unsigned kmask = K >= 32 ? 0xffffffffu : ((1u << K) - 1u);
    const unsigned zrow = ok ? ((unsigned)p.zmask[jj] & kmask) : 0u;
    unsigned zcol = 0u;     // column `lane` of Z over j = ballot of bit `lane` of every lane's row
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const unsigned bal = __ballot_sync(FULL, (zrow >> k) & 1u);
      if (lane == k) zcol = bal;
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      zc[q] = mk2((float)((zcol >> (2 * q)) & 1u), (float)((zcol >> (2 * q + 1)) & 1u));
      zr[q] = mk2((float)((zrow >> (2 * q)) & 1u), (float)((zrow >> (2 * q + 1)) & 1u));
    }
  }
  const float lw2 = lane < K ? G[gl.lw + i * K + lane] * kLog2e : kNegBig;
  float l1me2 = 0.f, zconst2 = 0.f, ninv2s2 = 0.f, inv_s2n = 0.f;
  if (NOISY) {
    const float e_i = G[gl.eps + i];
    l1me2 = log1pf(-e_i) * kLog2e;
    zconst2 = ((float)J * (-kHalfLog2Pi - logf(p.noisy_sd)) + logf(e_i)) * kLog2e;
    inv_s2n = 1.0f / (p.noisy_sd * p.noisy_sd);
    ninv2s2 = -0.5f * inv_s2n * kLog2e;
  }

  // ---------------- accumulators ----------------
  f2 sw0[NP0], sw1[NP1], swd0[NP0], swd1[NP1], dZ[16];
  f2 swdd = bc2(0.f);
  float gW = 0.f, dvm = 0.f, dvls = 0.f, nm = 0.f, sfr = 0.f, sfo = 0.f, lqy_l = 0.f, lmiss_l = 0.f;
  double ll2 = 0.0;
#pragma unroll
  for (int q = 0; q < NP0; ++q) sw0[q] = swd0[q] = bc2(0.f);
#pragma unroll
  for (int q = 0; q < NP1; ++q) sw1[q] = swd1[q] = bc2(0.f);
#pragma unroll
  for (int q = 0; q < 16; ++q) dZ[q] = bc2(0.f);

  // ---------------- stream the cells of this CTA, two per warp iteration ----------------
  // cell counters are 32-bit offsets inside this CTA's range; rows are 32-bit (< 2^31 rows)
  const int32_t* __restrict__ idx = p.idx + (HAS_IDX ? (cb + c_lo) : 0);
  const float* __restrict__ ybase = p.y + (size_t)row_base * J + lane;
  const int ncell = (int)(c_hi - c_lo);
  const int first_row = (int)c_lo;            // full batch: row = c_lo + t
  constexpr int STRIDE = kWarps * kCPW;
  int t = warp * kCPW;
  int row_cur[kCPW], row_nxt[kCPW];
  float y_cur[kCPW];
#pragma unroll
  for (int u = 0; u < kCPW; ++u) {
    row_cur[u] = row_nxt[u] = 0;
    y_cur[u] = 0.f;
    if (t + u < ncell) {
      row_cur[u] = HAS_IDX ? __ldg(idx + t + u) : first_row + t + u;
      if (ok) y_cur[u] = ld_stream_nv(ybase + (size_t)row_cur[u] * J);
    }
    if (HAS_IDX && t + STRIDE + u < ncell) row_nxt[u] = __ldg(idx + t + STRIDE + u);
  }

  for (; t < ncell; t += STRIDE) {
    float y_nxt[kCPW];
    int row_nn[kCPW];
#pragma unroll
    for (int u = 0; u < kCPW; ++u) {
      y_nxt[u] = 0.f;
      row_nn[u] = 0;
      if (!HAS_IDX) row_nxt[u] = first_row + t + STRIDE + u;
      if (t + STRIDE + u < ncell && ok) y_nxt[u] = ld_stream_nv(ybase + (size_t)row_nxt[u] * J);
      if (HAS_IDX && t + 2 * STRIDE + u < ncell) row_nn[u] = __ldg(idx + t + 2 * STRIDE + u);
    }

    // ---- per-cell state; every stage is written as a loop over the cells in flight so that
    // the two instruction streams are adjacent in program order and overlap each other's latency
    float facc[kCPW], validf[kCPW], yv[kCPW], e[kCPW], Dj[kCPW], rS0[kCPW], rS1[kCPW];
    bool miss[kCPW], anym[kCPW];
    f2 e0[kCPW][NP0], e1[kCPW][NP1], red[kCPW];
    const f2 nh = bc2(nh2);

#pragma unroll
    for (int u = 0; u < kCPW; ++u) {
      const bool valid = t + u < ncell;
      validf[u] = valid ? 1.f : 0.f;
      facc[u] = valid ? fac : 0.f;
      // ---- imputation of missing entries (vae.py:17-33)
      miss[u] = ok && (y_cur[u] != y_cur[u]);
      anym[u] = __any_sync(FULL, miss[u]);
      e[u] = 0.f;
      yv[u] = ok ? y_cur[u] : 0.f;
      if (anym[u] && miss[u]) {
        e[u] = p.noise_mode == CYTOF_NOISE_EXTERNAL
                   ? (p.eps ? __ldg(p.eps + (cb + c_lo + t + u) * J + lane) : 0.f)
                   : philox_normal(p.seed, elbo_step(p), grow_base + (unsigned long long)row_cur[u], philox_ctr(lane, i));
        yv[u] = fmaf(sd, e[u], vm);
      }
    }
    // ---- mixture log-densities, forward (Model.py:223-233), log2 domain, component pairs
    f2 v0[kCPW][NP0], v1[kCPW][NP1];
    float M0[kCPW], M1[kCPW];
#pragma unroll
    for (int u = 0; u < kCPW; ++u) {
      const f2 y2 = bc2(yv[u]);
      M0[u] = M1[u] = kNegBig;
#pragma unroll
      for (int q = 0; q < NP0; ++q) {
        const f2 d = sub2(y2, MU0(q));
        v0[u][q] = fma2(mul2(d, nh), d, CZ0(q));
        M0[u] = fmaxf(M0[u], lo2(v0[u][q]));
        if (2 * q + 1 < LM0) M0[u] = fmaxf(M0[u], hi2(v0[u][q]));
      }
#pragma unroll
      for (int q = 0; q < NP1; ++q) {
        const f2 d = sub2(y2, MU1(q));
        v1[u][q] = fma2(mul2(d, nh), d, CZ1(q));
        M1[u] = fmaxf(M1[u], lo2(v1[u][q]));
        if (2 * q + 1 < LM1) M1[u] = fmaxf(M1[u], hi2(v1[u][q]));
      }
    }
    float S0[kCPW], S1[kCPW];
#pragma unroll
    for (int u = 0; u < kCPW; ++u) {
      f2 s0 = bc2(0.f), s1 = bc2(0.f);
#pragma unroll
      for (int q = 0; q < NP0; ++q) {
        const f2 tt = sub2(v0[u][q], bc2(M0[u]));
        e0[u][q] = mk2(fast_ex2(lo2(tt)), (2 * q + 1 < LM0) ? fast_ex2(hi2(tt)) : 0.f);
        s0 = add2(s0, e0[u][q]);
      }
#pragma unroll
      for (int q = 0; q < NP1; ++q) {
        const f2 tt = sub2(v1[u][q], bc2(M1[u]));
        e1[u][q] = mk2(fast_ex2(lo2(tt)), (2 * q + 1 < LM1) ? fast_ex2(hi2(tt)) : 0.f);
        s1 = add2(s1, e1[u][q]);
      }
      S0[u] = lo2(s0) + hi2(s0);
      S1[u] = lo2(s1) + hi2(s1);
    }
#pragma unroll
    for (int u = 0; u < kCPW; ++u) {
      const float A0 = M0[u] + fast_lg2(S0[u]), A1 = M1[u] + fast_lg2(S1[u]);
      rS0[u] = fast_rcp(S0[u]);
      rS1[u] = fast_rcp(S1[u]);
      Dj[u] = okf * (A1 - A0);
      red[u] = mk2(okf * A0, yv[u] * yv[u]);     // {sum_j A0_j, sum_j y_j^2}
      sD[u * 32 + lane] = Dj[u];
    }
    // ---- warp reductions of {A0, y^2}, both cells interleaved
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < kCPW; ++u) red[u] = add2(red[u], shfl_xor2(red[u], o));
    }
    __syncwarp();

    // ---- Z-gated J->K sum + log W (Model.py:246-252): lane = feature k
    float f[kCPW], mx[kCPW], ex[kCPW], ssum[kCPW], lse2[kCPW], rinv[kCPW];
#pragma unroll
    for (int u = 0; u < kCPW; ++u) {
      f2 a = bc2(0.f), b = bc2(0.f);
      const ulonglong2* row = reinterpret_cast<const ulonglong2*>(sD + u * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const ulonglong2 d4 = row[q];
        f2 dx, dy;
        dx.v = d4.x;
        dy.v = d4.y;
        acc_fma2(a, zc[2 * q], dx);
        acc_fma2(b, zc[2 * q + 1], dy);
      }
      a = add2(a, b);
      f[u] = (lo2(a) + hi2(a)) + lw2;      // sum_j A0_j is common to all k: added after the softmax
    }
#pragma unroll
    for (int u = 0; u < kCPW; ++u) mx[u] = warp_max_redux_nv(f[u]);
#pragma unroll
    for (int u = 0; u < kCPW; ++u) ex[u] = fast_ex2(f[u] - mx[u]);
#pragma unroll
    for (int u = 0; u < kCPW; ++u) ssum[u] = warp_sum_unit_nv(ex[u]);
#pragma unroll
    for (int u = 0; u < kCPW; ++u) {
      lse2[u] = (mx[u] + lo2(red[u])) + fast_lg2(ssum[u]);
      rinv[u] = fast_rcp(ssum[u]);
    }
    // ---- noisy-cell class (Model.py:259-264)
    float rho[kCPW], L2[kCPW], omr[kCPW], fr[kCPW], gk[kCPW];
    if (NOISY) {
      float q2[kCPW], z2[kCPW], dd[kCPW], tq[kCPW], r1[kCPW];
#pragma unroll
      for (int u = 0; u < kCPW; ++u) {
        q2[u] = lse2[u] + l1me2;
        z2[u] = fmaf(ninv2s2, hi2(red[u]), zconst2);
        dd[u] = q2[u] - z2[u];
        tq[u] = fast_ex2(-fabsf(dd[u]));
      }
#pragma unroll
      for (int u = 0; u < kCPW; ++u) {
        r1[u] = fast_rcp(1.f + tq[u]);
        L2[u] = fmaxf(q2[u], z2[u]) + fast_lg2(1.f + tq[u]);
      }
#pragma unroll
      for (int u = 0; u < kCPW; ++u) {
        rho[u] = dd[u] >= 0.f ? r1[u] : tq[u] * r1[u];
        omr[u] = dd[u] >= 0.f ? tq[u] * r1[u] : r1[u];
      }
    } else {
#pragma unroll
      for (int u = 0; u < kCPW; ++u) { rho[u] = 1.f; omr[u] = 0.f; L2[u] = lse2[u]; }
    }
#pragma unroll
    for (int u = 0; u < kCPW; ++u) {
      ll2 += (double)(L2[u] * validf[u]);
      fr[u] = facc[u] * rho[u];
      sfr += fr[u];
      sfo = fmaf(facc[u], omr[u], sfo);
      gk[u] = fr[u] * rinv[u] * ex[u];       // g_k = fac rho softmax_k(f)
      gW += gk[u];
      sG[u * 32 + lane] = gk[u];
    }
    __syncwarp();

    // ---- backward: lane = marker j again
    float c1[kCPW], c0[kCPW];
#pragma unroll
    for (int u = 0; u < kCPW; ++u) {
      f2 ca = bc2(0.f), cb2 = bc2(0.f);
      const f2 D2 = bc2(Dj[u]);
      const ulonglong2* row = reinterpret_cast<const ulonglong2*>(sG + u * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const ulonglong2 g4 = row[q];
        f2 gx, gy;
        gx.v = g4.x;
        gy.v = g4.y;
        acc_fma2(ca, zr[2 * q], gx);
        acc_fma2(cb2, zr[2 * q + 1], gy);
        acc_fma2(dZ[2 * q], gx, D2);
        acc_fma2(dZ[2 * q + 1], gy, D2);
      }
      ca = add2(ca, cb2);
      c1[u] = okf * (lo2(ca) + hi2(ca));
      c0[u] = okf * fr[u] - c1[u];
    }
    // mixtures, backward: w_zl = c_z p_zl; statistics sum w, sum w d, sum w d^2 (d recomputed)
    f2 wds[kCPW];
#pragma unroll
    for (int u = 0; u < kCPW; ++u) {
      const f2 y2 = bc2(yv[u]);
      const f2 r0 = bc2(c0[u] * rS0[u]), r1 = bc2(c1[u] * rS1[u]);
      wds[u] = bc2(0.f);
#pragma unroll
      for (int q = 0; q < NP0; ++q) {
        const f2 d = sub2(y2, MU0(q));
        const f2 w = mul2(r0, e0[u][q]);
        const f2 wd = mul2(w, d);
        acc_add2(sw0[q], w);
        acc_add2(swd0[q], wd);
        acc_fma2(swdd, wd, d);
        acc_add2(wds[u], wd);
      }
#pragma unroll
      for (int q = 0; q < NP1; ++q) {
        const f2 d = sub2(y2, MU1(q));
        const f2 w = mul2(r1, e1[u][q]);
        const f2 wd = mul2(w, d);
        acc_add2(sw1[q], w);
        acc_add2(swd1[q], wd);
        acc_fma2(swdd, wd, d);
        acc_add2(wds[u], wd);
      }
    }
    // logistic missingness (Model.py:269-274) and d/dy chain into the imputation parameters
#pragma unroll
    for (int u = 0; u < kCPW; ++u) {
      if (anym[u]) {
        const float wdsum = lo2(wds[u]) + hi2(wds[u]);
        const float uu = fmaf(fmaf(b2, yv[u], b1), yv[u], b0);
        const float tm = fast_ex2(-fabsf(uu) * kLog2e);
        const float rr = fast_rcp(1.f + tm);
        const float ls = -(fmaxf(-uu, 0.f) + kLn2 * fast_lg2(1.f + tm));
        const float sneg = (uu > 0.f ? tm : 1.f) * rr;   // 1 - sigmoid(u)
        float dy = -wdsum * inv_sig2 + facc[u] * sneg * fmaf(2.f * b2, yv[u], b1);
        if (NOISY) dy = fmaf(-facc[u] * omr[u] * inv_s2n, yv[u], dy);
        if (miss[u]) {
          lmiss_l = fmaf(ls, validf[u], lmiss_l);
          dvm += dy;
          dvls = fmaf(dy * sd, e[u], dvls);
          nm += validf[u];
          lqy_l = fmaf(fmaf(-0.5f * e[u], e[u], -vls - kHalfLog2Pi), validf[u], lqy_l);
        }
      }
    }
    __syncwarp();   // sD / sG are rewritten by the next pair of cells

#pragma unroll
    for (int u = 0; u < kCPW; ++u) {
      row_cur[u] = row_nxt[u];
      if (HAS_IDX) row_nxt[u] = row_nn[u];
      y_cur[u] = y_nxt[u];
    }
  }

  // ---------------- CTA reduction (fixed order) and partial record ----------------
  const PartialLayout pl = partial_layout(J, K, L0, L1);
  float swd0f[2 * NP0], swd1f[2 * NP1];
#pragma unroll
  for (int q = 0; q < NP0; ++q) { swd0f[2 * q] = warp_sum(lo2(swd0[q])); swd0f[2 * q + 1] = warp_sum(hi2(swd0[q])); }
#pragma unroll
  for (int q = 0; q < NP1; ++q) { swd1f[2 * q] = warp_sum(lo2(swd1[q])); swd1f[2 * q + 1] = warp_sum(hi2(swd1[q])); }
  const float swdd_w = warp_sum(lo2(swdd) + hi2(swdd));
  lqy_l = warp_sum(lqy_l);
  lmiss_l = warp_sum(lmiss_l);
  // every warp writes its own record into its own shared-memory slice, then all threads add the
  // kWarps slices element-wise in a fixed order (reproducible, no serial rounds)
  float* mine = sred + warp * pl.total;
  __syncthreads();
  for (int t = lane; t < pl.total; t += 32) mine[t] = 0.f;
  __syncwarp();
  if (ok) {
#pragma unroll
    for (int l = 0; l < 2 * NP0; ++l) if (l < L0) mine[pl.sw0 + l * J + lane] = (l & 1) ? hi2(sw0[l / 2]) : lo2(sw0[l / 2]);
#pragma unroll
    for (int l = 0; l < 2 * NP1; ++l) if (l < L1) mine[pl.sw1 + l * J + lane] = (l & 1) ? hi2(sw1[l / 2]) : lo2(sw1[l / 2]);
#pragma unroll
    for (int k = 0; k < 32; ++k) if (k < K) mine[pl.dZ + k * J + lane] = (k & 1) ? hi2(dZ[k / 2]) : lo2(dZ[k / 2]);
    mine[pl.dvm + lane] = dvm;
    mine[pl.dvls + lane] = dvls;
    mine[pl.nm + lane] = nm;
  }
  if (lane < K) mine[pl.gW + lane] = gW;
  if (lane == 0) {
#pragma unroll
    for (int l = 0; l < 2 * NP0; ++l) if (l < L0) mine[pl.swd0 + l] = swd0f[l];
#pragma unroll
    for (int l = 0; l < 2 * NP1; ++l) if (l < L1) mine[pl.swd1 + l] = swd1f[l];
    mine[pl.sc + 0] = swdd_w;
    mine[pl.sc + 1] = sfr;
    mine[pl.sc + 2] = sfo;
    double* md = reinterpret_cast<double*>(mine + pl.dbl);
    md[0] = (double)fac * (ll2 * 0.6931471805599453 + (double)lmiss_l);
    md[1] = (double)fac * (double)p.scale * (double)lqy_l;
  }
  __syncthreads();
  float* out = p.partials + (size_t)blockIdx.x * pl.total;
  for (int t = threadIdx.x; t < pl.total; t += kThreads) {
    if (t >= pl.dbl && t < pl.dbl + 4) continue;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) a += sred[w * pl.total + t];
    out[t] = a;
  }
  if (threadIdx.x < 2) {
    double a = 0.0;
    for (int w = 0; w < kWarps; ++w) a += reinterpret_cast<const double*>(sred + w * pl.total + pl.dbl)[threadIdx.x];
    reinterpret_cast<double*>(out + pl.dbl)[threadIdx.x] = a;
  }
}

}  // namespace cytof
