// Fused forward + backward of the per-cell part of the CyTOF ELBO for sm_100a.
//
// Replaces (reference paths): cytofpy/model/vae.py:17-41 (imputation + log_qy),
// cytofpy/model/Model.py:210-278 (loglike) and the autograd backward of both.
//
// Mapping.  One warp owns one cell at a time and lane = marker j (J <= 32: one marker per
// lane, the 128-byte row of y is one coalesced request; J <= 64: two markers per lane).
// The mixture statistics that must be reduced over cells are (j,l)-indexed, so with
// lane = j every lane keeps its own accumulators in registers for the whole kernel and
// no per-cell cross-lane traffic is needed for them.  The Z-gated J->K sum switches the
// warp to lane = feature k through a per-warp shared-memory row (D_j broadcast to every
// lane), the logsumexp over K is a warp-shuffle reduction, and the responsibilities g_k go
// back through a second shared-memory row for the backward (c1_j, dZ_jk).  A CTA only sees
// cells of ONE sample, so all per-sample constants live in registers.
//
// Arithmetic is fp32 in the log2 domain (ex2/lg2 are the MUFU natives); ll is carried in
// fp64 per warp; per-CTA partial sums are combined in a fixed order in shared memory, and
// the finalise kernel sums the per-CTA records in fp64 in a fixed order => bitwise
// reproducible results for a given grid.
#include <string.h>

#include "common.cuh"
#include "elbo_kernel_mma.cuh"
#include "elbo_kernel_mma64.cuh"

namespace cytof {

// build-time tuning knobs (see tools/kbench.py): resident CTAs/SM the register allocator must
// allow for the one-marker-per-lane variants, and whether Z is kept as 0/1 floats (FFMA) or as
// bit masks (predicated FADD)
#ifndef CYTOF_MINBLOCKS
#define CYTOF_MINBLOCKS 1
#endif
#ifndef CYTOF_ZFLOAT
#define CYTOF_ZFLOAT 1
#endif

template <int LM0, int LM1, int JP, int KP, bool NOISY>
__global__ void __launch_bounds__(kThreads, (JP == 1 && KP == 1) ? CYTOF_MINBLOCKS : 1)
elbo_fwdbwd_kernel(const ElboParams p) {
  pdl_enter();
  constexpr bool ZF = (JP == 1 && KP == 1) && CYTOF_ZFLOAT;  // Z rows/columns as float registers
  constexpr int JW = JP * 32, KW = KP * 32;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sD = smem + warp * (JW + KW);
  float* sG = sD + JW;
  float* sred = smem + kWarps * (JW + KW);

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
  const float h2 = 0.5f * inv_sig2 * kLog2e;
  const float cbase = (-logf(sig) - kHalfLog2Pi) * kLog2e;
  const float fac = p.fac[i];
  const float b0 = G[gl.b + 3 * i], b1 = G[gl.b + 3 * i + 1], b2 = G[gl.b + 3 * i + 2];

  float mu0[LM0], mu1[LM1];
#pragma unroll
  for (int l = 0; l < LM0; ++l) mu0[l] = l < L0 ? G[gl.mu0 + l] : 0.f;
#pragma unroll
  for (int l = 0; l < LM1; ++l) mu1[l] = l < L1 ? G[gl.mu1 + l] : 0.f;

  bool ok[JP];
  float okf[JP], cz0[JP][LM0], cz1[JP][LM1], vm[JP], vls[JP], sd[JP];
  unsigned long long zrow[JP];
#pragma unroll
  for (int jp = 0; jp < JP; ++jp) {
    const int j = lane + 32 * jp;
    ok[jp] = j < J;
    okf[jp] = ok[jp] ? 1.f : 0.f;
    const int jj = ok[jp] ? j : 0;
#pragma unroll
    for (int l = 0; l < LM0; ++l)
      cz0[jp][l] = (ok[jp] && l < L0) ? fmaf(G[gl.le0 + (i * J + jj) * L0 + (l < L0 ? l : 0)], kLog2e, cbase)
                                      : (l == 0 ? 0.f : kNegBig);
#pragma unroll
    for (int l = 0; l < LM1; ++l)
      cz1[jp][l] = (ok[jp] && l < L1) ? fmaf(G[gl.le1 + (i * J + jj) * L1 + (l < L1 ? l : 0)], kLog2e, cbase)
                                      : (l == 0 ? 0.f : kNegBig);
    vm[jp] = G[gl.vm + i * J + jj];
    vls[jp] = G[gl.vls + i * J + jj];
    sd[jp] = expf(vls[jp]);
    const unsigned long long kmask = K >= 64 ? ~0ull : ((1ull << K) - 1ull);
    zrow[jp] = ok[jp] ? (p.zmask[jj] & kmask) : 0ull;
  }
  unsigned long long zcol[KP];
  float lw2[KP];
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) {
    const int k = lane + 32 * kp;
    unsigned long long m = 0;
    if (k < K)
      for (int j = 0; j < J; ++j) m |= ((p.zmask[j] >> k) & 1ull) << j;
    zcol[kp] = m;
    lw2[kp] = k < K ? G[gl.lw + i * K + k] * kLog2e : kNegBig;
  }
  float zcf[ZF ? JW : 1], zrf[ZF ? KW : 1];
  if constexpr (ZF) {
#pragma unroll
    for (int j = 0; j < JW; ++j) zcf[j] = (float)((zcol[0] >> j) & 1ull);
#pragma unroll
    for (int k = 0; k < KW; ++k) zrf[k] = (float)((zrow[0] >> k) & 1ull);
  }
  // noisy class constants (log2 domain)
  float l1me2 = 0.f, zconst2 = 0.f, inv2s2 = 0.f, inv_s2n = 0.f;
  if (NOISY) {
    const float e_i = G[gl.eps + i];
    l1me2 = log1pf(-e_i) * kLog2e;
    zconst2 = ((float)J * (-kHalfLog2Pi - logf(p.noisy_sd)) + logf(e_i)) * kLog2e;
    inv_s2n = 1.0f / (p.noisy_sd * p.noisy_sd);
    inv2s2 = 0.5f * inv_s2n * kLog2e;
  }

  // ---------------- accumulators ----------------
  float sw0[JP][LM0], sw1[JP][LM1], swd0[LM0], swd1[LM1], dZ[JP][KW], gW[KP];
  float dvm[JP], dvls[JP], nm[JP];
  float swdd = 0.f, sfr = 0.f, sfo = 0.f, lqy_l = 0.f, lmiss_l = 0.f;
  double ll2 = 0.0;
#pragma unroll
  for (int l = 0; l < LM0; ++l) swd0[l] = 0.f;
#pragma unroll
  for (int l = 0; l < LM1; ++l) swd1[l] = 0.f;
#pragma unroll
  for (int kp = 0; kp < KP; ++kp) gW[kp] = 0.f;
#pragma unroll
  for (int jp = 0; jp < JP; ++jp) {
    dvm[jp] = dvls[jp] = nm[jp] = 0.f;
#pragma unroll
    for (int l = 0; l < LM0; ++l) sw0[jp][l] = 0.f;
#pragma unroll
    for (int l = 0; l < LM1; ++l) sw1[jp][l] = 0.f;
#pragma unroll
    for (int k = 0; k < KW; ++k) dZ[jp][k] = 0.f;
  }

  // ---------------- stream the cells of this CTA ----------------
  const int32_t* __restrict__ idx = p.idx;
  RowSampler rsm;      // CYTOF_FLAG_SAMPLE_IN_KERNEL: rows from the sampler's keyed bijection instead of idx
  if (p.sampler_on) rsm = make_row_sampler(p, i);
  auto row_of = [&](long long c) -> long long {
    return row_base + (p.sampler_on ? (long long)rsm.row(c) : (idx ? (long long)__ldg(idx + cb + c) : c));
  };
  long long c = c_lo + warp;
  long long row_cur = 0, row_next = 0;
  float y_cur[JP];
#pragma unroll
  for (int jp = 0; jp < JP; ++jp) y_cur[jp] = 0.f;
  if (c < c_hi) {
    row_cur = row_of(c);
#pragma unroll
    for (int jp = 0; jp < JP; ++jp)
      if (ok[jp]) y_cur[jp] = ld_stream(p.y + row_cur * J + lane + 32 * jp);
    if (c + kWarps < c_hi) row_next = row_of(c + kWarps);
  }

  for (; c < c_hi; c += kWarps) {
    // prefetch: y of the next cell, row index of the one after
    float y_nxt[JP];
    long long row_nn = 0;
#pragma unroll
    for (int jp = 0; jp < JP; ++jp) y_nxt[jp] = 0.f;
    if (c + kWarps < c_hi) {
#pragma unroll
      for (int jp = 0; jp < JP; ++jp)
        if (ok[jp]) y_nxt[jp] = ld_stream(p.y + row_next * J + lane + 32 * jp);
      if (c + 2 * kWarps < c_hi) row_nn = row_of(c + 2 * kWarps);
    }

    // ---- imputation of missing entries (vae.py:17-33)
    bool miss[JP];
    float e[JP], yv[JP];
    bool anym = false;
#pragma unroll
    for (int jp = 0; jp < JP; ++jp) {
      miss[jp] = ok[jp] && (y_cur[jp] != y_cur[jp]);
      anym |= miss[jp];
      e[jp] = 0.f;
      yv[jp] = ok[jp] ? y_cur[jp] : 0.f;
    }
    anym = __any_sync(FULL, anym);
    if (anym) {
#pragma unroll
      for (int jp = 0; jp < JP; ++jp)
        if (miss[jp]) {
          const int j = lane + 32 * jp;
          e[jp] = p.noise_mode == CYTOF_NOISE_EXTERNAL
                      ? (p.eps ? __ldg(p.eps + (cb + c) * J + j) : 0.f)
                      : philox_normal(p.seed, elbo_step(p), grow_base + (unsigned long long)c, philox_ctr(j, i));   // keyed by the POSITION in the batch
          yv[jp] = fmaf(sd[jp], e[jp], vm[jp]);
        }
    }

    // ---- mixture log-densities, forward (Model.py:223-233), log2 domain
    float e0[JP][LM0], e1[JP][LM1], rS0[JP], rS1[JP], Dj[JP];
    float a0sum = 0.f, y2sum = 0.f;
#pragma unroll
    for (int jp = 0; jp < JP; ++jp) {
      float M0 = 0.f, M1 = 0.f;
#pragma unroll
      for (int l = 0; l < LM0; ++l) {
        const float d = yv[jp] - mu0[l];
        e0[jp][l] = fmaf(-h2 * d, d, cz0[jp][l]);
        M0 = l == 0 ? e0[jp][0] : fmaxf(M0, e0[jp][l]);
      }
#pragma unroll
      for (int l = 0; l < LM1; ++l) {
        const float d = yv[jp] - mu1[l];
        e1[jp][l] = fmaf(-h2 * d, d, cz1[jp][l]);
        M1 = l == 0 ? e1[jp][0] : fmaxf(M1, e1[jp][l]);
      }
      float S0 = 0.f, S1 = 0.f;
#pragma unroll
      for (int l = 0; l < LM0; ++l) {
        e0[jp][l] = fast_ex2(e0[jp][l] - M0);
        S0 += e0[jp][l];
      }
#pragma unroll
      for (int l = 0; l < LM1; ++l) {
        e1[jp][l] = fast_ex2(e1[jp][l] - M1);
        S1 += e1[jp][l];
      }
      const float A0 = M0 + fast_lg2(S0), A1 = M1 + fast_lg2(S1);
      rS0[jp] = fast_rcp(S0);
      rS1[jp] = fast_rcp(S1);
      Dj[jp] = okf[jp] * (A1 - A0);
      a0sum = fmaf(okf[jp], A0, a0sum);
      y2sum = fmaf(yv[jp], yv[jp], y2sum);
      sD[lane + 32 * jp] = Dj[jp];
    }
    a0sum = warp_sum(a0sum);
    if (NOISY) y2sum = warp_sum(y2sum);
    __syncwarp();

    // ---- Z-gated J->K sum + log W (Model.py:246-252): lane = feature k
    float f[KP], mx = kNegBig;
#pragma unroll
    for (int kp = 0; kp < KP; ++kp) {
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      if constexpr (ZF) {
#pragma unroll
        for (int q = 0; q < JW / 4; ++q) {
          const float4 d4 = *reinterpret_cast<const float4*>(sD + 4 * q);
          a[0] = fmaf(zcf[4 * q + 0], d4.x, a[0]);
          a[1] = fmaf(zcf[4 * q + 1], d4.y, a[1]);
          a[2] = fmaf(zcf[4 * q + 2], d4.z, a[2]);
          a[3] = fmaf(zcf[4 * q + 3], d4.w, a[3]);
        }
      } else {
#pragma unroll
        for (int q = 0; q < JW / 4; ++q) {
          const float4 d4 = *reinterpret_cast<const float4*>(sD + 4 * q);
          a[0] += ((zcol[kp] >> (4 * q + 0)) & 1ull) ? d4.x : 0.f;
          a[1] += ((zcol[kp] >> (4 * q + 1)) & 1ull) ? d4.y : 0.f;
          a[2] += ((zcol[kp] >> (4 * q + 2)) & 1ull) ? d4.z : 0.f;
          a[3] += ((zcol[kp] >> (4 * q + 3)) & 1ull) ? d4.w : 0.f;
        }
      }
      f[kp] = ((a[0] + a[1]) + (a[2] + a[3])) + a0sum + lw2[kp];
      mx = fmaxf(mx, f[kp]);
    }
    mx = warp_max_redux(mx);
    float ex[KP], s = 0.f;
#pragma unroll
    for (int kp = 0; kp < KP; ++kp) {
      ex[kp] = fast_ex2(f[kp] - mx);
      s += ex[kp];
    }
    s = warp_sum_unit(s);   // every term is in [0, KP]
    const float lse2 = mx + fast_lg2(s);
    const float rinv = fast_rcp(s);

    // ---- noisy-cell class (Model.py:259-264)
    float rho = 1.f, omr = 0.f, L2 = lse2;
    if (NOISY) {
      const float q2 = lse2 + l1me2;
      const float z2 = fmaf(-inv2s2, y2sum, zconst2);
      const float dd = q2 - z2;
      const float t = fast_ex2(-fabsf(dd));
      const float r1 = fast_rcp(1.f + t);
      L2 = fmaxf(q2, z2) + fast_lg2(1.f + t);
      rho = dd >= 0.f ? r1 : t * r1;
      omr = dd >= 0.f ? t * r1 : r1;
    }
    ll2 += (double)L2;
    const float fr = fac * rho;
    sfr += fr;
    sfo = fmaf(fac, omr, sfo);

    // ---- responsibilities g_k = fac * rho * softmax_k(f)
#pragma unroll
    for (int kp = 0; kp < KP; ++kp) {
      const float g = fr * rinv * ex[kp];
      gW[kp] += g;
      sG[lane + 32 * kp] = g;
    }
    __syncwarp();

    // ---- backward: lane = marker j again
#pragma unroll
    for (int jp = 0; jp < JP; ++jp) {
      float cc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < KW / 4; ++q) {
        const float4 g4 = *reinterpret_cast<const float4*>(sG + 4 * q);
        if constexpr (ZF) {
          cc[0] = fmaf(zrf[4 * q + 0], g4.x, cc[0]);
          cc[1] = fmaf(zrf[4 * q + 1], g4.y, cc[1]);
          cc[2] = fmaf(zrf[4 * q + 2], g4.z, cc[2]);
          cc[3] = fmaf(zrf[4 * q + 3], g4.w, cc[3]);
        } else {
          cc[0] += ((zrow[jp] >> (4 * q + 0)) & 1ull) ? g4.x : 0.f;
          cc[1] += ((zrow[jp] >> (4 * q + 1)) & 1ull) ? g4.y : 0.f;
          cc[2] += ((zrow[jp] >> (4 * q + 2)) & 1ull) ? g4.z : 0.f;
          cc[3] += ((zrow[jp] >> (4 * q + 3)) & 1ull) ? g4.w : 0.f;
        }
        dZ[jp][4 * q + 0] = fmaf(g4.x, Dj[jp], dZ[jp][4 * q + 0]);
        dZ[jp][4 * q + 1] = fmaf(g4.y, Dj[jp], dZ[jp][4 * q + 1]);
        dZ[jp][4 * q + 2] = fmaf(g4.z, Dj[jp], dZ[jp][4 * q + 2]);
        dZ[jp][4 * q + 3] = fmaf(g4.w, Dj[jp], dZ[jp][4 * q + 3]);
      }
      const float c1 = okf[jp] * ((cc[0] + cc[1]) + (cc[2] + cc[3]));
      const float c0 = okf[jp] * fr - c1;
      // mixtures, backward: w_zl = c_z p_zl; statistics sum w, sum w d, sum w d^2
      const float r0 = c0 * rS0[jp], r1 = c1 * rS1[jp];
      float wdsum = 0.f;
#pragma unroll
      for (int l = 0; l < LM0; ++l) {
        const float w = r0 * e0[jp][l];
        const float d = yv[jp] - mu0[l];
        const float wd = w * d;
        sw0[jp][l] += w;
        swd0[l] += wd;
        swdd = fmaf(wd, d, swdd);
        e0[jp][l] = wd;   // kept for d ll / d y (only consumed when the row has missing entries)
      }
#pragma unroll
      for (int l = 0; l < LM1; ++l) {
        const float w = r1 * e1[jp][l];
        const float d = yv[jp] - mu1[l];
        const float wd = w * d;
        sw1[jp][l] += w;
        swd1[l] += wd;
        swdd = fmaf(wd, d, swdd);
        e1[jp][l] = wd;
      }
      // logistic missingness (Model.py:269-274) and d/dy chain into the imputation parameters
      if (anym) {
#pragma unroll
        for (int l = 0; l < LM0; ++l) wdsum += e0[jp][l];
#pragma unroll
        for (int l = 0; l < LM1; ++l) wdsum += e1[jp][l];
        const float u = fmaf(fmaf(b2, yv[jp], b1), yv[jp], b0);
        const float t = fast_ex2(-fabsf(u) * kLog2e);
        const float rr = fast_rcp(1.f + t);
        const float ls = -(fmaxf(-u, 0.f) + kLn2 * fast_lg2(1.f + t));
        const float sneg = (u > 0.f ? t : 1.f) * rr;  // 1 - sigmoid(u)
        float dy = -wdsum * inv_sig2 + fac * sneg * fmaf(2.f * b2, yv[jp], b1);
        if (NOISY) dy = fmaf(-fac * omr * inv_s2n, yv[jp], dy);
        if (miss[jp]) {
          lmiss_l += ls;
          dvm[jp] += dy;
          dvls[jp] = fmaf(dy * sd[jp], e[jp], dvls[jp]);
          nm[jp] += 1.f;
          lqy_l += fmaf(-0.5f * e[jp], e[jp], -vls[jp] - kHalfLog2Pi);
        }
      }
    }
    __syncwarp();  // sD / sG are rewritten by the next cell

    row_cur = row_next;
    row_next = row_nn;
#pragma unroll
    for (int jp = 0; jp < JP; ++jp) y_cur[jp] = y_nxt[jp];
  }

  // ---------------- CTA reduction (fixed order) and partial record ----------------
  const PartialLayout pl = partial_layout(J, K, L0, L1);
#pragma unroll
  for (int l = 0; l < LM0; ++l) swd0[l] = warp_sum(swd0[l]);
#pragma unroll
  for (int l = 0; l < LM1; ++l) swd1[l] = warp_sum(swd1[l]);
  swdd = warp_sum(swdd);
  lqy_l = warp_sum(lqy_l);
  lmiss_l = warp_sum(lmiss_l);
  double* sdbl = reinterpret_cast<double*>(sred + pl.dbl);
  __syncthreads();
  for (int w = 0; w < kWarps; ++w) {
    if (warp == w) {
      auto put = [&](int off, float v) {
        if (w == 0) sred[off] = v; else sred[off] += v;
      };
#pragma unroll
      for (int jp = 0; jp < JP; ++jp) {
        const int j = lane + 32 * jp;
        if (ok[jp]) {
#pragma unroll
          for (int l = 0; l < LM0; ++l) if (l < L0) put(pl.sw0 + l * J + j, sw0[jp][l]);
#pragma unroll
          for (int l = 0; l < LM1; ++l) if (l < L1) put(pl.sw1 + l * J + j, sw1[jp][l]);
#pragma unroll
          for (int k = 0; k < KW; ++k) if (k < K) put(pl.dZ + k * J + j, dZ[jp][k]);
          put(pl.dvm + j, dvm[jp]);
          put(pl.dvls + j, dvls[jp]);
          put(pl.nm + j, nm[jp]);
        }
      }
#pragma unroll
      for (int kp = 0; kp < KP; ++kp) {
        const int k = lane + 32 * kp;
        if (k < K) put(pl.gW + k, gW[kp]);
      }
      if (lane == 0) {
#pragma unroll
        for (int l = 0; l < LM0; ++l) if (l < L0) put(pl.swd0 + l, swd0[l]);
#pragma unroll
        for (int l = 0; l < LM1; ++l) if (l < L1) put(pl.swd1 + l, swd1[l]);
        put(pl.sc + 0, swdd);
        put(pl.sc + 1, sfr);
        put(pl.sc + 2, sfo);
        put(pl.sc + 3, 0.f);
        const double llw = (double)fac * (ll2 * 0.6931471805599453 + (double)lmiss_l);
        const double lqw = (double)fac * (double)p.scale * (double)lqy_l;
        if (w == 0) { sdbl[0] = llw; sdbl[1] = lqw; } else { sdbl[0] += llw; sdbl[1] += lqw; }
      }
    }
    __syncthreads();
  }
  float* out = p.partials + (size_t)blockIdx.x * pl.total;
  for (int t = threadIdx.x; t < pl.total; t += kThreads) out[t] = sred[t];
}

// Finalise: the per-CTA partial records -> gradient block (+ ll, log_qy), fp64.  The records are [nblocks][pl.total]
// floats, so the kernel walks them in RECORD space: a CTA owns 32 consecutive offsets of one section (lane = offset:
// every load of a warp is one 128-byte line of one record), its 8 warps split the records of the section's range
// (warp w: b0 + w, b0 + w + 8, ...), and warp 0 adds the 8 partial sums in warp order -- a fixed order for a given
// grid, so the result is reproducible.  The handful of outputs that combine several sums with per-sample factors
// (dmu, dsig, deps, ll, log_qy) take one warp each (lanes over records, butterfly).  ~150 CTAs at cfg4 instead of
// one warp per output (570 CTAs reading one 32-byte sector per load).
struct FinPlan { int per_sample, w_s0, w_s1, w_k, w_j, w_z, n_small, n_units, n_ctas; };
__host__ __device__ inline FinPlan fin_plan(int I, int J, int K, int L0, int L1) {
  FinPlan f;
  f.w_s0 = (L0 * J + 31) / 32; f.w_s1 = (L1 * J + 31) / 32; f.w_k = (K + 31) / 32; f.w_j = (J + 31) / 32;
  f.per_sample = f.w_s0 + f.w_s1 + f.w_k + 3 * f.w_j;
  f.w_z = (J * K + 31) / 32;
  f.n_small = L0 + L1 + 2 * I + 2;
  f.n_units = I * f.per_sample + f.w_z;
  f.n_ctas = f.n_units + (f.n_small + 7) / 8;
  return f;
}

// One-shot all-reduce over NVLink / NVSwitch peer memory, fused into the finalise kernel.  Every
// rank owns a mailbox (cudaMalloc'ed by cytof_peer_alloc, mapped into the peers with CUDA IPC):
//   [0, 128)    uint64 flags[src]: sequence number of the last step rank `src` has delivered
//   [128, 132)  uint32 ticket: CTAs of the local finalise kernel that are done (local use)
//   [132, 136)  uint32 error: set when a wait of THIS rank timed out (sticky until the mailbox is freed)
//   [136, 140)  uint32 abort: set by ANY rank whose wait timed out (it stores it into every mailbox), so that all
//               ranks poison their results and raise together instead of diverging silently
//   [192, 224)  uint64 ns time stamps of CTA 0 in the last step: entry, own part stored, flags seen, sum done
//   [256, ...)  double slots[parity][src][nout]: the packed vector of rank `src` for steps of that parity
// Step s: every finalise thread stores its output straight into slot [s & 1][rank] of EVERY mailbox
// (its own included); the last CTA of the grid publishes flags[rank] = s to every peer with a
// system-scope release, waits (acquire, bounded) until all its own flags reach s and sums the
// world slots in rank order -- the same order on every rank, so the replicas stay bit-identical.
// Two parities suffice: no rank can start step s + 2 before every rank has finished reading step s.
struct PeerArgs {
  int world, rank;
  unsigned long long seq;
  const long long* seq_dev;    // non-null: seq = *seq_dev + 1
  int nout;
  int all_ctas;                // every CTA waits for the flags and sums its own outputs (grid co-resident), else the last CTA does
  unsigned char* mailbox[CYTOF_MAX_PEERS];
};
constexpr size_t kPeerHeaderBytes = 256;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
elbo_finalize_kernel(const ElboParams p, float* __restrict__ grad_block,
                     double* __restrict__ ll_lqy, double* __restrict__ packed,
                     const PeerArgs peer) {
  pdl_enter();
  const int I = p.I, J = p.J, K = p.K, L0 = p.L0, L1 = p.L1;
  const BlockLayout bl = block_layout(I, J, K, L0, L1);
  const PartialLayout pl = partial_layout(J, K, L0, L1);
  const GlobalsLayout gl = globals_layout(I, J, K, L0, L1);
  const FinPlan fp = fin_plan(I, J, K, L0, L1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* __restrict__ P = p.partials;
  const float* __restrict__ G = p.globals;
  __shared__ double s_part[8][32];
  int t = -1;            // the output this thread owns (-1: none)
  double r = 0.0;
  const int unit = (int)blockIdx.x;
  if (unit < fp.n_units) {
    // ---- 32 consecutive record offsets of one section, records split over the 8 warps
    int b0 = 0, b1 = p.nblocks, off = 0, len = 0, e = 0, kind = 6, i = 0;
    if (unit < I * fp.per_sample) {
      i = unit / fp.per_sample;
      int q = unit % fp.per_sample;
      b0 = p.blk_begin[i]; b1 = p.blk_begin[i + 1];
      if (q < fp.w_s0) { kind = 0; off = pl.sw0; len = L0 * J; }
      else if ((q -= fp.w_s0) < fp.w_s1) { kind = 1; off = pl.sw1; len = L1 * J; }
      else if ((q -= fp.w_s1) < fp.w_k) { kind = 2; off = pl.gW; len = K; }
      else if ((q -= fp.w_k) < fp.w_j) { kind = 3; off = pl.dvm; len = J; }
      else if ((q -= fp.w_j) < fp.w_j) { kind = 4; off = pl.dvls; len = J; }
      else { q -= fp.w_j; kind = 5; off = pl.nm; len = J; }
      e = 32 * q + lane;
    } else {
      off = pl.dZ; len = J * K;
      e = 32 * (unit - I * fp.per_sample) + lane;
    }
    const bool valid = e < len;
    const float* src = P + off + (valid ? e : 0);
    double a = 0.0;
    // 16 predicated loads in flight per lane: the 19 records per warp of a 148-CTA dZ section are 2 memory round
    // trips (were 5 with 4 in flight + a serial remainder); adding 0.0f for the masked ones keeps the order fixed
    for (int b = b0 + warp; b < b1; b += 128) {
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = (b + 8 * u < b1) ? __ldg(src + (size_t)(b + 8 * u) * pl.total) : 0.f;
#pragma unroll
      for (int u = 0; u < 16; ++u) a += (double)v[u];
    }
    s_part[warp][lane] = a;
    __syncthreads();
    if (warp == 0 && valid) {
#pragma unroll
      for (int w = 0; w < 8; ++w) r += s_part[w][lane];
      switch (kind) {
        case 0: t = bl.dle0 + (i * J + e % J) * L0 + e / J; break;               // e = l J + j
        case 1: t = bl.dle1 + (i * J + e % J) * L1 + e / J; break;
        case 2: t = bl.dlw + i * K + e; break;
        case 3: t = bl.dvm + i * J + e; break;
        case 4: t = bl.dvls + i * J + e; break;
        case 5: t = bl.dlq + i * J + e; r *= -(double)p.fac[i] * (double)p.scale; break;   // d log_qy / d vls = -fac scale #missing
        default: t = bl.dZ + (e % J) * K + e / J; r *= 0.6931471805599453; break;  // e = k J + j; dZ_jk = ln2 sum g_k D2_j
      }
    }
  } else {
    // ---- outputs that mix several sums: one warp each, lanes over the records.  Every lane issues all its loads
    // (records x up to two offsets, and the per-sample factor) before the first add: ONE memory round trip per output
    const int sidx = (unit - fp.n_units) * 8 + warp;
    auto butterfly = [&](double a) -> double {
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      return a;
    };
    if (sidx < fp.n_small) {
      if (sidx < L0 + L1) {             // dmu_z[l] = sum_i (1/sig_i^2) sum_b (sum w d)_b
        const bool z1 = sidx >= L0;
        const int l = z1 ? sidx - L0 : sidx;
        const int off = (z1 ? pl.swd1 : pl.swd0) + l;
        int i = 0;
        for (int bb = lane; bb < p.nblocks; bb += 256) {     // 8 records per lane in flight: one round trip up to 256 records
          float v[8], sg[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int b = bb + 32 * u;
            const bool in = b < p.nblocks;
            while (in && i + 1 < I && b >= p.blk_begin[i + 1]) ++i;
            v[u] = in ? __ldg(P + (size_t)b * pl.total + off) : 0.f;
            sg[u] = __ldg(G + gl.sig + i);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) r += (double)v[u] / ((double)sg[u] * (double)sg[u]);
        }
        r = butterfly(r);
        t = (z1 ? bl.dmu1 : bl.dmu0) + l;
      } else if (sidx < L0 + L1 + 2 * I) {
        // dsig_i = sum w d^2 / sig^3 - (J sum fac rho) / sig;  deps_i = -sum fac rho/(1-eps) + sum fac (1-rho)/eps
        const bool is_eps = sidx >= L0 + L1 + I;
        const int i = sidx - L0 - L1 - (is_eps ? I : 0);
        const int offa = pl.sc + (is_eps ? 2 : 0), offb = pl.sc + 1;
        const double par = (double)__ldg(G + (is_eps ? gl.eps : gl.sig) + i);
        double a = 0.0, c = 0.0;
        for (int b = p.blk_begin[i] + lane; b < p.blk_begin[i + 1]; b += 32) {
          a += (double)__ldg(P + (size_t)b * pl.total + offa);
          c += (double)__ldg(P + (size_t)b * pl.total + offb);
        }
        a = butterfly(a);
        c = butterfly(c);
        if (is_eps) { r = p.model_noisy ? -c / (1.0 - par) + a / par : 0.0; t = bl.deps + i; }
        else { r = a / (par * par * par) - (double)J * c / par; t = bl.dsig + i; }
      } else {                          // ll, log_qy
        const int which = sidx - (L0 + L1 + 2 * I);
        for (int bb = lane; bb < p.nblocks; bb += 256) {
          double v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            v[u] = (bb + 32 * u < p.nblocks) ? reinterpret_cast<const double*>(P + (size_t)(bb + 32 * u) * pl.total + pl.dbl)[which] : 0.0;
#pragma unroll
          for (int u = 0; u < 8; ++u) r += v[u];
        }
        r = butterfly(r);
        t = bl.total + which;
      }
      if (lane != 0) t = -1;
    }
  }
  const bool owner = t >= 0;
  if (peer.world <= 1) {
    if (owner) {
      if (t >= bl.total) {
        if (ll_lqy) ll_lqy[t - bl.total] = r;
      } else if (grad_block) {
        grad_block[t] = (float)r;
      }
      if (packed) packed[t] = r;
    }
    return;
  }
  // ---- fused one-shot all-reduce (see PeerArgs)
  const unsigned long long seq = peer.seq_dev ? (unsigned long long)(*peer.seq_dev + 1) : peer.seq;
  const int par = (int)(seq & 1ull);
  const size_t slot_off = kPeerHeaderBytes + ((size_t)par * peer.world + peer.rank) * (size_t)peer.nout * sizeof(double);
  if (owner) {
    for (int q = 0; q < peer.world; ++q)
      reinterpret_cast<double*>(peer.mailbox[q] + slot_off)[t] = r;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ int s_last;
  unsigned char* mine = peer.mailbox[peer.rank];
  unsigned* ticket = reinterpret_cast<unsigned*>(mine + 128);
  unsigned* errw = reinterpret_cast<unsigned*>(mine + 132);
  volatile unsigned* abortw = reinterpret_cast<volatile unsigned*>(mine + 136);
  unsigned long long* dbg = reinterpret_cast<unsigned long long*>(mine + 192);    // [4] ns time stamps of CTA 0 (cytof_peer_debug)
  const bool stamp = blockIdx.x == 0 && threadIdx.x == 0;
  auto now_ns = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
  if (stamp) dbg[0] = now_ns();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1u);
  __syncthreads();
  // bounded, abort-aware wait of thread q < world for the flag of rank q in this rank's mailbox
  auto wait_flags = [&]() {
    if ((int)threadIdx.x < peer.world) {
      const unsigned long long* f = reinterpret_cast<const unsigned long long*>(mine) + threadIdx.x;
      const long long t0 = clock64();
      while (ld_acquire_sys(f) < seq) {
        if (*abortw != 0u) break;                                    // another rank gave up: do not wait out the 2 s
        if (clock64() - t0 > 4000000000ll) {                         // ~2 s: a peer died; never hang the GPU
          *errw = 1u;
          for (int q = 0; q < peer.world; ++q) *reinterpret_cast<volatile unsigned*>(peer.mailbox[q] + 136) = 1u;
          __threadfence_system();
          break;
        }
        __nanosleep(64);
      }
    }
    __syncthreads();
  };
  const double* slots = reinterpret_cast<const double*>(mine + kPeerHeaderBytes) + (size_t)par * peer.world * peer.nout;
  const double kNaN = __longlong_as_double(0x7ff8000000000000ll), kInf = __longlong_as_double(0x7ff0000000000000ll);
  if (s_last) {      // this rank's vector is complete in every mailbox: publish the sequence number to every peer
    __threadfence_system();
    if ((int)threadIdx.x < peer.world) {
      if (threadIdx.x == 0) *ticket = 0u;
      st_release_sys(reinterpret_cast<unsigned long long*>(peer.mailbox[threadIdx.x]) + peer.rank, seq);
    }
  }
  if (peer.all_ctas) {
    // ---- every CTA of the (co-resident) grid waits for the world flags itself and sums ITS outputs over the
    // world slots in rank order: no serial tail in one CTA (with 8 ranks that tail read 36 k doubles from 256 threads)
    if (stamp) dbg[1] = now_ns();
    wait_flags();
    if (stamp) dbg[2] = now_ns();
    const bool bad = *reinterpret_cast<volatile unsigned*>(errw) != 0u || *abortw != 0u;
    if (owner) {
      double v[CYTOF_MAX_PEERS];
#pragma unroll
      for (int q = 0; q < CYTOF_MAX_PEERS; ++q) v[q] = q < peer.world ? __ldcg(slots + (size_t)q * peer.nout + t) : 0.0;
      double a = 0.0;
#pragma unroll
      for (int q = 0; q < CYTOF_MAX_PEERS; ++q) a += v[q];
      if (bad) a = (t == peer.nout - 2) ? kInf : kNaN;      // +inf in the ll slot: the mark global_adam_body reads as "the exchange failed"
      packed[t] = a;
    }
    if (stamp) dbg[3] = now_ns();
    return;
  }
  if (!s_last) return;
  if (stamp) dbg[1] = now_ns();
  wait_flags();
  if (stamp) dbg[2] = now_ns();
  const bool bad = *reinterpret_cast<volatile unsigned*>(errw) != 0u || *abortw != 0u;
  // one CTA sums world x nout doubles: all loads of two outputs (up to 32) are issued before the first add,
  // otherwise this tail is a chain of L2 round trips; the sum runs in rank order on every rank (x + 0.0 = x)
  for (int o = threadIdx.x; o < peer.nout; o += 2 * blockDim.x) {
    const int o2 = o + blockDim.x;
    double v[CYTOF_MAX_PEERS], w[CYTOF_MAX_PEERS];
#pragma unroll
    for (int q = 0; q < CYTOF_MAX_PEERS; ++q) {
      v[q] = q < peer.world ? __ldcg(slots + (size_t)q * peer.nout + o) : 0.0;
      w[q] = (q < peer.world && o2 < peer.nout) ? __ldcg(slots + (size_t)q * peer.nout + o2) : 0.0;
    }
    double a = 0.0, c = 0.0;
#pragma unroll
    for (int q = 0; q < CYTOF_MAX_PEERS; ++q) { a += v[q]; c += w[q]; }
    if (bad) {
      a = (o == peer.nout - 2) ? kInf : kNaN;
      c = (o2 == peer.nout - 2) ? kInf : kNaN;
    }
    packed[o] = a;
    if (o2 < peer.nout) packed[o2] = c;
  }
}

// y_out[C,J] = imputed minibatch (vae.py:33); eps_out[C,J] = the Philox draws.
__global__ void impute_emit_kernel(const ElboParams p, float* __restrict__ eps_out) {
  const long long C = p.cell_begin[p.I];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= C * p.J) return;
  const long long c = t / p.J;
  const int j = (int)(t % p.J);
  int i = 0;
  while (i + 1 < p.I && c >= p.cell_begin[i + 1]) ++i;
  const long long lc = c - p.cell_begin[i];
  const long long lrow = p.idx ? (long long)p.idx[c] : lc;
  float e = 0.f;
  if (p.noise_mode == CYTOF_NOISE_EXTERNAL) {
    if (p.eps) e = p.eps[c * p.J + j];
  } else {
    e = philox_normal(p.seed, elbo_step(p), (unsigned long long)(p.row_global[i] + lc), philox_ctr(j, i));   // position, not row
  }
  if (eps_out) { eps_out[t] = e; return; }
  const GlobalsLayout gl = globals_layout(p.I, p.J, p.K, p.L0, p.L1);
  float v = p.y[(p.row_base[i] + lrow) * p.J + j];
  if (v != v) v = fmaf(expf(p.globals[gl.vls + i * p.J + j]), e, p.globals[gl.vm + i * p.J + j]);
  p.y_out[t] = v;
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
struct Variant {
  int lm0, lm1, jp, kp;
  int smem_floats_per_warp, smem_floats_extra;
  KernelFn fn[8];  // [noisy + 2 * has_idx + 4 * no_missing] (the last bit only selects a different kernel in the MMA variants)
  bool records_alias;   // the final per-warp records reuse the main-loop shared memory
  int threads;          // CTA size (0 = kThreads)
  int smem_bytes;       // fixed dynamic shared memory size (0 = derived from the per-warp figures)
  int split;            // blocks of the grid plan (= partial records) per CTA (0 = 1): kMmaSplit warp groups in the MMA variants
};
static int variant_split(const Variant* v) { return v->split ? v->split : 1; }
static int variant_threads(const Variant* v) { return v->threads ? v->threads : kThreads; }
#define CYTOF_VARIANT(a, b, jp, kp) \
  { a, b, jp, kp, (jp) * 32 + (kp) * 32, 0, { elbo_fwdbwd_kernel<a, b, jp, kp, false>, elbo_fwdbwd_kernel<a, b, jp, kp, true>, \
                                           elbo_fwdbwd_kernel<a, b, jp, kp, false>, elbo_fwdbwd_kernel<a, b, jp, kp, true>, \
                                           elbo_fwdbwd_kernel<a, b, jp, kp, false>, elbo_fwdbwd_kernel<a, b, jp, kp, true>, \
                                           elbo_fwdbwd_kernel<a, b, jp, kp, false>, elbo_fwdbwd_kernel<a, b, jp, kp, true> } }
#define CYTOF_FN8(f, ...) { f(__VA_ARGS__, 0), f(__VA_ARGS__, 1), f(__VA_ARGS__, 2), f(__VA_ARGS__, 3), \
                            f(__VA_ARGS__, 4), f(__VA_ARGS__, 5), f(__VA_ARGS__, 6), f(__VA_ARGS__, 7) }
#define CYTOF_VARIANT_MMA(a, b, lidx) \
  { a, b, 1, 1, kMmaWarpFloats, kMmaZFloats, CYTOF_FN8(mma_kernel_fn, lidx), true, 0, 0, kMmaSplit }
#ifndef CYTOF_MMA
#define CYTOF_MMA 1   // 1: tensor-core kernel (elbo_kernel_mma.cuh) for J, K <= 32; 0: the first SIMT kernel (A/B builds)
#endif
// ordered by cost: the first variant that fits is used; (8,8) is the padded fallback
static const Variant kVariants[] = {
#if CYTOF_MMA
    CYTOF_VARIANT_MMA(3, 3, 0), CYTOF_VARIANT_MMA(5, 3, 1), CYTOF_VARIANT_MMA(5, 5, 2), CYTOF_VARIANT_MMA(8, 8, 3),
#else
    CYTOF_VARIANT(3, 3, 1, 1), CYTOF_VARIANT(4, 3, 1, 1), CYTOF_VARIANT(5, 3, 1, 1),
    CYTOF_VARIANT(5, 5, 1, 1), CYTOF_VARIANT(8, 8, 1, 1),
#endif
    CYTOF_VARIANT(3, 3, 2, 2), CYTOF_VARIANT(5, 5, 2, 2), CYTOF_VARIANT(8, 8, 2, 2),
};

// 32 < J <= 48, K <= 64 (cfg5): tensor-core kernel with CTA-cooperative
// dZ (elbo_kernel_mma64.cuh).  fn index = noisy + 2 * has_idx + 4 * (tail passes - 1).
#define CYTOF_VARIANT_MMA64(a, b, lidx, miss) \
  { a, b, 2, 2, kXWarpFloats, kXZFloats, CYTOF_FN8(mma64_kernel_fn, lidx, miss), true, kXThreads }
#ifndef CYTOF_MMA64
#define CYTOF_MMA64 1
#endif
// [0..1]: kernels without the imputation path (CYTOF_FLAG_NO_MISSING), [2..3]: with it
static const Variant kVariants64[] = { CYTOF_VARIANT_MMA64(5, 5, 0, 0), CYTOF_VARIANT_MMA64(8, 8, 1, 0),
                                       CYTOF_VARIANT_MMA64(5, 5, 0, 1), CYTOF_VARIANT_MMA64(8, 8, 1, 1) };
// 48 < J <= 64 with L0, L1 <= 5: the WIDE layout of the same kernel (NTP = 4); [0] without, [1] with the imputation path
#define CYTOF_VARIANT_MMA64W(a, b, lidx, miss) \
  { a, b, 2, 2, kXWarpFloatsW, kXZFloatsW, CYTOF_FN8(mma64_kernel_fn, lidx, miss), true, kXThreads }
static const Variant kVariants64W[] = { CYTOF_VARIANT_MMA64W(5, 5, 2, 0), CYTOF_VARIANT_MMA64W(5, 5, 2, 1) };
static const Variant* pick_variant64(const cytof_desc* d) {
  if (!CYTOF_MMA64) return nullptr;
  if (d->K > 64 || d->J > 64 || (d->J <= 32 && d->K <= 32)) return nullptr;     // J, K <= 32: the 8-cell-tile kernel
  if (d->J > 48) {      // WIDE layout; L = 8/8 at these widths does not fit the register file (-> first kernel)
    const Variant* w = &kVariants64W[(d->flags & CYTOF_FLAG_NO_MISSING) ? 0 : 1];
    return (w->lm0 >= d->L0 && w->lm1 >= d->L1) ? w : nullptr;
  }
  const int first = (d->flags & CYTOF_FLAG_NO_MISSING) ? 0 : 2;
  for (int q = first; q < first + 2; ++q)
    if (kVariants64[q].lm0 >= d->L0 && kVariants64[q].lm1 >= d->L1) return &kVariants64[q];
  return nullptr;
}

// J, K <= 32 with 16-byte rows: the CTA-wide tcgen05 kernel (elbo_kernel_umma.cuh), selected with CYTOF_UMMA=1 in the
// environment.  Parity-green (tests/test_gpu_umma_variant.py) but MEASURED SLOWER than the per-warp mma.sync kernel at
// cfg4 (0.486 vs 0.269 ms per 1 M cells; profiles/README.md "round 2": TMEM reads run at 32 B/clk/SM, SS-mode MMAs
// stream their operands from shared memory at 128 B/clk, and the forward -> backward round trip through two
// single-thread issuers and the cell warps is ~5,000 cycles), so it is not the default.
#define CYTOF_VARIANT_UMMA(a, b, lidx) { a, b, 1, 1, 0, 0, CYTOF_FN8(umma_kernel_fn, lidx), true, 512, -1 }
static const Variant kVariantsU[] = { CYTOF_VARIANT_UMMA(3, 3, 0), CYTOF_VARIANT_UMMA(5, 3, 1), CYTOF_VARIANT_UMMA(5, 5, 2) };
static bool umma_enabled() {
  static const bool on = [] { const char* e = getenv("CYTOF_UMMA"); return e && e[0] == '1'; }();
  return on;
}
static const Variant* pick_variant_umma(const cytof_desc* d, const float* y) {
  if (!umma_enabled() || d->J > 32 || (d->J & 3) || d->K > 32 || (reinterpret_cast<uintptr_t>(y) & 15)) return nullptr;
  for (const Variant& v : kVariantsU)
    if (v.lm0 >= d->L0 && v.lm1 >= d->L1) return &v;
  return nullptr;
}

static const Variant* pick_variant(int J, int K, int L0, int L1) {
  const int jp = J <= 32 ? 1 : 2, kp = K <= 32 ? 1 : 2;
  const int need = jp > kp ? jp : kp;
  for (const Variant& v : kVariants)
    if (v.jp == need && v.kp == need && v.lm0 >= L0 && v.lm1 >= L1) return &v;
  return nullptr;
}

static size_t fused_smem_bytes(const Variant* v, int J, int K, int L0, int L1) {
  if (v->smem_bytes) return (size_t)umma_smem_bytes();
  const PartialLayout pl = partial_layout(J, K, L0, L1);
  // the J,K<=32 kernel keeps one partial record per warp in shared memory for its final reduction
  const size_t nwarps = variant_threads(v) / 32;
  const size_t records = v->smem_floats_extra ? nwarps : 1;
  const size_t loop_floats = nwarps * v->smem_floats_per_warp + v->smem_floats_extra;
  if (v->records_alias) return sizeof(float) * (loop_floats > records * pl.total ? loop_floats : records * pl.total);
  return sizeof(float) * (loop_floats + records * pl.total);
}

static int g_sm_count[64] = {0};
static int device_sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (g_sm_count[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_sm_count[dev] = n;
  }
  return g_sm_count[dev];
}

constexpr int kMaxCtasPerSm = 8;   // upper bound used to size the workspace
constexpr int kMinCellsPerWarp = 4;

int max_grid_blocks() { return device_sm_count() * kMaxCtasPerSm + CYTOF_MAX_SAMPLES; }

// resident CTAs per SM of one kernel variant at a given dynamic shared-memory size (cached per
// device; re-evaluated when a different problem shape changes the shared-memory footprint)
struct OccEntry { size_t smem, smem_attr; int occ; };
static OccEntry g_occ[64][26][8] = {};     // [0..15] kVariants, [16..19] kVariants64, [20..22] kVariantsU, [24..25] kVariants64W
static int variant_occupancy(const Variant* v, int fi, size_t smem) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 1;
  const bool is64 = v >= kVariants64 && v < kVariants64 + 4;
  const bool isu = v >= kVariantsU && v < kVariantsU + 3;
  const bool isw = v >= kVariants64W && v < kVariants64W + 2;
  OccEntry& e = g_occ[dev][isw ? 24 + (int)(v - kVariants64W) : isu ? 20 + (int)(v - kVariantsU)
                                 : is64 ? 16 + (int)(v - kVariants64) : (int)(v - kVariants)][fi];
  if (e.occ == 0 || e.smem != smem) {
    if (smem > 48 * 1024 && smem > e.smem_attr) {   // opt in to large dynamic shared memory
      cudaFuncSetAttribute(v->fn[fi], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      e.smem_attr = smem;
    }
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, v->fn[fi], variant_threads(v), smem) != cudaSuccess || n < 1) n = 1;
    if (n > kMaxCtasPerSm) n = kMaxCtasPerSm;
    e.occ = n;
    e.smem = smem;
  }
  return e.occ;
}

int validate_desc(const cytof_desc* d) {
  if (!d) return CYTOF_EINVAL;
  if (d->I < 1 || d->I > CYTOF_MAX_SAMPLES || d->J < 1 || d->J > CYTOF_MAX_J || d->K < 1 ||
      d->K > CYTOF_MAX_K || d->L0 < 1 || d->L0 > CYTOF_MAX_L || d->L1 < 1 || d->L1 > CYTOF_MAX_L)
    return CYTOF_ESHAPE;
  if (d->cell_begin[0] != 0) return CYTOF_EINVAL;
  for (int i = 0; i < d->I; ++i)
    if (d->cell_begin[i + 1] < d->cell_begin[i]) return CYTOF_EINVAL;
  if (d->noise_mode != CYTOF_NOISE_EXTERNAL && d->noise_mode != CYTOF_NOISE_PHILOX) return CYTOF_EINVAL;
  if (d->flags & CYTOF_FLAG_SAMPLE_IN_KERNEL) {     // the minibatch of every sample must fit its sampling domain
    for (int i = 0; i < d->I; ++i)
      if (d->sampler_N[i] < d->cell_begin[i + 1] - d->cell_begin[i] || d->sampler_N[i] > 0x7fffffffLL) return CYTOF_EINVAL;
  }
  if (!pick_variant(d->J, d->K, d->L0, d->L1)) return CYTOF_ESHAPE;
  return CYTOF_OK;
}

void fill_params(const cytof_desc* d, ElboParams* p) {
  p->I = d->I; p->J = d->J; p->K = d->K; p->L0 = d->L0; p->L1 = d->L1;
  p->noise_mode = d->noise_mode;
  p->model_noisy = d->model_noisy;
  p->noisy_sd = d->noisy_sd;
  p->scale = d->scale;
  p->seed = d->seed;
  p->step = (uint32_t)d->step;
  p->step_dev = reinterpret_cast<const long long*>(d->step_dev);
  for (int i = 0; i <= d->I; ++i) p->cell_begin[i] = d->cell_begin[i];
  for (int i = 0; i < d->I; ++i) {
    p->row_base[i] = d->row_base[i];
    p->row_global[i] = d->row_global[i];
    p->fac[i] = (float)d->fac[i];
  }
  p->sampler_on = (d->flags & CYTOF_FLAG_SAMPLE_IN_KERNEL) ? 1 : 0;
  p->sampler_seed = d->sampler_seed;
  if (p->sampler_on) {
    p->sampler_step = d->sampler_step;
    for (int i = 0; i < d->I; ++i) {
      p->sampler_N[i] = (int)d->sampler_N[i];
      p->sampler_hbits[i] = (unsigned char)sampler_half_bits(d->sampler_N[i]);
    }
  }
}

// Splits the grid over samples so that a CTA never mixes samples: every non-empty sample gets
// one CTA, the rest go one at a time to the sample with the most cells per CTA (min-max
// optimal).  The total never exceeds SMs x resident CTAs/SM: exactly one wave, no tail.  A "block" of the plan is
// one partial record: a whole CTA, or one of the `split` warp groups of a CTA (MMA variants).
static int plan_grid(const cytof_desc* d, ElboParams* p, int occ, int split, int threads) {
  const long long C = d->cell_begin[d->I];
  const int cap = device_sm_count() * occ * split;
  const long long bw = threads / 32 / split;        // warps per block
  long long want = (C + bw * kMinCellsPerWarp - 1) / (bw * kMinCellsPerWarp);
  if (want > cap) want = cap;
  int n[CYTOF_MAX_SAMPLES];
  long long B[CYTOF_MAX_SAMPLES];
  int used = 0;
  for (int i = 0; i < d->I; ++i) {
    B[i] = d->cell_begin[i + 1] - d->cell_begin[i];
    n[i] = B[i] > 0 ? 1 : 0;
    used += n[i];
  }
  for (; used < want; ++used) {
    int best = -1;
    double load = 0.0;
    for (int i = 0; i < d->I; ++i) {
      if (n[i] == 0 || n[i] >= B[i]) continue;
      const double li = (double)B[i] / n[i];
      if (li > load) { load = li; best = i; }
    }
    if (best < 0) break;
    ++n[best];
  }
  int b = 0;
  for (int i = 0; i < d->I; ++i) {
    p->blk_begin[i] = b;
    b += n[i];
  }
  p->blk_begin[d->I] = b;
  p->nblocks = b;
  return b;
}

size_t workspace_bytes(const cytof_desc* d) {
  const PartialLayout pl = partial_layout(d->J, d->K, d->L0, d->L1);
  return sizeof(float) * (size_t)pl.total * (size_t)max_grid_blocks() + 256;
}

size_t peer_mailbox_bytes(const cytof_desc* d, int world) {
  const BlockLayout bl = block_layout(d->I, d->J, d->K, d->L0, d->L1);
  return kPeerHeaderBytes + (size_t)2 * world * (bl.total + 2) * sizeof(double);
}

int launch_fwdbwd(const cytof_desc* d, const float* y, const int32_t* idx, const float* eps,
                  const float* globals, const uint64_t* zmask, float* grad_block, double* ll_lqy,
                  double* packed, void* workspace, size_t workspace_bytes_, cudaStream_t stream,
                  cudaError_t* err, const cytof_peer_desc* peer) {
  if ((d->flags & CYTOF_FLAG_SAMPLE_IN_KERNEL) && idx) return CYTOF_EINVAL;     // one source of rows, not two
  const Variant* v64 = pick_variant64(d);
  const Variant* vu = v64 ? nullptr : pick_variant_umma(d, y);
  const Variant* v = v64 ? v64 : (vu ? vu : pick_variant(d->J, d->K, d->L0, d->L1));
  ElboParams p;
  fill_params(d, &p);
  p.y = y; p.idx = idx; p.eps = eps; p.globals = globals;
  p.zmask = reinterpret_cast<const unsigned long long*>(zmask);
  p.y_out = nullptr;
  uintptr_t wp = (reinterpret_cast<uintptr_t>(workspace) + 15) & ~(uintptr_t)15;
  p.partials = reinterpret_cast<float*>(wp);
  const size_t smem = fused_smem_bytes(v, d->J, d->K, d->L0, d->L1);
  const bool gathered = idx != nullptr || (d->flags & CYTOF_FLAG_SAMPLE_IN_KERNEL);
  const int fi = v64 ? (d->model_noisy ? 1 : 0) + (gathered ? 2 : 0) + ((d->J - 32 > 8 && d->J <= 48) ? 4 : 0)
                     : (d->model_noisy ? 1 : 0) + (gathered ? 2 : 0) + ((d->flags & CYTOF_FLAG_NO_MISSING) ? 4 : 0);
  const int split = variant_split(v);
  const int nblocks = plan_grid(d, &p, variant_occupancy(v, fi, smem), split, variant_threads(v));
  const PartialLayout pl = partial_layout(d->J, d->K, d->L0, d->L1);
  const size_t need = sizeof(float) * (size_t)pl.total * (size_t)nblocks + (wp - reinterpret_cast<uintptr_t>(workspace));
  if (need > workspace_bytes_) return CYTOF_EWORKSPACE;
  const BlockLayout bl = block_layout(d->I, d->J, d->K, d->L0, d->L1);
  if (nblocks > 0) {
    KernelFn fn = v->fn[fi];
    *err = launch_pdl(fn, dim3((nblocks + split - 1) / split), dim3(variant_threads(v)), smem, stream, p);
    if (*err != cudaSuccess) return CYTOF_ECUDA;
  }
  const int nout = bl.total + 2;
  PeerArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.world = 1;
  if (peer && peer->world > 1) {
    pa.world = peer->world;
    pa.rank = peer->rank;
    pa.seq = peer->seq;
    pa.seq_dev = reinterpret_cast<const long long*>(peer->seq_dev);
    pa.nout = nout;
    // all finalise CTAs co-resident (8 x 256 threads per SM): each of them may wait for the peers' flags
    static const bool serial_sum = [] { const char* e = getenv("CYTOF_PEER_SERIAL_SUM"); return e && e[0] == '1'; }();
    pa.all_ctas = !serial_sum && fin_plan(d->I, d->J, d->K, d->L0, d->L1).n_ctas <= device_sm_count() * 8;
    for (int q = 0; q < peer->world; ++q) pa.mailbox[q] = static_cast<unsigned char*>(peer->mailbox[q]);
  }
  *err = launch_pdl(elbo_finalize_kernel, dim3(fin_plan(d->I, d->J, d->K, d->L0, d->L1).n_ctas), dim3(256), 0, stream, p, grad_block, ll_lqy,
                    packed, pa);
  return *err == cudaSuccess ? CYTOF_OK : CYTOF_ECUDA;
}

int launch_emit(const cytof_desc* d, const float* y, const int32_t* idx, const float* eps,
                const float* globals, float* y_out, float* eps_out, cudaStream_t stream, cudaError_t* err) {
  ElboParams p;
  fill_params(d, &p);
  p.y = y; p.idx = idx; p.eps = eps; p.globals = globals; p.zmask = nullptr; p.partials = nullptr;
  p.y_out = y_out;
  p.nblocks = 0;
  const long long n = d->cell_begin[d->I] * (long long)d->J;
  if (n > 0) {
    impute_emit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(p, eps_out);
    *err = cudaGetLastError();
    if (*err != cudaSuccess) return CYTOF_ECUDA;
  }
  return CYTOF_OK;
}

}  // namespace cytof
