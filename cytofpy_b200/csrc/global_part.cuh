// Device side of the O(3.5k-scalar) "global part": parameter block, fp64 helpers and the phase bodies over a thread
// range (tid, nt), so that a phase can run on a whole grid or inside one CTA (the scans run in the LAST CTA of the
// element kernels).  See global_part.cu for what each phase computes and the reference lines it follows.
#pragma once
#include <math.h>

#include "common.cuh"

namespace cytof {


enum { B_DELTA0 = 0, B_DELTA1, B_EPS, B_SIG2, B_ETA0, B_ETA1, B_W, B_ALPHA, B_V, B_H, B_COUNT };

struct GlobalParams {
  int I, J, K, L0, L1;
  int model_noisy, use_stick_break;
  int noise_mode;      // CYTOF_NOISE_EXTERNAL: eps_in[R] supplied; CYTOF_NOISE_PHILOX: drawn here
  int sig2_family;     // 0 LogNormal(p0 = loc, p1 = scale), 1 Gamma(p0 = conc, p1 = rate)
  double tau;
  double delta0_a, delta0_b, delta1_a, delta1_b, sig2_p0, sig2_p1, alpha_a, alpha_b, eps_a, eps_b;
  double eta0_c[CYTOF_MAX_L], eta1_c[CYTOF_MAX_L], W_c[CYTOF_MAX_K];
  // normalising constants of the priors (pure functions of the hyper-parameters: host lgamma)
  double n_delta0, n_delta1, n_sig2, n_alpha, n_eps, n_eta0, n_eta1, n_W;
  unsigned long long seed, step;
  const long long* step_dev;   // non-null: step = *step_dev + 1 (CUDA-graph replay)
  // real-vector layout: offset of each block in the R-vector (noise / reals) and in theta
  int roff[B_COUNT + 1];   // roff[b]..roff[b+1]
  int toff[B_COUNT];       // theta offset of m; log_s at toff + size
  int n_global;            // = 2 R
  int vae_off;             // theta offset of the imputation parameters
  // Adam + loss scaling
  double grad_scale, lr, beta1, beta2, adam_eps;
  long long step_host;
};

__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }
__device__ __forceinline__ double log_sigmoid_d(double x) { return -(fmax(-x, 0.0) + log1p(exp(-fabs(x)))); }
__device__ __forceinline__ bool squashed(int b) { return b == B_EPS || b == B_ETA0 || b == B_ETA1 || b == B_W || b == B_V || b == B_H; }
__device__ __forceinline__ int block_of(const GlobalParams& p, int q) {
  int b = 0;
  while (b + 1 < B_COUNT && q >= p.roff[b + 1]) ++b;
  return b;
}

// fp64 standard normal from Philox (Box-Muller, 53-bit uniform)
__device__ __forceinline__ double philox_normal_d(unsigned long long seed, unsigned long long step, unsigned q) {
  uint32_t c[4] = {q, 0x5eed5eedu, (uint32_t)step, (uint32_t)(step >> 32)};
  philox4x32_10(c, (uint32_t)seed ^ 0x9e3779b9u, (uint32_t)(seed >> 32));
  const unsigned long long a = ((unsigned long long)c[0] << 21) ^ (unsigned long long)(c[1] >> 11);
  const double u1 = ((double)(a & ((1ull << 53) - 1)) + 0.5) * (1.0 / 9007199254740992.0);
  const double u2 = ((double)c[2] * 4294967296.0 + (double)c[3] + 0.5) * (1.0 / 18446744073709551616.0);
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];   // fixed order: reproducible
  return t;
}

// Which stick-breaking row family a real belongs to: n = row length, k = position in the row.
struct RowPos { int n, k; };
__device__ __forceinline__ RowPos row_pos(const GlobalParams& p, int b, int local) {
  const int n = b == B_ETA0 ? p.L0 - 1 : (b == B_ETA1 ? p.L1 - 1 : p.K - 1);
  return RowPos{n, local % n};
}

// ---------------------------------------------------------------------------------------------
// Structure of both kernels: (1) an element-parallel phase that evaluates every fp64
// transcendental (exp / log / sigmoid, ~10^2 cycles each) with one thread per scalar, then
// (2) short serial scans (cumsum, cumprod, stick-breaking prefix sums) that only add and
// multiply.  aux1 / aux2 carry the per-element results between the phases and to the backward.
//   eta0 / eta1 / W element:  aux1 = log z,   aux2 = log(1 - z)   (z = clipped sigmoid)
//   H element:                aux1 = h                            v element: aux1 = v, aux2 = log v
//   delta element:            aux1 = exp(real)
// (1) runs on kElemCtas CTAs (one SM's fp64 rate is the limit of a single CTA: ~10^5 DFMA per
// step); per-CTA partial sums of lp / lq go to cta_part[2 * blockIdx.x + {0,1}].
constexpr int kElemCtas = 24;
constexpr int kElemThreads = 256;

__device__ __forceinline__ void
global_fwd_scan_body(const GlobalParams& p, const double* aux1, const double* aux2, float* __restrict__ globals,
                     unsigned long long* __restrict__ zmask, const double* cta_part, double* __restrict__ scalars,
                     double* extra, const int staged, double* dyn);
__device__ __forceinline__ void
global_bwd_scan_body(const GlobalParams& p, const double* __restrict__ real, const double* aux1, const double* aux2,
                     const double* packed, double* __restrict__ greal, double* extra, long long* adam_step_dev,
                     const int do_update, const int staged, double* dyn);

// true in exactly one CTA of the grid: the one that finishes last (its threads then see everything the
// other CTAs wrote before their ticket); the ticket is left at 0 for the next launch
__device__ __forceinline__ bool last_cta_done(unsigned* ticket) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1u);
  __syncthreads();
  if (!s_last) return false;
  if (threadIdx.x == 0) *ticket = 0u;
  __threadfence();
  return true;
}

// (2) one CTA: serial scans (adds / multiplies only), Z bit mask, final lp / lq.  Runs either as its
// own kernel or -- the normal case -- as the tail of the LAST CTA of the element kernel (one launch
// less per step; the inputs are then always staged through shared memory with L2-coherent loads).
__device__ __forceinline__ void
global_fwd_scan_body(const GlobalParams& p, const double* aux1, const double* aux2,
                     float* __restrict__ globals, unsigned long long* __restrict__ zmask,
                     const double* cta_part, double* __restrict__ scalars /* lp, lq_global */,
                     double* extra /* out: cumprod(v)[MAX_K] | logit of it [MAX_K] for the backward pass */,
                     const int staged, double* dyn) {
  __shared__ double sh[32];
  __shared__ double s_vc[CYTOF_MAX_K];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int I = p.I, J = p.J, K = p.K, L0 = p.L0, L1 = p.L1;
  const GlobalsLayout gl = globals_layout(I, J, K, L0, L1);
  double lp = 0.0;
  // the scans below are serial chains of loads: from global memory every step costs an L2 round trip
  // (the kernel took 10 us for 1.7 k values); stage aux1 / aux2 in shared memory with all threads first
  if (staged) {
    const int R = p.roff[B_COUNT];
    for (int q = tid; q < R; q += nt) { dyn[q] = __ldcg(aux1 + q); dyn[R + q] = __ldcg(aux2 + q); }
    __syncthreads();
    aux1 = dyn;
    aux2 = dyn + R;
  }

  if (tid == 0) {   // mu0 = -cumsum(delta0)   (Model.py:223)
    double c = 0.0;
    for (int l = 0; l < L0; ++l) { c += aux1[p.roff[B_DELTA0] + l]; globals[gl.mu0 + l] = (float)(-c); }
  }
  if (tid == 32) {  // mu1 = +cumsum(delta1)   (Model.py:227)
    double c = 0.0;
    for (int l = 0; l < L1; ++l) { c += aux1[p.roff[B_DELTA1] + l]; globals[gl.mu1 + l] = (float)c; }
  }
  if (tid == 64) {  // v prior Beta(conc, 1) and cumprod(v)  (Model.py:236-239,292-298)
    const double alpha = aux1[p.roff[B_ALPHA]];
    const double conc = p.use_stick_break ? alpha : alpha / (double)K;
    double cp = 1.0, slv = 0.0;
    for (int k = 0; k < K; ++k) {
      const double v = aux1[p.roff[B_V] + k];
      slv += aux2[p.roff[B_V] + k];
      cp = p.use_stick_break ? cp * v : v;
      s_vc[k] = cp;
    }
    lp += (conc - 1.0) * slv + (double)K * log(conc);
  }
  {   // eta0 / eta1 / W rows: log y_k = log z_k + sum_{m<k} log(1-z_m); Dirichlet prior; log|det J|
    const int rows0 = I * J, rows1 = I * J, rowsW = I;
    for (int rr0 = tid - 96; rr0 < rows0 + rows1 + rowsW; rr0 += nt) {   // threads 96.. : keep the scan warps free
      if (rr0 < 0) continue;
      int n, goff, ro;
      const double* conc;
      double norm;
      if (rr0 < rows0) { n = L0 - 1; ro = p.roff[B_ETA0] + rr0 * n; conc = p.eta0_c; goff = gl.le0 + rr0 * L0; norm = p.n_eta0; }
      else if (rr0 < rows0 + rows1) { const int rr = rr0 - rows0; n = L1 - 1; ro = p.roff[B_ETA1] + rr * n; conc = p.eta1_c; goff = gl.le1 + rr * L1; norm = p.n_eta1; }
      else { const int rr = rr0 - rows0 - rows1; n = K - 1; ro = p.roff[B_W] + rr * n; conc = p.W_c; goff = gl.lw + rr * K; norm = p.n_W; }
      double acc = 0.0;
      for (int k = 0; k < n; ++k) {
        const double ly = aux1[ro + k] + acc;
        globals[goff + k] = (float)ly;
        lp += (conc[k] - 1.0) * ly + aux2[ro + k] + ly;
        acc += aux2[ro + k];
      }
      globals[goff + n] = (float)acc;
      lp += (conc[n] - 1.0) * acc + norm;
    }
  }
  __syncthreads();
  for (int k = nt - 1 - tid; k >= 0 && k < K; k += nt) {     // for the backward pass (the last threads: the first hold Z rows)
    extra[k] = s_vc[k];
    extra[CYTOF_MAX_K + k] = log(s_vc[k]) - log1p(-s_vc[k]);
  }
  for (int j = tid; j < J; j += nt) {       // Z = 1[v > H], thresholded in fp64 (Model.py:33)
    unsigned long long m = 0ull;
    for (int k = 0; k < K; ++k)
      if (s_vc[k] > aux1[p.roff[B_H] + j * K + k]) m |= 1ull << k;
    zmask[j] = m;
  }
  lp = block_sum(lp, sh);
  if (tid == 0) {
    double lq = 0.0;
    for (int c = 0; c < kElemCtas; ++c) { lp += __ldcg(cta_part + 2 * c); lq += __ldcg(cta_part + 2 * c + 1); }   // fixed order
    scalars[0] = lp;
    scalars[1] = lq;
  }
}

// (1) of the backward as a body: element-parallel part of d(ll + lp)/d real over the threads (tid, nt) -- a whole grid
// (global_bwd_elem_kernel) or one CTA (the tail of the finalise kernel, elbo_kernel.cu).  s_lvc: K doubles of shared memory.
__device__ __forceinline__ void
global_bwd_elem_body(const GlobalParams& p, const double* __restrict__ real, const double* __restrict__ aux1,
                     double* __restrict__ aux2, const double* G, double* __restrict__ greal, const double* extra,
                     double* s_lvc, const int tid, const int nt) {
  const int K = p.K;
  const int R = p.roff[B_COUNT];
  const BlockLayout bl = block_layout(p.I, p.J, p.K, p.L0, p.L1);
  // ---- (0) logit of cumprod(v): left in the stash by the forward scan
  for (int k = threadIdx.x; k < K; k += blockDim.x) s_lvc[k] = extra[CYTOF_MAX_K + k];
  __syncthreads();

  // ---- (1) element-parallel part of d(ll + lp)/d real
  for (int q = tid; q < R; q += nt) {
    const int b = block_of(p, q);
    const int local = q - p.roff[b];
    const double r = real[q];
    switch (b) {
      case B_SIG2: {
        double g = 0.5 * exp(0.5 * r) * G[bl.dsig + local] + 1.0;
        g += p.sig2_family == 0 ? (-1.0 - (r - p.sig2_p0) / (p.sig2_p1 * p.sig2_p1)) : ((p.sig2_p0 - 1.0) - p.sig2_p1 * exp(r));
        greal[q] = g;
      } break;
      case B_EPS: {
        const double e = sigmoid_d(r);
        greal[q] = e * (1.0 - e) * (G[bl.deps + local] + (p.eps_a - 1.0) / e - (p.eps_b - 1.0) / (1.0 - e)) + (1.0 - 2.0 * e);
      } break;
      case B_ETA0: case B_ETA1: case B_W:
        aux2[q] = exp(aux1[q]);          // z itself, for the row scans below
        break;
      case B_H: {
        // Z chain (Model.py:22-34): s = sigmoid(x / tau), x = logit(vc_k) - logit(h_jk);
        // dL/dx = G_Z s (1-s) / tau;  d x / d real_H = -1
        const int k = local % K;
        const double h = aux1[q];
        const double s = sigmoid_d((s_lvc[k] - (log(h) - log1p(-h))) / p.tau);
        const double a = G[bl.dZ + local] * s * (1.0 - s) / p.tau;
        aux2[q] = a;
        greal[q] = -a + (1.0 - 2.0 * h);
      } break;
      default: break;                    // delta0 / delta1 / alpha / v: scan kernel
    }
  }
}

// (2) one CTA: suffix scans (cumsum / cumprod / stick-breaking backward), adds and multiplies only
__device__ __forceinline__ void
global_bwd_scan_body(const GlobalParams& p, const double* __restrict__ real, const double* aux1,
                     const double* aux2, const double* packed,
                     double* __restrict__ greal, double* extra, long long* adam_step_dev, const int do_update,
                     const int staged, double* dyn) {
  __shared__ double s_vc[CYTOF_MAX_K], s_gvc[CYTOF_MAX_K], s_suf[CYTOF_MAX_K], s_red[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int I = p.I, J = p.J, K = p.K, L0 = p.L0, L1 = p.L1;
  const BlockLayout bl = block_layout(I, J, K, L0, L1);
  if (staged) {      // serial suffix scans: stage aux1 / aux2 / the gradient block in shared memory (see fwd scan)
    const int R = p.roff[B_COUNT], NB = bl.total + 2;
    for (int q = tid; q < R; q += nt) { dyn[q] = __ldcg(aux1 + q); dyn[R + q] = __ldcg(aux2 + q); }
    for (int q = tid; q < NB; q += nt) dyn[2 * R + q] = __ldcg(packed + q);
    __syncthreads();
    aux1 = dyn;
    aux2 = dyn + R;
    packed = dyn + 2 * R;
  }
  const double* G = packed;
  for (int k = tid; k < K; k += nt) s_vc[k] = extra[k];      // cumprod(v) from the forward scan
  __syncthreads();
  if (tid == 96) {   // Adam bias corrections of this step for global_adam_kernel (two pow() off its critical path)
    const long long step = adam_step_dev ? (*adam_step_dev + 1) : p.step_host;
    extra[2 * CYTOF_MAX_K] = 1.0 - pow(p.beta1, (double)step);
    extra[2 * CYTOF_MAX_K + 1] = sqrt(1.0 - pow(p.beta2, (double)step));
    // every reader of the device step counter of this step has finished (the Adam kernel takes its bias
    // corrections from `extra`): bump it here, so the Adam kernel needs no grid-wide rendezvous
    if (adam_step_dev && do_update) *adam_step_dev += 1;
  }

  if (tid == 0) {
    double suf = 0.0;
    for (int l = L0 - 1; l >= 0; --l) {       // mu0 = -cumsum(delta0)
      suf += G[bl.dmu0 + l];
      const double d = aux1[p.roff[B_DELTA0] + l];
      greal[p.roff[B_DELTA0] + l] = -d * suf + (p.delta0_a - 1.0) - p.delta0_b * d + 1.0;
    }
  }
  if (tid == 32) {
    double suf = 0.0;
    for (int l = L1 - 1; l >= 0; --l) {       // mu1 = +cumsum(delta1)
      suf += G[bl.dmu1 + l];
      const double d = aux1[p.roff[B_DELTA1] + l];
      greal[p.roff[B_DELTA1] + l] = d * suf + (p.delta1_a - 1.0) - p.delta1_b * d + 1.0;
    }
  }
  for (int k = tid - 64; k >= 0 && k < K; k += nt) {   // column sums of dL/dx over j (threads 64..64+K)
    double acc = 0.0;
    for (int j = 0; j < J; ++j) acc += aux2[p.roff[B_H] + j * K + k];
    s_gvc[k] = acc / (s_vc[k] * (1.0 - s_vc[k]));
  }
  {   // stick-breaking rows: with g_k = dL/dlog y_k (+ Dirichlet (c_k - 1)):
      //   dL/dx_m = g_m (1 - z_m) - z_m sum_{k>m} g_k + [1 - 2 z_m - z_m (n-1-m)]   (log|det J| part)
    const int rows0 = I * J, rows1 = I * J, rowsW = I;
    for (int rr0 = tid - 128; rr0 < rows0 + rows1 + rowsW; rr0 += nt) {   // threads 128.. : keep the scan warps free
      if (rr0 < 0) continue;
      int n, go, ro;
      const double* conc;
      if (rr0 < rows0) { n = L0 - 1; ro = p.roff[B_ETA0] + rr0 * n; conc = p.eta0_c; go = bl.dle0 + rr0 * L0; }
      else if (rr0 < rows0 + rows1) { const int rr = rr0 - rows0; n = L1 - 1; ro = p.roff[B_ETA1] + rr * n; conc = p.eta1_c; go = bl.dle1 + rr * L1; }
      else { const int rr = rr0 - rows0 - rows1; n = K - 1; ro = p.roff[B_W] + rr * n; conc = p.W_c; go = bl.dlw + rr * K; }
      double suf = G[go + n] + conc[n] - 1.0;
      for (int m = n - 1; m >= 0; --m) {
        const double gm = G[go + m] + conc[m] - 1.0, z = aux2[ro + m];
        greal[ro + m] = gm * (1.0 - z) - z * suf + (1.0 - 2.0 * z - z * (double)(n - 1 - m));
        suf += gm;
      }
    }
  }
  __syncthreads();
  // cumprod backward, v prior Beta(conc, 1), alpha: the suffix sum is the only serial part (adds only)
  if (tid == 0) {
    double suf = 0.0;
    for (int m = K - 1; m >= 0; --m) { suf += s_gvc[m] * s_vc[m]; s_suf[m] = suf; }
  }
  if (tid == 32) {
    double slv = 0.0;
    for (int m = K - 1; m >= 0; --m) slv += aux2[p.roff[B_V] + m];      // same order as the serial form
    s_red[0] = slv;
  }
  __syncthreads();
  const double alpha = aux1[p.roff[B_ALPHA]];
  const double conc = p.use_stick_break ? alpha : alpha / (double)K;
  for (int m = tid; m < K; m += nt) {
    const double v = aux1[p.roff[B_V] + m];
    const double gv = p.use_stick_break ? s_suf[m] / v : s_gvc[m];
    greal[p.roff[B_V] + m] = v * (1.0 - v) * (gv + (conc - 1.0) / v) + (1.0 - 2.0 * v);
  }
  if (tid == 64) {
    const double dconc = s_red[0] + (double)K / conc;
    const double dalpha = p.use_stick_break ? dconc : dconc / (double)K;
    greal[p.roff[B_ALPHA]] = alpha * dalpha + (p.alpha_a - 1.0) - p.alpha_b * alpha + 1.0;
  }
}

// (3) as a body over the threads (tid, nt): chain to (m, log_s), loss = -elbo / Nsum, NaN scrub (global blocks only), Adam
__device__ __forceinline__ void
global_adam_body(const GlobalParams& p, double* __restrict__ theta, const double* __restrict__ eps,
                 const double* packed, const double* __restrict__ scalars_in,
                 const double* greal, double* __restrict__ adam_m, double* __restrict__ adam_v,
                 double* __restrict__ grad_out, double* __restrict__ out_scalars,
                 int* flag, int do_update, const double* extra, const int tid, const int nt) {
  const int I = p.I, J = p.J, K = p.K, L0 = p.L0, L1 = p.L1;
  const int R = p.roff[B_COUNT];
  const BlockLayout bl = block_layout(I, J, K, L0, L1);
  const double* G = packed;
  const double bc1 = extra[2 * CYTOF_MAX_K], bc2s = extra[2 * CYTOF_MAX_K + 1];   // from the backward scan
  // +inf in the ll slot = the fused peer exchange gave up (elbo_finalize_kernel): this rank and, through the abort word,
  // every other rank skip the update and report code 2, so the replicas stop together instead of diverging
  const bool poisoned = G[bl.total] > 1.0e300;
  if (poisoned) do_update = 0;
  auto adam = [&](int t, double g, bool scrub) {
    g *= p.grad_scale;
    if (scrub && g != g) { g = 0.0; if (!poisoned) { if (flag) *flag = 1; out_scalars[4] = 1.0; } }   // same value from every writer
    if (grad_out) grad_out[t] = g;
    if (!do_update) return;
    const double mt = adam_m[t] + (1.0 - p.beta1) * (g - adam_m[t]);
    const double vt = p.beta2 * adam_v[t] + (1.0 - p.beta2) * g * g;
    adam_m[t] = mt;
    adam_v[t] = vt;
    theta[t] -= (p.lr / bc1) * (mt / (sqrt(vt) / bc2s + p.adam_eps));
  };
  for (int q = tid; q < R; q += nt) {
    const int b = block_of(p, q);
    const int sz = p.roff[b + 1] - p.roff[b], loc_i = p.toff[b] + (q - p.roff[b]);
    const double m = theta[loc_i], ls = theta[loc_i + sz], gr = greal[q], e = eps[q];
    double dloc, dsc, dlogsc;
    if (squashed(b)) {
      const double sm = sigmoid_d(m), ss = sigmoid_d(ls);
      dloc = 20.0 * sm * (1.0 - sm); dsc = 10.0 * ss * (1.0 - ss); dlogsc = 1.0 - ss;
    } else { dloc = 1.0; dsc = exp(ls); dlogsc = 1.0; }
    adam(loc_i, gr * dloc, true);                      // d elbo / d m
    adam(loc_i + sz, gr * e * dsc + dlogsc, true);     // d elbo / d log_s  (+ from -lq)
  }
  for (int q = tid; q < I * J; q += nt) {              // imputation parameters (not scrubbed, fit.py:40-45)
    const int i = q / J, j = q % J;
    adam(p.vae_off + i * 2 * J + j, G[bl.dvm + q], false);
    adam(p.vae_off + i * 2 * J + J + j, G[bl.dvls + q] - G[bl.dlq + q], false);
  }
  if (tid == 0) {
    const double ll = G[bl.total], lqy = G[bl.total + 1], lp = scalars_in[0], lq = scalars_in[1] + lqy;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    out_scalars[0] = poisoned ? nan : ll + lp - lq;
    out_scalars[1] = poisoned ? nan : ll;
    out_scalars[2] = lp;
    out_scalars[3] = poisoned ? nan : lq;
    if (poisoned) out_scalars[4] = 2.0;
  }
}


// host side (global_part.cu)
int global_make_params(const cytof_global_desc* g, GlobalParams* p);
size_t global_scan_smem_bytes(const void* fn, int which, size_t want);

}  // namespace cytof
