// The O(3.5k-scalar) "global part" of one ADVI step as two single-CTA fp64 kernels that
// bracket the fused per-cell kernel, replacing ~5,000 tiny ATen dispatches of the reference:
//
//   global_fwd_kernel      reference model_param.py:46-55 (reparameterised sample of every
//                          global block), :57-71 (exp / sigmoid / stick-breaking), Model.py:212,
//                          223,227,236-243 (sig, mu0, mu1, cumprod(v), Z = 1[v > H] in fp64),
//                          Model.py:289-304 (log prior + log|det J|), model_param.py:91-92
//                          (entropy term).  Emits the fp32 `globals` + `zmask` the cell kernel
//                          consumes and the scalars lp, lq_global.
//   global_bwd_adam_kernel the autograd backward of all of the above (closed forms below), the
//                          chain from the cell kernel's gradient block to every variational
//                          parameter, loss = -elbo / Nsum (fit.py:34), the NaN scrub of the
//                          global blocks (fit.py:24-30,44-45) and the Adam update (fit.py:47,83).
//
// theta layout (fp64, one flat buffer): for every block in the reference's registration order
//   delta0 | delta1 | [eps] | sig2 | eta0 | eta1 | W | alpha | v | H
// its means m[size] then its log-sds log_s[size] (= vp[0], vp[1] contiguous), followed by, per
// sample, the imputation mean[J] and log_sd[J] (Model.py:202-208).
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

__device__ double block_sum(double v, double* sh) {
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

__global__ void __launch_bounds__(kElemThreads)
global_fwd_elem_kernel(const GlobalParams p, const double* __restrict__ theta, const double* __restrict__ eps_in,
                       double* __restrict__ real, double* __restrict__ eps_out, double* __restrict__ aux1,
                       double* __restrict__ aux2, float* __restrict__ globals, double* __restrict__ cta_part,
                       unsigned* ticket, unsigned long long* zmask, double* scalars, double* extra) {
  pdl_enter();
  extern __shared__ double dyn[];
  __shared__ double sh[32];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  const int I = p.I, J = p.J, K = p.K, L0 = p.L0, L1 = p.L1;
  const int R = p.roff[B_COUNT];
  const GlobalsLayout gl = globals_layout(I, J, K, L0, L1);
  const double HL2PI = 0.9189385332046727;
  const double tiny = 2.2250738585072014e-308, one_m_eps = 1.0 - 2.220446049250313e-16;
  double lq = 0.0, lp = 0.0;

  // ---- (1) element-parallel: reparameterised draw (model_param.py:46-55), entropy term
  //      (:91-92), element-wise transforms and the element-wise parts of prior + log|det J|
  for (int q = tid; q < R; q += nt) {
    const int b = block_of(p, q);
    const int sz = p.roff[b + 1] - p.roff[b], local = q - p.roff[b], loc_i = p.toff[b] + local;
    const double m = theta[loc_i], ls = theta[loc_i + sz];
    double loc, sc;
    if (squashed(b)) { loc = 20.0 * sigmoid_d(m) - 10.0; sc = 10.0 * sigmoid_d(ls); }
    else { loc = m; sc = exp(ls); }
    const double e = p.noise_mode == CYTOF_NOISE_EXTERNAL ? eps_in[q] : philox_normal_d(p.seed, p.step_dev ? (unsigned long long)(*p.step_dev + 1) : p.step, (unsigned)q);
    const double r = loc + e * sc;
    real[q] = r;
    eps_out[q] = e;
    lq += -0.5 * e * e - log(sc) - HL2PI;
    switch (b) {
      case B_DELTA0: case B_DELTA1: {
        const double d = exp(r), a_ = b == B_DELTA0 ? p.delta0_a : p.delta1_a, b_ = b == B_DELTA0 ? p.delta0_b : p.delta1_b;
        aux1[q] = d;
        lp += (b == B_DELTA0 ? p.n_delta0 : p.n_delta1) + (a_ - 1.0) * r - b_ * d + r;
      } break;
      case B_EPS: {
        const double ee = sigmoid_d(r);
        globals[gl.eps + local] = (float)ee;
        lp += (p.eps_a - 1.0) * log(ee) + (p.eps_b - 1.0) * log1p(-ee) + p.n_eps + log(ee) + log1p(-ee);
      } break;
      case B_SIG2: {
        globals[gl.sig + local] = (float)exp(0.5 * r);
        if (p.sig2_family == 0)
          lp += -r + p.n_sig2 - (r - p.sig2_p0) * (r - p.sig2_p0) / (2.0 * p.sig2_p1 * p.sig2_p1) + r;
        else
          lp += p.n_sig2 + (p.sig2_p0 - 1.0) * r - p.sig2_p1 * exp(r) + r;
      } break;
      case B_ETA0: case B_ETA1: case B_W: {
        const RowPos rp = row_pos(p, b, local);
        double z = sigmoid_d(r - log((double)(rp.n - rp.k)));
        z = fmin(fmax(z, tiny), one_m_eps);
        aux1[q] = log(z);
        aux2[q] = log1p(-z);
      } break;
      case B_ALPHA: {
        const double alpha = exp(r);
        aux1[q] = alpha;
        lp += p.n_alpha + (p.alpha_a - 1.0) * r - p.alpha_b * alpha + r;
      } break;
      case B_V: {
        const double v = sigmoid_d(r), lv = log(v);
        aux1[q] = v;
        aux2[q] = lv;
        lp += lv + log1p(-v);          // log|det J|; the Beta(conc, 1) part needs alpha: phase 2
      } break;
      default: {                        // B_H: Uniform(0,1) prior contributes 0
        const double h = sigmoid_d(r);
        aux1[q] = h;
        lp += log(h) + log1p(-h);
      }
    }
  }
  if (!p.model_noisy)
    for (int i = tid; i < I; i += nt) globals[gl.eps + i] = 0.f;
  for (int q = tid; q < I * J; q += nt) {   // imputation parameters straight into globals
    globals[gl.vm + q] = (float)theta[p.vae_off + (q / J) * 2 * J + (q % J)];
    globals[gl.vls + q] = (float)theta[p.vae_off + (q / J) * 2 * J + J + (q % J)];
  }
  lp = block_sum(lp, sh);
  lq = block_sum(lq, sh);
  if (threadIdx.x == 0) { cta_part[2 * blockIdx.x] = lp; cta_part[2 * blockIdx.x + 1] = lq; }
  // fused tail: the last CTA to finish runs the scans (ticket == NULL: separate global_fwd_scan_kernel)
  if (ticket && last_cta_done(ticket)) global_fwd_scan_body(p, aux1, aux2, globals, zmask, cta_part, scalars, extra, 1, dyn);
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

__global__ void __launch_bounds__(1024, 1)
global_fwd_scan_kernel(const GlobalParams p, const double* aux1, const double* aux2, float* __restrict__ globals,
                       unsigned long long* __restrict__ zmask, const double* cta_part, double* __restrict__ scalars,
                       double* extra, const int staged) {
  pdl_enter();
  extern __shared__ double dyn[];
  global_fwd_scan_body(p, aux1, aux2, globals, zmask, cta_part, scalars, extra, staged, dyn);
}

// ---------------------------------------------------------------------------------------------
// packed: [grad block (BlockLayout, doubles) | ll | log_qy] — already summed over ranks.
// out_scalars: elbo, ll, lp, lq (lq includes log_qy), scrubbed.  grad_out (may be NULL): d loss / d theta.
__global__ void __launch_bounds__(kElemThreads)
global_bwd_elem_kernel(const GlobalParams p, const double* __restrict__ real, const double* __restrict__ aux1,
                       double* __restrict__ aux2, const double* __restrict__ packed, double* __restrict__ greal,
                       int* __restrict__ flag, unsigned* ticket, double* extra, long long* adam_step_dev,
                       double* out_scalars, const int do_update) {
  pdl_enter();
  extern __shared__ double dyn[];
  __shared__ double s_lvc[CYTOF_MAX_K];
  const int I = p.I, J = p.J, K = p.K, L0 = p.L0, L1 = p.L1;
  const int R = p.roff[B_COUNT];
  const BlockLayout bl = block_layout(I, J, K, L0, L1);
  const double* __restrict__ G = packed;
  if (blockIdx.x == 0 && threadIdx.x == 0) {      // the Adam kernel (next launch) raises both when it scrubs a NaN
    if (flag) *flag = 0;
    out_scalars[4] = 0.0;
  }

  // ---- (0) logit of cumprod(v): left in the stash by the forward scan
  for (int k = threadIdx.x; k < K; k += blockDim.x) s_lvc[k] = extra[CYTOF_MAX_K + k];
  __syncthreads();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  (void)I; (void)J; (void)L0; (void)L1;

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
  if (ticket && last_cta_done(ticket)) global_bwd_scan_body(p, real, aux1, aux2, packed, greal, extra, adam_step_dev, do_update, 1, dyn);
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

__global__ void __launch_bounds__(1024, 1)
global_bwd_scan_kernel(const GlobalParams p, const double* __restrict__ real, const double* aux1, const double* aux2,
                       const double* packed, double* __restrict__ greal, double* extra, long long* adam_step_dev,
                       const int do_update, const int staged) {
  pdl_enter();
  extern __shared__ double dyn[];
  global_bwd_scan_body(p, real, aux1, aux2, packed, greal, extra, adam_step_dev, do_update, staged, dyn);
}

// (3) element-parallel: chain to (m, log_s), loss = -elbo / Nsum, NaN scrub (global blocks only), Adam
__global__ void __launch_bounds__(kElemThreads)
global_adam_kernel(const GlobalParams p, double* __restrict__ theta, const double* __restrict__ eps,
                   const double* __restrict__ packed, const double* __restrict__ scalars_in,
                   const double* __restrict__ greal, double* __restrict__ adam_m, double* __restrict__ adam_v,
                   double* __restrict__ grad_out, double* __restrict__ out_scalars,
                   int* flag, int do_update, const double* extra) {
  pdl_enter();
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  const int I = p.I, J = p.J, K = p.K, L0 = p.L0, L1 = p.L1;
  const int R = p.roff[B_COUNT];
  const BlockLayout bl = block_layout(I, J, K, L0, L1);
  const double* __restrict__ G = packed;
  const double bc1 = extra[2 * CYTOF_MAX_K], bc2s = extra[2 * CYTOF_MAX_K + 1];   // from the backward scan
  auto adam = [&](int t, double g, bool scrub) {
    g *= p.grad_scale;
    if (scrub && g != g) { g = 0.0; if (flag) *flag = 1; out_scalars[4] = 1.0; }   // same value from every writer
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
    out_scalars[0] = ll + lp - lq;
    out_scalars[1] = ll;
    out_scalars[2] = lp;
    out_scalars[3] = lq;
  }
}

// ---------------------------------------------------------------------------------------------
static void fill_layout(GlobalParams* p) {
  const int I = p->I, J = p->J, K = p->K, L0 = p->L0, L1 = p->L1;
  const int sz[B_COUNT] = {L0, L1, p->model_noisy ? I : 0, I, I * J * (L0 - 1), I * J * (L1 - 1), I * (K - 1), 1, K, J * K};
  int r = 0, t = 0;
  for (int b = 0; b < B_COUNT; ++b) {
    p->roff[b] = r;
    p->toff[b] = t;
    r += sz[b];
    t += 2 * sz[b];
  }
  p->roff[B_COUNT] = r;
  p->n_global = t;
  p->vae_off = t;
}

int global_make_params(const cytof_global_desc* g, GlobalParams* p) {
  if (!g) return CYTOF_EINVAL;
  if (g->I < 1 || g->I > CYTOF_MAX_SAMPLES || g->J < 1 || g->J > CYTOF_MAX_J || g->K < 2 || g->K > CYTOF_MAX_K ||
      g->L0 < 2 || g->L0 > CYTOF_MAX_L || g->L1 < 2 || g->L1 > CYTOF_MAX_L)
    return CYTOF_ESHAPE;
  p->I = g->I; p->J = g->J; p->K = g->K; p->L0 = g->L0; p->L1 = g->L1;
  p->model_noisy = g->model_noisy; p->use_stick_break = g->use_stick_break;
  p->noise_mode = g->noise_mode; p->sig2_family = g->sig2_family;
  p->tau = g->tau;
  p->delta0_a = g->delta0[0]; p->delta0_b = g->delta0[1];
  p->delta1_a = g->delta1[0]; p->delta1_b = g->delta1[1];
  p->sig2_p0 = g->sig2[0]; p->sig2_p1 = g->sig2[1];
  p->alpha_a = g->alpha[0]; p->alpha_b = g->alpha[1];
  p->eps_a = g->eps[0]; p->eps_b = g->eps[1];
  for (int l = 0; l < CYTOF_MAX_L; ++l) { p->eta0_c[l] = g->eta0_conc[l]; p->eta1_c[l] = g->eta1_conc[l]; }
  for (int k = 0; k < CYTOF_MAX_K; ++k) p->W_c[k] = g->W_conc[k];
  p->n_delta0 = p->delta0_a * log(p->delta0_b) - lgamma(p->delta0_a);
  p->n_delta1 = p->delta1_a * log(p->delta1_b) - lgamma(p->delta1_a);
  p->n_alpha = p->alpha_a * log(p->alpha_b) - lgamma(p->alpha_a);
  p->n_sig2 = p->sig2_family == 0 ? (-log(p->sig2_p1) - 0.9189385332046727)
                                  : (p->sig2_p0 * log(p->sig2_p1) - lgamma(p->sig2_p0));
  p->n_eps = lgamma(p->eps_a + p->eps_b) - lgamma(p->eps_a) - lgamma(p->eps_b);
  {
    double s0 = 0, l0 = 0, s1 = 0, l1 = 0, sw = 0, lw = 0;
    for (int l = 0; l < p->L0; ++l) { s0 += p->eta0_c[l]; l0 += lgamma(p->eta0_c[l]); }
    for (int l = 0; l < p->L1; ++l) { s1 += p->eta1_c[l]; l1 += lgamma(p->eta1_c[l]); }
    for (int k = 0; k < p->K; ++k) { sw += p->W_c[k]; lw += lgamma(p->W_c[k]); }
    p->n_eta0 = lgamma(s0) - l0;
    p->n_eta1 = lgamma(s1) - l1;
    p->n_W = lgamma(sw) - lw;
  }
  p->seed = g->seed; p->step = g->step;
  p->step_dev = reinterpret_cast<const long long*>(g->step_dev);
  p->grad_scale = g->grad_scale; p->lr = g->lr; p->beta1 = g->beta1; p->beta2 = g->beta2; p->adam_eps = g->adam_eps;
  p->step_host = g->adam_step;
  fill_layout(p);
  return CYTOF_OK;
}

int global_sizes(const cytof_global_desc* g, int64_t out[4]) {
  GlobalParams p;
  const int rc = global_make_params(g, &p);
  if (rc != CYTOF_OK) return rc;
  out[0] = p.roff[B_COUNT];                 // R: number of global reals (= noise length)
  out[1] = p.n_global;                      // 2R: scrubbed prefix of theta
  out[2] = p.n_global + 2 * p.I * p.J;      // P: length of theta
  // stash: real | eps | greal | aux1 | aux2 | CTA partials | tickets (2 used) | cumprod(v) | its logit | Adam bias corrections
  out[3] = 5 * (int64_t)p.roff[B_COUNT] + 2 * kElemCtas + 2 + 2 * CYTOF_MAX_K + 2;
  return CYTOF_OK;
}

// dynamic shared memory for the staged scans: `want` bytes if they fit (opt-in above 48 KB, once per
// device and kernel), else 0 = separate scan kernels reading global memory
static size_t scan_smem_bytes(const void* fn, int which, size_t want) {
  static size_t granted[64][4] = {};
  static const bool unfused = [] { const char* e = getenv("CYTOF_GLOBAL_UNFUSED"); return e && e[0] == '1'; }();
  if (unfused || want > 200 * 1024) return 0;      // the environment switch lets the tests drive the fallback path
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return want <= 48 * 1024 ? want : 0;
  if (want > 48 * 1024 && want > granted[dev][which]) {
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want) != cudaSuccess) {
      (void)cudaGetLastError();
      return 0;
    }
    granted[dev][which] = want;
  }
  return want;
}

int launch_global_fwd(const cytof_global_desc* g, const double* theta, const double* eps_in, double* stash,
                      float* globals, uint64_t* zmask, double* scalars, cudaStream_t s, cudaError_t* err) {
  GlobalParams p;
  const int rc = global_make_params(g, &p);
  if (rc != CYTOF_OK) return rc;
  const int R = p.roff[B_COUNT];
  unsigned* tickets = reinterpret_cast<unsigned*>(stash + 5 * (size_t)R + 2 * kElemCtas);
  unsigned long long* zm = reinterpret_cast<unsigned long long*>(zmask);
  double* extra = stash + 5 * (size_t)R + 2 * kElemCtas + 2;
  const size_t want = sizeof(double) * 2 * (size_t)R;
  const size_t fused = scan_smem_bytes((const void*)global_fwd_elem_kernel, 0, want);
  const dim3 eg(kElemCtas), eb(kElemThreads);
  if (fused) {      // one launch: the last CTA of the element kernel runs the scans out of shared memory
    *err = launch_pdl(global_fwd_elem_kernel, eg, eb, fused, s, p, theta, eps_in, stash, stash + R, stash + 3 * R,
                      stash + 4 * R, globals, stash + 5 * R, tickets, zm, scalars, extra);
  } else {
    *err = launch_pdl(global_fwd_elem_kernel, eg, eb, 0, s, p, theta, eps_in, stash, stash + R, stash + 3 * R,
                      stash + 4 * R, globals, stash + 5 * R, nullptr, zm, scalars, extra);
    if (*err == cudaSuccess)
      *err = launch_pdl(global_fwd_scan_kernel, dim3(1), dim3(1024), 0, s, p, stash + 3 * R, stash + 4 * R, globals, zm,
                        stash + 5 * R, scalars, extra, 0);
  }
  return *err == cudaSuccess ? CYTOF_OK : CYTOF_ECUDA;
}
int launch_global_bwd(const cytof_global_desc* g, double* theta, double* stash, const double* packed,
                      const double* scalars_in, double* adam_m, double* adam_v, double* grad_out,
                      double* out_scalars, int64_t* step_dev, int32_t* flag, int do_update, cudaStream_t s,
                      cudaError_t* err) {
  GlobalParams p;
  const int rc = global_make_params(g, &p);
  if (rc != CYTOF_OK) return rc;
  const int R = p.roff[B_COUNT];
  unsigned* tickets = reinterpret_cast<unsigned*>(stash + 5 * (size_t)R + 2 * kElemCtas);
  const BlockLayout blh = block_layout(p.I, p.J, p.K, p.L0, p.L1);
  double* extra = stash + 5 * (size_t)R + 2 * kElemCtas + 2;
  long long* sd = reinterpret_cast<long long*>(step_dev);
  const size_t want = sizeof(double) * (2 * (size_t)R + (size_t)blh.total + 2);
  const size_t fused = scan_smem_bytes((const void*)global_bwd_elem_kernel, 1, want);
  const dim3 eg(kElemCtas), eb(kElemThreads);
  if (fused) {
    *err = launch_pdl(global_bwd_elem_kernel, eg, eb, fused, s, p, stash, stash + 3 * R, stash + 4 * R, packed,
                      stash + 2 * R, flag, tickets + 1, extra, sd, out_scalars, do_update);
  } else {
    *err = launch_pdl(global_bwd_elem_kernel, eg, eb, 0, s, p, stash, stash + 3 * R, stash + 4 * R, packed,
                      stash + 2 * R, flag, nullptr, extra, sd, out_scalars, do_update);
    if (*err == cudaSuccess)
      *err = launch_pdl(global_bwd_scan_kernel, dim3(1), dim3(1024), 0, s, p, stash, stash + 3 * R, stash + 4 * R, packed,
                        stash + 2 * R, extra, sd, do_update, 0);
  }
  if (*err == cudaSuccess)
    *err = launch_pdl(global_adam_kernel, eg, eb, 0, s, p, theta, stash + R, packed, scalars_in, stash + 2 * R, adam_m,
                      adam_v, grad_out, out_scalars, flag, do_update, extra);
  return *err == cudaSuccess ? CYTOF_OK : CYTOF_ECUDA;
}

}  // namespace cytof
