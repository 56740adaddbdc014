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
#include "global_part.cuh"

namespace cytof {

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
  if (blockIdx.x == 0 && threadIdx.x == 0) {      // the Adam kernel (next launch) raises both when it scrubs a NaN
    if (flag) *flag = 0;
    out_scalars[4] = 0.0;
  }

  global_bwd_elem_body(p, real, aux1, aux2, packed, greal, extra, s_lvc, blockIdx.x * blockDim.x + threadIdx.x,
                       gridDim.x * blockDim.x);
  if (ticket && last_cta_done(ticket)) global_bwd_scan_body(p, real, aux1, aux2, packed, greal, extra, adam_step_dev, do_update, 1, dyn);
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
  global_adam_body(p, theta, eps, packed, scalars_in, greal, adam_m, adam_v, grad_out, out_scalars, flag, do_update, extra,
                   blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
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
size_t global_scan_smem_bytes(const void* fn, int which, size_t want) {
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
  const size_t fused = global_scan_smem_bytes((const void*)global_fwd_elem_kernel, 0, want);
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
  const size_t fused = global_scan_smem_bytes((const void*)global_bwd_elem_kernel, 1, want);
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
