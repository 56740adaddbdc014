// Posterior cluster-label sampler: forward half of the per-cell likelihood without gradient,
// emitting for every cell the (K+1) class probabilities and one categorical draw.
// Replaces reference cytofpy/model/post_proc/lam_post.py:6-55 (label 0 = noisy class,
// labels 1..K = features).  One warp per cell, lane = marker j / lane = feature k; J, K <= 64.
#include "common.cuh"

namespace cytof {

__global__ void __launch_bounds__(256, 2)
label_sample_kernel(const ElboParams p, const float* __restrict__ y, int32_t* __restrict__ lam_out,
                    float* __restrict__ probs_out) {
  __shared__ float sDm[8][64];
  __shared__ float sPm[8][65];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int I = p.I, J = p.J, K = p.K, L0 = p.L0, L1 = p.L1;
  const GlobalsLayout gl = globals_layout(I, J, K, L0, L1);
  const float* __restrict__ G = p.globals;
  const long long C = p.cell_begin[I];
  float* sD = sDm[warp];
  float* sP = sPm[warp];
  for (long long c = (long long)blockIdx.x * 8 + warp; c < C; c += (long long)gridDim.x * 8) {
    int i = 0;
    while (i + 1 < I && c >= p.cell_begin[i + 1]) ++i;
    const float sig = G[gl.sig + i], inv2s2 = 0.5f / (sig * sig), cbase = -logf(sig) - kHalfLog2Pi;
    float a0sum = 0.f, y2sum = 0.f;
    for (int j = lane; j < 64; j += 32) {
      float D = 0.f;
      if (j < J) {
        const float yv = y[c * J + j];
        float A[2];
        for (int z = 0; z < 2; ++z) {
          const int L = z ? L1 : L0;
          const float* mu = G + (z ? gl.mu1 : gl.mu0);
          const float* le = G + (z ? gl.le1 : gl.le0) + (size_t)(i * J + j) * L;
          float M = -INFINITY;
          for (int l = 0; l < L; ++l) { const float d = yv - mu[l]; M = fmaxf(M, le[l] + cbase - d * d * inv2s2); }
          float S = 0.f;
          for (int l = 0; l < L; ++l) { const float d = yv - mu[l]; S += expf(le[l] + cbase - d * d * inv2s2 - M); }
          A[z] = M + logf(S);
        }
        D = A[1] - A[0];
        a0sum += A[0];
        y2sum += yv * yv;
      }
      sD[j] = D;
    }
    a0sum = warp_sum(a0sum);
    y2sum = warp_sum(y2sum);
    __syncwarp();
    // log_pi[0] = log eps + sum_j log N(y; 0, s);  log_pi[k+1] = log1p(-eps) + log W_k + f_k
    const float e_i = p.model_noisy ? G[gl.eps + i] : 0.f;
    const float lp0 = p.model_noisy ? logf(e_i) + (float)J * (-kHalfLog2Pi - logf(p.noisy_sd)) - y2sum / (2.f * p.noisy_sd * p.noisy_sd)
                                    : -INFINITY;
    const float l1me = p.model_noisy ? log1pf(-e_i) : 0.f;
    float v[3], mx = lp0;          // lane handles classes 1+lane and 1+lane+32; class 0 is uniform
    for (int kp = 0; kp < 2; ++kp) {
      const int k = lane + 32 * kp;
      float acc = -INFINITY;
      if (k < K) {
        acc = a0sum + G[gl.lw + i * K + k] + l1me;
        for (int j = 0; j < J; ++j) if ((p.zmask[j] >> k) & 1ull) acc += sD[j];
      }
      v[kp] = acc;
      mx = fmaxf(mx, acc);
    }
    mx = warp_max(mx);
    const float p0 = expf(lp0 - mx);
    float e0 = expf(v[0] - mx), e1 = expf(v[1] - mx);
    const float tot = warp_sum(e0 + e1) + p0;
    const float inv = 1.f / tot;
    if (lane == 0) sP[0] = p0 * inv;
    if (lane < K) sP[1 + lane] = e0 * inv;
    if (lane + 32 < K) sP[33 + lane] = e1 * inv;
    __syncwarp();
    if (probs_out)
      for (int t = lane; t <= K; t += 32) probs_out[c * (K + 1) + t] = sP[t];
    if (lane == 0) {   // inverse-CDF draw, u keyed by (seed, step, global row)
      uint32_t ctr[4] = {(uint32_t)(p.row_global[i] + (c - p.cell_begin[i])), (uint32_t)((unsigned long long)(p.row_global[i] + (c - p.cell_begin[i])) >> 32),
                         0x1ab31u + (uint32_t)i, elbo_step(p)};
      philox4x32_10(ctr, (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
      const float u = ((float)(ctr[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
      float cdf = 0.f;
      int pick = K;
      for (int t = 0; t <= K; ++t) { cdf += sP[t]; if (u <= cdf) { pick = t; break; } }
      if (!p.model_noisy && pick == 0) pick = 1;
      lam_out[c] = pick;
    }
    __syncwarp();
  }
}

void fill_params(const cytof_desc* d, ElboParams* p);
int validate_desc(const cytof_desc* d);

int launch_labels(const cytof_desc* d, const float* y, const float* globals, const uint64_t* zmask,
                  int32_t* lam_out, float* probs_out, cudaStream_t stream, cudaError_t* err) {
  ElboParams p;
  fill_params(d, &p);
  p.y = y; p.idx = nullptr; p.eps = nullptr; p.globals = globals;
  p.zmask = reinterpret_cast<const unsigned long long*>(zmask);
  p.partials = nullptr; p.y_out = nullptr; p.nblocks = 0;
  const long long C = d->cell_begin[d->I];
  if (C > 0) {
    long long nb = (C + 7) / 8;
    if (nb > 148 * 8) nb = 148 * 8;
    label_sample_kernel<<<(unsigned)nb, 256, 0, stream>>>(p, y, lam_out, probs_out);
    *err = cudaGetLastError();
    if (*err != cudaSuccess) return CYTOF_ECUDA;
  }
  return CYTOF_OK;
}

}  // namespace cytof
