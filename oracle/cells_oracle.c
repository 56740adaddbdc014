/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * Plain-C fp64 restatement of the O(cells) part of CytofPy's ELBO and of its analytic
 * gradient ("gradient block"), one cell at a time, scalar loops only.  Same inputs and
 * outputs as the CUDA entry point cytof_elbo_fwdbwd_f32 (include/cytof_b200.h) except that
 * everything is double.  It exists so that the GPU tests can check the kernel against an
 * fp64 answer at sizes where the numpy oracle (oracle/cells_closed_form.py) is too slow.
 *
 * Follows (paths relative to /root/reference):
 *   imputation                 cytofpy/model/vae.py:17-41
 *   mixtures, Z gate, LSE_k    cytofpy/model/Model.py:223-255
 *   noisy class                cytofpy/model/Model.py:259-266
 *   missingness                cytofpy/model/Model.py:36-38,269-274
 *   N_i/B_i weighting          cytofpy/model/Model.py:257,276
 * Pinning: tests/test_oracle_golden.py compares it with the numpy oracle on the golden
 * fixtures recorded from the live reference (oracle/gen_golden.py).
 *
 * Build: gcc -O2 -fPIC -shared -o oracle/_build/libcells_oracle.so oracle/cells_oracle.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXJ 64
#define MAXK 64
#define MAXL 8

static double lse(const double* a, int n) {
  double m = a[0], s = 0.0;
  for (int i = 1; i < n; ++i) if (a[i] > m) m = a[i];
  for (int i = 0; i < n; ++i) s += exp(a[i] - m);
  return m + log(s);
}
static double log_sigmoid(double u) { return -(fmax(-u, 0.0) + log1p(exp(-fabs(u)))); }

/* y [rows,J] (NaN = missing), idx [C] or NULL, eps [C,J] or NULL (treated as 0),
 * per-sample arrays of length I; Z [J,K] in {0,1}; block layout = cytof_grad_block_layout. */
int cells_oracle_fwdbwd(int I, int J, int K, int L0, int L1, int model_noisy, double noisy_sd,
                        double scale, const int64_t* cell_begin, const int64_t* row_base,
                        const double* fac, const double* y, const int64_t* idx, const double* eps,
                        const double* mu0, const double* mu1, const double* sig,
                        const double* logeta0, const double* logeta1, const double* logW,
                        const double* Z, const double* eps_noisy, const double* b, const double* vm,
                        const double* vls, double* block, double* ll_lqy) {
  if (J > MAXJ || K > MAXK || L0 > MAXL || L1 > MAXL) return -2;
  const double LOG2PI = log(2.0 * M_PI);
  const int64_t o_dmu0 = 0, o_dmu1 = L0, o_dsig = o_dmu1 + L1, o_dle0 = o_dsig + I,
                o_dle1 = o_dle0 + (int64_t)I * J * L0, o_dlw = o_dle1 + (int64_t)I * J * L1,
                o_dZ = o_dlw + (int64_t)I * K, o_deps = o_dZ + (int64_t)J * K, o_dvm = o_deps + I,
                o_dvls = o_dvm + (int64_t)I * J, o_dlq = o_dvls + (int64_t)I * J,
                total = o_dlq + (int64_t)I * J;
  memset(block, 0, sizeof(double) * (size_t)total);
  double ll = 0.0, lqy = 0.0;
  for (int i = 0; i < I; ++i) {
    const double s = sig[i], s2 = s * s;
    for (int64_t c = cell_begin[i]; c < cell_begin[i + 1]; ++c) {
      const int64_t row = row_base[i] + (idx ? idx[c] : c - cell_begin[i]);
      double yv[MAXJ], e[MAXJ], A0[MAXJ], A1[MAXJ], p0[MAXJ][MAXL], p1[MAXJ][MAXL], f[MAXK], g[MAXK];
      int miss[MAXJ];
      double sumA0 = 0.0, sumy2 = 0.0, cell = 0.0;
      for (int j = 0; j < J; ++j) {
        const double v = y[row * J + j];
        miss[j] = isnan(v);
        e[j] = (miss[j] && eps) ? eps[c * J + j] : 0.0;
        yv[j] = miss[j] ? vm[i * J + j] + exp(vls[i * J + j]) * e[j] : v;
        double a[MAXL];
        for (int l = 0; l < L0; ++l) {
          const double d = yv[j] - mu0[l];
          a[l] = logeta0[((int64_t)i * J + j) * L0 + l] - log(s) - 0.5 * LOG2PI - d * d / (2 * s2);
        }
        A0[j] = lse(a, L0);
        for (int l = 0; l < L0; ++l) p0[j][l] = exp(a[l] - A0[j]);
        for (int l = 0; l < L1; ++l) {
          const double d = yv[j] - mu1[l];
          a[l] = logeta1[((int64_t)i * J + j) * L1 + l] - log(s) - 0.5 * LOG2PI - d * d / (2 * s2);
        }
        A1[j] = lse(a, L1);
        for (int l = 0; l < L1; ++l) p1[j][l] = exp(a[l] - A1[j]);
        sumA0 += A0[j];
        sumy2 += yv[j] * yv[j];
      }
      for (int k = 0; k < K; ++k) {
        double acc = sumA0 + logW[i * K + k];
        for (int j = 0; j < J; ++j) acc += Z[j * K + k] * (A1[j] - A0[j]);
        f[k] = acc;
      }
      const double lk = lse(f, K);
      double rho = 1.0, omr = 0.0, L = lk;
      if (model_noisy) {
        const double ei = eps_noisy[i];
        const double q = lk + log1p(-ei);
        const double z = J * (-0.5 * LOG2PI - log(noisy_sd)) - sumy2 / (2 * noisy_sd * noisy_sd) + log(ei);
        const double pair[2] = {q, z};
        L = lse(pair, 2);
        rho = exp(q - L);
        omr = exp(z - L);
        block[o_deps + i] += fac[i] * (-rho / (1 - ei) + omr / ei);
      }
      cell = L;
      for (int k = 0; k < K; ++k) {
        g[k] = fac[i] * rho * exp(f[k] - lk);
        block[o_dlw + i * K + k] += g[k];
      }
      for (int j = 0; j < J; ++j) {
        double c1 = 0.0;
        for (int k = 0; k < K; ++k) {
          c1 += g[k] * Z[j * K + k];
          block[o_dZ + j * K + k] += g[k] * (A1[j] - A0[j]);
        }
        const double c0 = fac[i] * rho - c1;
        double dy = 0.0;
        for (int l = 0; l < L0; ++l) {
          const double w = c0 * p0[j][l], t = (yv[j] - mu0[l]) / s2;
          block[o_dle0 + ((int64_t)i * J + j) * L0 + l] += w;
          block[o_dmu0 + l] += w * t;
          block[o_dsig + i] += w * (t * t * s - 1.0 / s);
          dy -= w * t;
        }
        for (int l = 0; l < L1; ++l) {
          const double w = c1 * p1[j][l], t = (yv[j] - mu1[l]) / s2;
          block[o_dle1 + ((int64_t)i * J + j) * L1 + l] += w;
          block[o_dmu1 + l] += w * t;
          block[o_dsig + i] += w * (t * t * s - 1.0 / s);
          dy -= w * t;
        }
        if (miss[j]) {
          const double u = b[i * 3] + b[i * 3 + 1] * yv[j] + b[i * 3 + 2] * yv[j] * yv[j];
          cell += log_sigmoid(u);
          if (model_noisy) dy += fac[i] * omr * (-yv[j] / (noisy_sd * noisy_sd));
          dy += fac[i] * exp(log_sigmoid(-u)) * (b[i * 3 + 1] + 2 * b[i * 3 + 2] * yv[j]);
          block[o_dvm + i * J + j] += dy;
          block[o_dvls + i * J + j] += dy * exp(vls[i * J + j]) * e[j];
          block[o_dlq + i * J + j] += -fac[i] * scale;
          lqy += fac[i] * scale * (-0.5 * e[j] * e[j] - vls[i * J + j] - 0.5 * LOG2PI);
        }
      }
      ll += fac[i] * cell;
    }
  }
  ll_lqy[0] = ll;
  ll_lqy[1] = lqy;
  return 0;
}
