/* cytof_b200.h — C ABI of the B200-native CyTOF ELBO hot path.
 *
 * The reference (luiarthur/CytofPy) is pure Python and has NO FFI seam: its
 * hot path is Python methods of cytofpy.model.Model.  This header is therefore
 * a new boundary placed *behind* those methods; each entry point names the
 * reference code it replaces (paths relative to the reference root):
 *
 *   cytof_elbo_fwdbwd_f32   cytofpy/model/vae.py:17-41   (reparameterised imputation, log_qy)
 *                           cytofpy/model/Model.py:210-278 (loglike: mixtures, Z-gated J->K sum,
 *                               logsumexp over K, noisy-cell class, logistic missingness)
 *                           + the autograd backward of both that fit.py:36 triggers
 *   cytof_impute_emit_f32   cytofpy/model/vae.py:17-33   (materialise imputed y for
 *                               Model.sample_params, Model.py:332-337)
 *   cytof_noise_emit_f32    the Normal.rsample draw of vae.py:33 (test hook for Philox mode)
 *   cytof_adam_step_f64     torch.optim.Adam as used by fit.py:83 + NaN scrub fit.py:12-30,44-45
 *   cytof_sample_minibatch  np.random.choice(N_i, mb, replace=False), fit.py:96-102
 *
 * Conventions (all entry points):
 *   - every data pointer is a DEVICE pointer owned by the caller; the library
 *     never allocates, frees or synchronises; work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - returns 0 on success or a negative CYTOF_E* code (never throws/exits);
 *     cytof_strerror() maps codes to text;
 *   - no global mutable state except a per-device cache of immutable device
 *     attributes; re-entrant across streams/devices with distinct workspaces.
 */
#ifndef CYTOF_B200_H
#define CYTOF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CYTOF_ABI_VERSION 2
/* Compiled limits (the reference has none, Model.py:210-278).  Every shape inside them is correct; which kernel
 * serves it decides the speed (per 1 M cells on one B200, profiles/README.md):
 *   J <= 32, K <= 32                  tensor-core kernel, mma.sync TF32 (cfg1-cfg4: 0.27 ms at J=32, K=30, L=5/5)
 *   J <= 48, K <= 64 otherwise        tensor-core kernel with dZ in tensor memory (cfg5: 0.85 ms at J=40, K=50, L=8/8;
 *                                     J <= 32 with K > 32 runs it with the marker lanes >= J masked: 0.66 ms at J=32, K=50)
 *   48 < J <= 64, L0,L1 <= 5          the same kernel in its 64-marker layout (0.86 ms at J=64, K=64, L=5/5)
 *   48 < J <= 64 with L0 or L1 > 5    the first SIMT kernel: ~3x slower per cell than the tensor-core kernels */
#define CYTOF_MAX_SAMPLES 32 /* I */
#define CYTOF_MAX_J 64       /* markers  */
#define CYTOF_MAX_K 64       /* features */
#define CYTOF_MAX_L 8        /* mixture components per z */

#define CYTOF_OK 0
#define CYTOF_EINVAL (-1)     /* bad descriptor / null pointer */
#define CYTOF_ESHAPE (-2)     /* sizes outside the compiled limits */
#define CYTOF_EWORKSPACE (-3) /* workspace too small */
#define CYTOF_ECUDA (-4)      /* CUDA launch error (see cytof_last_cuda_error) */
#define CYTOF_EALIGN (-5)     /* pointer not aligned as required */

#define CYTOF_NOISE_EXTERNAL 0 /* eps[C,J] supplied (parity mode; only missing entries are read) */
#define CYTOF_NOISE_PHILOX 1   /* in-kernel Philox4x32-10 keyed by (seed, step, sample, row_global[i] + position of the cell in the
                                 * sample's batch, marker): the global row for full batches, the minibatch slot for gathered ones */

/* cytof_desc.flags: the caller's promise that y contains no NaN (the reference computes
 * m = isnan(y) once in fit.py:66; a data set whose mask is all-false may say so).  Lets the
 * library pick kernels without the imputation path; a NaN in y then yields a NaN ELBO. */
#define CYTOF_FLAG_NO_MISSING 1
/* cytof_desc.flags: the per-cell kernel draws the minibatch rows ITSELF (idx must be NULL): cell c of sample i reads the
 * row cytof_sample_minibatch* would have written to idx[c] for (sampler_seed, step, sample i, N = sampler_N[i],
 * m = cell_begin[i+1] - cell_begin[i]); step = *step_dev + 1 when step_dev is set, else sampler_step.  A minibatch step
 * then needs neither a sampler launch nor an index buffer (reference fit.py:96-102 is a host permutation per step). */
#define CYTOF_FLAG_SAMPLE_IN_KERNEL 2

/* One ELBO evaluation ("step") over C = sum_i B_i cells.
 * The step's cell list is sample-major: cells [cell_begin[i], cell_begin[i+1]) belong to sample i.
 * Cell c of sample i reads row  row_base[i] + (idx ? idx[c] : c - cell_begin[i])  of y
 * (CYTOF_FLAG_SAMPLE_IN_KERNEL: row_base[i] + perm_i(c - cell_begin[i]), see the flag). */
typedef struct cytof_desc {
  int32_t I, J, K, L0, L1;
  int32_t model_noisy; /* Model.py:259-266 on/off */
  int32_t noise_mode;  /* CYTOF_NOISE_* */
  int32_t flags;       /* CYTOF_FLAG_* (0 = none) */
  float noisy_sd; /* sqrt(priors['noisy_var']), Model.py:148 */
  float scale;    /* vae.py:41 */
  uint64_t seed;  /* Philox key   */
  uint64_t step;  /* Philox counter word (fit iteration) */
  int64_t cell_begin[CYTOF_MAX_SAMPLES + 1];
  int64_t row_base[CYTOF_MAX_SAMPLES];   /* first row of sample i inside y */
  int64_t row_global[CYTOF_MAX_SAMPLES]; /* global id of that row (rank offset; Philox counter) */
  double fac[CYTOF_MAX_SAMPLES];         /* N_i / B_i with GLOBAL counts, Model.py:257 */
  /* CUDA-graph replay: if non-NULL (int64, device) the Philox counter word is *step_dev + 1,
   * read on the device, and `step` is ignored.  cytof_global_bwd_adam_f64 bumps the counter at
   * the end of a step, so one captured graph advances the noise / sampler / Adam step by itself. */
  const int64_t* step_dev;
  /* CYTOF_FLAG_SAMPLE_IN_KERNEL only (ignored otherwise) */
  uint64_t sampler_seed;
  uint64_t sampler_step;
  int64_t sampler_N[CYTOF_MAX_SAMPLES]; /* rows of sample i the minibatch is drawn from (this rank's shard) */
} cytof_desc;

/* `globals` (float, device): the transformed global sample of this step, packed as
 *   mu0[L0] | mu1[L1] | sig[I] | logeta0[I,J,L0] | logeta1[I,J,L1] | logW[I,K] |
 *   eps[I] | b[I,3] | vm[I,J] | vls[I,J]
 * (mu0 = -cumsum(delta0), mu1 = cumsum(delta1), sig = sqrt(sig2): Model.py:212,223,227;
 *  eps ignored when !model_noisy; b = missingness coefficients b0,b1,b2: Model.py:139-141;
 *  vm, vls = imputation mean / log sd: vae.py:12-13).
 * cytof_globals_layout writes the 10 section offsets (+ total in off[10]). */
int cytof_globals_layout(const cytof_desc* d, int64_t off[11]);

/* `grad_block` (float, device, output): d ll / d(each globals entry) plus d log_qy / d vls:
 *   dmu0[L0] | dmu1[L1] | dsig[I] | dlogeta0[I,J,L0] | dlogeta1[I,J,L1] | dlogW[I,K] |
 *   dZ[J,K] | deps[I] | dvm[I,J] | dvls[I,J] | dlqy_dvls[I,J]
 * cytof_grad_block_layout writes the 11 section offsets (+ total in off[11]). */
int cytof_grad_block_layout(const cytof_desc* d, int64_t off[12]);

size_t cytof_workspace_bytes(const cytof_desc* d);

/* Fused forward + backward of the per-cell part of the ELBO.
 *   y        [rows, J] row-major fp32, NaN = missing (fit.py:66)
 *   idx      [C] int32 sample-local row indices, or NULL for "all rows in order"
 *   eps      [C, J] fp32 standard-normal draws (CYTOF_NOISE_EXTERNAL) or NULL (Philox)
 *   zmask    [J] uint64, bit k of word j = 1[v_k > H_jk]  (thresholded in fp64 by the caller)
 *   ll_lqy   [2] double: ll (Model.py:276 summed over i) and log_qy (vae.py:41 summed over i)
 * Launches: 1 fused kernel + 1 finalise kernel. */
int cytof_elbo_fwdbwd_f32(const cytof_desc* d, const float* y, const int32_t* idx, const float* eps,
                          const float* globals, const uint64_t* zmask, float* grad_block,
                          double* ll_lqy, void* workspace, size_t workspace_bytes, void* stream);

/* Same computation; results as ONE fp64 vector packed[total + 2] = grad block | ll | log_qy
 * (the buffer a multi-GPU run all-reduces and cytof_global_bwd_adam_f64 consumes). */
int cytof_elbo_fwdbwd_packed(const cytof_desc* d, const float* y, const int32_t* idx, const float* eps,
                             const float* globals, const uint64_t* zmask, double* packed,
                             void* workspace, size_t workspace_bytes, void* stream);

/* ---- multi-GPU: the step's only exchange, fused into the finalise kernel -------------------
 * The reference has no communication at all; with cells sharded over the GPUs of one
 * node the packed vector [grad block | ll | log_qy] must be summed over ranks between the per-cell
 * kernel and the backward of the global part.  Instead of a separate NCCL all-reduce the finalise
 * kernel stores its outputs directly into a mailbox of every peer (NVLink / NVSwitch P2P stores
 * through CUDA-IPC mappings), publishes a sequence flag and sums the `world` slots in rank order.
 *   cytof_peer_mailbox_bytes  size of one rank's mailbox
 *   cytof_peer_alloc/free     cudaMalloc'ed, zeroed mailbox (the one exception to "never allocates":
 *                             IPC handles need a whole allocation, not a caching-allocator slice)
 *   cytof_ipc_get_handle      64-byte cudaIpcMemHandle_t of a mailbox, to be sent to the peers
 *   cytof_ipc_open_handle     maps a peer's mailbox (enables peer access lazily); cytof_ipc_close
 *   cytof_elbo_fwdbwd_allreduce   = cytof_elbo_fwdbwd_packed, but `packed` receives the SUM over
 *                             ranks; peer->seq must be 1, 2, 3, ... on every rank in lock step.
 * A wait that exceeds ~2 s (a dead peer) fills `packed` with NaN instead of hanging the GPU, and raises an abort
 * word in EVERY rank's mailbox: from then on all ranks get NaN with +inf in the ll slot, which
 * cytof_global_bwd_adam_f64 reads as "the exchange failed" (no update, out_scalars[4] = 2), so the replicas fail
 * together instead of diverging.  The words are cleared by cytof_peer_alloc. */
#define CYTOF_MAX_PEERS 16
typedef struct cytof_peer_desc {
  int32_t world, rank;
  uint64_t seq;
  void* mailbox[CYTOF_MAX_PEERS]; /* [rank] = own allocation, others = IPC mappings */
  const int64_t* seq_dev;         /* non-NULL: seq = *seq_dev + 1, read on the device (CUDA graphs) */
} cytof_peer_desc;
size_t cytof_peer_mailbox_bytes(const cytof_desc* d, int32_t world);
int cytof_peer_alloc(size_t bytes, void** ptr);
int cytof_peer_free(void* ptr);
int cytof_ipc_get_handle(void* ptr, unsigned char handle[64]);
int cytof_ipc_open_handle(const unsigned char handle[64], void** ptr);
int cytof_ipc_close(void* ptr);
/* Debug: globaltimer (ns) stamps CTA 0 of the last exchange left in the OWN mailbox: kernel entry, own outputs
 * stored, all flags seen, sum done.  Synchronous copy; call after the stream has drained. */
int cytof_peer_debug(const void* mailbox, uint64_t out[4]);
int cytof_elbo_fwdbwd_allreduce(const cytof_desc* d, const float* y, const int32_t* idx, const float* eps,
                                const float* globals, const uint64_t* zmask, double* packed,
                                void* workspace, size_t workspace_bytes, const cytof_peer_desc* peer,
                                void* stream);

/* Writes the imputed minibatch y_out[C, J] (vae.py:33) for the same descriptor/noise. */
int cytof_impute_emit_f32(const cytof_desc* d, const float* y, const int32_t* idx, const float* eps,
                          const float* globals, float* y_out, void* stream);

/* Writes the standard-normal draws eps_out[C, J] that CYTOF_NOISE_PHILOX uses (test hook). */
int cytof_noise_emit_f32(const cytof_desc* d, const int32_t* idx, float* eps_out, void* stream);

/* Posterior cluster labels (post_proc/lam_post.py:6-55): for each of the C cells of y_imp[C, J]
 * (already imputed, no NaN; sample-major as in cell_begin) writes the class probabilities
 * probs_out[C, K+1] (nullable; class 0 = noisy cell, 1..K = features) and one categorical draw
 * lam_out[C] keyed by (seed, step, global row).  Forward half of the likelihood, no gradient. */
int cytof_label_sample_f32(const cytof_desc* d, const float* y_imp, const float* globals,
                           const uint64_t* zmask, int32_t* lam_out, float* probs_out, void* stream);

/* Adam over a flat fp64 buffer.  n_scrub leading entries get NaN gradients replaced by 0
 * (fit.py:24-30 scrubs the 10 global blocks but not the imputation parameters); *flag (int32,
 * device, may be NULL) is set to 1 if anything was scrubbed.  grad_scale multiplies the raw
 * gradient first (fit.py:34: loss = -elbo / Nsum).  step_count is the 1-based Adam step; if
 * step_dev (int64, device) is non-NULL the step is *step_dev + 1 instead and the counter is
 * incremented on the stream, so the call can be captured once in a CUDA graph and replayed. */
int cytof_adam_step_f64(double* param, const double* grad, double* exp_avg, double* exp_avg_sq,
                        int64_t n, int64_t n_scrub, double grad_scale, double lr, double beta1,
                        double beta2, double eps, int64_t step_count, int64_t* step_dev,
                        int32_t* flag, void* stream);

/* Without-replacement minibatch: writes m distinct int32 indices in [0, N) to idx_out — the
 * first m values of a keyed pseudo-random bijection of [0, N) (cycle-walking Feistel network,
 * key = (seed, step, sample)), O(m) work and no scratch.  m >= N writes 0..N-1 (fit.py:98-99). */
int cytof_sample_minibatch(int64_t N, int64_t m, uint64_t seed, uint64_t step, int32_t sample,
                           int32_t* idx_out, void* stream);
/* The draws of ALL samples in one launch: sample i gets min(m[i], N[i]) indices (0..N-1 in order when
 * m[i] >= N[i]) at idx_out + sum_{i' < i} min(m, N); same keyed bijection as cytof_sample_minibatch
 * (identical indices).  step_dev non-NULL: the step is *step_dev + 1, read on the device. */
int cytof_sample_minibatch_multi(int32_t I, const int64_t* N, const int64_t* m, uint64_t seed, uint64_t step,
                                 const int64_t* step_dev, int32_t* idx_out, void* stream);
/* Same draw with the step read on the device (*step_dev + 1): capturable in a CUDA graph. */
int cytof_sample_minibatch_dev(int64_t N, int64_t m, uint64_t seed, const int64_t* step_dev,
                               int32_t sample, int32_t* idx_out, void* stream);

/* ---- the O(3.5k-scalar) global part of a step as two single-CTA fp64 kernels ----------------
 * Replaces: model_param.py:46-92 (sample / transform / log|det J| / entropy of every global
 * block), Model.py:212,223,227,236-243 (sig, mu0, mu1, cumprod(v), Z = 1[v > H] in fp64),
 * Model.py:289-304 (log prior), their autograd backward, and fit.py:32-48 (loss = -elbo/Nsum,
 * NaN scrub of the global blocks, Adam).  Priors are the families default_priors uses
 * (Model.py:92-111): Gamma for delta0/delta1/alpha, LogNormal or Gamma for sig2, Dirichlet for
 * eta0/eta1/W, Beta for eps, Uniform(0,1) for H, Beta(alpha,1) / Beta(alpha/K,1) for v.
 *
 * theta (fp64, device): for each block, in the order delta0 | delta1 | [eps] | sig2 | eta0 | eta1 |
 * W | alpha | v | H, its m[size] then log_s[size]; then per sample mean[J], log_sd[J]. */
typedef struct cytof_global_desc {
  int32_t I, J, K, L0, L1;
  int32_t model_noisy, use_stick_break;
  int32_t noise_mode;   /* CYTOF_NOISE_EXTERNAL: eps_in[R] given; CYTOF_NOISE_PHILOX: (seed, step) */
  int32_t sig2_family;  /* 0 = LogNormal(loc, scale), 1 = Gamma(conc, rate) */
  int32_t reserved0;
  double tau;
  double delta0[2], delta1[2], sig2[2], alpha[2], eps[2]; /* (conc, rate) / (loc, scale) / (a, b) */
  double eta0_conc[CYTOF_MAX_L], eta1_conc[CYTOF_MAX_L], W_conc[CYTOF_MAX_K];
  uint64_t seed, step;
  double grad_scale; /* -1 / Nsum (fit.py:34) */
  double lr, beta1, beta2, adam_eps;
  int64_t adam_step; /* 1-based; ignored when step_dev != NULL */
  const int64_t* step_dev; /* non-NULL: `step` (Philox counter of the global draw) = *step_dev + 1 */
} cytof_global_desc;

/* out[0] = R (global reals = noise length), out[1] = 2R (scrubbed prefix of theta),
 * out[2] = P (length of theta), out[3] = stash length in doubles. */
int cytof_global_sizes(const cytof_global_desc* g, int64_t out[4]);

/* Draw + transform + prior + entropy.  Writes globals (fp32; the b section is left untouched),
 * zmask[J], scalars[2] = {lp, lq_global} and the stash consumed by the backward call. */
int cytof_global_fwd_f64(const cytof_global_desc* g, const double* theta, const double* eps_in,
                         double* stash, float* globals, uint64_t* zmask, double* scalars, void* stream);

/* Backward of the global part from packed (= output of cytof_elbo_fwdbwd_packed, summed over
 * ranks), loss scaling, NaN scrub (sets *flag), Adam update of theta when do_update != 0.
 * grad_out (nullable): d loss / d theta.  out_scalars[5] = elbo, ll, lp, lq, scrubbed (0 / 1; 2 = `packed`
 * carries the failed-exchange mark of cytof_elbo_fwdbwd_allreduce: nothing was updated, elbo .. lq are NaN). */
int cytof_global_bwd_adam_f64(const cytof_global_desc* g, double* theta, double* stash,
                              const double* packed, const double* scalars_in, double* adam_m,
                              double* adam_v, double* grad_out, double* out_scalars,
                              int64_t* step_dev, int32_t* flag, int32_t do_update, void* stream);

const char* cytof_strerror(int code);
const char* cytof_last_cuda_error(void);
int cytof_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CYTOF_B200_H */
