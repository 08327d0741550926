/* pvae_b200.h - C ABI of the B200-native P-VAE training hot path.
 *
 * The reference (hadivafaii/PoissonVAE) is pure Python/PyTorch and has no FFI; the surface this
 * library replaces is a set of Python methods.  Each entry point below names the reference code it
 * replaces (file:line relative to the reference checkout).  INTEGRATION.md shows the ctypes binding
 * a reference maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer (fp32 unless stated) owned by the caller; the library never
 *    allocates or frees caller-visible memory and never synchronises the device;
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it (graph-capturable);
 *  - matrices are row-major and contiguous: (B, K) means element (b, k) at [b * K + k];
 *  - return value: 0 on success, a negative PVAE_ERR_* otherwise; pvae_last_error() returns a
 *    thread-local message for the last failure.
 *
 * RNG contract (tests/test_sampler_gpu.py pins it against oracle/pvae_oracle.py)
 *  Philox4x32-10, key = (lo32(seed) ^ 0x50564145, hi32(seed)),
 *  counter = (latent_quad, event, lo32(off), hi32(off)) where off = offset + (*offset_dev if given),
 *  latent_quad = ((row_offset + b) * K + k) >> 2 and word ((row_offset + b) * K + k) & 3 of the
 *  block is the draw for event `event` of latent (b, k):  u = ((word >> 9) + 0.5) * 2^-23,
 *  e = -logf(u).  The stream depends on the GLOBAL sample index only, so 1/2/4/8-GPU runs that
 *  shard the batch draw identical numbers.  Callers advance `offset` by 1 per sampling call.
 */
#ifndef PVAE_B200_H_
#define PVAE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVAE_OK 0
#define PVAE_ERR_INVALID (-1)     /* bad argument (shape, null pointer, unknown indicator, temp < 0) */
#define PVAE_ERR_UNSUPPORTED (-2) /* valid request this build does not cover */
#define PVAE_ERR_CUDA (-3)        /* a CUDA call failed; see pvae_last_error() */

/* base/distributions.py:294-300 */
#define PVAE_IND_SIGMOID 0
#define PVAE_IND_COSINE 1
#define PVAE_IND_LINEAR 2
#define PVAE_IND_CUBIC 3
#define PVAE_IND_QUINTIC 4

/* exit_logit values.  The chain of a latent stops once the remaining events cannot change the fp32 result:
 * PVAE_EXIT_LOSSLESS stops (a) where the reference's fp32 term is exactly 0 (sigmoid: logit < -88.73,
 * finite-support indicators: logit <= -1) and (b) past t > 1, once every term is below 2^-34 of its running
 * sum - the at most n_exp < 2^9 remaining, monotonically shrinking terms then add up to less than half an ulp
 * of the sum (bit-identical to the full chain for chains finished thread-per-quad, within 1 ulp otherwise;
 * pvae_set_absorb_exit(0) keeps only (a)).  PVAE_EXIT_NEVER runs all n_exp events; c > 0 stops at logit < -c. */
#define PVAE_EXIT_LOSSLESS 0.0f
#define PVAE_EXIT_NEVER (-1.0f)

int pvae_version(void);
/* Programmatic dependent launch for every kernel of the library (default on): a kernel may be scheduled
 * while its predecessor in the stream drains and waits (griddepcontrol.wait) before touching its data. */
int pvae_set_pdl(int on);
/* Tuning knob (process-wide, default on): every kernel asks for the all-shared L1 / shared-memory split, so that an SM never
 * re-partitions its L1 between two kernels of a step; 0 leaves the driver's per-kernel default. */
int pvae_set_carveout(int on);
/* Number of kernels this library has launched in this process so far (all streams, all entry points). */
int64_t pvae_launch_count(void);
const char* pvae_last_error(void);

/* Rows of the (P, K) partial-sum workspaces the backward entry points need for batch size B. */
int64_t pvae_num_partials(int64_t B);

/* ---- fused encoder-output -> relaxed spike counts (+ KL) ------------------------------------
 * Replaces, for one batch: PoissonVAE.infer main/vae.py:393-415 (softclamp_upper x2 or softclamp,
 * base/distributions.py:233-242), Poisson.__init__ rate = exp(.)+eps base/distributions.py:24-25,
 * Poisson.rsample base/distributions.py:55-81 and PoissonVAE.loss_kl main/vae.py:446-450.
 *   h (B,K) = fc_enc(x) WITHOUT bias; b_enc (K) optional bias; log_r (K) prior log-rates.
 *   z (B,K) out.  Optional outputs (NULL to skip): log_dr (B,K), rate (B,K), kl (B,K),
 *   kl_rowsum (B) = kl.sum(1), rate_max (1 float, must be pre-set by the caller, updated with max),
 *   dz_acc (B,K) = sum_n f'(logit_n) t_n, accumulated in the same walk of the chain: handing it to
 *   pvae_sampler_bwd turns the backward into one elementwise pass (12 -> 16 B/latent of L2-resident traffic
 *   instead of a second walk of the chain).
 *   temp > 0: relaxed counts with `indicator`; temp == 0: hard counts #{t_n <= 1} (the reference
 *   draws torch.poisson there, base/distributions.py:56-57,83-85 - same law, different stream). */
int pvae_sampler_fwd(const float* h, const float* log_r, const float* b_enc, float* z, float* log_dr,
                     float* rate, float* kl, float* kl_rowsum, float* rate_max, float* dz_acc, int64_t B, int64_t K,
                     int64_t row_offset, int n_exp, float temp, int indicator, int exc_only,
                     float exit_logit, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                     void* stream);

/* ---- backward of the above: from dz_acc if given, else recomputing the chain from the Philox counters
 * Replaces autograd through base/distributions.py:55-81 and main/vae.py:393-415,446-450
 * (closed form: SURVEY.md section 8 row a13).  Must be called with the same h/log_r/b_enc,
 * shape, n_exp, temp, indicator, exc_only, exit_logit, seed and offset as the forward.
 *   g_z (B,K) upstream gradient wrt z; g_log_dr (B,K) optional upstream gradient wrt the returned
 *   log_dr; dz_acc (B,K) optional, as written by the forward (NULL: recompute the chain); kl_scale = beta / B_global fuses the gradient of beta * mean_b(sum_k kl) (0 to skip).
 *   g_h (B,K) out; g_log_r (K) and g_b_enc (K, optional) out - overwritten, or added to if
 *   accumulate > 0; accumulate < 0: the column partials are LEFT in partial_ws for pvae_adamax_from_partials
 *   (pvae_sampler_bwd_partial_rows(B) rows of K floats for log_r, and the same for b_enc starting at row
 *   pvae_sampler_bwd_bias_offset_rows(B)) and g_log_r / g_b_enc are not written;
 *   partial_ws: workspace of 2 * pvae_num_partials(B) * K floats. */
int pvae_sampler_bwd(const float* h, const float* log_r, const float* b_enc, const float* g_z,
                     const float* g_log_dr, const float* dz_acc, float kl_scale, float* g_h, float* g_log_r, float* g_b_enc,
                     float* partial_ws, int accumulate, int64_t B, int64_t K, int64_t row_offset,
                     int n_exp, float temp, int indicator, int exc_only, float exit_logit,
                     uint64_t seed, uint64_t offset, const uint64_t* offset_dev, void* stream);

/* ---- bare Poisson(log_rate).rsample() / hard counts -----------------------------------------
 * Replaces Poisson.__init__ + rsample base/distributions.py:6-30,55-81 for a caller-supplied
 * log_rate of n_rows x K (PoissonVAE.sample main/vae.py:431-444 and standalone use).
 * clamp < 0 means "no clamp" (reference clamp=None), otherwise softclamp_upper(log_rate, clamp). */
int pvae_rsample_fwd(const float* log_rate, float* z, float* rate, float* dz_acc, int64_t B, int64_t K, int64_t row_offset,
                     int n_exp, float temp, int indicator, float clamp, float exit_logit, uint64_t seed,
                     uint64_t offset, const uint64_t* offset_dev, void* stream);
int pvae_rsample_bwd(const float* log_rate, const float* g_z, const float* dz_acc, float* g_log_rate, int64_t B, int64_t K,
                     int64_t row_offset, int n_exp, float temp, int indicator, float clamp, float exit_logit,
                     uint64_t seed, uint64_t offset, const uint64_t* offset_dev, void* stream);

/* ---- PoissonVAE.loss_kl main/vae.py:446-450 as a standalone elementwise op -------------------
 * kl (B,K) = exp(log_r) * (1 + exp(log_dr) * (log_dr - 1)); backward gives g_log_dr (B,K) and
 * g_log_r (K) from g_kl (B,K). */
int pvae_kl_fwd(const float* log_dr, const float* log_r, float* kl, int64_t B, int64_t K, void* stream);
/* kl_diag of main/train_vae.py:348 (torch.mean(kl, dim=0), the per-latent KL the reference logs and counts active latents
 * from, :442-445) straight from the encoder output h (B,K) the fused step leaves in its workspace: kl_diag (K).  Meant for
 * logging steps.  ws: pvae_kl_diag_ws_floats(B, K) floats. */
int64_t pvae_kl_diag_ws_floats(int64_t B, int64_t K);
int pvae_kl_diag(const float* h, const float* log_r, const float* b_enc, float* kl_diag, float* ws, int64_t B, int64_t K, int exc_only,
                 void* stream);
int pvae_kl_bwd(const float* log_dr, const float* log_r, const float* g_kl, float* g_log_dr, float* g_log_r,
                float* partial_ws, int accumulate, int64_t B, int64_t K, void* stream);

/* ---- dense contractions on the tensor cores -------------------------------------------------
 * C (M,N) = A' * B'^T with A' (M,Kd), B' (N,Kd); fp32 in HBM, TF32 tcgen05.mma, fp32 accumulation in TMEM.
 * precise != 0 ("3xTF32"): every operand is split in shared memory into its TF32-representable part and the
 * exact remainder and hi*hi + lo*hi + hi*lo is accumulated - products good to ~2^-22, i.e. fp32-grade;
 * precise == 0: single pass, operands cut to 10 mantissa bits (relative error up to 2^-10 per product).
 * a_kmajor != 0: A is stored (M,Kd) row-major; a_kmajor == 0: A is stored (Kd,M) row-major (same for B/N),
 * so no transposed copy is ever made.  Replaces F.linear in Linear.forward base/common.py:409-414 as used by
 * fc_enc (main/vae.py:51) and fc_dec (main/vae.py:59-60), and the dgrad / wgrad GEMMs autograd derives:
 *   forward  y = x W^T      : A = x (kmajor), B = W (kmajor)
 *   dgrad    g_x = g_y W    : A = g_y (kmajor), B = W (Kd = out features, stored (Kd,N) -> b_kmajor = 0)
 *   wgrad    g_W = g_y^T x  : A = g_y (stored (Kd = batch, M) -> a_kmajor = 0), B = x (b_kmajor = 0)
 * M, N, Kd must be multiples of 4 and pointers 16-byte aligned.  splits > 1 cuts Kd into that many
 * ranges (one CTA each) whose partial tiles are summed in a fixed order; splitk_ws then needs
 * pvae_gemm_ws_floats(M, N, splits) floats. */
int64_t pvae_gemm_ws_floats(int64_t M, int64_t N, int splits);
/* The 3xTF32 product with the operands' low parts supplied by the caller instead of computed per stage in shared memory:
 * A_lo / B_lo have the shape and layout of A / B and hold x - trunc_tf32(x) (pvae_tf32_lo, or written by the kernel that
 * produced the operand).  The main loop is then pure TMA -> tcgen05.mma; same MMAs on the same bits as precise = 1, hence
 * bit-identical results.  pvae_tf32_lo: lo[i] = x[i] - trunc_tf32(x[i]) (exact), n a multiple of 4. */
int pvae_gemm_tf32_planes(const float* A, const float* A_lo, const float* B, const float* B_lo, float* C, int64_t M, int64_t N, int64_t Kd,
                          int a_kmajor, int b_kmajor, int splits, float* splitk_ws, void* stream);
int pvae_tf32_lo(const float* x, float* lo, int64_t n, void* stream);
/* Tuning knob (process-wide): 0 = automatic tile choice, 1 = 128-wide tiles with one CTA per SM,
 * 2 = 64-wide tiles with two co-resident CTAs per SM. */
int pvae_set_gemm_variant(int v);
int pvae_gemm_tf32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t Kd, int a_kmajor, int b_kmajor,
                   int precise, int splits, float* splitk_ws, void* stream);
/* Decoder GEMM of the training step with BaseVAE.loss_recon (main/vae.py:68-72) and its gradient fused into the
 * epilogue: y = A B^T (A (M,Kd), B (N,Kd), both K-major) is never stored; instead G (M,N) = g_scale * (y - X) and
 * row_part (M, P) with P = pvae_gemm_residual_parts(N): row_part[m, p] = sum over the p-th 32-column half tile of (y - X)^2. */
int64_t pvae_gemm_residual_parts(int64_t N);
int pvae_gemm_tf32_residual(const float* A, const float* B, const float* X, float* G, float* row_part, int64_t M, int64_t N,
                            int64_t Kd, float g_scale, int precise, void* stream);

/* ---- training-step pieces around the sampler and the GEMMs -----------------------------------
 * pvae_recon_residual: BaseVAE.loss_recon main/vae.py:68-72 and its gradient in one pass:
 *   recon (B) = sum_d (y - x)^2,  g_y (B,D) = g_scale * (y - x)  (g_scale = 2 / B_global is
 *   d mean_b(recon) / d y).  g_y or recon may be NULL.  D must be a multiple of 4.
 * pvae_loss_assemble: main/train_vae.py:347-351: out3[0] = sum_b(recon + beta * kl_rowsum) / B_global,
 *   out3[1] = sum recon / B_global, out3[2] = sum kl_rowsum / B_global (shards: sum over ranks); recon is
 *   (B, recon_parts): per-sample squared error, possibly as recon_parts partial sums (1 for pvae_recon_residual).
 * pvae_grad_norm: nn.utils.clip_grad_norm_ main/train_vae.py:365-377 without the host sync:
 *   out2[0] = ||g||_2, out2[1] = min(1, max_norm / (||g|| + 1e-6)) (1 if max_norm <= 0);
 *   ws: pvae_grad_norm_ws_floats() floats.
 * pvae_adamax_step: base/adamax.py:34-100 on a flat buffer of n parameters; `step` >= 1 is the
 *   1-based update count (bias correction 1 - beta1^step); clip_coef = out2 of pvae_grad_norm or NULL. */
int pvae_recon_residual(const float* y, const float* x, float* g_y, float* recon, int64_t B, int64_t D, float g_scale,
                        void* stream);
/* pvae_recon_residual on a decoder output that arrives as n_parts split-K partial sums (n_parts x (B, D), as left in the
 * workspace by pvae_gemm_tf32 with C = NULL), summed in order. */
int pvae_recon_residual_parts(const float* y_parts, int n_parts, const float* x, float* g_y, float* recon, int64_t B, int64_t D,
                              float g_scale, void* stream);
int pvae_loss_assemble(const float* recon, const float* kl_rowsum, float* out3, int64_t B, float beta, int64_t B_global,
                       int recon_parts, void* stream);
int64_t pvae_grad_norm_ws_floats(void);
int pvae_grad_norm(const float* g, int64_t n, float max_norm, float* ws, float* out2, void* stream);
/* Gradient reduction fused with Adamax (single GPU, no clipping): a parameter segment [offset, offset + length) of the flat
 * buffer takes its gradient from `n_partials` partial sums - columnar = 1: rows of `length` floats (column partials of
 * the sampler backward; reduced like its own column-sum kernel), columnar = 0: split-K partials `stride` floats apart
 * (pvae_gemm_tf32 called with C = NULL leaves them in the workspace), partials = NULL: g_out already holds it.  One
 * kernel reduces (bit-identical to the stand-alone reductions), stores the gradient into g_out and updates p / exp_avg /
 * exp_inf as pvae_adamax_step does.  p = NULL: reduction only (data parallel: the gradient exchange applies the update). */
typedef struct pvae_grad_segment {
  int64_t offset, length;
  const float* partials;
  int n_partials;
  int columnar;
  int64_t stride;
  /* columnar segments only, optional (NULL): a second set of n_partials2 row partials whose column sums are added with
   * weight scale2 (the KL part of g_log_r: per-block KL column sums the forward sampler left, times beta / B) */
  const float* partials2;
  int n_partials2;
  float scale2;
} pvae_grad_segment;
int pvae_adamax_from_partials(float* p, float* g_out, float* exp_avg, float* exp_inf, const pvae_grad_segment* segs, int n_segs,
                              float lr, int step, float beta1, float beta2, float eps, float weight_decay, void* stream);
int pvae_gemm_effective_splits(int64_t Kd, int splits);
int64_t pvae_sampler_bwd_partial_rows(int64_t B);
int64_t pvae_sampler_bwd_bias_offset_rows(int64_t B);
/* Adamax with every step-dependent scalar on the device, for CUDA-graph replay of a whole training step: `state` is 4 ints on
 * the device: [optimiser steps taken so far, -, Philox offset (uint64 in ints 2..3)]; lr_sched[i] is the learning rate of
 * optimiser step i + 1 (clamped at n_sched - 1).  A one-thread kernel enqueued right behind the update advances the step
 * counter by 1 and the Philox offset by philox_increment, so the sampler launches of the next replay (which read the offset
 * through their offset_dev argument) draw fresh numbers without any host-side parameter.  rmax_step / rmax_ring / ring_n
 * (optional): the same kernel moves the step's rate.max() from the fixed word the captured sampler max-reduces into to slot
 * (steps taken so far) % ring_n of the epoch's ring and zeroes the word, like the eager path's per-step slot
 * (`r_max.update(dist.rate.max())`, main/train_vae.py:388). */
int pvae_adamax_step_dev(float* p, const float* g, float* exp_avg, float* exp_inf, int64_t n, const float* lr_sched, int n_sched,
                         int* state, float beta1, float beta2, float eps, float weight_decay, const float* clip_coef,
                         uint64_t philox_increment, float* rmax_step, float* rmax_ring, int ring_n, void* stream);
/* Gather n_tensors device arrays into one flat buffer: dst[offsets[i] .. offsets[i+1]) = srcs[i] (zeros where srcs[i] is
 * NULL).  srcs / offsets are HOST arrays (offsets has n_tensors + 1 entries).  One launch per 96 tensors: the per-parameter
 * gradients autograd hands out become the flat buffer the all-reduce / clip / Adamax kernels work on. */
int pvae_gather(const void* const* srcs, const int64_t* offsets, int n_tensors, float* dst, void* stream);
int pvae_adamax_step(float* p, const float* g, float* exp_avg, float* exp_inf, int64_t n, float lr, int step, float beta1,
                     float beta2, float eps, float weight_decay, const float* clip_coef, void* stream);

/* ---- data parallel: gradient all-reduce over NVLink peer memory fused with the Adamax update -------------------
 * One kernel per step replaces NCCL all-reduce + pvae_adamax_step: every rank publishes a step token to its peers'
 * flag arrays, waits for theirs, then sums all ranks' gradient buffers (P2P loads, fixed rank order -> bit-identical
 * sums everywhere) and updates p / exp_avg / exp_inf in the same pass.
 *   peer_grads[r], peer_flags[r]: HOST arrays of `world` DEVICE pointers - rank r's symmetric gradient buffer for this
 *   step's parity (n_ext floats: n gradients + a tail of (n_ext - n) statistics that is only summed into tail_out) and
 *   rank r's flag array (>= world 32-bit words, zero before the first step).  token: the 1-based step count, equal on
 *   all ranks and increasing by 1 per call; gradients must alternate between two buffers by step parity (the kernel has
 *   no trailing barrier).  g_sum_out: optional copy of the summed gradients.  No gradient clipping on this path.
 *   Ranks must sit on different GPUs: the kernels of a step wait on one another. */
int pvae_allreduce_adamax(float* p, float* exp_avg, float* exp_inf, int64_t n, int64_t n_ext, const void* const* peer_grads,
                          void* const* peer_flags, int world, int rank, uint32_t token, float* tail_out, float* g_sum_out,
                          float lr, int step, float beta1, float beta2, float eps, float weight_decay, void* stream);

/* ---- one call per training step of the lin|lin model ------------------------------------------
 * Replaces the forward + loss + backward part of TrainerVAE.iteration main/train_vae.py:342-362 (model forward
 * main/vae.py:384-391, losses :68-72 / :446-450, loss assembly main/train_vae.py:347-351, autograd backward) for
 * this rank's B samples of a global batch of B_global (row_offset = global index of the first local sample).
 *   x (B,D) input patches; params / grads: flat [log_r (K) | W_enc (K,D) | W_dec (D,K) | b_enc (K) if has_bias];
 *   grads is overwritten with d loss / d params where loss = mean over the GLOBAL batch (sum the buffers of all
 *   ranks to get the global gradient); loss3 = this rank's share of [loss, mse, kl] (sum over ranks);
 *   ws: pvae_lin_step_ws_floats(B, K, D) floats of scratch; rate_max: optional running max of the rates;
 *   precise: GEMM mode (see pvae_gemm_tf32); events: NULL or 4 cudaEvent_t (6 with pvae_lin_step's update = 2: entries 4 / 5
 *   bracket the exchange bucket that stays on the critical path) recorded before/after the sampler
 *   forward (0,1) and before/after the sampler backward (2,3) - used by bench.py for the roofline timing. */
int64_t pvae_lin_step_ws_floats(int64_t B, int64_t K, int64_t D);
/* Offsets (in floats from the start of ws) of the step's intermediates, for parity tests that read them back after a step:
 * out7 = [h (B,K), z (B,K), dz_acc (B,K), g_y (B,D), g_z (B,K), g_h (B,K), kl_rowsum (B)]. */
int pvae_lin_step_ws_offsets(int64_t B, int64_t K, int64_t D, int64_t* out7);
/* pvae_lin_step_fwd_bwd followed by the optimiser in the same call: the wgrad split partials and the sampler's column
 * partials are reduced inside ONE Adamax kernel (pvae_adamax_from_partials) instead of three reduction kernels + Adamax.
 * Single GPU (B_global == B), no gradient clipping.  grads receives the gradient as before; results are bit-identical to
 * pvae_lin_step_fwd_bwd + pvae_adamax_step. */
int pvae_lin_step_train(const float* x, float* params, float* grads, float* exp_avg, float* exp_inf, float* ws, float* loss3,
                        float* rate_max, int64_t B, int64_t K, int64_t D, int has_bias, int n_exp, float temp, float beta,
                        int indicator, int exc_only, float exit_logit, int precise, uint64_t seed, uint64_t offset,
                        const uint64_t* offset_dev, void* const* events, float lr, int step, float beta1, float beta2, float eps,
                        float weight_decay, void* stream);
/* Tuning knob (process-wide, default on): run the decoder wgrad and the loss assembly of pvae_lin_step_fwd_bwd on
 * an internal side stream (forked from / joined back into the caller's stream with events). */
int pvae_set_step_overlap(int on);
/* Tuning knob (process-wide, default off): coefficient mode of the step's sampler - the forward also leaves d z / d u, the
 * encoder clamp's derivative and the KL gradient term per latent (+ per-block KL column sums), and the backward becomes a
 * streaming kernel (4 loads, a few FMAs) instead of a second evaluation of the prologue.  Same gradients to ~1e-7.
 * Measured slower at config 2 (+16 MB of working set: 120.9 vs 112 us per step), kept for A/B. */
int pvae_set_step_coef(int on);
/* Debugging aid (process-wide): events = HOST array of n cudaEvent_t (kept alive by the caller) recorded on the step's main
 * stream at: 0 start, 1 after the optional x split, 2 encoder GEMM, 3 sampler forward, 4 decoder GEMM, 5 residual, 6 dgrad,
 * 7 sampler backward, 8 encoder wgrad, 9 gradient reduction + Adamax (single-GPU paths).  Per-kernel live durations, at
 * the price of the programmatic overlap between the kernels.  n = 0 switches it off. */
int pvae_set_step_trace(void* const* events, int n);
/* ---- fused forward after the encoder GEMM: rsample + KL + decoder GEMM + reconstruction residual in ONE kernel ---------
 * (csrc/fused_fwd.cu; BASELINE config 2 "fused rsample+KL+decoder GEMM").  Replaces, in one launch: PoissonVAE.infer +
 * Poisson.rsample (main/vae.py:393-415, base/distributions.py:55-81), loss_kl's per-sample sums (main/vae.py:446-450),
 * decode (main/vae.py:59-60: y = z W_dec^T on tcgen05, z handed to the tensor core through shared memory k-block by
 * k-block) and loss_recon + its gradient (main/vae.py:68-72).
 *   h (B,K) encoder output; w_dec (D,K) and its TF32 low plane w_dec_lo; x (B,D).
 *   outputs: z, z_lo (optional), dz_acc (B,K) as pvae_sampler_fwd; g_y (B,D) = g_scale (y - x) and g_y_lo (optional);
 *   recon (B) = sum_d (y - x)^2; kl_rowsum (B, optional); rate_max (optional running max).
 * Shapes: K a multiple of 256 and <= 1024, D in {128, 256}, B >= 12 rows per SM (pvae_fused_fwd_supported); temp > 0.
 * z / dz_acc of a chain finished by a single lane equal pvae_sampler_fwd's bit for bit, a chain finished by a lane group
 * within a few ulp. */
int pvae_fused_fwd_supported(int64_t B, int64_t K, int64_t D);
int pvae_fused_fwd(const float* h, const float* log_r, const float* b_enc, const float* w_dec, const float* w_dec_lo, const float* x,
                   float* z, float* z_lo, float* dz_acc, float* g_y, float* g_y_lo, float* recon, float* kl_rowsum, float* rate_max,
                   int64_t B, int64_t K, int64_t D, int64_t row_offset, int n_exp, float temp, int indicator, int exc_only,
                   float exit_logit, float g_scale, uint64_t seed, uint64_t offset, const uint64_t* offset_dev, void* stream);
/* Tuning knob (process-wide, default OFF): 1 = the step uses pvae_fused_fwd where it is supported and the operand planes are
 * kept (params_lo given); 0 = sampler, decoder GEMM and residual as three kernels.  Off because it measured slower at config
 * 2 (98 vs 64 us for the three kernels, DESIGN.md section 4.6 has the phase timeline). */
int pvae_set_step_fused(int on);
/* Timing experiments on pvae_fused_fwd (results are WRONG unless 0): 1 = no W_dec loads / MMAs, 2 = hi x hi product only. */
int pvae_set_fused_debug(int mode);
/* Debugging aid: ts_dev = 1024 device words that CTA 0 of pvae_fused_fwd fills with %globaltimer stamps of its phases
 * (tools/fused_timeline.py decodes them); NULL switches it off. */
int pvae_set_fused_timeline(uint64_t* ts_dev);
/* The same for pvae_gemm_tf32 (tools/gemm_timeline.py): tile (0,0,0)'s pipeline stamps + extremes over all CTAs. */
int pvae_set_gemm_timeline(uint64_t* ts_dev);
/* Tuning knob (process-wide, default on): pvae_lin_step_fwd_bwd reduces the wgrad split partials and the sampler's column
 * partials with ONE kernel instead of two split-K reductions + one or two column-sum kernels (bit-identical results). */
int pvae_set_fuse_reduce(int on);
/* Tuning knob (process-wide): K-splits of the decoder GEMM of the training step.  1: one fused GEMM + residual epilogue;
 * 2 (or more): split-K partials + a residual kernel that sums them (halves the GEMM's serial K loop); 0 (default):
 * automatic - two splits when K >= 512 and both halves of every tile fit one wave of CTAs. */
int pvae_set_dec_splits(int n);
int pvae_lin_step_fwd_bwd(const float* x, const float* params, float* grads, float* ws, float* loss3, float* rate_max,
                          int64_t B, int64_t B_global, int64_t row_offset, int64_t K, int64_t D, int has_bias, int n_exp,
                          float temp, float beta, int indicator, int exc_only, float exit_logit, int precise,
                          uint64_t seed, uint64_t offset, const uint64_t* offset_dev, void* const* events, void* stream);
/* The same step driven by a descriptor the caller keeps from step to step (only x, the schedule scalars and the Philox
 * offset change), with the options the positional entry points above do not carry:
 *   params_lo  the parameters' TF32 low plane (flat, same layout as params; pvae_tf32_lo once, then kept up to date by the
 *              optimiser kernels of this call).  Non-NULL switches the step's GEMMs to operand planes (pvae_gemm_tf32_planes):
 *              the sampler, the residual kernel and the sampler backward also write the low planes of z, g_y and g_h.
 *   x_lo       low plane of x (pvae_tf32_lo at upload time), or NULL: computed by the step into ws.
 *   update     0: gradients only, reduced into grads (data parallel over NCCL: the caller all-reduces, then pvae_adamax_step);
 *              1: Adamax fused (pvae_lin_step_train; single GPU: B_global == B);
 *              2: data parallel over NVLink peer memory, exchange + Adamax fused into the step (fields at the end).
 *   state / lr_sched / n_sched / philox_increment / rmax_ring / ring_n (update = 1): step-dependent scalars on the device
 *              for CUDA-graph replay, see pvae_adamax_step_dev; rate_max is then the fixed word moved to the ring. */
typedef struct pvae_lin_step_desc {
  int64_t B, B_global, row_offset, K, D;
  int has_bias, n_exp, indicator, exc_only, precise;
  float exit_logit, temp, beta;
  const float* x;
  const float* x_lo;
  float* params;
  float* params_lo;
  float* grads;
  float* ws;
  float* loss3;
  float* rate_max;
  uint64_t seed, offset;
  const uint64_t* offset_dev;
  void* const* events;
  int update;
  int step;
  float* exp_avg;
  float* exp_inf;
  float lr, beta1, beta2, eps, weight_decay;
  const float* lr_sched;
  int n_sched;
  int ring_n;
  int* state;
  uint64_t philox_increment;
  float* rmax_ring;
  /* update = 2: data parallel, the gradient sum over NVLink peer memory fused with Adamax (see pvae_allreduce_adamax) in
   * two buckets - W_dec's gradient and the loss statistics are exchanged on the side stream under the rest of the
   * backward, only log_r / W_enc / b_enc stay on the critical path.  grads must be this rank's symmetric buffer of this
   * step's parity (n + 4 floats; = peer_grads[rank]); loss3 receives the statistics summed over the ranks; flag arrays
   * need 64 words (bucket b uses words [16 b, 16 b + world), word 63 reports a timed-out wait). */
  int world, rank;
  uint32_t token;
  const void* const* peer_grads;
  void* const* peer_flags;
} pvae_lin_step_desc;
/* sizeof(pvae_lin_step_desc), for bindings that mirror the struct */
int64_t pvae_lin_step_desc_size(void);
int pvae_lin_step(const pvae_lin_step_desc* desc, void* stream);

/* ---- conv encoder / decoder ----------------------------------------------
 * Activations are NHWC fp32.  Replaces Conv2D.forward base/common.py:313-323 (F.conv2d on the weight-normalised
 * filter, `_normalize` :426-431) and the data / weight gradients autograd derives from it, as used by ConvLayer
 * :195-242, FactorizedReduce / StrideReduce :75-126, Cell :145-192, ResConvPool main/vae.py:682-706.
 *
 * pvae_conv_weight_prep: weight (Co, Ci, KH, KW) [torch layout] + lognorm (Co, NULL = no normalisation) ->
 *   w_fwd [co][tap][ci] and (optional) w_dgrad [ci][tap][co], tap = kh * KW + kw.  groups > 1 (= ntaps): the four
 *   1x1 stride-2 filters of FactorizedReduce stored back to back as (Co, Ci); output channel co only sees tap
 *   co / (Co / groups) of a 2x2 window.  Co_pad / Ci_pad >= Co / Ci: channel counts of the packed matrices (the tensor-core
 *   kernels want multiples of 32; the 16-channel decoder layers are zero-padded to 32); bias_pad (Co_pad, optional) receives
 *   the zero-padded bias.  pvae_conv_wgrad's Ci / Co are the padded tensor dimensions, Ci_real / Co_real the parameter's.
 *   with_lo = 1: w_fwd / w_dgrad have TWICE the rows; the second half holds x - trunc_tf32(x) of the first, the low parts of
 *   the 3xTF32 operand split, so that pvae_conv_igemm (w_has_lo = 1) loads them with TMA instead of splitting the filter
 *   tile in shared memory at every K step (the activation tile is still split in the kernel).  Same bits either way.
 * pvae_conv_weight_grad: g_wp (ntaps * Ci, Co) [rows (tap, ci), as written by pvae_conv_wgrad] -> g_weight in the
 *   torch layout and g_lognorm, through the normalisation.
 * pvae_conv_igemm: out[n, oh*out_step+out_off_h, ow*out_step+out_off_w, :] = epilogue(sum_taps in[n, oh*stride+dy,
 *   ow*stride+dx, :] . w_packed[:, widx, :] + bias) over the (OH, OW) grid; `out` is an (NI, OHf, OWf, Co) tensor.
 *   Out-of-range input coordinates read zeros (padding).  Epilogue, in this order, each optional: += add_pre,
 *   *= silu'(dsilu_src), += add_post (all shaped like `out`); out_act receives silu(out).  A forward convolution is
 *   one call; the data gradient of a stride-1 convolution is one call on g_out with w_dgrad and mirrored offsets;
 *   of a stride-2 convolution one call per parity class of the input grid (out_step 2, the class's taps).
 *   wtaps = number of taps in w_packed's row.  Ci, Co multiples of 32.  precise: 3xTF32 (1) or single-pass TF32 (0).
 * pvae_conv_wgrad: g_w[co, (widx, ci)] = sum_{n, oh, ow} in[n, oh*stride+dy, ow*stride+dx, ci] g_out[n, oh, ow, co] on the
 *   tensor cores (split over pixel blocks, partials in ws, summed in split order), then pushed through the weight
 *   normalisation: g_weight (torch layout of `weight`), g_lognorm (Co), and g_bias (Co) = sum over pixels of g_out (an
 *   all-ones operand slab in the same GEMM).  ws: pvae_conv_wgrad_ws_floats(Ci, Co, ntaps) floats. */
int pvae_conv_weight_prep(const float* weight, const float* lognorm, const float* bias, float* w_fwd, float* w_dgrad,
                          float* bias_pad, int Co, int Ci, int Co_pad, int Ci_pad, int ntaps, int groups, int with_lo,
                          void* stream);
int pvae_conv_weight_grad(const float* g_wp, const float* weight, const float* lognorm, float* g_weight, float* g_lognorm,
                          int Co, int Ci, int ntaps, int groups, void* stream);
int pvae_conv_igemm(const float* in, const float* w_packed, float* out, float* out_act, const float* bias,
                    const float* add_pre, const float* dsilu_src, const float* add_post, int NI, int IH, int IW, int Ci,
                    int Co, int OH, int OW, int OHf, int OWf, int out_step, int out_off_h, int out_off_w, int stride,
                    int ntaps, const int* tap_dy, const int* tap_dx, const int* tap_widx, int wtaps, int w_has_lo,
                    int precise, void* stream);
int64_t pvae_conv_wgrad_ws_floats(int Ci, int Co, int ntaps);
int pvae_conv_wgrad(const float* in, const float* g_out, const float* weight, const float* lognorm, float* g_weight,
                    float* g_lognorm, float* g_bias, float* ws, int NI, int IH, int IW, int Ci, int Co, int OH, int OW,
                    int stride, int ntaps, const int* tap_dy, const int* tap_dx, const int* tap_widx, int groups,
                    int Ci_real, int Co_real, int precise, void* stream);

/* Memory-bound pieces around the convolutions (csrc/conv_misc.cu).  All reductions are two fixed-order stages;
 * `ws` arguments need pvae_reduce_ws_floats(number of outputs) floats.
 *   pvae_stem_fwd / pvae_stem_wgrad: 3x3 weight-normalised convolution of the single-channel input, main/vae.py:211-218
 *     (pad 1 or 0); g_wpb (10, Co): rows 0..8 = tap-major gradient of the normalised filter (feed to
 *     pvae_conv_weight_grad with Ci = 1), row 9 = bias gradient.
 *   pvae_se_fwd / pvae_se_bwd: SELayer base/common.py:129-142 fused with the cell's residual, Cell.forward :187-192:
 *     y = skip + eps_res * c * sigmoid(W2 relu(W1 mean_hw(c) + b1) + b2); saves pool (N,C), hid (N,hd), gate (N,C);
 *     backward returns g_c and the per-image pre-activation gradients d2 (N,C), d1 (N,hd) (g_skip = g_y).
 *   pvae_avgpool_fwd, pvae_pool_silu_bwd: ResConvPool main/vae.py:682-706 skip path and its backward fused with silu'.
 *   pvae_bias_act, pvae_relu_bwd: bias + ReLU + dropout (Philox mask) of ResDenseLayer main/vae.py:657-679.
 *   pvae_layernorm_fwd / bwd: nn.LayerNorm over the last dimension of a + b. */
int64_t pvae_reduce_ws_floats(int64_t n_out);
int pvae_silu_fwd(const float* x, float* out, int64_t n, void* stream);
int pvae_add(const float* a, const float* b, float* out, int64_t n, void* stream);
int pvae_stem_fwd(const float* x, const float* weight, const float* lognorm, const float* bias, float* out, float* out_act,
                  int N, int IH, int IW, int Co, int pad, void* stream);
int pvae_stem_wgrad(const float* x, const float* g_out, float* g_wpb, float* ws, int N, int IH, int IW, int Co, int pad,
                    void* stream);
int pvae_colsum(const float* x, float* out, float* ws, int64_t R, int C, void* stream);
int pvae_small_tn(const float* a, const float* b, float* out, float* ws, int64_t R, int Ca, int Cb, void* stream);
int pvae_se_fwd(const float* c, const float* skip, const float* W1, const float* b1, const float* W2, const float* b2,
                float eps_res, float* y, float* y_act, float* pool, float* hid, float* gate, int N, int H, int W, int C,
                int C_real, int hd, int skip_up, void* stream);
int pvae_se_bwd(const float* g_y, const float* c, const float* gate, const float* hid, const float* W1, const float* W2,
                float eps_res, float* g_c, float* d2, float* d1, int N, int HW, int C, int C_real, int hd, void* stream);
/* Parameter gradients of the squeeze-excitation layer from the per-image vectors pvae_se_fwd / pvae_se_bwd produced, in one
 * launch (partial sums per block of images, combined in block order by the block that finishes last):
 * g_w2 (C_real, hd) = d2^T hid, g_b2 = colsum d2, g_w1 (hd, C_real) = d1^T pool, g_b1 = colsum d1.
 * ws: pvae_reduce_ws_floats(2 C_real hd + C_real + hd) floats.  Launches of this kernel on one device must be
 * stream-ordered (it keeps one ticket counter). */
int pvae_se_param_grads(const float* d2, const float* hid, const float* d1, const float* pool, float* g_w2, float* g_b2,
                        float* g_w1, float* g_b1, float* ws, int N, int C_real, int hd, void* stream);
/* Decoder pieces (main/vae.py:760-789, base/common.py:35-47,211-217): nearest-neighbour resampling with ATen's index
 * map (scale_factor 2 of the 'up' cells, size=14 for MNIST) and its gather backward (optionally times silu'(src));
 * (N, C, P) <-> (N, P, C) permutation of fc_dec's output (+ bias, + silu copy); the final 1x1 convolution to one channel
 * (g_wb: C + 1 floats, gradient of the normalised filter then of the bias).  C = channels in memory, C_real = the
 * parameters' (16-channel layers are stored zero-padded to 32); pool / gate / d2 are (N, C_real). */
int pvae_upsample_fwd(const float* x, float* out, int N, int IH, int IW, int OH, int OW, int C, void* stream);
int pvae_upsample_bwd(const float* g_out, const float* dsilu_src, float* g_in, int N, int IH, int IW, int OH, int OW, int C,
                      void* stream);
int pvae_permute_cp(const float* in, const float* bias, float* out, float* out_act, int64_t N, int C, int P, int to_nhwc,
                    void* stream);
int pvae_head_fwd(const float* x, const float* weight, const float* lognorm, const float* bias, float* y, int64_t R, int C,
                  int C_real, void* stream);
int pvae_head_bwd(const float* x, const float* g_y, const float* weight, const float* lognorm, float* g_x, float* g_wb, float* ws,
                  int64_t R, int C, int C_real, void* stream);
int pvae_avgpool_fwd(const float* x, float* pool, int N, int HW, int C, void* stream);
int pvae_pool_silu_bwd(const float* g_a, const float* x, const float* g_pool, float* g_x, int N, int HW, int C, void* stream);
int pvae_bias_act(float* y, const float* bias, const float* add, int64_t R, int C, int relu, float drop_p, uint64_t seed,
                  uint64_t offset, void* stream);
int pvae_relu_bwd(float* g, const float* y_saved, int64_t n, float drop_p, void* stream);
int pvae_layernorm_fwd(const float* a, const float* b, const float* gamma, const float* beta, float* out, float* xhat,
                       float* rstd, int64_t R, int C, float eps, void* stream);
int pvae_layernorm_bwd(const float* g, const float* xhat, const float* rstd, const float* gamma, float* g_s,
                       float* g_gamma_beta, float* ws, int64_t R, int C, void* stream);

/* ---- one call per residual cell (csrc/conv_net.cu) -----------------------------------------------------------------
 * Cell.forward base/common.py:187-192 (ConvLayer :233-242, SELayer :138-142, FactorizedReduce / StrideReduce :75-126, the
 * up-sampling skip :35-47) and its whole backward, every kernel enqueued from C++.  x / y are NHWC maps with channel
 * counts padded to multiples of 32; x_act = silu(x) if the producer already has it (NULL: computed here); y_act = silu(y).
 * `saved` (pvae_cell_saved_floats(c, x_act == NULL) floats) carries the forward's intermediates to the backward; `scratch`
 * (pvae_cell_scratch_floats) is only used inside pvae_cell_bwd, which writes the g_* gradients of the descriptor and g_x.
 * skip_w / skip_ln / skip_b: PVAE_CELL_DOWN_FACTORIZED: the four 1x1 filters back to back as (Co, Ci), (Co), (Co);
 * PVAE_CELL_DOWN_STRIDE / PVAE_CELL_UP_DEC: the 1x1 filter (Co, Ci, 1, 1). */
enum { PVAE_CELL_NORMAL_PRE = 0, PVAE_CELL_DOWN_FACTORIZED = 1, PVAE_CELL_DOWN_STRIDE = 2, PVAE_CELL_NORMAL_DEC = 3, PVAE_CELL_UP_DEC = 4 };
typedef struct pvae_cell {
  int kind;
  int N, H, W; /* input map */
  int Ci, Co;  /* channel counts of the parameters */
  float eps_res;
  int precise;
  const float *w1, *ln1, *b1; /* first 3x3 convolution (the only one of a decoder cell) */
  const float *w2, *ln2, *b2; /* second 3x3 convolution (encoder cells) */
  const float *se_w1, *se_b1, *se_w2, *se_b2;
  const float *skip_w, *skip_ln, *skip_b;
  float *g_w1, *g_ln1, *g_b1, *g_w2, *g_ln2, *g_b2, *g_se_w1, *g_se_b1, *g_se_w2, *g_se_b2, *g_skip_w, *g_skip_ln, *g_skip_b;
} pvae_cell;
int64_t pvae_cell_saved_floats(const pvae_cell* c, int own_act);
int64_t pvae_cell_scratch_floats(const pvae_cell* c);
int pvae_cell_fwd(const pvae_cell* c, const float* x, const float* x_act, float* y, float* y_act, float* saved, void* stream);
int pvae_cell_bwd(const pvae_cell* c, const float* x, const float* x_act, const float* g_y, float* g_x, float* saved,
                  float* scratch, void* stream);
/* Tuning knob (process-wide, default on): pvae_cell_bwd runs the parameter-gradient kernels on an internal side stream
 * (forked from / joined back into the caller's stream with events) next to the data-gradient chain. */
int pvae_set_cell_overlap(int on);
/* Tuning knob (A/B measurements): pipeline depth / CTAs per SM of the narrow convolution tiles.  0 = default. */
int pvae_set_conv_variant(int v);

/* Tuning knob (process-wide): number of events each quad of latents walks thread-per-quad before a
 * still-unfinished chain is handed to a whole warp (32 events per step).  Default 8.  Results do
 * not depend on the launch geometry, but the float summation order of a chain depends on this value. */
int pvae_set_phase_a_events(int n);
int pvae_set_absorb_exit(int on);
/* Tuning knob (process-wide): forward sampler schedule.  1: every lane of a warp takes the next quad of the block's rows x
 * 128 pool the moment its own chain ends (work-efficient under heavy-tailed chain lengths); 0: one row at a time in lock
 * step - bit-identical results; 2: one resident wave of blocks, every block pooling all quads of its rows
 * (sampler_pool_kernel; 37.8 vs 41.7 us at config 2).  A chain that exceeds the phase-A cap is finished by a lane group:
 * when the quad's three output slots exist (z, z_lo, dz_acc: the training forward with operand planes) its state is parked
 * there and the chain continues exactly as in schedules 0 / 1 - bit-identical; otherwise the lane group walks it again from
 * its first event, which may move the last bits (~1e-6 relative).  3 (default): schedule 2 where it is bit-identical to 1,
 * schedule 1 elsewhere. */
int pvae_set_pool(int on);
/* Tuning knob (A/B): grid of the one-wave schedule as a multiple of the resident wave (default 1). */
int pvae_set_wave_factor(float f);
/* Tuning knob (process-wide): rows (samples) per thread block of the sampler kernels, 1..4; 0 = automatic.
 * Changes pvae_num_partials(); results of the forward are unaffected. */
int pvae_set_rows_per_block(int n);

/* ---- parity harness only: dump the uniforms / exponentials the kernels consume ---------------
 * U, E: (n_exp, B, K).  Either may be NULL.
 * pvae_debug_log_check adds to *mismatches_dev the number of the 2^23 possible uniforms for which the
 * kernels' -log(u) differs by a bit from libdevice's -logf(u) (expected 0). */
int pvae_debug_log_check(uint64_t* mismatches_dev, void* stream);
int pvae_debug_draws(float* U, float* E, int64_t B, int64_t K, int64_t row_offset, int n_exp, uint64_t seed,
                     uint64_t offset, const uint64_t* offset_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PVAE_B200_H_ */
