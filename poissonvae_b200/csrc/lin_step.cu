// One C call per training step of the lin|lin P-VAE: enqueues, back to back on one stream, every kernel of
// forward + loss + backward (closed form) for this rank's slice of the batch.
//
//   h   = x W_enc^T                                   tcgen05 GEMM            main/vae.py:44-51
//   z, dz_acc = rsample(softclamps(h, log_r)), kl sums fused sampler           main/vae.py:393-415, base/distributions.py:55-81
//   y   = z W_dec^T; g_y = 2 (y - x) / B_global,      tcgen05 GEMM with the residual fused into its epilogue
//   recon partial sums                                                        main/vae.py:59-60, 68-72
//   loss = sum_b(recon + beta kl) / B_global          loss kernel             main/train_vae.py:347-351
//   g_W_dec = g_y^T z (split-K), g_z = g_y W_dec      tcgen05 GEMMs
//   g_h, g_log_r (+ g_b_enc)                          sampler backward from the dz_acc the forward stored (elementwise)
//   g_W_enc = g_h^T x (split-K)                       tcgen05 GEMM
// The caller then all-reduces the flat gradient buffer (data parallel) and calls pvae_grad_norm /
// pvae_adamax_step.  Parameter and gradient buffers are flat: [log_r (K) | W_enc (K,D) | W_dec (D,K) | b_enc (K)].
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pvae_b200.h"
#include "pvae_host.h"

namespace {

// Side stream for the work that is off the critical path (decoder wgrad, loss assembly): created once per device.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr, dgrad = nullptr;
};

SideStream* side_stream() {
  static SideStream per_device[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& s = per_device[dev];
  if (!s.stream) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.dgrad, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &s;
}

void* const* g_trace = nullptr;  // pvae_set_step_trace: CUDA events recorded between the main-stream launches of a step
int g_trace_n = 0;
int g_overlap = 1;
int g_fused = 0;  // pvae_set_step_fused: sampler + decoder GEMM + residual in one kernel where supported (measured slower: off)
int g_coef = 0;  // pvae_set_step_coef: the forward sampler leaves the backward's coefficients (streaming backward); measured slower
int g_dec_splits = 0;   // K-splits of the decoder GEMM (pvae_set_dec_splits); 0 = automatic
int g_fuse_reduce = 1;  // pvae_lin_step_fwd_bwd: reduce all gradient partials with one kernel (bit-identical to the three)

struct Ws {
  float *h, *z, *dz, *y, *g_y, *g_z, *g_h, *recon, *klrow, *partials, *splitk, *splitk2, *ydec;
  float *z_lo, *g_y_lo, *g_h_lo, *x_lo;  // TF32 low planes of the GEMM operands the step produces
  float *c_ddr, *c_ck, *klcol;           // the forward sampler's coefficients for the streaming backward (dz holds A)
  int64_t total;
};

int wgrad_splits(int64_t m, int64_t n, int64_t kd) {
  const int64_t tiles = ((m + 127) / 128) * ((n + 127) / 128);
  if (tiles >= 64 || kd < 512) return 1;
  int64_t s = 128 / tiles;
  if (s > kd / 256) s = kd / 256;
  return static_cast<int>(s < 1 ? 1 : s);
}

Ws carve(float* base, int64_t B, int64_t K, int64_t D) {
  Ws w;
  int64_t o = 0;
  auto take = [&](int64_t n) {
    float* p = base ? base + o : nullptr;
    o += (n + 3) & ~int64_t(3);  // keep every buffer 16-byte aligned
    return p;
  };
  w.h = take(B * K), w.z = take(B * K), w.dz = take(B * K), w.y = take(B * D), w.g_y = take(B * D);
  w.recon = take(B * pvae_gemm_residual_parts(D)), w.klrow = take(B);
  w.partials = take(2 * pvae_num_partials(B) * K);
  const int64_t s1 = pvae_gemm_ws_floats(D, K, wgrad_splits(D, K, B)), s2 = pvae_gemm_ws_floats(K, D, wgrad_splits(K, D, B));
  w.splitk = take(s1);   // decoder wgrad (may run on the side stream, concurrently with ...)
  w.splitk2 = take(s2);  // ... the encoder wgrad
  w.ydec = take(std::max<int64_t>(8 * B * D, B * K));  // split-K partials of the decoder output (pvae_set_dec_splits, at most 8)
  // The step's working set is what decides whether its intermediates stay in the L2 (126 MB): buffers whose lives do not
  // overlap share storage.  g_z (dgrad output) reuses the decoder partials the residual kernel has consumed; the sampler
  // backward writes g_h over the g_z it has just read and g_h's low plane over dz_acc (same index, same thread).
  w.g_z = w.ydec, w.g_h = w.g_z, w.g_h_lo = w.dz;
  w.z_lo = take(B * K), w.g_y_lo = take(B * D), w.x_lo = take(B * D);
  w.c_ddr = take(B * K), w.c_ck = take(B * K), w.klcol = take(pvae_num_partials(B) * K);
  w.total = o;
  return w;
}

}  // namespace

extern "C" int pvae_set_step_overlap(int on) {
  g_overlap = on != 0;
  return PVAE_OK;
}

extern "C" int64_t pvae_lin_step_ws_floats(int64_t B, int64_t K, int64_t D) {
  if (B <= 0 || K <= 0 || D <= 0) return 0;
  return carve(nullptr, B, K, D).total;
}

// where the step's intermediates sit in the workspace (floats from its start): parity tests read z, h, ... back
extern "C" int pvae_lin_step_ws_offsets(int64_t B, int64_t K, int64_t D, int64_t* out7) {
  if (B <= 0 || K <= 0 || D <= 0 || !out7) return pvae::fail(PVAE_ERR_INVALID, "pvae_lin_step_ws_offsets: bad argument");
  float* const base = reinterpret_cast<float*>(uintptr_t(1) << 20);  // any non-null base: only differences are used
  const Ws w = carve(base, B, K, D);
  const float* f[7] = {w.h, w.z, w.dz, w.g_y, w.g_z, w.g_h, w.klrow};
  for (int i = 0; i < 7; ++i) out7[i] = f[i] - base;
  return PVAE_OK;
}

// What the optimiser part of a step does with the gradient partials.
static int lin_step_core(const pvae_lin_step_desc& d, void* stream) {
  using pvae::fail;
  const int64_t B = d.B, B_global = d.B_global, K = d.K, D = d.D;
  if (B <= 0 || B_global < B || K <= 0 || D <= 0) return fail(PVAE_ERR_INVALID, "pvae_lin_step: bad shape");
  if (B % 4 || K % 4 || D % 4)
    return fail(PVAE_ERR_UNSUPPORTED, "pvae_lin_step: B=%lld K=%lld D=%lld must be multiples of 4", (long long)B, (long long)K, (long long)D);
  if (!d.x || !d.params || !d.grads || !d.ws || !d.loss3) return fail(PVAE_ERR_INVALID, "pvae_lin_step: null pointer");
  if (!(d.temp > 0.f)) return fail(PVAE_ERR_INVALID, "pvae_lin_step: training needs temp > 0, got %g", d.temp);
  if (d.update < 0 || d.update > 2) return fail(PVAE_ERR_INVALID, "pvae_lin_step: update mode %d", d.update);
  const bool dp_peer = d.update == 2;
  if (dp_peer && (!d.exp_avg || !d.exp_inf || d.step < 1 || !d.peer_grads || !d.peer_flags || d.world < 2 || d.rank < 0 || d.rank >= d.world ||
                  d.peer_grads[d.rank] != d.grads || d.state))
    return fail(PVAE_ERR_INVALID, "pvae_lin_step: bad exchange argument (grads must be this rank's symmetric buffer)");
  if (d.update == 1 && (!d.exp_avg || !d.exp_inf || (d.step < 1 && !d.state) || B_global != B))
    return fail(PVAE_ERR_INVALID, "pvae_lin_step: bad optimiser argument (the fused update is single-GPU: B_global == B)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const Ws w = carve(d.ws, B, K, D);
  const int has_bias = d.has_bias, precise = d.precise;
  const float *log_r = d.params, *w_enc = d.params + K, *w_dec = w_enc + K * D;
  const float* b_enc = has_bias ? w_dec + D * K : nullptr;
  float *g_log_r = d.grads, *g_w_enc = d.grads + K, *g_w_dec = g_w_enc + K * D;
  float* g_b_enc = has_bias ? g_w_dec + D * K : nullptr;
  // Operand planes: when the caller keeps the parameters' TF32 low plane (params_lo, maintained by the optimiser kernels),
  // every producer of a GEMM operand also writes that operand's low plane (z, g_y, g_h; x once per batch), and the GEMMs
  // load hi and lo tiles by TMA instead of splitting each stage in shared memory.  Same MMAs on the same bits.
  const bool planes = precise && d.params_lo != nullptr;
  const float *w_enc_lo = planes ? d.params_lo + K : nullptr, *w_dec_lo = planes ? d.params_lo + K + K * D : nullptr;
  float *z_lo = planes ? w.z_lo : nullptr, *g_y_lo = planes ? w.g_y_lo : nullptr, *g_h_lo = planes ? w.g_h_lo : nullptr;
  const bool coef = g_coef != 0;
  const float* x_lo = planes ? d.x_lo : nullptr;
  auto mark = [&](int i) {
    if (d.events && d.events[i]) cudaEventRecord(static_cast<cudaEvent_t>(d.events[i]), s);
  };
  int trace_i = 0;
  auto trace = [&]() {  // debugging aid: per-kernel live durations (defeats the programmatic overlap between the kernels)
    if (g_trace && trace_i < g_trace_n) cudaEventRecord(static_cast<cudaEvent_t>(g_trace[trace_i]), s);
    ++trace_i;
  };
  trace();
  auto gemm = [&](const float* A, const float* A_lo, const float* Bm, const float* B_lo, float* C, int64_t M, int64_t N, int64_t Kd, int ak,
                  int bk, int splits, float* sws, cudaStream_t st) {
    return pvae::gemm_impl(A, planes ? A_lo : nullptr, Bm, planes ? B_lo : nullptr, C, M, N, Kd, ak, bk, precise, splits, sws, nullptr, nullptr,
                           nullptr, nullptr, 0.f, 0, st);
  };
  int rc;
  // ---- forward
  if (planes && x_lo == nullptr) {  // the caller did not keep x's low plane (pvae_tf32_lo at upload time): make it now
    if ((rc = pvae_tf32_lo(d.x, w.x_lo, B * D, s))) return rc;
    x_lo = w.x_lo;
  }
  trace();  // 1: after the optional x split
  if ((rc = gemm(d.x, x_lo, w_enc, w_enc_lo, w.h, B, K, D, 1, 1, 1, nullptr, s))) return rc;
  trace();  // 2: encoder GEMM
  mark(0);
  const bool fused = g_fused && planes && !coef && pvae::fused_fwd_supported(B, K, D);
  int recon_parts = static_cast<int>(pvae_gemm_residual_parts(D));
  const float g_scale = 2.0f / static_cast<float>(B_global);
  if (fused) {
    // one kernel: rsample + KL row sums + decoder GEMM (z through shared memory) + residual; y is complete in TMEM
    if ((rc = pvae::fused_fwd_impl(w.h, log_r, b_enc, w_dec, w_dec_lo, d.x, w.z, z_lo, w.dz, w.g_y, g_y_lo, w.recon, w.klrow, d.rate_max, B, K,
                                   D, d.row_offset, d.n_exp, d.temp, d.indicator, d.exc_only, d.exit_logit, g_scale, d.seed, d.offset,
                                   d.offset_dev, s)))
      return rc;
    mark(1);
    trace();  // 3: fused forward
    trace();  // 4, 5: (no separate decoder GEMM / residual)
    trace();
    recon_parts = 1;
  } else {
  if (coef) {
    // coefficient mode: the forward leaves d z / d u, the clamp derivative and the KL gradient term per latent and the
    // KL's column sums per block; the backward is then a streaming kernel instead of a second evaluation of the prologue
    if ((rc = pvae::sampler_fwd_coef_impl(w.h, log_r, b_enc, w.z, z_lo, w.klrow, d.rate_max, w.dz, w.c_ddr, w.c_ck, w.klcol, B, K,
                                          d.row_offset, d.n_exp, d.temp, d.indicator, d.exc_only, d.exit_logit, d.seed, d.offset,
                                          d.offset_dev, s)))
      return rc;
  } else if ((rc = pvae::sampler_fwd_impl(w.h, log_r, b_enc, w.z, z_lo, nullptr, nullptr, nullptr, w.klrow, d.rate_max, w.dz, B, K,
                                          d.row_offset, d.n_exp, d.temp, d.indicator, d.exc_only, d.exit_logit, d.seed, d.offset,
                                          d.offset_dev, s))) {
    return rc;
  }
  mark(1);
  trace();  // 3: sampler forward
  // decoder GEMM with the residual fused into its epilogue: g_y and per-tile squared errors, y itself is never stored
  // automatic: the decoder's K loop (K / 32 blocks) is the longest serial stretch of any GEMM of the step; when two
  // K-halves of every 128 x 128 tile still fit one wave of CTAs, split it (measured at config 2: 115.0 -> 111.6 us per
  // step; four splits: 116.9)
  int want = g_dec_splits;
  if (want == 0) want = (K >= 512 && ((B + 127) / 128) * ((D + 127) / 128) * 2 <= 160) ? 2 : 1;
  const int dsp = want > 1 ? pvae_gemm_effective_splits(K, want) : 1;
  if (dsp > 1) {  // split-K partials (in the not-yet-used wgrad workspace) + one kernel that sums them into the residual
    if ((rc = gemm(w.z, z_lo, w_dec, w_dec_lo, nullptr, B, D, K, 1, 1, dsp, w.ydec, s))) return rc;
    trace();  // 4: decoder GEMM
    if ((rc = pvae::recon_parts_impl(w.ydec, dsp, d.x, w.g_y, g_y_lo, w.recon, B, D, g_scale, s))) return rc;
    trace();  // 5: residual
    recon_parts = 1;
  } else if ((rc = pvae::gemm_impl(w.z, z_lo, w_dec, w_dec_lo, nullptr, B, D, K, 1, 1, precise, 1, nullptr, d.x, w.g_y, g_y_lo, w.recon, g_scale,
                                   64, s))) {
    return rc;
  }
  }
  // ---- backward.  Off the critical path (decoder wgrad, loss assembly) goes to a side stream so that it fills
  // the SMs the dgrad GEMM leaves idle and overlaps the elementwise sampler backward.
  SideStream* side = g_overlap ? side_stream() : nullptr;
  cudaStream_t s2 = side ? side->stream : s;
  if (side) {
    if ((rc = pvae::cuda_ok("fork", cudaEventRecord(side->fork, s)))) return rc;
    if ((rc = pvae::cuda_ok("fork", cudaStreamWaitEvent(s2, side->fork, 0)))) return rc;
  }
  const int64_t n = K + 2 * K * D + (has_bias ? K : 0);
  // data parallel over peer memory: this rank's share of [loss, mse, kl] rides behind its gradients (floats n .. n + 3 of
  // the symmetric buffer) and is summed over the ranks by the first exchange bucket into loss3
  if ((rc = pvae_loss_assemble(w.recon, w.klrow, dp_peer ? d.grads + n : d.loss3, B, d.beta, B_global, recon_parts, s2))) return rc;
  // fused update: the split partials stay in the workspace (C = NULL) and are reduced inside the Adamax kernel
  const int sp_dec = wgrad_splits(D, K, B), sp_enc = wgrad_splits(K, D, B);
  const bool fuse = (d.update >= 1 || g_fuse_reduce) && sp_dec > 1 && sp_enc > 1;
  if ((rc = gemm(w.g_y, g_y_lo, w.z, z_lo, fuse ? nullptr : g_w_dec, D, K, B, 0, 0, sp_dec, w.splitk, s2))) return rc;
  const float kl_scale = d.beta / static_cast<float>(B_global);
  const int prow = static_cast<int>(coef ? pvae::sampler_bwd_coef_rows(B) : pvae_sampler_bwd_partial_rows(B));
  const int64_t bias_row = coef ? prow : pvae_sampler_bwd_bias_offset_rows(B);
  pvae_grad_segment seg[4] = {};
  seg[0] = pvae_grad_segment{0, K, w.partials, prow, 1, K, nullptr, 0, 0.f};
  if (coef)  // g_log_r = sum_b g_u (the backward's column partials) + beta / B_global * sum_b kl (the forward's column sums)
    seg[0].partials2 = w.klcol, seg[0].n_partials2 = static_cast<int>(pvae_num_partials(B)), seg[0].scale2 = kl_scale;
  seg[1] = pvae_grad_segment{K, K * D, w.splitk2, pvae_gemm_effective_splits(B, sp_enc), 0, K * D, nullptr, 0, 0.f};
  seg[2] = pvae_grad_segment{K + K * D, D * K, w.splitk, pvae_gemm_effective_splits(B, sp_dec), 0, D * K, nullptr, 0, 0.f};
  if (has_bias) seg[3] = pvae_grad_segment{K + 2 * K * D, K, w.partials + bias_row * K, prow, 1, K, nullptr, 0, 0.f};
  pvae::AdamaxDev dev;
  dev.lr_sched = d.lr_sched, dev.n_sched = d.n_sched, dev.state = d.state;
  // the data gradient first: it READS W_dec, which exchange bucket 0 below updates in place
  if ((rc = gemm(w.g_y, g_y_lo, w_dec, w_dec_lo, w.g_z, B, K, D, 1, 0, 1, nullptr, s))) return rc;
  if (dp_peer) {
    // Exchange bucket 0 = W_dec's gradient (+ the loss statistics), complete as soon as the decoder wgrad is: its sum over
    // the ranks and its Adamax update run on the side stream UNDER the sampler backward and the encoder wgrad of the main
    // stream - but only after the dgrad GEMM, the last reader of W_dec, has finished (an event; without it the update raced
    // with that GEMM: 2-rank parameters 5e-3 off the single-GPU ones).  Only bucket 1 (log_r, W_enc, b_enc) stays on the
    // critical path.
    if (fuse && (rc = pvae::adamax_from_partials_impl(nullptr, nullptr, d.grads, nullptr, nullptr, seg + 2, 1, 0.f, 1, 0.f, 0.f, 0.f, 0.f,
                                                      pvae::AdamaxDev{}, s2)))
      return rc;
    if (side) {  // only the exchange (it writes W_dec) waits for the dgrad GEMM; the reduction of the wgrad partials does not
      if ((rc = pvae::cuda_ok("dgrad event", cudaEventRecord(side->dgrad, s)))) return rc;
      if ((rc = pvae::cuda_ok("dgrad event", cudaStreamWaitEvent(s2, side->dgrad, 0)))) return rc;
    }
    const int64_t rs[1] = {K + K * D}, rl[1] = {D * K};
    if ((rc = pvae::peer_exchange_impl(d.params, planes ? d.params_lo : nullptr, d.exp_avg, d.exp_inf, rs, rl, 1, n, d.peer_grads, d.peer_flags,
                                       0, d.world, d.rank, d.token, d.loss3, nullptr, d.lr, d.step, d.beta1, d.beta2, d.eps, d.weight_decay,
                                       s2)))
      return rc;
  }
  if (side && (rc = pvae::cuda_ok("join", cudaEventRecord(side->join, s2)))) return rc;
  trace();  // 6: dgrad
  mark(2);
  if (coef) {
    if ((rc = pvae::sampler_bwd_coef_impl(w.g_z, w.dz, w.c_ddr, w.c_ck, kl_scale, w.g_h, g_h_lo, w.partials, has_bias, w.klcol,
                                          pvae_num_partials(B), fuse ? nullptr : g_log_r, g_b_enc, B, K, s)))
      return rc;
  } else if ((rc = pvae::sampler_bwd_impl(w.h, log_r, b_enc, w.g_z, nullptr, w.dz, kl_scale, w.g_h, g_h_lo, g_log_r, g_b_enc, w.partials,
                                          fuse ? -1 : 0, B, K, d.row_offset, d.n_exp, d.temp, d.indicator, d.exc_only, d.exit_logit, d.seed,
                                          d.offset, d.offset_dev, s))) {
    return rc;
  }
  mark(3);
  trace();  // 7: sampler backward
  if ((rc = gemm(w.g_h, g_h_lo, d.x, x_lo, fuse ? nullptr : g_w_enc, K, D, B, 0, 0, sp_enc, w.splitk2, s))) return rc;
  trace();  // 8: encoder wgrad
  if (dp_peer) {  // bucket 1: everything but W_dec
    if (fuse) {
      pvae_grad_segment sb[3] = {seg[0], seg[1], seg[3]};
      if ((rc = pvae::adamax_from_partials_impl(nullptr, nullptr, d.grads, nullptr, nullptr, sb, has_bias ? 3 : 2, 0.f, 1, 0.f, 0.f, 0.f, 0.f,
                                                pvae::AdamaxDev{}, s)))
        return rc;
    }
    const int64_t rs[2] = {0, K + 2 * K * D}, rl[2] = {K + K * D, K};
    mark(4);  // update == 2: the descriptor's `events` has six entries; 4 / 5 bracket the exchange bucket on the critical path
    if ((rc = pvae::peer_exchange_impl(d.params, planes ? d.params_lo : nullptr, d.exp_avg, d.exp_inf, rs, rl, has_bias ? 2 : 1, -1,
                                       d.peer_grads, d.peer_flags, 1, d.world, d.rank, d.token, nullptr, nullptr, d.lr, d.step, d.beta1,
                                       d.beta2, d.eps, d.weight_decay, s)))
      return rc;
    mark(5);
    if (side && (rc = pvae::cuda_ok("join", cudaStreamWaitEvent(s, side->join, 0)))) return rc;
    return PVAE_OK;
  }
  if (side && (rc = pvae::cuda_ok("join", cudaStreamWaitEvent(s, side->join, 0)))) return rc;
  if (d.update == 0 && !fuse) return PVAE_OK;
  if (!fuse) {  // small shapes: the wgrad GEMMs did not split, the gradients are already complete in grads
    if (d.state) {
      if ((rc = pvae::adamax_step_dev_impl(d.params, planes ? d.params_lo : nullptr, d.grads, d.exp_avg, d.exp_inf, n, dev, d.beta1, d.beta2,
                                           d.eps, d.weight_decay, nullptr, s)))
        return rc;
      return pvae::advance_state(d.state, d.philox_increment, d.rate_max, d.rmax_ring, d.ring_n, s);
    }
    return pvae::adamax_step_impl(d.params, planes ? d.params_lo : nullptr, d.grads, d.exp_avg, d.exp_inf, n, d.lr, d.step, d.beta1, d.beta2,
                                  d.eps, d.weight_decay, nullptr, s);
  }
  const bool upd = d.update == 1;
  if ((rc = pvae::adamax_from_partials_impl(upd ? d.params : nullptr, upd && planes ? d.params_lo : nullptr, d.grads, d.exp_avg, d.exp_inf, seg,
                                            has_bias ? 4 : 3, d.lr, d.step, d.beta1, d.beta2, d.eps, d.weight_decay, dev, s)))
    return rc;
  trace();  // 9: join + gradient reduction / Adamax
  if (upd && d.state)  // graph replay: advance the device-side step counter / Philox offset, move rate.max() to its slot
    return pvae::advance_state(d.state, d.philox_increment, d.rate_max, d.rmax_ring, d.ring_n, s);
  return PVAE_OK;
}

extern "C" int pvae_set_step_fused(int on) {
  g_fused = on != 0;
  return PVAE_OK;
}

extern "C" int pvae_set_step_coef(int on) {
  g_coef = on != 0;
  return PVAE_OK;
}

extern "C" int pvae_set_step_trace(void* const* events, int n) {
  g_trace = n > 0 ? events : nullptr;
  g_trace_n = n > 0 ? n : 0;
  return PVAE_OK;
}

extern "C" int64_t pvae_lin_step_desc_size(void) { return static_cast<int64_t>(sizeof(pvae_lin_step_desc)); }

extern "C" int pvae_lin_step(const pvae_lin_step_desc* d, void* stream) {
  if (!d) return pvae::fail(PVAE_ERR_INVALID, "pvae_lin_step: null descriptor");
  return lin_step_core(*d, stream);
}

static pvae_lin_step_desc make_desc(const float* x, float* params, float* grads, float* ws, float* loss3, float* rate_max, int64_t B,
                                    int64_t B_global, int64_t row_offset, int64_t K, int64_t D, int has_bias, int n_exp, float temp,
                                    float beta, int indicator, int exc_only, float exit_logit, int precise, uint64_t seed,
                                    uint64_t offset, const uint64_t* offset_dev, void* const* events) {
  pvae_lin_step_desc d = {};
  d.B = B, d.B_global = B_global, d.row_offset = row_offset, d.K = K, d.D = D;
  d.has_bias = has_bias, d.n_exp = n_exp, d.indicator = indicator, d.exc_only = exc_only, d.precise = precise;
  d.exit_logit = exit_logit, d.temp = temp, d.beta = beta;
  d.x = x, d.params = params, d.grads = grads, d.ws = ws, d.loss3 = loss3, d.rate_max = rate_max;
  d.seed = seed, d.offset = offset, d.offset_dev = offset_dev, d.events = events;
  return d;
}

extern "C" int pvae_lin_step_fwd_bwd(const float* x, const float* params, float* grads, float* ws, float* loss3,
                                     float* rate_max, int64_t B, int64_t B_global, int64_t row_offset, int64_t K, int64_t D,
                                     int has_bias, int n_exp, float temp, float beta, int indicator, int exc_only,
                                     float exit_logit, int precise, uint64_t seed, uint64_t offset,
                                     const uint64_t* offset_dev, void* const* events, void* stream) {
  const pvae_lin_step_desc d = make_desc(x, const_cast<float*>(params), grads, ws, loss3, rate_max, B, B_global, row_offset, K, D, has_bias,
                                         n_exp, temp, beta, indicator, exc_only, exit_logit, precise, seed, offset, offset_dev, events);
  return lin_step_core(d, stream);
}

extern "C" int pvae_set_dec_splits(int n) {
  if (n < 0 || n > 8) return pvae::fail(PVAE_ERR_INVALID, "pvae_set_dec_splits: %d not in [0, 8]", n);
  g_dec_splits = n;
  return PVAE_OK;
}

extern "C" int pvae_set_fuse_reduce(int on) {
  g_fuse_reduce = on != 0;
  return PVAE_OK;
}

extern "C" int pvae_lin_step_train(const float* x, float* params, float* grads, float* exp_avg, float* exp_inf, float* ws,
                                   float* loss3, float* rate_max, int64_t B, int64_t K, int64_t D, int has_bias, int n_exp,
                                   float temp, float beta, int indicator, int exc_only, float exit_logit, int precise, uint64_t seed,
                                   uint64_t offset, const uint64_t* offset_dev, void* const* events, float lr, int step, float beta1,
                                   float beta2, float eps, float weight_decay, void* stream) {
  if (!params || !exp_avg || !exp_inf || step < 1) return pvae::fail(PVAE_ERR_INVALID, "pvae_lin_step_train: bad optimiser argument");
  pvae_lin_step_desc d = make_desc(x, params, grads, ws, loss3, rate_max, B, B, 0, K, D, has_bias, n_exp, temp, beta, indicator, exc_only,
                                   exit_logit, precise, seed, offset, offset_dev, events);
  d.update = 1, d.exp_avg = exp_avg, d.exp_inf = exp_inf;
  d.lr = lr, d.beta1 = beta1, d.beta2 = beta2, d.eps = eps, d.weight_decay = weight_decay, d.step = step;
  return lin_step_core(d, stream);
}
