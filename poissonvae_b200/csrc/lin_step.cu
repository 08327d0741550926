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
  cudaEvent_t fork = nullptr, join = nullptr;
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
  }
  return &s;
}

int g_overlap = 1;
int g_dec_splits = 0;   // K-splits of the decoder GEMM (pvae_set_dec_splits); 0 = automatic
int g_fuse_reduce = 1;  // pvae_lin_step_fwd_bwd: reduce all gradient partials with one kernel (bit-identical to the three)

struct Ws {
  float *h, *z, *dz, *y, *g_y, *g_z, *g_h, *recon, *klrow, *partials, *splitk, *splitk2, *ydec;
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
  w.h = take(B * K), w.z = take(B * K), w.dz = take(B * K), w.y = take(B * D), w.g_y = take(B * D), w.g_z = take(B * K), w.g_h = take(B * K);
  w.recon = take(B * pvae_gemm_residual_parts(D)), w.klrow = take(B);
  w.partials = take(2 * pvae_num_partials(B) * K);
  const int64_t s1 = pvae_gemm_ws_floats(D, K, wgrad_splits(D, K, B)), s2 = pvae_gemm_ws_floats(K, D, wgrad_splits(K, D, B));
  w.splitk = take(s1);   // decoder wgrad (may run on the side stream, concurrently with ...)
  w.splitk2 = take(s2);  // ... the encoder wgrad
  w.ydec = take(8 * B * D);  // split-K partials of the decoder output (pvae_set_dec_splits, at most 8)
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

namespace {
struct Update {  // optimiser fused into the step (pvae_lin_step_train); p == nullptr: gradients only
  bool fuse_reduce = false;  // gradients only, but all partials reduced by ONE kernel instead of three
  float *p = nullptr, *exp_avg = nullptr, *exp_inf = nullptr;
  float lr = 0.f, beta1 = 0.f, beta2 = 0.f, eps = 0.f, weight_decay = 0.f;
  int step = 0;
};
}  // namespace

static int lin_step_impl(const float* x, const float* params, float* grads, float* ws, float* loss3, float* rate_max, int64_t B,
                         int64_t B_global, int64_t row_offset, int64_t K, int64_t D, int has_bias, int n_exp, float temp,
                         float beta, int indicator, int exc_only, float exit_logit, int precise, uint64_t seed, uint64_t offset,
                         const uint64_t* offset_dev, void* const* events, void* stream, const Update& up) {
  using pvae::fail;
  if (B <= 0 || B_global < B || K <= 0 || D <= 0) return fail(PVAE_ERR_INVALID, "pvae_lin_step_fwd_bwd: bad shape");
  if (B % 4 || K % 4 || D % 4)
    return fail(PVAE_ERR_UNSUPPORTED, "pvae_lin_step_fwd_bwd: B=%lld K=%lld D=%lld must be multiples of 4", (long long)B,
                (long long)K, (long long)D);
  if (!x || !params || !grads || !ws || !loss3) return fail(PVAE_ERR_INVALID, "pvae_lin_step_fwd_bwd: null pointer");
  if (!(temp > 0.f)) return fail(PVAE_ERR_INVALID, "pvae_lin_step_fwd_bwd: training needs temp > 0, got %g", temp);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const Ws w = carve(ws, B, K, D);
  const float *log_r = params, *w_enc = params + K, *w_dec = w_enc + K * D;
  const float* b_enc = has_bias ? w_dec + D * K : nullptr;
  float *g_log_r = grads, *g_w_enc = grads + K, *g_w_dec = g_w_enc + K * D;
  float* g_b_enc = has_bias ? g_w_dec + D * K : nullptr;
  auto mark = [&](int i) {
    if (events && events[i]) cudaEventRecord(static_cast<cudaEvent_t>(events[i]), s);
  };
  int rc;
  // ---- forward
  if ((rc = pvae_gemm_tf32(x, w_enc, w.h, B, K, D, 1, 1, precise, 1, nullptr, s))) return rc;
  mark(0);
  if ((rc = pvae_sampler_fwd(w.h, log_r, b_enc, w.z, nullptr, nullptr, nullptr, w.klrow, rate_max, w.dz, B, K, row_offset, n_exp,
                             temp, indicator, exc_only, exit_logit, seed, offset, offset_dev, s)))
    return rc;
  mark(1);
  // decoder GEMM with the residual fused into its epilogue: g_y and per-tile squared errors, y itself is never stored
  int recon_parts = static_cast<int>(pvae_gemm_residual_parts(D));
  // automatic: the decoder's K loop (K / 32 blocks of ~1 us each in 3xTF32 mode) is the longest serial stretch of any GEMM
  // of the step; when two K-halves of every 128 x 128 tile still fit one wave of CTAs, split it (measured at config 2:
  // 115.0 -> 111.6 us per step; four splits: 116.9)
  int want = g_dec_splits;
  if (want == 0) want = (K >= 512 && ((B + 127) / 128) * ((D + 127) / 128) * 2 <= 160) ? 2 : 1;
  const int dsp = want > 1 ? pvae_gemm_effective_splits(K, want) : 1;
  if (dsp > 1) {  // split-K partials (in the not-yet-used wgrad workspace) + one kernel that sums them into the residual
    if ((rc = pvae_gemm_tf32(w.z, w_dec, nullptr, B, D, K, 1, 1, precise, dsp, w.ydec, s))) return rc;
    if ((rc = pvae_recon_residual_parts(w.ydec, dsp, x, w.g_y, w.recon, B, D, 2.0f / static_cast<float>(B_global), s))) return rc;
    recon_parts = 1;
  } else if ((rc = pvae_gemm_tf32_residual(w.z, w_dec, x, w.g_y, w.recon, B, D, K, 2.0f / static_cast<float>(B_global), precise, s))) {
    return rc;
  }
  // ---- backward.  Off the critical path (decoder wgrad, loss assembly) goes to a side stream so that it fills
  // the SMs the dgrad GEMM leaves idle and overlaps the elementwise sampler backward.
  SideStream* side = g_overlap ? side_stream() : nullptr;
  cudaStream_t s2 = side ? side->stream : s;
  if (side) {
    if ((rc = pvae::cuda_ok("fork", cudaEventRecord(side->fork, s)))) return rc;
    if ((rc = pvae::cuda_ok("fork", cudaStreamWaitEvent(s2, side->fork, 0)))) return rc;
  }
  if ((rc = pvae_loss_assemble(w.recon, w.klrow, loss3, B, beta, B_global, recon_parts, s2))) return rc;
  // fused update: the split partials stay in the workspace (C = NULL) and are reduced inside the Adamax kernel
  const int sp_dec = wgrad_splits(D, K, B), sp_enc = wgrad_splits(K, D, B);
  const bool fuse = (up.p != nullptr || up.fuse_reduce) && sp_dec > 1 && sp_enc > 1;
  if ((rc = pvae_gemm_tf32(w.g_y, w.z, fuse ? nullptr : g_w_dec, D, K, B, 0, 0, precise, sp_dec, w.splitk, s2))) return rc;
  if (side && (rc = pvae::cuda_ok("join", cudaEventRecord(side->join, s2)))) return rc;
  if ((rc = pvae_gemm_tf32(w.g_y, w_dec, w.g_z, B, K, D, 1, 0, precise, 1, nullptr, s))) return rc;
  mark(2);
  if ((rc = pvae_sampler_bwd(w.h, log_r, b_enc, w.g_z, nullptr, w.dz, beta / static_cast<float>(B_global), w.g_h, g_log_r, g_b_enc,
                             w.partials, fuse ? -1 : 0, B, K, row_offset, n_exp, temp, indicator, exc_only, exit_logit, seed, offset,
                             offset_dev, s)))
    return rc;
  mark(3);
  if ((rc = pvae_gemm_tf32(w.g_h, x, fuse ? nullptr : g_w_enc, K, D, B, 0, 0, precise, sp_enc, w.splitk2, s))) return rc;
  if (side && (rc = pvae::cuda_ok("join", cudaStreamWaitEvent(s, side->join, 0)))) return rc;
  if (up.p == nullptr && !fuse) return PVAE_OK;
  const int64_t n = K + 2 * K * D + (has_bias ? K : 0);
  if (!fuse)
    return pvae_adamax_step(up.p, grads, up.exp_avg, up.exp_inf, n, up.lr, up.step, up.beta1, up.beta2, up.eps, up.weight_decay, nullptr, s);
  const int prow = static_cast<int>(pvae_sampler_bwd_partial_rows(B));
  pvae_grad_segment seg[4];
  seg[0] = pvae_grad_segment{0, K, w.partials, prow, 1, K};
  seg[1] = pvae_grad_segment{K, K * D, w.splitk2, pvae_gemm_effective_splits(B, sp_enc), 0, K * D};
  seg[2] = pvae_grad_segment{K + K * D, D * K, w.splitk, pvae_gemm_effective_splits(B, sp_dec), 0, D * K};
  if (has_bias) seg[3] = pvae_grad_segment{K + 2 * K * D, K, w.partials + pvae_sampler_bwd_bias_offset_rows(B) * K, prow, 1, K};
  return pvae_adamax_from_partials(up.p, grads, up.exp_avg, up.exp_inf, seg, has_bias ? 4 : 3, up.lr, up.step, up.beta1, up.beta2, up.eps,
                                   up.weight_decay, s);
}

extern "C" int pvae_lin_step_fwd_bwd(const float* x, const float* params, float* grads, float* ws, float* loss3,
                                     float* rate_max, int64_t B, int64_t B_global, int64_t row_offset, int64_t K, int64_t D,
                                     int has_bias, int n_exp, float temp, float beta, int indicator, int exc_only,
                                     float exit_logit, int precise, uint64_t seed, uint64_t offset,
                                     const uint64_t* offset_dev, void* const* events, void* stream) {
  Update up;
  up.fuse_reduce = g_fuse_reduce != 0;
  return lin_step_impl(x, params, grads, ws, loss3, rate_max, B, B_global, row_offset, K, D, has_bias, n_exp, temp, beta, indicator,
                       exc_only, exit_logit, precise, seed, offset, offset_dev, events, stream, up);
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
  Update up;
  up.p = params, up.exp_avg = exp_avg, up.exp_inf = exp_inf;
  up.lr = lr, up.beta1 = beta1, up.beta2 = beta2, up.eps = eps, up.weight_decay = weight_decay, up.step = step;
  return lin_step_impl(x, params, grads, ws, loss3, rate_max, B, B, 0, K, D, has_bias, n_exp, temp, beta, indicator, exc_only,
                       exit_logit, precise, seed, offset, offset_dev, events, stream, up);
}
