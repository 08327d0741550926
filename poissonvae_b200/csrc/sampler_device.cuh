// Device-side pieces of the Poisson sampler shared by the stand-alone sampler kernels (sampler.cu) and the fused
// sampler + decoder kernel (fused_fwd.cu): launch parameters, one soft-threshold term, the division by the rate, the
// prologue (softclamps, rate, clamp derivatives).  Reference semantics: base/distributions.py:5-92, main/vae.py:393-415.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pvae_device.cuh"

namespace pvae {

// kFwdSoft: z only.  kFwdBoth: z and dz_acc = sum_n f'(l_n) t_n in one walk of the chain (training forward).
// kBwd: backward that recomputes the chain from the Philox counters.  kBwdSaved: backward from a stored dz_acc.
enum Mode : int { kFwdSoft = 0, kFwdHard = 1, kBwd = 2, kFwdBoth = 3, kBwdSaved = 4 };

struct SamplerParams {
  const float* h;         // (B,K) encoder output, or log_rate in RAW mode
  const float* log_r;     // (K)
  const float* b_enc;     // (K) or null
  const float* g_z;       // (B,K) bwd
  const float* g_log_dr;  // (B,K) bwd, optional
  const float* dz_in;     // (B,K) kBwdSaved: sum_n f'(l_n) t_n written by the forward
  float* dz_out;          // (B,K) kFwdBoth
  float* z;
  float* z_lo;            // (B,K) optional: z - trunc_tf32(z), the low plane the decoder / wgrad GEMMs load by TMA
  float* g_h_lo;          // (B,K) optional, backward: the same for g_h (encoder wgrad)
  // Coefficient mode of the training forward (kFwdBoth, all three non-null): the forward leaves in dz_out, instead of the
  // raw sum_n f'(l_n) t_n, the finished factor A = d z / d u (that sum times exp(lambda) / (temp rate) times the rate clamp's
  // derivative), in c_ddr the encoder clamp's derivative d delta / d h and in c_ck the KL term e^{log_r} delta e^{delta};
  // klcol receives per-block column sums of the KL (P,K).  The backward is then four loads and a handful of FMAs per latent
  // (sampler_bwd_coef_kernel) instead of a second evaluation of the prologue (~125 instructions, 16 of them MUFU).
  float* c_ddr;
  float* c_ck;
  float* klcol;
  const float* c_a;       // backward (coef kernel): A
  float* log_dr;
  float* rate;
  float* kl;
  float* kl_rowsum;
  float* rate_max;
  float* g_h;
  float* part_log_r;  // (P,K)
  float* part_b;      // (P,K) or null
  long long B, K, row_offset;
  int rows_per_block;  // rows per tile (<= kMaxRows)
  int rows_block;      // rows per block (a multiple of tiles)
  int n_exp;
  int phase_a;        // events walked thread-per-quad before a quad is parked for phase B
  int pool;           // forward modes: lanes take the next quad of the warp's pool as soon as theirs ends
  float inv_temp;     // 1 / temp
  float sig_a;        // log2(e) / temp: sigmoid((1 - t) / temp) = 1 / (1 + 2^(sig_a * (t - 1)))
  float t_exit;       // leave a chain once every arrival time exceeds this (+inf: never)
  float absorb;       // leave a chain once every term < absorb * its running sum (0: never)
  float clamp_dr, clamp_rate;  // RAW: clamp_rate < 0 means no clamp
  float kl_scale;
  int exc_only;
  uint64_t offset;
  const uint64_t* offset_dev;
  PhiloxKey ks;
};

// store a quad of z / g_h and, when the GEMMs want it, its TF32 low plane (x - trunc_tf32(x), exact)
__device__ __forceinline__ void stg4_planes(float* hi, float* lo, long long off, const float4 v) {
  stg4(hi + off, v);
  if (lo != nullptr)
    stg4(lo + off, make_float4(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u), v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u),
                               v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u), v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u)));
}

// ---- one soft-threshold term -------------------------------------------------------------------
template <int IND, bool GRAD>
__device__ __forceinline__ void soft_term(const SamplerParams& p, float t, float& f, float& df) {
  if (IND == kSigmoid) {
    const float w = ex2_approx(fmaf(t, p.sig_a, -p.sig_a));  // exp(-(1 - t) / temp)
    f = rcp_approx(1.f + w);
    if (GRAD) df = fmaf(-f, f, f);
  } else {
    indicator_eval<IND, GRAD>(fmaf(-t, p.inv_temp, p.inv_temp), f, df);
  }
}

// x = e / rate.  Soft path: Markstein's correction on a Newton-refined reciprocal (equals div.rn except in
// rare last-bit cases, 3 FMA-pipe ops; both phases use the same reciprocal, so a chain that moves from
// phase A to phase B keeps dividing the same way).  Hard path: IEEE division, bit-exact by definition.
__device__ __forceinline__ float recip_for_div(float rate) {
  const float r = rcp_approx(rate);
  return fmaf(r, fmaf(-rate, r, 1.f), r);
}

template <bool EXACT>
__device__ __forceinline__ float div_by_rate(float e, float rate, float inv_rate) {
  if (EXACT) return __fdiv_rn(e, rate);
  const float q = e * inv_rate;
  return fmaf(fmaf(-q, rate, e), inv_rate, q);
}

// ---- phase B: a group of kGroup lanes finishes one parked quad, kGroup events per step -----------
// 32 / kGroup quads are walked by a warp at once.  Lane gl of a group draws event n0 + gl; the arrival
// times come from a prefix sum over the group's lanes (shfl.up).  A group that finishes its quad takes
// the next one from the block's queue, so long and short chains share a warp without waiting for each
// other.  Every lane runs every step (the shuffles need them); idle groups contribute exact zeros.
constexpr int kGroup = 8;

// ---- prologue: encoder output -> log_dr, u, rate (and the clamp derivatives for the backward) ---
struct Latent {
  float log_dr;  // delta
  float explam;  // exp(lambda)
  float rate;    // exp(lambda) + eps
  float d_rate;  // d lambda / d u            (softclamp_upper'(u, 5))
  float d_dr;    // d delta / d (h + bias)    (softclamp'(h, 10) / softclamp_upper'(h, 10))
};

template <bool RAW, bool GRAD>
__device__ __forceinline__ Latent prologue(const SamplerParams& p, float hv, float lr, float bias) {
  Latent q;
  float lam;
  if (RAW) {
    q.log_dr = 0.f;
    q.d_dr = 1.f;
    if (p.clamp_rate >= 0.f) {
      const Softplus s = softplus20_fast<GRAD>(p.clamp_rate - hv);
      lam = p.clamp_rate - s.value;
      q.d_rate = s.grad;
    } else {
      lam = hv;
      q.d_rate = 1.f;
    }
  } else {
    const float hb = hv + bias;
    if (p.exc_only) {  // softclamp(h, 10, 0) = softplus(h) - softplus(h - 10), base/distributions.py:233-234
      const Softplus a = softplus20_fast<GRAD>(hb), b = softplus20_fast<GRAD>(hb - p.clamp_dr);
      q.log_dr = a.value - b.value;
      q.d_dr = a.grad - b.grad;
    } else {  // softclamp_upper(h, 10) = 10 - softplus(10 - h), base/distributions.py:241-242
      const Softplus a = softplus20_fast<GRAD>(p.clamp_dr - hb);
      q.log_dr = p.clamp_dr - a.value;
      q.d_dr = a.grad;
    }
    const Softplus c = softplus20_fast<GRAD>(p.clamp_rate - (lr + q.log_dr));  // main/vae.py:409
    lam = p.clamp_rate - c.value;
    q.d_rate = c.grad;
  }
  q.explam = expf(lam);
  q.rate = q.explam + kF32Eps;  // base/distributions.py:24-25
  return q;
}

// d z / d u = (sum_n f'(l_n) t_n) * coef_factor: 1 / temp, exp(lambda) / rate (= 1 - eps / rate), and the rate clamp's derivative
__device__ __forceinline__ float coef_factor(const SamplerParams& p, const Latent& q) {
  return p.inv_temp * (q.explam * recip_for_div(q.rate)) * q.d_rate;
}


// Host: chain-control fields of SamplerParams (temperature constants, early-exit thresholds, phase split) - sampler.cu
void sampler_set_chain(SamplerParams& p, int n_exp, float temp, int indicator, float exit_logit);

}  // namespace pvae
