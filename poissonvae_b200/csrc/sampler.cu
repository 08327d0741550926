// Fused Poisson reparameterised sampler + KL kernels for sm_100a.
//
// One thread owns a quad of four consecutive latents of one sample (128-bit loads of the encoder
// output, 128-bit stores of the counts) and walks their arrival chains in lock step: one
// Philox4x32-10 block per event feeds the four chains, the running arrival times stay in
// registers and the (n_exp, B, K) tensors of the reference never exist.  A block owns a tile of
// rows and all columns of those rows, so per-sample KL sums and per-latent (batch) gradient sums
// are produced without atomics, in a fixed order.
//
// Reference semantics: base/distributions.py:5-92 (Poisson), main/vae.py:393-415 (infer),
// main/vae.py:446-450 (loss_kl); closed-form backward SURVEY.md section 8 row a13.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pvae_b200.h"
#include "pvae_device.cuh"
#include "pvae_host.h"

namespace pvae {

constexpr int kThreads = 128;          // threads per block
constexpr int kColsPerPass = 4 * kThreads;  // latents of one row a block covers per pass
constexpr int kMaxRowsPerBlock = 8;

enum Mode : int { kFwdSoft = 0, kFwdHard = 1, kBwd = 2 };

struct SamplerParams {
  // inputs
  const float* h;      // (B,K) encoder output, or log_rate in RAW mode
  const float* log_r;  // (K)
  const float* b_enc;  // (K) or null
  const float* g_z;    // (B,K) bwd
  const float* g_log_dr;  // (B,K) bwd, optional
  // outputs
  float* z;
  float* log_dr;
  float* rate;
  float* kl;
  float* kl_rowsum;
  float* rate_max;
  float* g_h;
  float* part_log_r;  // (P,K)
  float* part_b;      // (P,K) or null
  long long B, K, row_offset;
  int rows_per_block;
  int n_exp;
  float temp, inv_temp, exit_logit;  // exit_logit < 0: never leave early
  float clamp_dr, clamp_rate;        // RAW: clamp_rate < 0 means no clamp
  float kl_scale;
  int exc_only;
  uint64_t offset;
  const uint64_t* offset_dev;
  PhiloxKey ks;
};

struct Quad {
  float v[4];
};

// ---- the arrival chain of four latents ---------------------------------------------------------
// Returns, per latent, acc_f = sum_n f(logit_n) (relaxed count) or the hard count, and if GRAD
// acc_g = sum_n f'(logit_n) * t_n.
template <int MODE, int IND>
__device__ __forceinline__ void run_chain(const SamplerParams& p, uint32_t quad, uint32_t off_lo, uint32_t off_hi,
                                          const float (&rate)[4], float (&acc_f)[4], float (&acc_g)[4]) {
  float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) acc_f[j] = acc_g[j] = 0.f;
  const float inv_temp = p.inv_temp;
  const float exit_logit = p.exit_logit;
  const bool may_exit = exit_logit >= 0.f;
  for (int n = 0; n < p.n_exp; ++n) {
    const uint4 r = philox4x32_10(quad, static_cast<uint32_t>(n), off_lo, off_hi, p.ks);
    const uint32_t bits[4] = {r.x, r.y, r.z, r.w};
    float lmax = -3.0e38f, tmin = 3.0e38f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float e = exponential_from_bits(bits[j]);
      t[j] = __fadd_rn(t[j], __fdiv_rn(e, rate[j]));
      if (MODE == kFwdHard) {
        acc_f[j] += (t[j] <= 1.f) ? 1.f : 0.f;
        tmin = fminf(tmin, t[j]);
      } else {
        const float l = (1.f - t[j]) * inv_temp;
        float f, df;
        indicator_eval<IND, MODE == kBwd>(l, f, df);
        if (MODE == kBwd) acc_g[j] = fmaf(df, t[j], acc_g[j]);
        else acc_f[j] += f;
        lmax = fmaxf(lmax, l);
      }
    }
    if (MODE == kFwdHard) {
      if (may_exit && tmin > 1.f) break;
    } else {
      if (may_exit && lmax < -exit_logit) break;
    }
  }
}

// ---- prologue: encoder output -> log_dr, u, log-rate, rate -------------------------------------
struct Latent {
  float hb;      // h + bias
  float log_dr;  // delta
  float u;       // log_r + delta
  float explam;  // exp(lambda)
  float rate;    // exp(lambda) + eps
};

template <bool RAW>
__device__ __forceinline__ Latent prologue(const SamplerParams& p, float hv, float lr, float bias) {
  Latent q;
  if (RAW) {
    q.hb = hv;
    q.log_dr = 0.f;
    q.u = hv;
    const float lam = p.clamp_rate >= 0.f ? softclamp_upper(hv, p.clamp_rate) : hv;
    q.explam = expf(lam);
  } else {
    q.hb = hv + bias;
    q.log_dr = p.exc_only ? softclamp_band(q.hb, p.clamp_dr, 0.f) : softclamp_upper(q.hb, p.clamp_dr);
    q.u = lr + q.log_dr;
    q.explam = expf(softclamp_upper(q.u, p.clamp_rate));
  }
  q.rate = q.explam + kF32Eps;
  return q;
}

template <int MODE, int IND, bool RAW>
__global__ void __launch_bounds__(kThreads) sampler_kernel(const SamplerParams p) {
  __shared__ float s_rowsum[kMaxRowsPerBlock][kThreads / 32];
  __shared__ float s_max[kThreads / 32];

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const long long row0 = static_cast<long long>(blockIdx.x) * p.rows_per_block;
  const int nrows = static_cast<int>(min(static_cast<long long>(p.rows_per_block), p.B - row0));
  uint64_t off = p.offset;
  if (p.offset_dev != nullptr) off += *p.offset_dev;
  const uint32_t off_lo = static_cast<uint32_t>(off), off_hi = static_cast<uint32_t>(off >> 32);

  float rmax = 0.f;
  const bool want_rowsum = MODE != kBwd && !RAW && p.kl_rowsum != nullptr;
  if (want_rowsum) {
    if (tid < kMaxRowsPerBlock * (kThreads / 32)) (&s_rowsum[0][0])[tid] = 0.f;
    __syncthreads();
  }

  // every thread runs every pass so that the warp-level reductions below see full warps
  for (long long k0 = 0; k0 < p.K; k0 += kColsPerPass) {
    const long long k = k0 + 4 * tid;
    const bool act = k < p.K;
    float lr[4] = {0.f, 0.f, 0.f, 0.f}, bias[4] = {0.f, 0.f, 0.f, 0.f}, elr[4] = {0.f, 0.f, 0.f, 0.f};
    if (!RAW && act) {
      const float4 v = ldg4(p.log_r + k);
      lr[0] = v.x, lr[1] = v.y, lr[2] = v.z, lr[3] = v.w;
      if (p.b_enc != nullptr) {
        const float4 b = ldg4(p.b_enc + k);
        bias[0] = b.x, bias[1] = b.y, bias[2] = b.z, bias[3] = b.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) elr[j] = expf(lr[j]);
    }
    float col_lr[4] = {0.f, 0.f, 0.f, 0.f}, col_b[4] = {0.f, 0.f, 0.f, 0.f};

#pragma unroll 1
    for (int r = 0; r < nrows; ++r) {
      float klsum = 0.f;
      if (act) {
      const long long row = row0 + r;
      const long long base = row * p.K + k;
      const float4 hv4 = ldg4(p.h + base);
      const float hv[4] = {hv4.x, hv4.y, hv4.z, hv4.w};
      Latent q[4];
      float rate[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        q[j] = prologue<RAW>(p, hv[j], lr[j], bias[j]);
        rate[j] = q[j].rate;
      }
      const uint32_t quad = static_cast<uint32_t>(((p.row_offset + row) * p.K + k) >> 2);
      float acc_f[4], acc_g[4];
      run_chain<MODE, IND>(p, quad, off_lo, off_hi, rate, acc_f, acc_g);

      if (MODE != kBwd) {
        stg4(p.z + base, make_float4(acc_f[0], acc_f[1], acc_f[2], acc_f[3]));
        if (p.rate != nullptr) stg4(p.rate + base, make_float4(rate[0], rate[1], rate[2], rate[3]));
        if (p.rate_max != nullptr) rmax = fmaxf(fmaxf(fmaxf(rate[0], rate[1]), fmaxf(rate[2], rate[3])), rmax);
        if (!RAW) {
          if (p.log_dr != nullptr)
            stg4(p.log_dr + base, make_float4(q[0].log_dr, q[1].log_dr, q[2].log_dr, q[3].log_dr));
          if (p.kl != nullptr || p.kl_rowsum != nullptr) {
            float klv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float d = q[j].log_dr;
              klv[j] = elr[j] * (1.f + expf(d) * (d - 1.f));  // main/vae.py:448-449
            }
            if (p.kl != nullptr) stg4(p.kl + base, make_float4(klv[0], klv[1], klv[2], klv[3]));
            klsum = (klv[0] + klv[1]) + (klv[2] + klv[3]);
          }
        }
      } else {
        const float4 gz4 = ldg4(p.g_z + base);
        const float gz[4] = {gz4.x, gz4.y, gz4.z, gz4.w};
        float gld[4] = {0.f, 0.f, 0.f, 0.f};
        if (!RAW && p.g_log_dr != nullptr) {
          const float4 g = ldg4(p.g_log_dr + base);
          gld[0] = g.x, gld[1] = g.y, gld[2] = g.z, gld[3] = g.w;
        }
        float gh[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // d z / d lambda = (sum_n f'(l_n) t_n) / temp * exp(lambda) / rate
          const float dz = acc_g[j] * p.inv_temp * (q[j].explam / q[j].rate);
          if (RAW) {
            gh[j] = gz[j] * dz * (p.clamp_rate >= 0.f ? softclamp_upper_grad(q[j].u, p.clamp_rate) : 1.f);
          } else {
            const float d = q[j].log_dr;
            const float g_u = gz[j] * dz * softclamp_upper_grad(q[j].u, p.clamp_rate);
            const float ed = expf(d);
            const float g_d = g_u + p.kl_scale * elr[j] * d * ed + gld[j];
            const float dcl = p.exc_only ? softclamp_band_grad(q[j].hb, p.clamp_dr, 0.f)
                                         : softclamp_upper_grad(q[j].hb, p.clamp_dr);
            gh[j] = g_d * dcl;
            col_lr[j] += g_u + p.kl_scale * (elr[j] * (1.f + ed * (d - 1.f)));
            col_b[j] += gh[j];
          }
        }
        stg4(p.g_h + base, make_float4(gh[0], gh[1], gh[2], gh[3]));
      }
      }  // act
      if (want_rowsum) {
        __syncwarp();
        const float s = warp_sum(klsum);
        if (lane == 0) s_rowsum[r][warp] += s;
      }
    }
    if (MODE == kBwd && !RAW && act) {
      const long long pbase = static_cast<long long>(blockIdx.x) * p.K + k;
      stg4(p.part_log_r + pbase, make_float4(col_lr[0], col_lr[1], col_lr[2], col_lr[3]));
      if (p.part_b != nullptr) stg4(p.part_b + pbase, make_float4(col_b[0], col_b[1], col_b[2], col_b[3]));
    }
  }

  if (MODE != kBwd) {
    if (want_rowsum) {
      __syncthreads();
      if (tid < nrows) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) s += s_rowsum[tid][w];
        p.kl_rowsum[row0 + tid] = s;
      }
    }
    if (p.rate_max != nullptr) {
      const float m = warp_max(rmax);
      if (lane == 0) s_max[warp] = m;
      __syncthreads();
      if (tid == 0) {
        float mm = s_max[0];
#pragma unroll
        for (int w = 1; w < kThreads / 32; ++w) mm = fmaxf(mm, s_max[w]);
        atomicMax(reinterpret_cast<int*>(p.rate_max), __float_as_int(mm));  // rates are > 0
      }
    }
  }
}

// ---- fixed-order column reduction of the (P,K) partials ----------------------------------------
constexpr int kRedX = 32, kRedY = 8;
__global__ void __launch_bounds__(kRedX* kRedY) colsum_finalize_kernel(const float* __restrict__ part, float* __restrict__ out,
                                                                      long long P, long long K, int accumulate) {
  __shared__ float s[kRedY][kRedX + 1];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const long long k = static_cast<long long>(blockIdx.x) * kRedX + tx;
  float a = 0.f;
  if (k < K)
    for (long long r = ty; r < P; r += kRedY) a += part[r * K + k];
  s[ty][tx] = a;
  __syncthreads();
  if (ty == 0 && k < K) {
    float v = s[0][tx];
#pragma unroll
    for (int y = 1; y < kRedY; ++y) v += s[y][tx];
    out[k] = accumulate ? out[k] + v : v;
  }
}

// ---- standalone KL ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kl_fwd_kernel(const float* __restrict__ log_dr, const float* __restrict__ log_r,
                                                     float* __restrict__ kl, long long n4, long long K) {
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += gridDim.x * 256ll) {
    const long long e = 4 * i;
    const float4 d = ldg4(log_dr + e);
    const float4 l = ldg4(log_r + (e % K));
    float4 o;
    o.x = expf(l.x) * (1.f + expf(d.x) * (d.x - 1.f));
    o.y = expf(l.y) * (1.f + expf(d.y) * (d.y - 1.f));
    o.z = expf(l.z) * (1.f + expf(d.z) * (d.z - 1.f));
    o.w = expf(l.w) * (1.f + expf(d.w) * (d.w - 1.f));
    stg4(kl + e, o);
  }
}

__global__ void __launch_bounds__(kThreads) kl_bwd_kernel(const float* __restrict__ log_dr, const float* __restrict__ log_r,
                                                          const float* __restrict__ g_kl, float* __restrict__ g_log_dr,
                                                          float* __restrict__ part, long long B, long long K, int rows_per_block) {
  const long long row0 = static_cast<long long>(blockIdx.x) * rows_per_block;
  const int nrows = static_cast<int>(min(static_cast<long long>(rows_per_block), B - row0));
  for (long long k = 4 * threadIdx.x; k < K; k += kColsPerPass) {
    const float4 l4 = ldg4(log_r + k);
    const float el[4] = {expf(l4.x), expf(l4.y), expf(l4.z), expf(l4.w)};
    float col[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < nrows; ++r) {
      const long long base = (row0 + r) * K + k;
      const float4 d4 = ldg4(log_dr + base), g4 = ldg4(g_kl + base);
      const float d[4] = {d4.x, d4.y, d4.z, d4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w};
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float ed = expf(d[j]);
        o[j] = g[j] * el[j] * d[j] * ed;                     // d kl / d log_dr = exp(log_r) * d * exp(d)
        col[j] += g[j] * el[j] * (1.f + ed * (d[j] - 1.f));  // d kl / d log_r  = kl
      }
      stg4(g_log_dr + base, make_float4(o[0], o[1], o[2], o[3]));
    }
    stg4(part + static_cast<long long>(blockIdx.x) * K + k, make_float4(col[0], col[1], col[2], col[3]));
  }
}

// ---- parity harness: dump the draws -------------------------------------------------------------
__global__ void debug_draws_kernel(float* U, float* E, long long B, long long K, long long row_offset, int n_exp,
                                   uint64_t offset, const uint64_t* offset_dev, const PhiloxKey ks) {
  uint64_t off = offset;
  if (offset_dev != nullptr) off += *offset_dev;
  const long long total = static_cast<long long>(n_exp) * B * K;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long n = i / (B * K), bk = i % (B * K);
    const unsigned long long gi = static_cast<unsigned long long>(row_offset * K + bk);
    const uint4 r = philox4x32_10(static_cast<uint32_t>(gi >> 2), static_cast<uint32_t>(n), static_cast<uint32_t>(off),
                                  static_cast<uint32_t>(off >> 32), ks);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    const uint32_t bits = w[gi & 3];
    if (U != nullptr) U[i] = uniform_from_bits(bits);
    if (E != nullptr) E[i] = exponential_from_bits(bits);
  }
}

// ---- host-side launch --------------------------------------------------------------------------
static int rows_per_block_for(long long B) {
  // enough blocks for several waves over 148 SMs, at most kMaxRowsPerBlock rows per block
  const long long target_blocks = 148ll * 12;
  long long r = B / target_blocks;
  if (r < 1) r = 1;
  if (r > kMaxRowsPerBlock) r = kMaxRowsPerBlock;
  return static_cast<int>(r);
}

template <int MODE, bool RAW>
static cudaError_t launch_ind(const SamplerParams& p, int indicator, dim3 grid, cudaStream_t s) {
  switch (indicator) {
    case kSigmoid: sampler_kernel<MODE, kSigmoid, RAW><<<grid, kThreads, 0, s>>>(p); break;
    case kCosine: sampler_kernel<MODE, kCosine, RAW><<<grid, kThreads, 0, s>>>(p); break;
    case kLinear: sampler_kernel<MODE, kLinear, RAW><<<grid, kThreads, 0, s>>>(p); break;
    case kCubic: sampler_kernel<MODE, kCubic, RAW><<<grid, kThreads, 0, s>>>(p); break;
    case kQuintic: sampler_kernel<MODE, kQuintic, RAW><<<grid, kThreads, 0, s>>>(p); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

static float resolve_exit(float exit_logit, int indicator, bool hard) {
  if (exit_logit < 0.f) return -1.f;                       // never
  if (hard) return 0.f;                                    // value unused, >= 0 enables the exit
  if (exit_logit == 0.f) return indicator == kSigmoid ? 88.73f : 1.0f;  // lossless
  return exit_logit;
}

static int check_common(const char* fn, long long B, long long K, long long row_offset, int n_exp, float temp,
                        int indicator) {
  if (B < 0 || K <= 0 || row_offset < 0) return fail(PVAE_ERR_INVALID, "%s: bad shape B=%lld K=%lld row_offset=%lld", fn, B, K, row_offset);
  if (K % 4 != 0) return fail(PVAE_ERR_UNSUPPORTED, "%s: K=%lld must be a multiple of 4", fn, K);
  if (n_exp < 0) return fail(PVAE_ERR_INVALID, "%s: n_exp=%d must be >= 0", fn, n_exp);
  if (!(temp >= 0.f)) return fail(PVAE_ERR_INVALID, "%s: must be non-neg: temp=%g", fn, temp);  // base/distributions.py:15
  if (indicator < 0 || indicator > 4) return fail(PVAE_ERR_INVALID, "%s: unknown indicator %d", fn, indicator);  // :16
  if (((row_offset + B) * K) >> 2 > 0xFFFFFFFFll) return fail(PVAE_ERR_UNSUPPORTED, "%s: more than 2^34 latents", fn);
  return PVAE_OK;
}

static int finalize_cols(const float* part, float* out, long long P, long long K, int accumulate, cudaStream_t s) {
  dim3 grid(static_cast<unsigned>((K + kRedX - 1) / kRedX)), block(kRedX, kRedY);
  colsum_finalize_kernel<<<grid, block, 0, s>>>(part, out, P, K, accumulate);
  return cuda_ok("colsum_finalize", cudaGetLastError());
}

}  // namespace pvae

using namespace pvae;

extern "C" int64_t pvae_num_partials(int64_t B) {
  if (B <= 0) return 0;
  const int r = rows_per_block_for(B);
  return (B + r - 1) / r;
}

extern "C" int pvae_sampler_fwd(const float* h, const float* log_r, const float* b_enc, float* z, float* log_dr,
                                float* rate, float* kl, float* kl_rowsum, float* rate_max, int64_t B, int64_t K,
                                int64_t row_offset, int n_exp, float temp, int indicator, int exc_only,
                                float exit_logit, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                                void* stream) {
  if (int rc = check_common("pvae_sampler_fwd", B, K, row_offset, n_exp, temp, indicator)) return rc;
  if (B == 0) return PVAE_OK;
  if (!h || !log_r || !z) return fail(PVAE_ERR_INVALID, "pvae_sampler_fwd: null h/log_r/z");
  SamplerParams p{};
  p.h = h, p.log_r = log_r, p.b_enc = b_enc, p.z = z, p.log_dr = log_dr, p.rate = rate, p.kl = kl;
  p.kl_rowsum = kl_rowsum, p.rate_max = rate_max;
  p.B = B, p.K = K, p.row_offset = row_offset, p.rows_per_block = rows_per_block_for(B);
  p.n_exp = n_exp, p.temp = temp, p.inv_temp = temp > 0.f ? 1.f / temp : 0.f;
  p.exit_logit = resolve_exit(exit_logit, indicator, temp == 0.f);
  p.clamp_dr = 10.f, p.clamp_rate = 5.f, p.exc_only = exc_only;  // main/vae.py:403-409
  p.offset = offset, p.offset_dev = offset_dev;
  philox_key_schedule(seed, p.ks);
  dim3 grid(static_cast<unsigned>((B + p.rows_per_block - 1) / p.rows_per_block));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = temp == 0.f ? launch_ind<kFwdHard, false>(p, kSigmoid, grid, s)
                              : launch_ind<kFwdSoft, false>(p, indicator, grid, s);
  return cuda_ok("pvae_sampler_fwd", e);
}

extern "C" int pvae_sampler_bwd(const float* h, const float* log_r, const float* b_enc, const float* g_z,
                                const float* g_log_dr, float kl_scale, float* g_h, float* g_log_r, float* g_b_enc,
                                float* partial_ws, int accumulate, int64_t B, int64_t K, int64_t row_offset,
                                int n_exp, float temp, int indicator, int exc_only, float exit_logit, uint64_t seed,
                                uint64_t offset, const uint64_t* offset_dev, void* stream) {
  if (int rc = check_common("pvae_sampler_bwd", B, K, row_offset, n_exp, temp, indicator)) return rc;
  if (temp == 0.f) return fail(PVAE_ERR_INVALID, "pvae_sampler_bwd: hard counts (temp == 0) have no gradient");
  if (B == 0) return PVAE_OK;
  if (!h || !log_r || !g_z || !g_h || !g_log_r || !partial_ws)
    return fail(PVAE_ERR_INVALID, "pvae_sampler_bwd: null h/log_r/g_z/g_h/g_log_r/partial_ws");
  SamplerParams p{};
  p.h = h, p.log_r = log_r, p.b_enc = b_enc, p.g_z = g_z, p.g_log_dr = g_log_dr, p.g_h = g_h;
  p.B = B, p.K = K, p.row_offset = row_offset, p.rows_per_block = rows_per_block_for(B);
  const long long P = (B + p.rows_per_block - 1) / p.rows_per_block;
  p.part_log_r = partial_ws;
  p.part_b = g_b_enc ? partial_ws + P * K : nullptr;
  p.n_exp = n_exp, p.temp = temp, p.inv_temp = 1.f / temp;
  p.exit_logit = resolve_exit(exit_logit, indicator, false);
  p.clamp_dr = 10.f, p.clamp_rate = 5.f, p.exc_only = exc_only, p.kl_scale = kl_scale;
  p.offset = offset, p.offset_dev = offset_dev;
  philox_key_schedule(seed, p.ks);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = cuda_ok("pvae_sampler_bwd", launch_ind<kBwd, false>(p, indicator, dim3(static_cast<unsigned>(P)), s))) return rc;
  if (int rc = finalize_cols(p.part_log_r, g_log_r, P, K, accumulate, s)) return rc;
  if (g_b_enc)
    if (int rc = finalize_cols(p.part_b, g_b_enc, P, K, accumulate, s)) return rc;
  return PVAE_OK;
}

extern "C" int pvae_rsample_fwd(const float* log_rate, float* z, float* rate, int64_t B, int64_t K, int64_t row_offset,
                                int n_exp, float temp, int indicator, float clamp, float exit_logit, uint64_t seed,
                                uint64_t offset, const uint64_t* offset_dev, void* stream) {
  if (int rc = check_common("pvae_rsample_fwd", B, K, row_offset, n_exp, temp, indicator)) return rc;
  if (B == 0) return PVAE_OK;
  if (!log_rate || !z) return fail(PVAE_ERR_INVALID, "pvae_rsample_fwd: null log_rate/z");
  SamplerParams p{};
  p.h = log_rate, p.z = z, p.rate = rate;
  p.B = B, p.K = K, p.row_offset = row_offset, p.rows_per_block = rows_per_block_for(B);
  p.n_exp = n_exp, p.temp = temp, p.inv_temp = temp > 0.f ? 1.f / temp : 0.f;
  p.exit_logit = resolve_exit(exit_logit, indicator, temp == 0.f);
  p.clamp_rate = clamp;
  p.offset = offset, p.offset_dev = offset_dev;
  philox_key_schedule(seed, p.ks);
  dim3 grid(static_cast<unsigned>((B + p.rows_per_block - 1) / p.rows_per_block));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = temp == 0.f ? launch_ind<kFwdHard, true>(p, kSigmoid, grid, s)
                              : launch_ind<kFwdSoft, true>(p, indicator, grid, s);
  return cuda_ok("pvae_rsample_fwd", e);
}

extern "C" int pvae_rsample_bwd(const float* log_rate, const float* g_z, float* g_log_rate, int64_t B, int64_t K,
                                int64_t row_offset, int n_exp, float temp, int indicator, float clamp, float exit_logit,
                                uint64_t seed, uint64_t offset, const uint64_t* offset_dev, void* stream) {
  if (int rc = check_common("pvae_rsample_bwd", B, K, row_offset, n_exp, temp, indicator)) return rc;
  if (temp == 0.f) return fail(PVAE_ERR_INVALID, "pvae_rsample_bwd: hard counts (temp == 0) have no gradient");
  if (B == 0) return PVAE_OK;
  if (!log_rate || !g_z || !g_log_rate) return fail(PVAE_ERR_INVALID, "pvae_rsample_bwd: null pointer");
  SamplerParams p{};
  p.h = log_rate, p.g_z = g_z, p.g_h = g_log_rate;
  p.B = B, p.K = K, p.row_offset = row_offset, p.rows_per_block = rows_per_block_for(B);
  p.n_exp = n_exp, p.temp = temp, p.inv_temp = 1.f / temp;
  p.exit_logit = resolve_exit(exit_logit, indicator, false);
  p.clamp_rate = clamp;
  p.offset = offset, p.offset_dev = offset_dev;
  philox_key_schedule(seed, p.ks);
  dim3 grid(static_cast<unsigned>((B + p.rows_per_block - 1) / p.rows_per_block));
  return cuda_ok("pvae_rsample_bwd", launch_ind<kBwd, true>(p, indicator, grid, static_cast<cudaStream_t>(stream)));
}

extern "C" int pvae_kl_fwd(const float* log_dr, const float* log_r, float* kl, int64_t B, int64_t K, void* stream) {
  if (B < 0 || K <= 0 || K % 4 != 0) return fail(PVAE_ERR_INVALID, "pvae_kl_fwd: bad shape B=%lld K=%lld", (long long)B, (long long)K);
  if (B == 0) return PVAE_OK;
  if (!log_dr || !log_r || !kl) return fail(PVAE_ERR_INVALID, "pvae_kl_fwd: null pointer");
  const long long n4 = B * K / 4;
  const unsigned grid = static_cast<unsigned>(std::min<long long>((n4 + 255) / 256, 148ll * 16));
  kl_fwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(log_dr, log_r, kl, n4, K);
  return cuda_ok("pvae_kl_fwd", cudaGetLastError());
}

extern "C" int pvae_kl_bwd(const float* log_dr, const float* log_r, const float* g_kl, float* g_log_dr, float* g_log_r,
                           float* partial_ws, int accumulate, int64_t B, int64_t K, void* stream) {
  if (B < 0 || K <= 0 || K % 4 != 0) return fail(PVAE_ERR_INVALID, "pvae_kl_bwd: bad shape B=%lld K=%lld", (long long)B, (long long)K);
  if (B == 0) return PVAE_OK;
  if (!log_dr || !log_r || !g_kl || !g_log_dr || !g_log_r || !partial_ws) return fail(PVAE_ERR_INVALID, "pvae_kl_bwd: null pointer");
  const int rpb = rows_per_block_for(B);
  const long long P = (B + rpb - 1) / rpb;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  kl_bwd_kernel<<<static_cast<unsigned>(P), kThreads, 0, s>>>(log_dr, log_r, g_kl, g_log_dr, partial_ws, B, K, rpb);
  if (int rc = cuda_ok("pvae_kl_bwd", cudaGetLastError())) return rc;
  return finalize_cols(partial_ws, g_log_r, P, K, accumulate, s);
}

extern "C" int pvae_debug_draws(float* U, float* E, int64_t B, int64_t K, int64_t row_offset, int n_exp, uint64_t seed,
                                uint64_t offset, const uint64_t* offset_dev, void* stream) {
  if (B < 0 || K <= 0 || n_exp < 0 || row_offset < 0) return fail(PVAE_ERR_INVALID, "pvae_debug_draws: bad shape");
  const long long total = static_cast<long long>(n_exp) * B * K;
  if (total == 0) return PVAE_OK;
  PhiloxKey ks;
  philox_key_schedule(seed, ks);
  const unsigned grid = static_cast<unsigned>(std::min<long long>((total + 255) / 256, 148ll * 32));
  debug_draws_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(U, E, B, K, row_offset, n_exp, offset, offset_dev, ks);
  return cuda_ok("pvae_debug_draws", cudaGetLastError());
}
