// Small fused kernels of the training step around the sampler and the GEMMs (sm_100a):
// reconstruction residual + per-sample squared error, loss assembly, gradient norm / clip
// coefficient and the Adamax update over the flat parameter buffer.
//
// Reference semantics: main/vae.py:68-72 (loss_recon), main/train_vae.py:346-377 (loss assembly,
// clip_grad_norm_), base/adamax.py:34-100 (Adamax "fast").
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pvae_b200.h"
#include "pvae_device.cuh"
#include "pvae_host.h"

namespace pvae {

// ---- residual: g_y = (2 / B_global) * (y - x), recon[b] = sum_d (y - x)^2 -----------------------
// One warp per sample; D is a multiple of 4.
// y_parts > 1: y arrives as split-K partial sums y_parts * (B, D) apart (pvae_gemm_tf32 with C = NULL), summed here in order
__global__ void __launch_bounds__(256) recon_kernel(const float* __restrict__ y, const float* __restrict__ x,
                                                    float* __restrict__ g_y, float* __restrict__ g_y_lo,
                                                    float* __restrict__ recon, long long B, long long D, float g_scale, int y_parts) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long b = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
    float acc = 0.f;
    for (long long d = 4 * lane; d < D; d += 128) {
      float4 yv = ldg4(y + b * D + d);
      const float4 xv = ldg4(x + b * D + d);
      for (int s = 1; s < y_parts; ++s) {
        const float4 w = ldg4(y + s * B * D + b * D + d);
        yv.x += w.x, yv.y += w.y, yv.z += w.z, yv.w += w.w;
      }
      const float4 r = make_float4(yv.x - xv.x, yv.y - xv.y, yv.z - xv.z, yv.w - xv.w);
      acc += (r.x * r.x + r.y * r.y) + (r.z * r.z + r.w * r.w);
      if (g_y != nullptr) {
        const float4 gv = make_float4(g_scale * r.x, g_scale * r.y, g_scale * r.z, g_scale * r.w);
        stg4(g_y + b * D + d, gv);
        if (g_y_lo != nullptr)  // TF32 low plane for the dgrad / wgrad GEMMs (x - trunc_tf32(x), exact)
          stg4(g_y_lo + b * D + d, make_float4(gv.x - __uint_as_float(__float_as_uint(gv.x) & 0xFFFFE000u),
                                               gv.y - __uint_as_float(__float_as_uint(gv.y) & 0xFFFFE000u),
                                               gv.z - __uint_as_float(__float_as_uint(gv.z) & 0xFFFFE000u),
                                               gv.w - __uint_as_float(__float_as_uint(gv.w) & 0xFFFFE000u)));
      }
    }
    acc = warp_sum(acc);
    if (lane == 0 && recon != nullptr) recon[b] = acc;
  }
}

// ---- loss assembly: out[0] = sum_b(recon + beta * kl_row) / B_global, out[1] = sum recon / Bg,
// out[2] = sum kl_row / Bg.  Single block, fixed order (deterministic).
__global__ void __launch_bounds__(1024) loss_kernel(const float* __restrict__ recon, const float* __restrict__ kl_row,
                                                    float* __restrict__ out, long long B, float beta, float inv_bg,
                                                    int recon_parts) {
  __shared__ float s_r[32], s_k[32];
  pdl_launch_dependents();
  pdl_wait();
  float r = 0.f, k = 0.f;
  for (long long b = threadIdx.x; b < B; b += 1024) {
    float rb = 0.f;  // recon may arrive as per-column-tile partial sums (fused decoder epilogue)
    for (int j = 0; j < recon_parts; ++j) rb += recon[b * recon_parts + j];
    r += rb;
    k += kl_row[b];
  }
  r = warp_sum(r), k = warp_sum(k);
  if ((threadIdx.x & 31) == 0) s_r[threadIdx.x >> 5] = r, s_k[threadIdx.x >> 5] = k;
  __syncthreads();
  if (threadIdx.x < 32) {
    r = warp_sum(s_r[threadIdx.x]), k = warp_sum(s_k[threadIdx.x]);
    if (threadIdx.x == 0) {
      out[0] = (r + beta * k) * inv_bg;
      out[1] = r * inv_bg;
      out[2] = k * inv_bg;
    }
  }
}

// ---- gradient L2 norm over the flat buffer, two fixed-order stages -----------------------------
constexpr int kNormBlocks = 148;
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ part) {
  __shared__ float s[8];
  pdl_launch_dependents();
  pdl_wait();
  float a = 0.f;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) a = fmaf(g[i], g[i], a);
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s[w];
    part[blockIdx.x] = t;
  }
}
// out[0] = ||g||, out[1] = clip coefficient min(1, max_norm / (||g|| + 1e-6)) (torch clip_grad_norm_)
__global__ void norm_finalize_kernel(const float* __restrict__ part, int nparts, float max_norm, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < nparts; ++i) t += part[i];
    const float nrm = sqrtf(t);
    out[0] = nrm;
    out[1] = max_norm > 0.f ? fminf(1.f, max_norm / (nrm + 1e-6f)) : 1.f;
  }
}

// ---- Adamax (base/adamax.py:79-88): m = b1 m + (1-b1) g; u = max(b2 u, |g| + eps); p -= clr m / u
__global__ void __launch_bounds__(256) adamax_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                     float* __restrict__ u, long long n, float clr, float beta1, float beta2,
                                                     float eps, float weight_decay, const float* __restrict__ clip_coef,
                                                     float* __restrict__ p_lo) {
  pdl_launch_dependents();
  pdl_wait();
  const float cc = clip_coef != nullptr ? clip_coef[1] : 1.f;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    float pi = p[i], mi = m[i], ui = u[i];
    const float lo = adamax_apply(pi, mi, ui, cc == 1.f ? g[i] : g[i] * cc, clr, beta1, beta2, eps, weight_decay);
    m[i] = mi;
    u[i] = ui;
    p[i] = pi;
    if (p_lo != nullptr) p_lo[i] = lo;
  }
}

// ---- gradient reduction fused with Adamax (single GPU, no clipping) --------------------------------------------
// The step's gradients arrive as partial sums: split-K partials of the two wgrad GEMMs and per-block column partials
// of the sampler backward.  Instead of three reduction kernels followed by Adamax, ONE kernel reduces each parameter's
// partials (in the order the stand-alone reductions use, so the result is bit-identical), stores the gradient and
// applies the update.  blockDim = (32, 32).
constexpr int kMaxSegments = 8;
constexpr int kDenseRows = 8;  // rows of the 32 x 32 block that work on a dense segment (256 threads x 4 elements)
struct SegTable {
  long long offset[kMaxSegments], length[kMaxSegments], stride[kMaxSegments];
  const float* partials[kMaxSegments];
  const float* partials2[kMaxSegments];  // columnar segments: a second set of row partials, added with weight scale2
  float scale2[kMaxSegments];
  int n_partials[kMaxSegments], n_partials2[kMaxSegments], columnar[kMaxSegments], first_block[kMaxSegments + 1];
  int n;
};
// fixed-order column sum of (P, len) row partials for column k by the 32 x 32 block (the order of colsum_finalize_kernel)
__device__ __forceinline__ float block_colsum(const float* __restrict__ part, long long P, long long len, long long k, bool in_range,
                                              float (*s)[33]) {
  const int tx = threadIdx.x, ty = threadIdx.y;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (in_range) {
    long long r = ty;
    for (; r + 3 * 32 < P; r += 4 * 32) {
      a0 += __ldg(part + r * len + k);
      a1 += __ldg(part + (r + 32) * len + k);
      a2 += __ldg(part + (r + 64) * len + k);
      a3 += __ldg(part + (r + 96) * len + k);
    }
    for (; r < P; r += 32) a0 += __ldg(part + r * len + k);
  }
  s[ty][tx] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  float v = 0.f;
  if (ty == 0 && in_range) {
    v = s[0][tx];
#pragma unroll
    for (int y = 1; y < 32; ++y) v += s[y][tx];
  }
  __syncthreads();
  return v;
}
__device__ __forceinline__ void adamax_update(float* p, float* m, float* u, float* p_lo, long long i, float gi, float clr, float beta1,
                                              float beta2, float eps, float weight_decay) {
  float pi = p[i], mi = m[i], ui = u[i];
  const float lo = adamax_apply(pi, mi, ui, gi, clr, beta1, beta2, eps, weight_decay);
  m[i] = mi;
  u[i] = ui;
  p[i] = pi;
  if (p_lo != nullptr) p_lo[i] = lo;
}
__global__ void __launch_bounds__(1024) adamax_partials_kernel(const __grid_constant__ SegTable t, float* __restrict__ p,
                                                               float* __restrict__ g_out, float* __restrict__ m,
                                                               float* __restrict__ u, float clr, float beta1, float beta2, float eps,
                                                               float weight_decay, float* __restrict__ p_lo,
                                                               const float* __restrict__ lr_sched, int n_sched,
                                                               const int* __restrict__ state) {
  __shared__ float s[32][33];
  pdl_launch_dependents();
  pdl_wait();
  if (state != nullptr) {  // CUDA-graph replay: step count and learning rate live on the device (see adamax_dev_kernel)
    const int step = state[0] + 1;
    const float lr = lr_sched[min(step - 1, n_sched - 1)];
    clr = static_cast<float>(static_cast<double>(lr) / (1.0 - pow(static_cast<double>(beta1), step)));
  }
  int seg = 0;
  while (seg + 1 < t.n && static_cast<int>(blockIdx.x) >= t.first_block[seg + 1]) ++seg;
  const int blk = blockIdx.x - t.first_block[seg];
  const float* __restrict__ part = t.partials[seg];
  const long long off = t.offset[seg], len = t.length[seg];
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (t.columnar[seg]) {  // (n_partials, len) row partials: the reduction of colsum_finalize_kernel (sampler.cu)
    const long long k = static_cast<long long>(blk) * 32 + tx;
    float v = block_colsum(part, t.n_partials[seg], len, k, k < len, s);
    if (t.partials2[seg] != nullptr) {
      const float v2 = block_colsum(t.partials2[seg], t.n_partials2[seg], len, k, k < len, s);
      v = fmaf(t.scale2[seg], v2, v);
    }
    if (ty == 0 && k < len) {
      g_out[off + k] = v;
      if (p != nullptr) adamax_update(p, m, u, p_lo, off + k, v, clr, beta1, beta2, eps, weight_decay);
    }
    return;
  }
  // dense split-K partials (n_partials, stride): the sum of splitk_reduce_kernel (gemm_tc.cu), four elements per thread;
  // only the first kDenseRows rows of the block's 32 x 32 threads work, so that a weight matrix spreads over 4x as many
  // blocks (128 instead of 32 for 131072 elements: the kernel is latency-bound, it needs the SMs, not the threads)
  if (ty >= kDenseRows) return;
  const long long i4 = static_cast<long long>(blk) * (32 * kDenseRows) + ty * 32 + tx;
  if (4 * i4 >= len) return;
  const int splits = t.n_partials[seg];
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (part == nullptr) {
    a = *reinterpret_cast<const float4*>(g_out + off + 4 * i4);  // already reduced
  } else {
    const float4* w = reinterpret_cast<const float4*>(part);
    const long long n4 = t.stride[seg] >> 2;
    int sp = 0;
    for (; sp + 8 <= splits; sp += 8) {
      float4 v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = __ldg(w + (sp + q) * n4 + i4);
#pragma unroll
      for (int q = 0; q < 8; ++q) a.x += v[q].x, a.y += v[q].y, a.z += v[q].z, a.w += v[q].w;
    }
    for (; sp < splits; ++sp) {
      const float4 b = __ldg(w + sp * n4 + i4);
      a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    }
    *reinterpret_cast<float4*>(g_out + off + 4 * i4) = a;
  }
  if (p == nullptr) return;  // reduction only (data parallel: the exchange kernel applies the update)
  const long long i = off + 4 * i4;
  adamax_update(p, m, u, p_lo, i, a.x, clr, beta1, beta2, eps, weight_decay);
  adamax_update(p, m, u, p_lo, i + 1, a.y, clr, beta1, beta2, eps, weight_decay);
  adamax_update(p, m, u, p_lo, i + 2, a.z, clr, beta1, beta2, eps, weight_decay);
  adamax_update(p, m, u, p_lo, i + 3, a.w, clr, beta1, beta2, eps, weight_decay);
}

// ---- Adamax with the step-dependent scalars on the device (CUDA-graph replay of a training step) -------------
// state[0] = number of optimiser steps taken so far (int), state[2..3] = Philox offset of the step (uint64).
// lr_sched[i] = learning rate of optimiser step i + 1.  The kernel reads the counter, the advance kernel that follows
// it in the stream bumps the counter and the Philox offset - so a captured graph replays step after step with no
// host-side parameter at all.
__global__ void __launch_bounds__(256) adamax_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ u, long long n, const float* __restrict__ lr_sched,
                                                         int n_sched, const int* __restrict__ state, float beta1, float beta2,
                                                         float eps, float weight_decay, const float* __restrict__ clip_coef,
                                                         float* __restrict__ p_lo) {
  pdl_launch_dependents();
  pdl_wait();
  const int step = state[0] + 1;
  const float lr = lr_sched[min(step - 1, n_sched - 1)];
  const float clr = static_cast<float>(static_cast<double>(lr) / (1.0 - pow(static_cast<double>(beta1), step)));  // base/adamax.py:75-76
  const float cc = clip_coef != nullptr ? clip_coef[1] : 1.f;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) {
    float pi = p[i], mi = m[i], ui = u[i];
    const float lo = adamax_apply(pi, mi, ui, cc == 1.f ? g[i] : g[i] * cc, clr, beta1, beta2, eps, weight_decay);
    m[i] = mi;
    u[i] = ui;
    p[i] = pi;
    if (p_lo != nullptr) p_lo[i] = lo;
  }
}
__global__ void advance_state_kernel(int* __restrict__ state, unsigned long long philox_increment, float* __restrict__ rmax_step,
                                     float* __restrict__ rmax_ring, int ring_n) {
  pdl_wait();  // every reader of this step's counters has been launched before us and the predecessor has finished
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (rmax_step != nullptr && rmax_ring != nullptr && ring_n > 0) {
      // the replayed sampler max-reduces rate.max() into one fixed word: move it to this step's slot of the epoch ring
      // (the eager path writes slot gstep % ring_n directly, main/train_vae.py:388) and reset the word
      rmax_ring[state[0] % ring_n] = *rmax_step;
      *rmax_step = 0.f;
    }
    state[0] += 1;
    *reinterpret_cast<unsigned long long*>(state + 2) += philox_increment;
  }
}

// ---- TF32 low plane of an array: lo = x - trunc_tf32(x) (exact); what the plane-fed GEMMs load next to x itself ----
__global__ void __launch_bounds__(256) tf32_lo_kernel(const float* __restrict__ x, float* __restrict__ lo, long long n4) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += gridDim.x * 256ll) {
    const float4 v = ldg4(x + 4 * i);
    stg4(lo + 4 * i, make_float4(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u), v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u),
                                 v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u), v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u)));
  }
}

// ---- gather up to kGatherMax gradient tensors into the flat buffer (one launch instead of one copy each) ----
constexpr int kGatherMax = 96;
struct GatherTable {
  const float* src[kGatherMax];
  long long off[kGatherMax + 1];  // destination offsets in floats; off[n] = end
  int n;
};
__global__ void __launch_bounds__(256) gather_kernel(const __grid_constant__ GatherTable t, float* __restrict__ dst) {
  pdl_launch_dependents();
  pdl_wait();
  // blockIdx.y = tensor; blocks of a tensor stride over it
  const int i = blockIdx.y;
  if (i >= t.n) return;
  const long long n = t.off[i + 1] - t.off[i];
  const float* __restrict__ s = t.src[i];
  float* __restrict__ d = dst + t.off[i];
  for (long long k = blockIdx.x * 256ll + threadIdx.x; k < n; k += gridDim.x * 256ll) d[k] = s != nullptr ? s[k] : 0.f;
}

}  // namespace pvae

using namespace pvae;

extern "C" int pvae_adamax_from_partials(float* p, float* g_out, float* exp_avg, float* exp_inf, const pvae_grad_segment* segs,
                                         int n_segs, float lr, int step, float beta1, float beta2, float eps, float weight_decay,
                                         void* stream) {
  return adamax_from_partials_impl(p, nullptr, g_out, exp_avg, exp_inf, segs, n_segs, lr, step, beta1, beta2, eps, weight_decay, AdamaxDev{},
                                   stream);
}

int pvae::adamax_from_partials_impl(float* p, float* p_lo, float* g_out, float* exp_avg, float* exp_inf, const pvae_grad_segment* segs,
                                    int n_segs, float lr, int step, float beta1, float beta2, float eps, float weight_decay,
                                    const AdamaxDev& dev, void* stream) {
  if (!g_out || !segs || n_segs < 1 || n_segs > kMaxSegments || (p && (!exp_avg || !exp_inf || (step < 1 && !dev.state))))
    return fail(PVAE_ERR_INVALID, "pvae_adamax_from_partials: bad argument");
  if (!p || dev.state) step = 1;  // p == NULL: reduce the partials into g_out only; device-side counters: computed in the kernel
  SegTable t = {};
  t.n = n_segs;
  int blocks = 0;
  for (int i = 0; i < n_segs; ++i) {
    const pvae_grad_segment& g = segs[i];
    if (g.length <= 0 || g.offset < 0 || (g.partials && g.n_partials < 1) || (!g.columnar && (g.length % 4 || g.offset % 4 || g.stride % 4)))
      return fail(PVAE_ERR_INVALID, "pvae_adamax_from_partials: bad segment %d", i);
    if (g.columnar && !g.partials) return fail(PVAE_ERR_INVALID, "pvae_adamax_from_partials: a columnar segment needs partials");
    t.offset[i] = g.offset, t.length[i] = g.length, t.stride[i] = g.stride, t.partials[i] = g.partials;
    t.n_partials[i] = g.n_partials, t.columnar[i] = g.columnar, t.first_block[i] = blocks;
    t.partials2[i] = g.columnar ? g.partials2 : nullptr, t.n_partials2[i] = g.n_partials2, t.scale2[i] = g.scale2;
    if (t.partials2[i] && g.n_partials2 < 1) return fail(PVAE_ERR_INVALID, "pvae_adamax_from_partials: bad second source of segment %d", i);
    blocks += static_cast<int>(g.columnar ? (g.length + 31) / 32 : (g.length / 4 + 32 * kDenseRows - 1) / (32 * kDenseRows));
  }
  t.first_block[n_segs] = blocks;
  const double bias_correction = 1.0 - pow(static_cast<double>(beta1), step);  // base/adamax.py:75-76
  const float clr = static_cast<float>(static_cast<double>(lr) / bias_correction);
  return cuda_ok("pvae_adamax_from_partials", launch_pdl(adamax_partials_kernel, dim3(blocks), dim3(32, 32), 0,
                                                         static_cast<cudaStream_t>(stream), t, p, g_out, exp_avg, exp_inf, clr, beta1,
                                                         beta2, eps, weight_decay, p_lo, dev.lr_sched, dev.n_sched, dev.state));
}

int pvae::advance_state(int* state, uint64_t philox_increment, float* rmax_step, float* rmax_ring, int ring_n, void* stream) {
  // plain launch (no programmatic overlap): it must not run before the kernels that read the counters have finished
  advance_state_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(state, static_cast<unsigned long long>(philox_increment), rmax_step,
                                                                        rmax_ring, ring_n);
  count_launch();
  return cuda_ok("advance_state_kernel", cudaGetLastError());
}

extern "C" int pvae_tf32_lo(const float* x, float* lo, int64_t n, void* stream) {
  if (n < 0 || n % 4 || (n > 0 && (!x || !lo))) return fail(PVAE_ERR_INVALID, "pvae_tf32_lo: bad argument (n must be a multiple of 4)");
  if (n == 0) return PVAE_OK;
  const unsigned grid = static_cast<unsigned>(std::min<long long>((n / 4 + 255) / 256, 148ll * 8));
  return cuda_ok("pvae_tf32_lo", launch_pdl(tf32_lo_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), x, lo,
                                            static_cast<long long>(n / 4)));
}

int pvae::adamax_step_dev_impl(float* p, float* p_lo, const float* g, float* exp_avg, float* exp_inf, int64_t n, const AdamaxDev& dev,
                               float beta1, float beta2, float eps, float weight_decay, const float* clip_coef, void* stream) {
  if (n <= 0 || !p || !g || !exp_avg || !exp_inf || !dev.lr_sched || dev.n_sched < 1 || !dev.state)
    return fail(PVAE_ERR_INVALID, "pvae_adamax_step_dev: bad argument");
  const unsigned grid = static_cast<unsigned>(std::min<long long>((n + 255) / 256, 148ll * 8));
  return cuda_ok("pvae_adamax_step_dev", launch_pdl(adamax_dev_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), p, g,
                                                    exp_avg, exp_inf, static_cast<long long>(n), dev.lr_sched, dev.n_sched, dev.state, beta1,
                                                    beta2, eps, weight_decay, clip_coef, p_lo));
}

extern "C" int pvae_adamax_step_dev(float* p, const float* g, float* exp_avg, float* exp_inf, int64_t n, const float* lr_sched,
                                    int n_sched, int* state, float beta1, float beta2, float eps, float weight_decay,
                                    const float* clip_coef, uint64_t philox_increment, float* rmax_step, float* rmax_ring,
                                    int ring_n, void* stream) {
  if (reinterpret_cast<uintptr_t>(state) & 7) return fail(PVAE_ERR_INVALID, "pvae_adamax_step_dev: state must be 8-byte aligned");
  AdamaxDev dev;
  dev.lr_sched = lr_sched, dev.n_sched = n_sched, dev.state = state;
  if (int rc = adamax_step_dev_impl(p, nullptr, g, exp_avg, exp_inf, n, dev, beta1, beta2, eps, weight_decay, clip_coef, stream)) return rc;
  return advance_state(state, philox_increment, rmax_step, rmax_ring, ring_n, stream);
}

extern "C" int pvae_gather(const void* const* srcs, const int64_t* offsets, int n_tensors, float* dst, void* stream) {
  if (!srcs || !offsets || !dst || n_tensors < 0) return fail(PVAE_ERR_INVALID, "pvae_gather: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int base = 0; base < n_tensors; base += kGatherMax) {
    GatherTable t;
    t.n = std::min(kGatherMax, n_tensors - base);
    long long biggest = 1;
    for (int i = 0; i < t.n; ++i) {
      t.src[i] = static_cast<const float*>(srcs[base + i]);
      t.off[i] = offsets[base + i];
      biggest = std::max<long long>(biggest, offsets[base + i + 1] - offsets[base + i]);
    }
    t.off[t.n] = offsets[base + t.n];
    const unsigned gx = static_cast<unsigned>(std::min<long long>((biggest + 2047) / 2048, 64));
    if (int rc = cuda_ok("gather_kernel", launch_pdl(gather_kernel, dim3(gx, t.n), dim3(256), 0, s, t, dst))) return rc;
  }
  return PVAE_OK;
}

extern "C" int pvae_recon_residual_parts(const float* y_parts, int n_parts, const float* x, float* g_y, float* recon, int64_t B,
                                         int64_t D, float g_scale, void* stream) {
  return recon_parts_impl(y_parts, n_parts, x, g_y, nullptr, recon, B, D, g_scale, stream);
}

int pvae::recon_parts_impl(const float* y_parts, int n_parts, const float* x, float* g_y, float* g_y_lo, float* recon, int64_t B, int64_t D,
                           float g_scale, void* stream) {
  if (B <= 0 || D <= 0 || D % 4 != 0 || n_parts < 1 || !y_parts || !x)
    return fail(PVAE_ERR_INVALID, "pvae_recon_residual_parts: bad argument");
  const unsigned grid = static_cast<unsigned>(std::min<long long>((B + 7) / 8, 148ll * 8));
  return cuda_ok("pvae_recon_residual_parts", launch_pdl(recon_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream),
                                                         y_parts, x, g_y, g_y_lo, recon, static_cast<long long>(B),
                                                         static_cast<long long>(D), g_scale, n_parts));
}

extern "C" int pvae_recon_residual(const float* y, const float* x, float* g_y, float* recon, int64_t B, int64_t D,
                                   float g_scale, void* stream) {
  if (B < 0 || D <= 0 || D % 4 != 0) return fail(PVAE_ERR_INVALID, "pvae_recon_residual: bad shape B=%lld D=%lld", (long long)B, (long long)D);
  if (B == 0) return PVAE_OK;
  if (!y || !x) return fail(PVAE_ERR_INVALID, "pvae_recon_residual: null y/x");
  const unsigned grid = static_cast<unsigned>(std::min<long long>((B + 7) / 8, 148ll * 8));
  return cuda_ok("pvae_recon_residual", launch_pdl(recon_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), y, x,
                                                   g_y, static_cast<float*>(nullptr), recon, static_cast<long long>(B),
                                                   static_cast<long long>(D), g_scale, 1));
}

extern "C" int pvae_loss_assemble(const float* recon, const float* kl_rowsum, float* out3, int64_t B, float beta,
                                  int64_t B_global, int recon_parts, void* stream) {
  if (B <= 0 || B_global <= 0 || recon_parts < 1 || !recon || !kl_rowsum || !out3)
    return fail(PVAE_ERR_INVALID, "pvae_loss_assemble: bad argument");
  return cuda_ok("pvae_loss_assemble", launch_pdl(loss_kernel, dim3(1), dim3(1024), 0, static_cast<cudaStream_t>(stream), recon,
                                                  kl_rowsum, out3, static_cast<long long>(B), beta,
                                                  1.f / static_cast<float>(B_global), recon_parts));
}

extern "C" int64_t pvae_grad_norm_ws_floats(void) { return kNormBlocks; }

extern "C" int pvae_grad_norm(const float* g, int64_t n, float max_norm, float* ws, float* out2, void* stream) {
  if (n <= 0 || !g || !ws || !out2) return fail(PVAE_ERR_INVALID, "pvae_grad_norm: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = cuda_ok("pvae_grad_norm", launch_pdl(sumsq_partial_kernel, dim3(kNormBlocks), dim3(256), 0, s, g,
                                                    static_cast<long long>(n), ws)))
    return rc;
  return cuda_ok("pvae_grad_norm", launch_pdl(norm_finalize_kernel, dim3(1), dim3(32), 0, s, ws, kNormBlocks, max_norm, out2));
}

extern "C" int pvae_adamax_step(float* p, const float* g, float* exp_avg, float* exp_inf, int64_t n, float lr, int step,
                                float beta1, float beta2, float eps, float weight_decay, const float* clip_coef,
                                void* stream) {
  return adamax_step_impl(p, nullptr, g, exp_avg, exp_inf, n, lr, step, beta1, beta2, eps, weight_decay, clip_coef, stream);
}

int pvae::adamax_step_impl(float* p, float* p_lo, const float* g, float* exp_avg, float* exp_inf, int64_t n, float lr, int step, float beta1,
                           float beta2, float eps, float weight_decay, const float* clip_coef, void* stream) {
  if (n <= 0 || !p || !g || !exp_avg || !exp_inf || step < 1) return fail(PVAE_ERR_INVALID, "pvae_adamax_step: bad argument");
  const double bias_correction = 1.0 - pow(static_cast<double>(beta1), step);  // base/adamax.py:75-76
  const float clr = static_cast<float>(static_cast<double>(lr) / bias_correction);
  const unsigned grid = static_cast<unsigned>(std::min<long long>((n + 255) / 256, 148ll * 8));
  return cuda_ok("pvae_adamax_step", launch_pdl(adamax_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), p, g,
                                                exp_avg, exp_inf, static_cast<long long>(n), clr, beta1, beta2, eps,
                                                weight_decay, clip_coef, p_lo));
}
