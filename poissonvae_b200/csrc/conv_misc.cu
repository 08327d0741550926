// Memory-bound pieces of the conv encoder / decoder around the implicit-GEMM convolutions (sm_100a):
// single-channel stem convolution, SiLU, squeeze-excitation gate + residual, pooling, bias / ReLU / dropout,
// LayerNorm, batch-column sums.  NHWC fp32, 128-bit accesses, fixed-order (deterministic) reductions.
//
// Reference semantics: stem main/vae.py:211-218; SELayer base/common.py:129-142; Cell.forward :187-192;
// ResConvPool main/vae.py:682-706; ResDenseLayer :657-679 (nn.LayerNorm eps 1e-5, Dropout 0.1).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pvae_b200.h"
#include "pvae_device.cuh"
#include "pvae_host.h"

namespace pvae {

__device__ __forceinline__ float sigmoid_m(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float silu_m(float x) { return x * sigmoid_m(x); }
__device__ __forceinline__ float dsilu_m(float x) {
  const float s = sigmoid_m(x);
  return s * (1.f + x * (1.f - s));
}

// ---- SiLU ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) silu_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, long long n4) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += gridDim.x * 256ll) {
    const float4 v = ldg4(x + 4 * i);
    stg4(out + 4 * i, make_float4(silu_m(v.x), silu_m(v.y), silu_m(v.z), silu_m(v.w)));
  }
}

// out = (a (+ b)) optionally: dst may alias a
__global__ void __launch_bounds__(256) add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                                  long long n4) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += gridDim.x * 256ll) {
    const float4 u = ldg4(a + 4 * i), v = ldg4(b + 4 * i);
    stg4(out + 4 * i, make_float4(u.x + v.x, u.y + v.y, u.z + v.z, u.w + v.w));
  }
}

// ---- stem: 3x3 convolution of a single-channel image, weight-normalised, to Co channels -----------------
__global__ void __launch_bounds__(256) stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ weight,
                                                       const float* __restrict__ lognorm, const float* __restrict__ bias,
                                                       float* __restrict__ out, float* __restrict__ out_act, int N, int IH, int IW,
                                                       int OH, int OW, int Co, int pad) {
  extern __shared__ float s_w[];  // [9][Co] normalised, then [Co] bias
  pdl_launch_dependents();
  pdl_wait();
  for (int co = threadIdx.x; co < Co; co += 256) {
    float ss = 0.f;
    for (int t = 0; t < 9; ++t) ss = fmaf(weight[co * 9 + t], weight[co * 9 + t], ss);
    const float scale = expf(lognorm[co]) / (sqrtf(ss) + kF32Eps);
    for (int t = 0; t < 9; ++t) s_w[t * Co + co] = scale * weight[co * 9 + t];
    s_w[9 * Co + co] = bias != nullptr ? bias[co] : 0.f;
  }
  __syncthreads();
  const int c4n = Co >> 2;
  const long long total = static_cast<long long>(N) * OH * OW * c4n;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int c4 = static_cast<int>(i % c4n);
    const long long pix = i / c4n;
    const int ow = static_cast<int>(pix % OW), oh = static_cast<int>((pix / OW) % OH);
    const long long n = pix / (static_cast<long long>(OW) * OH);
    float4 acc = *reinterpret_cast<const float4*>(s_w + 9 * Co + 4 * c4);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int ih = oh + t / 3 - pad, iw = ow + t % 3 - pad;
      if (ih >= 0 && ih < IH && iw >= 0 && iw < IW) {
        const float xv = __ldg(x + (n * IH + ih) * IW + iw);
        const float4 w = *reinterpret_cast<const float4*>(s_w + t * Co + 4 * c4);
        acc.x = fmaf(xv, w.x, acc.x), acc.y = fmaf(xv, w.y, acc.y), acc.z = fmaf(xv, w.z, acc.z), acc.w = fmaf(xv, w.w, acc.w);
      }
    }
    stg4(out + pix * Co + 4 * c4, acc);
    if (out_act != nullptr) stg4(out_act + pix * Co + 4 * c4, make_float4(silu_m(acc.x), silu_m(acc.y), silu_m(acc.z), silu_m(acc.w)));
  }
}

// partial[b][t][co] = sum over block b's pixels of g[pix, co] * x[pix (+) tap t]   (t = 9: bias gradient)
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                         float* __restrict__ partial, int N, int IH, int IW, int OH, int OW, int Co,
                                                         int pad, long long pix_per_block) {
  extern __shared__ float s_red[];  // [groups][10][Co]
  pdl_launch_dependents();
  pdl_wait();
  const int groups = 256 / Co;  // Co in {16, 32, 64, 128, 256}
  const int co = threadIdx.x % Co, grp = threadIdx.x / Co;
  const long long total = static_cast<long long>(N) * OH * OW;
  const long long p0 = blockIdx.x * pix_per_block, p1 = min(total, p0 + pix_per_block);
  float acc[10];
#pragma unroll
  for (int t = 0; t < 10; ++t) acc[t] = 0.f;
  for (long long pix = p0 + grp; pix < p1; pix += groups) {
    const int ow = static_cast<int>(pix % OW), oh = static_cast<int>((pix / OW) % OH);
    const long long n = pix / (static_cast<long long>(OW) * OH);
    const float gv = __ldg(g + pix * Co + co);
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int ih = oh + t / 3 - pad, iw = ow + t % 3 - pad;
      if (ih >= 0 && ih < IH && iw >= 0 && iw < IW) acc[t] = fmaf(gv, __ldg(x + (n * IH + ih) * IW + iw), acc[t]);
    }
    acc[9] += gv;
  }
#pragma unroll
  for (int t = 0; t < 10; ++t) s_red[(grp * 10 + t) * Co + co] = acc[t];
  __syncthreads();
  for (int i = threadIdx.x; i < 10 * Co; i += 256) {
    float s = 0.f;
    for (int gI = 0; gI < groups; ++gI) s += s_red[gI * 10 * Co + i];
    partial[static_cast<long long>(blockIdx.x) * 10 * Co + i] = s;
  }
}

// out[i] = sum_p part[p * n + i]: 32 outputs x 8 slices of p per block, slices combined in a fixed order
__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ part, float* __restrict__ out, int P, long long n,
                                                           float scale) {
  __shared__ float s_red[8][33];
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
  for (long long base = blockIdx.x * 32ll; base < n; base += gridDim.x * 32ll) {
    const long long i = base + lane;
    float s = 0.f;
    if (i < n) {
      int p = q;
      for (; p + 24 < P; p += 32) {  // four independent loads in flight
        const float v0 = part[p * n + i], v1 = part[(p + 8) * n + i], v2 = part[(p + 16) * n + i], v3 = part[(p + 24) * n + i];
        s += (v0 + v1) + (v2 + v3);
      }
      for (; p < P; p += 8) s += part[p * n + i];
    }
    s_red[q][lane] = s;
    __syncthreads();
    if (q == 0 && i < n) {
      float t = s_red[0][lane];
#pragma unroll
      for (int k = 1; k < 8; ++k) t += s_red[k][lane];
      out[i] = t * scale;
    }
    __syncthreads();
  }
}

// ---- column sums of a (R, C) matrix: part[b][c] over block b's rows ------------------------------------
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ x, float* __restrict__ part, long long R, int C,
                                                             long long rows_per_block) {
  __shared__ float4 s_red[256];
  pdl_launch_dependents();
  pdl_wait();
  const int c4n = C >> 2;
  const long long r0 = blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
  for (int cbase = 0; cbase < c4n; cbase += 256) {
    const int cols = min(256, c4n - cbase);
    const int phases = 256 / cols;
    const int c4 = cbase + threadIdx.x % cols, ph = threadIdx.x / cols;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ph < phases) {
      long long r = r0 + ph;
      for (; r + 3ll * phases < r1; r += 4ll * phases) {  // four loads in flight per thread
        const float4 v0 = ldg4(x + r * C + 4 * c4), v1 = ldg4(x + (r + phases) * C + 4 * c4),
                     v2 = ldg4(x + (r + 2ll * phases) * C + 4 * c4), v3 = ldg4(x + (r + 3ll * phases) * C + 4 * c4);
        a.x += (v0.x + v1.x) + (v2.x + v3.x), a.y += (v0.y + v1.y) + (v2.y + v3.y);
        a.z += (v0.z + v1.z) + (v2.z + v3.z), a.w += (v0.w + v1.w) + (v2.w + v3.w);
      }
      for (; r < r1; r += phases) {
        const float4 v = ldg4(x + r * C + 4 * c4);
        a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
      }
    }
    s_red[threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.x < cols) {
      float4 s = s_red[threadIdx.x];
      for (int q = 1; q < phases; ++q) {
        const float4 v = s_red[q * cols + threadIdx.x];
        s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
      }
      stg4(part + static_cast<long long>(blockIdx.x) * C + 4 * (cbase + threadIdx.x), s);
    }
    __syncthreads();
  }
}

// out (Ca, Cb) partials of sum_n a[n, i] b[n, j] over a block of rows
__global__ void __launch_bounds__(256) small_tn_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                               float* __restrict__ part, long long R, int Ca, int Cb,
                                                               long long rows_per_block) {
  pdl_launch_dependents();
  pdl_wait();
  const long long r0 = blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
  for (int e = threadIdx.x; e < Ca * Cb; e += 256) {
    const int i = e / Cb, j = e % Cb;
    float s = 0.f;
    long long r = r0;
    for (; r + 4 <= r1; r += 4) {  // eight independent loads in flight
      const float a0 = __ldg(a + r * Ca + i), a1 = __ldg(a + (r + 1) * Ca + i), a2 = __ldg(a + (r + 2) * Ca + i),
                  a3 = __ldg(a + (r + 3) * Ca + i);
      const float b0 = __ldg(b + r * Cb + j), b1 = __ldg(b + (r + 1) * Cb + j), b2 = __ldg(b + (r + 2) * Cb + j),
                  b3 = __ldg(b + (r + 3) * Cb + j);
      s = fmaf(a0, b0, s), s = fmaf(a1, b1, s), s = fmaf(a2, b2, s), s = fmaf(a3, b3, s);
    }
    for (; r < r1; ++r) s = fmaf(__ldg(a + r * Ca + i), __ldg(b + r * Cb + j), s);
    part[static_cast<long long>(blockIdx.x) * Ca * Cb + e] = s;
  }
}

// ---- per-image channel means (shared by the SE gate and the pooling skip) ------------------------------
// returns in s_out[c] (shared) the sum over HW of v1[p, c] * (v2 ? v2[p, c] : 1) for image n; blockDim = 256
__device__ __forceinline__ void image_channel_sums(const float* __restrict__ v1, const float* __restrict__ v2, int HW, int C,
                                                   float* s_red /* 1024 floats */, float* s_out /* C floats */) {
  const int c4n = C >> 2;  // <= 64
  const int phases = 256 / c4n;
  const int c4 = threadIdx.x % c4n, ph = threadIdx.x / c4n;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ph < phases)
    for (int p = ph; p < HW; p += phases) {
      const float4 u = ldg4(v1 + static_cast<long long>(p) * C + 4 * c4);
      if (v2 != nullptr) {
        const float4 w = ldg4(v2 + static_cast<long long>(p) * C + 4 * c4);
        a.x = fmaf(u.x, w.x, a.x), a.y = fmaf(u.y, w.y, a.y), a.z = fmaf(u.z, w.z, a.z), a.w = fmaf(u.w, w.w, a.w);
      } else {
        a.x += u.x, a.y += u.y, a.z += u.z, a.w += u.w;
      }
    }
  reinterpret_cast<float4*>(s_red)[threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x < c4n) {
    float4 s = reinterpret_cast<float4*>(s_red)[threadIdx.x];
    for (int q = 1; q < phases; ++q) {
      const float4 v = reinterpret_cast<float4*>(s_red)[q * c4n + threadIdx.x];
      s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
    }
    reinterpret_cast<float4*>(s_out)[threadIdx.x] = s;
  }
  __syncthreads();
}

// ---- squeeze-excitation gate + residual: y = skip + eps * c * sigmoid(W2 relu(W1 mean_hw(c) + b1) + b2) -----
// C = channels in memory (multiple of 32), Cr <= C = channels of the parameters (the rest is zero padding).
// skip_up: skip is a (N, H/2, W/2, C) map read with nearest-neighbour 2x up-sampling (decoder 'up' cells).
__global__ void __launch_bounds__(256) se_fwd_kernel(const float* __restrict__ c, const float* __restrict__ skip,
                                                     const float* __restrict__ W1, const float* __restrict__ b1,
                                                     const float* __restrict__ W2, const float* __restrict__ b2, float eps_res,
                                                     float* __restrict__ y, float* __restrict__ y_act, float* __restrict__ pool,
                                                     float* __restrict__ hid, float* __restrict__ gate, int H, int W, int C, int Cr,
                                                     int hd, int skip_up) {
  __shared__ __align__(16) float s_red[1024];
  __shared__ __align__(16) float s_pool[256];
  __shared__ float s_hid[32];
  __shared__ __align__(16) float s_gate[256];
  pdl_launch_dependents();
  pdl_wait();
  const long long n = blockIdx.x;
  const int HW = H * W;
  const float* cn = c + n * HW * C;
  image_channel_sums(cn, nullptr, HW, C, s_red, s_pool);
  if (threadIdx.x < C) {
    s_pool[threadIdx.x] *= 1.f / static_cast<float>(HW);
    if (threadIdx.x < Cr) pool[n * Cr + threadIdx.x] = s_pool[threadIdx.x];
  }
  __syncthreads();
  if (threadIdx.x < hd) {
    float a = b1[threadIdx.x];
    for (int k = 0; k < Cr; ++k) a = fmaf(W1[threadIdx.x * Cr + k], s_pool[k], a);
    a = fmaxf(a, 0.f);
    s_hid[threadIdx.x] = a;
    hid[n * hd + threadIdx.x] = a;
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float a = 0.f;
    if (threadIdx.x < Cr) {
      a = b2[threadIdx.x];
      for (int j = 0; j < hd; ++j) a = fmaf(W2[threadIdx.x * hd + j], s_hid[j], a);
      a = sigmoid_m(a);
      gate[n * Cr + threadIdx.x] = a;
    }
    s_gate[threadIdx.x] = a;
  }
  __syncthreads();
  const int c4n = C >> 2;
  for (int i = threadIdx.x; i < HW * c4n; i += 256) {
    const int c4 = i % c4n;
    const long long off = n * HW * C + 4ll * i;
    long long soff = off;
    if (skip_up) {
      const int pix = i / c4n, h = pix / W, w = pix % W;
      soff = ((n * (H >> 1) + (h >> 1)) * (W >> 1) + (w >> 1)) * C + 4ll * c4;
    }
    const float4 cv = ldg4(c + off), sv = ldg4(skip + soff);
    const float4 gt = reinterpret_cast<const float4*>(s_gate)[c4];
    const float4 r = make_float4(fmaf(eps_res * cv.x, gt.x, sv.x), fmaf(eps_res * cv.y, gt.y, sv.y),
                                 fmaf(eps_res * cv.z, gt.z, sv.z), fmaf(eps_res * cv.w, gt.w, sv.w));
    stg4(y + off, r);
    if (y_act != nullptr) stg4(y_act + off, make_float4(silu_m(r.x), silu_m(r.y), silu_m(r.z), silu_m(r.w)));
  }
}

// backward: g_c = eps g_y gate + g_pool / HW; d2 = d loss / d (pre-sigmoid), d1 = d loss / d (pre-ReLU) per image
__global__ void __launch_bounds__(256) se_bwd_kernel(const float* __restrict__ g_y, const float* __restrict__ c,
                                                     const float* __restrict__ gate, const float* __restrict__ hid,
                                                     const float* __restrict__ W1, const float* __restrict__ W2, float eps_res,
                                                     float* __restrict__ g_c, float* __restrict__ d2, float* __restrict__ d1, int HW,
                                                     int C, int Cr, int hd) {
  __shared__ __align__(16) float s_red[1024];
  __shared__ __align__(16) float s_gs[256];
  __shared__ float s_d1[32];
  __shared__ __align__(16) float s_gpool[256];
  __shared__ __align__(16) float s_gate[256];
  pdl_launch_dependents();
  pdl_wait();
  const long long n = blockIdx.x;
  image_channel_sums(g_y + n * HW * C, c + n * HW * C, HW, C, s_red, s_gs);
  if (threadIdx.x < C) {
    float v = 0.f, gt = 0.f;
    if (threadIdx.x < Cr) {
      gt = gate[n * Cr + threadIdx.x];
      v = eps_res * s_gs[threadIdx.x] * gt * (1.f - gt);
      d2[n * Cr + threadIdx.x] = v;
    }
    s_gs[threadIdx.x] = v;
    s_gate[threadIdx.x] = gt;
  }
  __syncthreads();
  if (threadIdx.x < hd) {
    float a = 0.f;
    for (int k = 0; k < Cr; ++k) a = fmaf(W2[k * hd + threadIdx.x], s_gs[k], a);
    a = hid[n * hd + threadIdx.x] > 0.f ? a : 0.f;
    s_d1[threadIdx.x] = a;
    d1[n * hd + threadIdx.x] = a;
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float a = 0.f;
    if (threadIdx.x < Cr)
      for (int j = 0; j < hd; ++j) a = fmaf(W1[j * Cr + threadIdx.x], s_d1[j], a);
    s_gpool[threadIdx.x] = a / static_cast<float>(HW);
  }
  __syncthreads();
  const int c4n = C >> 2;
  for (int i = threadIdx.x; i < HW * c4n; i += 256) {
    const int c4 = i % c4n;
    const long long off = n * HW * C + 4ll * i;
    const float4 gv = ldg4(g_y + off);
    const float4 gt = reinterpret_cast<const float4*>(s_gate)[c4];
    const float4 gp = reinterpret_cast<const float4*>(s_gpool)[c4];
    stg4(g_c + off, make_float4(fmaf(eps_res * gv.x, gt.x, gp.x), fmaf(eps_res * gv.y, gt.y, gp.y),
                                fmaf(eps_res * gv.z, gt.z, gp.z), fmaf(eps_res * gv.w, gt.w, gp.w)));
  }
}

// ---- pooling skip of ResConvPool: pool[n, c] = mean_hw x[n, :, c] ---------------------------------------
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ pool, int HW, int C) {
  __shared__ __align__(16) float s_red[1024];
  __shared__ __align__(16) float s_out[256];
  pdl_launch_dependents();
  pdl_wait();
  const long long n = blockIdx.x;
  image_channel_sums(x + n * HW * C, nullptr, HW, C, s_red, s_out);
  if (threadIdx.x < C) pool[n * C + threadIdx.x] = s_out[threadIdx.x] / static_cast<float>(HW);
}

// g_x[n, p, c] = g_a[n, p, c] * silu'(x[n, p, c]) + g_pool[n, c] / HW
__global__ void __launch_bounds__(256) pool_silu_bwd_kernel(const float* __restrict__ g_a, const float* __restrict__ x,
                                                            const float* __restrict__ g_pool, float* __restrict__ g_x, long long N,
                                                            int HW, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int c4n = C >> 2;
  const float inv = 1.f / static_cast<float>(HW);
  const long long total = N * HW * c4n;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int c4 = static_cast<int>(i % c4n);
    const long long n = i / (static_cast<long long>(HW) * c4n);
    const float4 ga = ldg4(g_a + 4 * i), xv = ldg4(x + 4 * i), gp = ldg4(g_pool + n * C + 4 * c4);
    stg4(g_x + 4 * i, make_float4(fmaf(ga.x, dsilu_m(xv.x), gp.x * inv), fmaf(ga.y, dsilu_m(xv.y), gp.y * inv),
                                  fmaf(ga.z, dsilu_m(xv.z), gp.z * inv), fmaf(ga.w, dsilu_m(xv.w), gp.w * inv)));
  }
}

// ---- y = act(y + bias (+ add)) * dropout, in place on a (R, C) matrix -----------------------------------
__global__ void __launch_bounds__(256) bias_act_kernel(float* __restrict__ y, const float* __restrict__ bias,
                                                       const float* __restrict__ add, long long R, int C, int relu, float drop_p,
                                                       uint64_t seed, uint64_t offset) {
  pdl_launch_dependents();
  pdl_wait();
  const int c4n = C >> 2;
  PhiloxKey ks;
  if (drop_p > 0.f) philox_key_schedule(seed, ks);
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < R * c4n; i += gridDim.x * 256ll) {
    const int c4 = static_cast<int>(i % c4n);
    float4 v = ldg4(y + 4 * i);
    if (bias != nullptr) {
      const float4 b = ldg4(bias + 4 * c4);
      v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
    }
    if (add != nullptr) {
      const float4 a = ldg4(add + 4 * i);
      v.x += a.x, v.y += a.y, v.z += a.z, v.w += a.w;
    }
    if (relu) v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
    if (drop_p > 0.f) {
      const uint4 bits = philox4x32_10(static_cast<uint32_t>(i), static_cast<uint32_t>(i >> 32), static_cast<uint32_t>(offset),
                                       static_cast<uint32_t>(offset >> 32), ks);
      v.x = uniform_from_bits(bits.x) >= drop_p ? v.x * keep_scale : 0.f;
      v.y = uniform_from_bits(bits.y) >= drop_p ? v.y * keep_scale : 0.f;
      v.z = uniform_from_bits(bits.z) >= drop_p ? v.z * keep_scale : 0.f;
      v.w = uniform_from_bits(bits.w) >= drop_p ? v.w * keep_scale : 0.f;
    }
    stg4(y + 4 * i, v);
  }
}

// g = (y_saved > 0) ? g * keep_scale : 0   (y_saved = relu(.) * dropout mask / (1 - p): positive iff active and kept)
__global__ void __launch_bounds__(256) relu_bwd_kernel(float* __restrict__ g, const float* __restrict__ y_saved, long long n4,
                                                       float keep_scale) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += gridDim.x * 256ll) {
    const float4 gv = ldg4(g + 4 * i), yv = ldg4(y_saved + 4 * i);
    stg4(g + 4 * i, make_float4(yv.x > 0.f ? gv.x * keep_scale : 0.f, yv.y > 0.f ? gv.y * keep_scale : 0.f,
                                yv.z > 0.f ? gv.z * keep_scale : 0.f, yv.w > 0.f ? gv.w * keep_scale : 0.f));
  }
}

// ---- LayerNorm over the last dimension of s = a + b (nn.LayerNorm, eps 1e-5, biased variance) ----------
// one warp per row; C a multiple of 32, C <= 256
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ out, float* __restrict__ xhat, float* __restrict__ rstd,
                                                            long long R, int C, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int per = C >> 5;
  const long long warps = (static_cast<long long>(gridDim.x) * 256) >> 5;
  for (long long r = (blockIdx.x * 256ll + threadIdx.x) >> 5; r < R; r += warps) {
    float v[8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = 0.f;
      if (j < per) v[j] = a[r * C + lane + 32 * j] + (b != nullptr ? b[r * C + lane + 32 * j] : 0.f);
      s += v[j];
    }
    const float mean = warp_sum(s) / static_cast<float>(C);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < per) q = fmaf(v[j] - mean, v[j] - mean, q);
    const float rs = 1.f / sqrtf(warp_sum(q) / static_cast<float>(C) + eps);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j >= per) break;
      const int cidx = lane + 32 * j;
      const float xh = (v[j] - mean) * rs;
      xhat[r * C + cidx] = xh;
      out[r * C + cidx] = fmaf(xh, gamma[cidx], beta[cidx]);
    }
    if (lane == 0) rstd[r] = rs;
  }
}

// g_s = rstd (g gamma - mean(g gamma) - xhat mean(g gamma xhat)); part[b] = [sum g xhat (C) | sum g (C)] over block b's rows
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ xhat,
                                                            const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                            float* __restrict__ g_s, float* __restrict__ part, long long R, int C,
                                                            long long rows_per_block) {
  extern __shared__ float s_acc[];  // [8 warps][2][C]
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per = C >> 5;
  const long long r0 = blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
  float dg[8], db[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) dg[j] = 0.f, db[j] = 0.f;
  for (long long r = r0 + warp; r < r1; r += 8) {
    float gv[8], xh[8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j >= per) break;
      const int cidx = lane + 32 * j;
      gv[j] = g[r * C + cidx];
      xh[j] = xhat[r * C + cidx];
      const float gg = gv[j] * gamma[cidx];
      s1 += gg;
      s2 = fmaf(gg, xh[j], s2);
      dg[j] = fmaf(gv[j], xh[j], dg[j]);
      db[j] += gv[j];
    }
    s1 = warp_sum(s1) / static_cast<float>(C);
    s2 = warp_sum(s2) / static_cast<float>(C);
    const float rs = rstd[r];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j >= per) break;
      const int cidx = lane + 32 * j;
      g_s[r * C + cidx] = rs * (gv[j] * gamma[cidx] - s1 - xh[j] * s2);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (j >= per) break;
    s_acc[(warp * 2 + 0) * C + lane + 32 * j] = dg[j];
    s_acc[(warp * 2 + 1) * C + lane + 32 * j] = db[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += 256) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += s_acc[w * 2 * C + i];
    part[static_cast<long long>(blockIdx.x) * 2 * C + i] = s;
  }
}

// ---- nearest-neighbour resampling (nn.Upsample(mode='nearest'): scale_factor 2 in the decoder's 'up' cells,
// size=14 from 16x16 for MNIST, main/vae.py:779-780).  Source index as ATen computes it: min(floorf(dst * scale), in - 1),
// scale = (float) in / out.
__device__ __forceinline__ int nearest_src(int dst, float scale, int in) { return min(static_cast<int>(floorf(dst * scale)), in - 1); }

__global__ void __launch_bounds__(256) upsample_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, long long N, int IH,
                                                           int IW, int OH, int OW, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int c4n = C >> 2;
  const float sh = static_cast<float>(IH) / static_cast<float>(OH), sw = static_cast<float>(IW) / static_cast<float>(OW);
  const long long total = N * OH * OW * c4n;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int c4 = static_cast<int>(i % c4n);
    const long long pix = i / c4n;
    const int ow = static_cast<int>(pix % OW), oh = static_cast<int>((pix / OW) % OH);
    const long long n = pix / (static_cast<long long>(OW) * OH);
    const int ih = nearest_src(oh, sh, IH), iw = nearest_src(ow, sw, IW);
    stg4(out + 4 * i, ldg4(x + ((n * IH + ih) * IW + iw) * C + 4 * c4));
  }
}

// g_in[n, ih, iw, :] = (sum of g_out over the output pixels that read (ih, iw)) * silu'(dsilu_src) : a gather, no atomics
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const float* __restrict__ g_out, const float* __restrict__ dsilu_src,
                                                           float* __restrict__ g_in, long long N, int IH, int IW, int OH, int OW,
                                                           int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int c4n = C >> 2;
  const float sh = static_cast<float>(IH) / static_cast<float>(OH), sw = static_cast<float>(IW) / static_cast<float>(OW);
  const long long total = N * IH * IW * c4n;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int c4 = static_cast<int>(i % c4n);
    const long long pix = i / c4n;
    const int iw = static_cast<int>(pix % IW), ih = static_cast<int>((pix / IW) % IH);
    const long long n = pix / (static_cast<long long>(IW) * IH);
    // candidate output range: a little wider than ih / scale, filtered by the exact forward map
    const int oh_lo = max(0, static_cast<int>(floorf(ih / sh)) - 1), oh_hi = min(OH - 1, static_cast<int>(floorf((ih + 1) / sh)) + 1);
    const int ow_lo = max(0, static_cast<int>(floorf(iw / sw)) - 1), ow_hi = min(OW - 1, static_cast<int>(floorf((iw + 1) / sw)) + 1);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      if (nearest_src(oh, sh, IH) != ih) continue;
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        if (nearest_src(ow, sw, IW) != iw) continue;
        const float4 g = ldg4(g_out + ((n * OH + oh) * OW + ow) * C + 4 * c4);
        a.x += g.x, a.y += g.y, a.z += g.z, a.w += g.w;
      }
    }
    if (dsilu_src != nullptr) {
      const float4 xv = ldg4(dsilu_src + 4 * i);
      a.x *= dsilu_m(xv.x), a.y *= dsilu_m(xv.y), a.z *= dsilu_m(xv.z), a.w *= dsilu_m(xv.w);
    }
    stg4(g_in + 4 * i, a);
  }
}

// ---- (N, C, P) <-> (N, P, C): fc_dec's output viewed as (N, C, 2, 2) NCHW (main/vae.py:56-57) into NHWC, + bias -----
__global__ void __launch_bounds__(256) permute_cp_kernel(const float* __restrict__ in, const float* __restrict__ bias,
                                                         float* __restrict__ out, float* __restrict__ out_act, long long N, int C,
                                                         int P, int to_nhwc) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = N * C * P;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    // i indexes the OUTPUT (coalesced stores)
    const long long n = i / (static_cast<long long>(C) * P);
    const int r = static_cast<int>(i % (static_cast<long long>(C) * P));
    int c, p2;
    if (to_nhwc) p2 = r / C, c = r % C; else c = r / P, p2 = r % P;
    const int src = to_nhwc ? c * P + p2 : p2 * C + c;
    float v = __ldg(in + n * C * P + src);
    if (bias != nullptr) v += __ldg(bias + (to_nhwc ? src : r));
    out[i] = v;
    if (out_act != nullptr) out_act[i] = silu_m(v);
  }
}

// ---- decoder head: weight-normalised 1x1 convolution to ONE channel (main/vae.py:782-789) -------------------
// y[pix] = sum_{c < Cr} x[pix, c] wn[c] + b; x has C >= Cr channels in memory
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ x, const float* __restrict__ weight,
                                                       const float* __restrict__ lognorm, const float* __restrict__ bias,
                                                       float* __restrict__ y, long long R, int C, int Cr) {
  __shared__ float s_w[256];
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x < 32) {
    float ss = 0.f;
    for (int k = threadIdx.x; k < Cr; k += 32) ss = fmaf(weight[k], weight[k], ss);
    ss = warp_sum(ss);
    const float scale = expf(lognorm[0]) / (sqrtf(ss) + kF32Eps);
    for (int k = threadIdx.x; k < C; k += 32) s_w[k] = k < Cr ? scale * weight[k] : 0.f;
  }
  __syncthreads();
  const float b = bias != nullptr ? bias[0] : 0.f;
  const int c4n = Cr >> 2;
  for (long long r = blockIdx.x * 256ll + threadIdx.x; r < R; r += gridDim.x * 256ll) {
    float a = b;
    for (int k = 0; k < c4n; ++k) {
      const float4 v = ldg4(x + r * C + 4 * k);
      a = fmaf(v.x, s_w[4 * k], a), a = fmaf(v.y, s_w[4 * k + 1], a), a = fmaf(v.z, s_w[4 * k + 2], a), a = fmaf(v.w, s_w[4 * k + 3], a);
    }
    y[r] = a;
  }
}

// g_x[pix, c] = g_y[pix] wn[c]; part[b][c] = sum over block b's pixels of g_y x[pix, c] (c < C), part[b][C] = sum g_y
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g_y,
                                                       const float* __restrict__ weight, const float* __restrict__ lognorm,
                                                       float* __restrict__ g_x, float* __restrict__ part, long long R, int C, int Cr,
                                                       long long rows_per_block) {
  __shared__ float s_w[256];
  __shared__ float s_red[8][260];
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x < 32) {
    float ss = 0.f;
    for (int k = threadIdx.x; k < Cr; k += 32) ss = fmaf(weight[k], weight[k], ss);
    ss = warp_sum(ss);
    const float scale = expf(lognorm[0]) / (sqrtf(ss) + kF32Eps);
    for (int k = threadIdx.x; k < C; k += 32) s_w[k] = k < Cr ? scale * weight[k] : 0.f;
  }
  __syncthreads();
  // C <= 256 a multiple of 32: thread -> channel c = t % C, row phase t / C
  const int c = threadIdx.x % C, ph = threadIdx.x / C, phases = 256 / C;
  const long long r0 = blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
  float acc = 0.f, accb = 0.f;
  for (long long r = r0 + ph; r < r1; r += phases) {
    const float g = __ldg(g_y + r);
    acc = fmaf(g, __ldg(x + r * C + c), acc);
    accb += g;
    g_x[r * C + c] = g * s_w[c];
  }
  s_red[ph][c] = acc;
  if (c == 0) s_red[ph][256] = accb;
  __syncthreads();
  if (threadIdx.x <= C) {
    const int k = threadIdx.x == C ? 256 : threadIdx.x;
    float t = 0.f;
    for (int q = 0; q < phases; ++q) t += s_red[q][k];
    part[static_cast<long long>(blockIdx.x) * (C + 1) + threadIdx.x] = t;
  }
}

// ---- squeeze-excitation parameter gradients in ONE launch ------------------------------------------------------------
// g_w2 (Cr, hd) = d2^T hid, g_b2 (Cr) = colsum d2, g_w1 (hd, Cr) = d1^T pool, g_b1 (hd) = colsum d1 over the N images.
// Stage 1: every block sums its chunk of images into ws[block][n_out]; stage 2: the block that finishes last (ticket
// counter) adds the partials in block order - deterministic, and one launch instead of eight.
__device__ unsigned int g_se_ticket = 0;
__global__ void __launch_bounds__(256) se_param_grads_kernel(const float* __restrict__ d2, const float* __restrict__ hid,
                                                             const float* __restrict__ d1, const float* __restrict__ pool,
                                                             float* __restrict__ g_w2, float* __restrict__ g_b2,
                                                             float* __restrict__ g_w1, float* __restrict__ g_b1,
                                                             float* __restrict__ ws, int N, int Cr, int hd, int rows_per_block) {
  __shared__ bool s_last;
  pdl_launch_dependents();
  pdl_wait();
  const int n_w = Cr * hd, n_out = 2 * n_w + Cr + hd;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(N, r0 + rows_per_block);
  float* mine = ws + static_cast<long long>(blockIdx.x) * n_out;
  for (int e = threadIdx.x; e < n_out; e += 256) {
    float acc = 0.f;
    if (e < n_w) {  // g_w2[c][j]
      const int c = e / hd, j = e % hd;
      for (int r = r0; r < r1; ++r) acc = fmaf(__ldg(d2 + r * Cr + c), __ldg(hid + r * hd + j), acc);
    } else if (e < 2 * n_w) {  // g_w1[j][c]
      const int j = (e - n_w) / Cr, c = (e - n_w) % Cr;
      for (int r = r0; r < r1; ++r) acc = fmaf(__ldg(d1 + r * hd + j), __ldg(pool + r * Cr + c), acc);
    } else if (e < 2 * n_w + Cr) {  // g_b2[c]
      const int c = e - 2 * n_w;
      for (int r = r0; r < r1; ++r) acc += __ldg(d2 + r * Cr + c);
    } else {  // g_b1[j]
      const int j = e - 2 * n_w - Cr;
      for (int r = r0; r < r1; ++r) acc += __ldg(d1 + r * hd + j);
    }
    mine[e] = acc;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&g_se_ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int e = threadIdx.x; e < n_out; e += 256) {
    float t = 0.f;
    int b = 0;
    for (; b + 4 <= static_cast<int>(gridDim.x); b += 4) {
      const float v0 = ws[static_cast<long long>(b) * n_out + e], v1 = ws[static_cast<long long>(b + 1) * n_out + e],
                  v2 = ws[static_cast<long long>(b + 2) * n_out + e], v3 = ws[static_cast<long long>(b + 3) * n_out + e];
      t += v0, t += v1, t += v2, t += v3;
    }
    for (; b < static_cast<int>(gridDim.x); ++b) t += ws[static_cast<long long>(b) * n_out + e];
    if (e < n_w) g_w2[e] = t;
    else if (e < 2 * n_w) g_w1[e - n_w] = t;
    else if (e < 2 * n_w + Cr) g_b2[e - 2 * n_w] = t;
    else g_b1[e - 2 * n_w - Cr] = t;
  }
  if (threadIdx.x == 0) g_se_ticket = 0;  // ready for the next launch (launches of this kernel are stream-ordered)
}

static unsigned grid_for(long long work_items, int per_block = 256, int max_blocks = 148 * 8) {
  return static_cast<unsigned>(std::max<long long>(1, std::min<long long>((work_items + per_block - 1) / per_block, max_blocks)));
}

}  // namespace pvae

using namespace pvae;

#define PVAE_STREAM static_cast<cudaStream_t>(stream)

extern "C" int pvae_silu_fwd(const float* x, float* out, int64_t n, void* stream) {
  if (!x || !out || n <= 0 || n % 4) return fail(PVAE_ERR_INVALID, "pvae_silu_fwd: bad argument");
  return cuda_ok("silu_fwd_kernel", launch_pdl(silu_fwd_kernel, dim3(grid_for(n / 4)), dim3(256), 0, PVAE_STREAM, x, out,
                                               static_cast<long long>(n / 4)));
}

extern "C" int pvae_add(const float* a, const float* b, float* out, int64_t n, void* stream) {
  if (!a || !b || !out || n <= 0 || n % 4) return fail(PVAE_ERR_INVALID, "pvae_add: bad argument");
  return cuda_ok("add_kernel", launch_pdl(add_kernel, dim3(grid_for(n / 4)), dim3(256), 0, PVAE_STREAM, a, b, out,
                                          static_cast<long long>(n / 4)));
}

extern "C" int pvae_stem_fwd(const float* x, const float* weight, const float* lognorm, const float* bias, float* out,
                             float* out_act, int N, int IH, int IW, int Co, int pad, void* stream) {
  if (!x || !weight || !lognorm || !out || N < 1 || Co % 4 || Co > 256 || pad < 0 || pad > 1)
    return fail(PVAE_ERR_INVALID, "pvae_stem_fwd: bad argument");
  const int OH = IH + 2 * pad - 2, OW = IW + 2 * pad - 2;
  const long long total = static_cast<long long>(N) * OH * OW * (Co / 4);
  return cuda_ok("stem_fwd_kernel", launch_pdl(stem_fwd_kernel, dim3(grid_for(total)), dim3(256), 10 * Co * sizeof(float), PVAE_STREAM,
                                               x, weight, lognorm, bias, out, out_act, N, IH, IW, OH, OW, Co, pad));
}

static int partial_blocks(long long rows, long long min_rows_per_block, long long* rows_per_block) {
  long long blocks = std::max<long long>(1, std::min<long long>(296, (rows + min_rows_per_block - 1) / min_rows_per_block));
  *rows_per_block = (rows + blocks - 1) / blocks;
  return static_cast<int>((rows + *rows_per_block - 1) / *rows_per_block);
}

extern "C" int64_t pvae_reduce_ws_floats(int64_t n_out) { return 296 * n_out; }

// g_wpb (10, Co): rows 0..8 the tap-major gradient of the normalised stem weight, row 9 the bias gradient;
// ws >= pvae_reduce_ws_floats(10 * Co)
extern "C" int pvae_stem_wgrad(const float* x, const float* g_out, float* g_wpb, float* ws, int N, int IH, int IW, int Co, int pad,
                               void* stream) {
  if (!x || !g_out || !g_wpb || !ws || N < 1 || Co < 16 || 256 % Co) return fail(PVAE_ERR_INVALID, "pvae_stem_wgrad: bad argument");
  const int OH = IH + 2 * pad - 2, OW = IW + 2 * pad - 2;
  long long ppb;
  const int blocks = partial_blocks(static_cast<long long>(N) * OH * OW, 256, &ppb);
  const size_t smem = static_cast<size_t>(256 / Co) * 10 * Co * sizeof(float);
  if (int rc = cuda_ok("stem_wgrad_kernel", launch_pdl(stem_wgrad_kernel, dim3(blocks), dim3(256), smem, PVAE_STREAM, x, g_out, ws, N,
                                                       IH, IW, OH, OW, Co, pad, ppb)))
    return rc;
  return cuda_ok("sum_partials_kernel", launch_pdl(sum_partials_kernel, dim3(grid_for(10 * Co, 32)), dim3(256), 0, PVAE_STREAM,
                                                   static_cast<const float*>(ws), g_wpb, blocks, 10ll * Co, 1.f));
}

extern "C" int pvae_colsum(const float* x, float* out, float* ws, int64_t R, int C, void* stream) {
  if (!x || !out || !ws || R < 1 || C < 4 || C % 4) return fail(PVAE_ERR_INVALID, "pvae_colsum: bad argument");
  long long rpb;
  const int blocks = partial_blocks(R, 64, &rpb);
  if (int rc = cuda_ok("colsum_partial_kernel", launch_pdl(colsum_partial_kernel, dim3(blocks), dim3(256), 0, PVAE_STREAM, x, ws,
                                                           static_cast<long long>(R), C, rpb)))
    return rc;
  return cuda_ok("sum_partials_kernel", launch_pdl(sum_partials_kernel, dim3(grid_for(C, 32)), dim3(256), 0, PVAE_STREAM,
                                                   static_cast<const float*>(ws), out, blocks, static_cast<long long>(C), 1.f));
}

extern "C" int pvae_small_tn(const float* a, const float* b, float* out, float* ws, int64_t R, int Ca, int Cb, void* stream) {
  if (!a || !b || !out || !ws || R < 1 || Ca < 1 || Cb < 1) return fail(PVAE_ERR_INVALID, "pvae_small_tn: bad argument");
  long long rpb;
  const int blocks = partial_blocks(R, 32, &rpb);
  if (int rc = cuda_ok("small_tn_partial_kernel", launch_pdl(small_tn_partial_kernel, dim3(blocks), dim3(256), 0, PVAE_STREAM, a, b, ws,
                                                             static_cast<long long>(R), Ca, Cb, rpb)))
    return rc;
  return cuda_ok("sum_partials_kernel", launch_pdl(sum_partials_kernel, dim3(grid_for(static_cast<long long>(Ca) * Cb, 32)), dim3(256), 0,
                                                   PVAE_STREAM, static_cast<const float*>(ws), out, blocks,
                                                   static_cast<long long>(Ca) * Cb, 1.f));
}

extern "C" int pvae_se_fwd(const float* c, const float* skip, const float* W1, const float* b1, const float* W2, const float* b2,
                           float eps_res, float* y, float* y_act, float* pool, float* hid, float* gate, int N, int H, int W, int C,
                           int C_real, int hd, int skip_up, void* stream) {
  if (!c || !skip || !W1 || !b1 || !W2 || !b2 || !y || !pool || !hid || !gate) return fail(PVAE_ERR_INVALID, "pvae_se_fwd: null pointer");
  if (N < 1 || H < 1 || W < 1 || C < 16 || C > 256 || C % 4 || 256 % (C / 4) || hd < 1 || hd > 32 || C_real < 4 || C_real > C ||
      C_real % 4 || (skip_up && ((H | W) & 1)))
    return fail(PVAE_ERR_UNSUPPORTED, "pvae_se_fwd: C=%d C_real=%d hd=%d H=%d W=%d", C, C_real, hd, H, W);
  return cuda_ok("se_fwd_kernel", launch_pdl(se_fwd_kernel, dim3(N), dim3(256), 0, PVAE_STREAM, c, skip, W1, b1, W2, b2, eps_res, y,
                                             y_act, pool, hid, gate, H, W, C, C_real, hd, skip_up));
}

extern "C" int pvae_se_bwd(const float* g_y, const float* c, const float* gate, const float* hid, const float* W1, const float* W2,
                           float eps_res, float* g_c, float* d2, float* d1, int N, int HW, int C, int C_real, int hd, void* stream) {
  if (!g_y || !c || !gate || !hid || !W1 || !W2 || !g_c || !d2 || !d1) return fail(PVAE_ERR_INVALID, "pvae_se_bwd: null pointer");
  if (N < 1 || HW < 1 || C < 16 || C > 256 || C % 4 || 256 % (C / 4) || hd < 1 || hd > 32 || C_real < 4 || C_real > C || C_real % 4)
    return fail(PVAE_ERR_UNSUPPORTED, "pvae_se_bwd: C=%d C_real=%d hd=%d", C, C_real, hd);
  return cuda_ok("se_bwd_kernel", launch_pdl(se_bwd_kernel, dim3(N), dim3(256), 0, PVAE_STREAM, g_y, c, gate, hid, W1, W2, eps_res,
                                             g_c, d2, d1, HW, C, C_real, hd));
}

// ws: gridDim * (2 Cr hd + Cr + hd) floats, at most 128 blocks: pvae_reduce_ws_floats(2 Cr hd + Cr + hd) is enough
extern "C" int pvae_se_param_grads(const float* d2, const float* hid, const float* d1, const float* pool, float* g_w2, float* g_b2,
                                   float* g_w1, float* g_b1, float* ws, int N, int C_real, int hd, void* stream) {
  if (!d2 || !hid || !d1 || !pool || !g_w2 || !g_b2 || !g_w1 || !g_b1 || !ws || N < 1 || C_real < 1 || hd < 1)
    return fail(PVAE_ERR_INVALID, "pvae_se_param_grads: bad argument");
  const int blocks = std::max(1, std::min(128, (N + 31) / 32));
  const int rows = (N + blocks - 1) / blocks;
  return cuda_ok("se_param_grads_kernel", launch_pdl(se_param_grads_kernel, dim3((N + rows - 1) / rows), dim3(256), 0, PVAE_STREAM, d2,
                                                     hid, d1, pool, g_w2, g_b2, g_w1, g_b1, ws, N, C_real, hd, rows));
}

extern "C" int pvae_upsample_fwd(const float* x, float* out, int N, int IH, int IW, int OH, int OW, int C, void* stream) {
  if (!x || !out || N < 1 || IH < 1 || IW < 1 || OH < 1 || OW < 1 || C < 4 || C % 4) return fail(PVAE_ERR_INVALID, "pvae_upsample_fwd: bad argument");
  const long long total = static_cast<long long>(N) * OH * OW * (C / 4);
  return cuda_ok("upsample_fwd_kernel", launch_pdl(upsample_fwd_kernel, dim3(grid_for(total)), dim3(256), 0, PVAE_STREAM, x, out,
                                                   static_cast<long long>(N), IH, IW, OH, OW, C));
}

extern "C" int pvae_upsample_bwd(const float* g_out, const float* dsilu_src, float* g_in, int N, int IH, int IW, int OH, int OW, int C,
                                 void* stream) {
  if (!g_out || !g_in || N < 1 || IH < 1 || IW < 1 || OH < 1 || OW < 1 || C < 4 || C % 4) return fail(PVAE_ERR_INVALID, "pvae_upsample_bwd: bad argument");
  const long long total = static_cast<long long>(N) * IH * IW * (C / 4);
  return cuda_ok("upsample_bwd_kernel", launch_pdl(upsample_bwd_kernel, dim3(grid_for(total)), dim3(256), 0, PVAE_STREAM, g_out,
                                                   dsilu_src, g_in, static_cast<long long>(N), IH, IW, OH, OW, C));
}

extern "C" int pvae_permute_cp(const float* in, const float* bias, float* out, float* out_act, int64_t N, int C, int P, int to_nhwc,
                               void* stream) {
  if (!in || !out || N < 1 || C < 1 || P < 1) return fail(PVAE_ERR_INVALID, "pvae_permute_cp: bad argument");
  return cuda_ok("permute_cp_kernel", launch_pdl(permute_cp_kernel, dim3(grid_for(N * C * P)), dim3(256), 0, PVAE_STREAM, in, bias, out,
                                                 out_act, static_cast<long long>(N), C, P, to_nhwc));
}

extern "C" int pvae_head_fwd(const float* x, const float* weight, const float* lognorm, const float* bias, float* y, int64_t R, int C,
                             int C_real, void* stream) {
  if (!x || !weight || !lognorm || !y || R < 1 || C < 4 || C > 256 || C % 4 || C_real < 4 || C_real > C || C_real % 4)
    return fail(PVAE_ERR_INVALID, "pvae_head_fwd: bad argument");
  return cuda_ok("head_fwd_kernel", launch_pdl(head_fwd_kernel, dim3(grid_for(R)), dim3(256), 0, PVAE_STREAM, x, weight, lognorm, bias, y,
                                               static_cast<long long>(R), C, C_real));
}

// g_wb (C + 1): gradient of the normalised filter (first C_real entries meaningful) and, last, of the bias;
// ws >= pvae_reduce_ws_floats(C + 1)
extern "C" int pvae_head_bwd(const float* x, const float* g_y, const float* weight, const float* lognorm, float* g_x, float* g_wb,
                             float* ws, int64_t R, int C, int C_real, void* stream) {
  if (!x || !g_y || !weight || !lognorm || !g_x || !g_wb || !ws || R < 1 || C < 32 || C > 256 || 256 % C || C_real < 4 || C_real > C)
    return fail(PVAE_ERR_INVALID, "pvae_head_bwd: bad argument");
  long long rpb;
  const int blocks = partial_blocks(R, 256, &rpb);
  if (int rc = cuda_ok("head_bwd_kernel", launch_pdl(head_bwd_kernel, dim3(blocks), dim3(256), 0, PVAE_STREAM, x, g_y, weight, lognorm,
                                                     g_x, ws, static_cast<long long>(R), C, C_real, rpb)))
    return rc;
  return cuda_ok("sum_partials_kernel", launch_pdl(sum_partials_kernel, dim3(grid_for(C + 1, 32)), dim3(256), 0, PVAE_STREAM,
                                                   static_cast<const float*>(ws), g_wb, blocks, static_cast<long long>(C + 1), 1.f));
}

extern "C" int pvae_avgpool_fwd(const float* x, float* pool, int N, int HW, int C, void* stream) {
  if (!x || !pool || N < 1 || HW < 1 || C < 16 || C > 256 || C % 4 || 256 % (C / 4))
    return fail(PVAE_ERR_UNSUPPORTED, "pvae_avgpool_fwd: C=%d", C);
  return cuda_ok("avgpool_fwd_kernel", launch_pdl(avgpool_fwd_kernel, dim3(N), dim3(256), 0, PVAE_STREAM, x, pool, HW, C));
}

extern "C" int pvae_pool_silu_bwd(const float* g_a, const float* x, const float* g_pool, float* g_x, int N, int HW, int C,
                                  void* stream) {
  if (!g_a || !x || !g_pool || !g_x || N < 1 || HW < 1 || C % 4) return fail(PVAE_ERR_INVALID, "pvae_pool_silu_bwd: bad argument");
  const long long total = static_cast<long long>(N) * HW * (C / 4);
  return cuda_ok("pool_silu_bwd_kernel", launch_pdl(pool_silu_bwd_kernel, dim3(grid_for(total)), dim3(256), 0, PVAE_STREAM, g_a, x,
                                                    g_pool, g_x, static_cast<long long>(N), HW, C));
}

extern "C" int pvae_bias_act(float* y, const float* bias, const float* add, int64_t R, int C, int relu, float drop_p, uint64_t seed,
                             uint64_t offset, void* stream) {
  if (!y || R < 1 || C < 4 || C % 4 || drop_p < 0.f || drop_p >= 1.f) return fail(PVAE_ERR_INVALID, "pvae_bias_act: bad argument");
  return cuda_ok("bias_act_kernel", launch_pdl(bias_act_kernel, dim3(grid_for(R * (C / 4))), dim3(256), 0, PVAE_STREAM, y, bias, add,
                                               static_cast<long long>(R), C, relu, drop_p, seed, offset));
}

extern "C" int pvae_relu_bwd(float* g, const float* y_saved, int64_t n, float drop_p, void* stream) {
  if (!g || !y_saved || n < 4 || n % 4 || drop_p < 0.f || drop_p >= 1.f) return fail(PVAE_ERR_INVALID, "pvae_relu_bwd: bad argument");
  return cuda_ok("relu_bwd_kernel", launch_pdl(relu_bwd_kernel, dim3(grid_for(n / 4)), dim3(256), 0, PVAE_STREAM, g, y_saved,
                                               static_cast<long long>(n / 4), 1.f / (1.f - drop_p)));
}

extern "C" int pvae_layernorm_fwd(const float* a, const float* b, const float* gamma, const float* beta, float* out, float* xhat,
                                  float* rstd, int64_t R, int C, float eps, void* stream) {
  if (!a || !gamma || !beta || !out || !xhat || !rstd || R < 1 || C < 32 || C % 32 || C > 256)
    return fail(PVAE_ERR_INVALID, "pvae_layernorm_fwd: bad argument");
  return cuda_ok("layernorm_fwd_kernel", launch_pdl(layernorm_fwd_kernel, dim3(grid_for(R, 8)), dim3(256), 0, PVAE_STREAM, a, b, gamma,
                                                    beta, out, xhat, rstd, static_cast<long long>(R), C, eps));
}

// g_gamma_beta: [g_gamma (C) | g_beta (C)]; ws >= pvae_reduce_ws_floats(2 * C)
extern "C" int pvae_layernorm_bwd(const float* g, const float* xhat, const float* rstd, const float* gamma, float* g_s,
                                  float* g_gamma_beta, float* ws, int64_t R, int C, void* stream) {
  if (!g || !xhat || !rstd || !gamma || !g_s || !g_gamma_beta || !ws || R < 1 || C < 32 || C % 32 || C > 256)
    return fail(PVAE_ERR_INVALID, "pvae_layernorm_bwd: bad argument");
  long long rpb;
  const int blocks = partial_blocks(R, 32, &rpb);
  if (int rc = cuda_ok("layernorm_bwd_kernel", launch_pdl(layernorm_bwd_kernel, dim3(blocks), dim3(256), 16 * C * sizeof(float),
                                                          PVAE_STREAM, g, xhat, rstd, gamma, g_s, ws, static_cast<long long>(R), C, rpb)))
    return rc;
  return cuda_ok("sum_partials_kernel", launch_pdl(sum_partials_kernel, dim3(grid_for(2 * C, 32)), dim3(256), 0, PVAE_STREAM,
                                                   static_cast<const float*>(ws), g_gamma_beta, blocks, 2ll * C, 1.f));
}
