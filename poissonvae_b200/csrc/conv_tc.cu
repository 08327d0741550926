// Implicit-GEMM convolutions of the conv encoder / decoder on tcgen05.
//
// Activations are NHWC fp32 in HBM.  A convolution is D[pixel, co] = sum_{tap, ci} X[pixel (+) tap, ci] W[co, tap, ci]:
//   * operand A (activations) is never unfolded in memory: for every (tap, 32-channel block) the TMA engine loads
//     a {32 ch, BW, BH, BI} box of the 4-D tensor (C, W, H, N) at the tap's spatial offset - out-of-bounds
//     coordinates are zero-filled by the hardware (= the padding), a traversal stride of 2 gives the strided
//     convolutions - straight into the 128-byte-swizzled K-major layout the tensor core reads (128 pixels x 32 ch);
//   * operand B (weights) is a small packed matrix [co][tap][ci] produced once per step by the weight-norm kernel;
//   * tcgen05.mma.kind::tf32 accumulates in TMEM; "3xTF32" operand splitting as in gemm_tc.cu keeps fp32-grade
//     products; the epilogue fuses bias, SiLU of the output, the SiLU derivative and residual additions.
// The same kernel computes the data gradient (a convolution of g_out with the [ci][tap][co] packing; a strided
// convolution's dgrad is launched once per output parity class with that class's taps).
// The weight gradient is a second kernel: K runs over pixels, both operands are channel-contiguous (MN-major,
// 32-byte swizzle atoms), the tap shift is again applied by the TMA coordinates, split-K over pixel blocks.
//
// Replaces: Conv2D.forward base/common.py:313-323 (F.conv2d), `_normalize` :426-431, and what autograd derives
// from them, as used by ConvLayer :195-242, FactorizedReduce/StrideReduce :75-126, Cell :145-192.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pvae_b200.h"
#include "pvae_device.cuh"
#include "pvae_host.h"
#include "tc_common.cuh"

namespace pvae {

constexpr int kMaxTaps = 16;
struct ConvTap {
  short dy, dx;  // input offset of this tap: ih = oh * stride + dy
  short widx;    // tap index inside the packed weight
  short pad;
};

struct ConvParams {
  int Ci, Co;                // channels of the A tensor / of the output
  int OH, OW, NI;            // output grid covered by this launch, images
  int lw, lh;                // log2 of the pixel tile's BW, BH (BW * BH * BI = 128)
  int tiles_w, tiles_h;
  int stride;
  int ntaps;
  int w_lo_row;              // > 0: the packed filter carries its TF32 low parts in rows [w_lo_row, 2 w_lo_row) (prep kernel)
  ConvTap taps[kMaxTaps];
  // output pixel (n, oh, ow) -> ((n * OHf + oh * os + ooh) * OWf + ow * os + oow) * Co
  int OHf, OWf, os, ooh, oow;
  float* out;
  float* out_act;          // nullable: silu(out)
  const float* bias;       // nullable (Co)
  const float* add_pre;    // nullable: v += add_pre[idx] before the SiLU-derivative factor
  const float* dsilu_src;  // nullable: v *= silu'(dsilu_src[idx])
  const float* add_post;   // nullable: v += add_post[idx]
};

// MUFU-based sigmoid (ex2.approx + rcp.approx, relative error ~3e-7): the epilogue evaluates it for every output element
// while the tensor pipe idles, so its instruction count matters
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }
__device__ __forceinline__ float dsilu_f(float x) {
  const float s = sigmoid_f(x);
  return s * (1.f + x * (1.f - s));
}

template <int BN, bool SPLIT, int STAGES>
struct ConvCfg {
  static constexpr uint32_t kABytes = 128 * 32 * 4;
  static constexpr uint32_t kBBytes = BN * 32 * 4;
  static constexpr uint32_t kTileBytes = kABytes + kBBytes;
  static constexpr uint32_t kStageBytes = SPLIT ? 2 * kTileBytes : kTileBytes;
  static constexpr int kSmem = STAGES * kStageBytes + 1024 + 256;
  static constexpr int kCtasPerSm = kSmem <= 110 * 1024 ? 2 : 1;
};

// ---------------------------------------------------------------------------------------------------------
// forward / data-gradient kernel: 128 output pixels x BN output channels per CTA
// ---------------------------------------------------------------------------------------------------------
template <int BN, bool SPLIT, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, ConvCfg<BN, SPLIT, STAGES>::kCtasPerSm)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BN, SPLIT, STAGES>;
  constexpr uint32_t kABytes = Cfg::kABytes, kTileBytes = Cfg::kTileBytes, kStageBytes = Cfg::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* conv_bar = empty_bar + STAGES;
  uint64_t* accum_bar = conv_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.y;
  const int tw = tile % p.tiles_w, th = (tile / p.tiles_w) % p.tiles_h, tg = tile / (p.tiles_w * p.tiles_h);
  const int ow0 = tw << p.lw, oh0 = th << p.lh, img0 = tg << (7 - p.lw - p.lh);
  const int co0 = blockIdx.x * BN;
  const int cblocks = p.Ci >> 5;
  const int num_kb = p.ntaps * cblocks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&conv_bar[s], kGemmThreads - 64);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t kTmemCols = (SPLIT ? 2 * BN : BN) < 32 ? 32 : (SPLIT ? 2 * BN : BN);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      int t = 0, cb = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
        uint8_t* sa = smem + s * kStageBytes;
        const bool w_lo = SPLIT && p.w_lo_row > 0;  // the filter's low parts come precomputed: only A is split in the kernel
        mbar_expect_tx(&full_bar[s], kTileBytes + (w_lo ? Cfg::kBBytes : 0u));
        const ConvTap tap = p.taps[t];
        tma_load_4d(&tma_a, &full_bar[s], sa, cb * 32, ow0 * p.stride + tap.dx, oh0 * p.stride + tap.dy, img0);
        tma_load_2d(&tma_b, &full_bar[s], sa + kABytes, tap.widx * p.Ci + cb * 32, co0);
        if (w_lo) tma_load_2d(&tma_b, &full_bar[s], sa + kTileBytes + kABytes, tap.widx * p.Ci + cb * 32, p.w_lo_row + co0);
        if (++cb == cblocks) cb = 0, ++t;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, BN, false, false);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(SPLIT ? &conv_bar[s] : &full_bar[s], (kb / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(smem + s * kStageBytes), sb = sa + kABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = make_smem_desc(sa + k * 32, 16, 1024, 2);
          const uint64_t db = make_smem_desc(sb + k * 32, 16, 1024, 2);
          umma_tf32(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          if (SPLIT) {
            const uint64_t dal = make_smem_desc(sa + kTileBytes + k * 32, 16, 1024, 2);
            const uint64_t dbl = make_smem_desc(sb + kTileBytes + k * 32, 16, 1024, 2);
            umma_tf32(tmem_base + BN, dal, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_tf32(tmem_base + BN, da, dbl, idesc, 1u);
          }
        }
        umma_commit(&empty_bar[s]);
      }
      if (num_kb > 0) umma_commit(accum_bar);
    }
  } else {
    if (SPLIT) {
      const int ct = threadIdx.x - 64;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&full_bar[s], (kb / STAGES) & 1);
        const uint4* tile_hi = reinterpret_cast<const uint4*>(smem + s * kStageBytes);
        uint4* tile_lo = reinterpret_cast<uint4*>(smem + s * kStageBytes + kTileBytes);
#pragma unroll 4
        const int n16 = static_cast<int>((p.w_lo_row > 0 ? kABytes : kTileBytes) / 16);
        for (int i = ct; i < n16; i += kGemmThreads - 64) tile_lo[i] = tf32_lo4(tile_hi[i]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&conv_bar[s])) : "memory");
      }
    }
    // ===== epilogue =====
    if (num_kb > 0) {
      mbar_wait(accum_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int quarter = warp & 3;
    const int sub = lane >> 3, col = 4 * (lane & 7);
    const int bw_mask = (1 << p.lw) - 1, bh_mask = (1 << p.lh) - 1;
    float* stg = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 33);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      if (num_kb > 0) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c0);
        tmem_ld_32x32(taddr, v);
        if (SPLIT) {
          uint32_t w[32];
          tmem_ld_32x32(taddr + BN, w);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
        } else {
          tmem_ld_wait();
        }
      } else {  // a parity class no tap reaches: the convolution term is zero, the epilogue still applies
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(v[j]);
      __syncwarp();
      const int co = co0 + c0 + col;
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.bias != nullptr && co < p.Co) bv = ldg4(p.bias + co);
#pragma unroll
      for (int rr = 0; rr < 32; rr += 4) {
        const int m = quarter * 32 + rr + sub;  // pixel of the tile
        const int ow = ow0 + (m & bw_mask), oh = oh0 + ((m >> p.lw) & bh_mask), n = img0 + (m >> (p.lw + p.lh));
        if (ow < p.OW && oh < p.OH && n < p.NI && co < p.Co) {
          const long long idx =
              ((static_cast<long long>(n) * p.OHf + oh * p.os + p.ooh) * p.OWf + ow * p.os + p.oow) * p.Co + co;
          const float* src = stg + (rr + sub) * 33 + col;
          float4 r = make_float4(src[0] + bv.x, src[1] + bv.y, src[2] + bv.z, src[3] + bv.w);
          if (p.add_pre != nullptr) {
            const float4 a = ldg4(p.add_pre + idx);
            r.x += a.x, r.y += a.y, r.z += a.z, r.w += a.w;
          }
          if (p.dsilu_src != nullptr) {
            const float4 x = ldg4(p.dsilu_src + idx);
            r.x *= dsilu_f(x.x), r.y *= dsilu_f(x.y), r.z *= dsilu_f(x.z), r.w *= dsilu_f(x.w);
          }
          if (p.add_post != nullptr) {
            const float4 a = ldg4(p.add_post + idx);
            r.x += a.x, r.y += a.y, r.z += a.z, r.w += a.w;
          }
          stg4(p.out + idx, r);
          if (p.out_act != nullptr) stg4(p.out_act + idx, make_float4(silu_f(r.x), silu_f(r.y), silu_f(r.z), silu_f(r.w)));
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// ---------------------------------------------------------------------------------------------------------
// weight-gradient kernel: out[z][(tap, ci), co] = sum over this split's pixels of X[pixel (+) tap, ci] G[pixel, co]
// 128 rows = four (tap, 32-channel block) slabs per CTA, BN output channels, K = 32 pixels per stage
// ---------------------------------------------------------------------------------------------------------
struct WgradParams {
  int Ci, Co;
  int OH, OW, NI;
  int lw, lh;  // log2 of the pixel block's PW, PH (PW * PH * PI = 32)
  int tiles_w, tiles_h;
  int stride;
  int ntaps;
  ConvTap taps[kMaxTaps];
  int num_kb, kb_per_split;
  float* out;  // (splits, ntaps * Ci + 1, Co): rows (tap, ci), last row = sum over pixels of g_out (bias gradient)
};

template <int BN, bool SPLIT, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, ConvCfg<BN, SPLIT, STAGES>::kCtasPerSm)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_g,
                  const __grid_constant__ WgradParams p) {
  using Cfg = ConvCfg<BN, SPLIT, STAGES>;
  constexpr uint32_t kABytes = Cfg::kABytes, kTileBytes = Cfg::kTileBytes, kStageBytes = Cfg::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* conv_bar = empty_bar + STAGES;
  uint64_t* accum_bar = conv_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.x * BN;
  const int cblocks = p.Ci >> 5;
  const int npairs = p.ntaps * cblocks;
  const int pair0 = blockIdx.y * 4;
  const int nslabs = max(0, min(4, npairs - pair0));  // slabs the TMA engine fills
  const int ones_slab = npairs - pair0;               // in [0, 4): this row block also carries the all-ones slab whose
                                                      // output row is sum_pixels g_out = the bias gradient
  const int kb_begin = blockIdx.z * p.kb_per_split;
  const int num_kb = max(0, min(p.num_kb, kb_begin + p.kb_per_split) - kb_begin);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_g)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&conv_bar[s], kGemmThreads - 64);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t kTmemCols = (SPLIT ? 2 * BN : BN) < 32 ? 32 : (SPLIT ? 2 * BN : BN);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // slabs the TMA engine does not fill: the all-ones slab (constant for the whole kernel; its lo plane comes out
  // as zeros from the splitter) and unused ones (zeros; their rows are dropped)
  if (nslabs < 4) {
    const uint32_t one = __float_as_uint(1.f);
    for (int s = 0; s < STAGES; ++s) {
      uint4* z = reinterpret_cast<uint4*>(smem + s * kStageBytes + nslabs * 4096);
      for (int i = threadIdx.x; i < (4 - nslabs) * 256; i += kGemmThreads) {
        const uint32_t v = (nslabs + i / 256 == ones_slab) ? one : 0u;
        z[i] = make_uint4(v, v, v, v);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = static_cast<uint32_t>(nslabs + BN / 32) * 4096u;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
        uint8_t* sa = smem + s * kStageBytes;
        uint8_t* sb = sa + kABytes;
        mbar_expect_tx(&full_bar[s], tx);
        const int kbg = kb_begin + kb;
        const int tw = kbg % p.tiles_w, th = (kbg / p.tiles_w) % p.tiles_h, tg = kbg / (p.tiles_w * p.tiles_h);
        const int ow0 = tw << p.lw, oh0 = th << p.lh, img0 = tg << (5 - p.lw - p.lh);
        for (int j = 0; j < nslabs; ++j) {
          const int pair = pair0 + j;
          const ConvTap tap = p.taps[pair / cblocks];
          tma_load_4d(&tma_x, &full_bar[s], sa + j * 4096, (pair % cblocks) * 32, ow0 * p.stride + tap.dx,
                      oh0 * p.stride + tap.dy, img0);
        }
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) tma_load_4d(&tma_g, &full_bar[s], sb + j * 4096, co0 + 32 * j, ow0, oh0, img0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, BN, true, true);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(SPLIT ? &conv_bar[s] : &full_bar[s], (kb / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(smem + s * kStageBytes), sb = sa + kABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = make_smem_desc(sa + k * 1024, 4096, 512, 1);
          const uint64_t db = make_smem_desc(sb + k * 1024, 4096, 512, 1);
          umma_tf32(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          if (SPLIT) {
            const uint64_t dal = make_smem_desc(sa + kTileBytes + k * 1024, 4096, 512, 1);
            const uint64_t dbl = make_smem_desc(sb + kTileBytes + k * 1024, 4096, 512, 1);
            umma_tf32(tmem_base + BN, dal, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_tf32(tmem_base + BN, da, dbl, idesc, 1u);
          }
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(accum_bar);
    }
  } else {
    if (SPLIT) {
      const int ct = threadIdx.x - 64;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&full_bar[s], (kb / STAGES) & 1);
        const uint4* tile_hi = reinterpret_cast<const uint4*>(smem + s * kStageBytes);
        uint4* tile_lo = reinterpret_cast<uint4*>(smem + s * kStageBytes + kTileBytes);
#pragma unroll 4
        for (int i = ct; i < static_cast<int>(kTileBytes / 16); i += kGemmThreads - 64) tile_lo[i] = tf32_lo4(tile_hi[i]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&conv_bar[s])) : "memory");
      }
    }
    const int quarter = warp & 3;  // rows [32 quarter, +32) = slab `quarter`
    const int pair = pair0 + quarter;
    float* obase = p.out + static_cast<long long>(blockIdx.z) * (p.ntaps * p.Ci + 1) * p.Co;
    if (num_kb > 0) {
      mbar_wait(accum_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int sub = lane >> 3, col = 4 * (lane & 7);
    float* stg = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 33);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      if (num_kb > 0) {
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c0);
        tmem_ld_32x32(taddr, v);
        if (SPLIT) {
          uint32_t w[32];
          tmem_ld_32x32(taddr + BN, w);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
        } else {
          tmem_ld_wait();
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(v[j]);
      __syncwarp();
      const int co = co0 + c0 + col;
      if (pair < npairs && co < p.Co) {
        const int row0 = p.taps[pair / cblocks].widx * p.Ci + (pair % cblocks) * 32;
#pragma unroll
        for (int rr = 0; rr < 32; rr += 4) {
          const float* src = stg + (rr + sub) * 33 + col;
          stg4(obase + static_cast<long long>(row0 + rr + sub) * p.Co + co, make_float4(src[0], src[1], src[2], src[3]));
        }
      } else if (pair == npairs && co < p.Co && sub == 0) {  // bias row (all 32 rows of the ones slab are equal)
        const float* src = stg + col;
        stg4(obase + static_cast<long long>(p.ntaps * p.Ci) * p.Co + co, make_float4(src[0], src[1], src[2], src[3]));
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// ---------------------------------------------------------------------------------------------------------
// weight normalisation (base/common.py:426-431): w = exp(lognorm) * weight / (||weight||_co + eps), written in the
// two packings the tensor-core kernels read: w_fwd[co][tap][ci] and w_dgrad[ci][tap][co].
// `groups` > 1: FactorizedReduce (base/common.py:98-126) - output channel co only sees tap co / (Co / groups) of a
// 2x2 stride-2 window; weight is then (Co, Ci) and the other taps are packed as zeros.
// One block per output channel.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv_weight_prep_kernel(const float* __restrict__ weight, const float* __restrict__ lognorm,
                                                               const float* __restrict__ bias, float* __restrict__ w_fwd,
                                                               float* __restrict__ w_dgrad, float* __restrict__ bias_pad, int Co,
                                                               int Ci, int CoP, int CiP, int ntaps, int groups, int with_lo) {
  __shared__ float s_part[4];
  pdl_launch_dependents();
  pdl_wait();
  const int co = blockIdx.x;  // < CoP; channels >= Co are padding (the tensor-core kernels want multiples of 32): zeros
  const bool real = co < Co;
  const int wtaps = groups > 1 ? 1 : ntaps;  // taps stored in `weight`
  const int n = Ci * wtaps;
  const float* w = weight + static_cast<long long>(real ? co : 0) * n;  // (ci, tap) order, torch (Co, Ci, KH, KW)
  float ss = 0.f;
  for (int i = threadIdx.x; i < n; i += 128) ss = fmaf(w[i], w[i], ss);
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = ss;
  __syncthreads();
  const float nrm = sqrtf((s_part[0] + s_part[1]) + (s_part[2] + s_part[3]));
  const float scale = !real ? 0.f : (lognorm != nullptr ? expf(lognorm[co]) / (nrm + kF32Eps) : 1.f);
  const int my_tap = groups > 1 ? co / (Co / groups) : -1;
  for (int i = threadIdx.x; i < CiP * ntaps; i += 128) {
    const int t = i / CiP, ci = i % CiP;
    float v = 0.f;
    if (real && ci < Ci) {
      if (groups > 1)
        v = t == my_tap ? scale * w[ci] : 0.f;
      else
        v = scale * w[ci * ntaps + t];
    }
    w_fwd[(static_cast<long long>(co) * ntaps + t) * CiP + ci] = v;
    if (w_dgrad != nullptr) w_dgrad[(static_cast<long long>(ci) * ntaps + t) * CoP + co] = v;
    if (with_lo) {  // low part of the 3xTF32 split, x - trunc_tf32(x), as rows [rows, 2 rows) of the same matrices
      const float lo = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      w_fwd[(static_cast<long long>(CoP + co) * ntaps + t) * CiP + ci] = lo;
      if (w_dgrad != nullptr) w_dgrad[(static_cast<long long>(CiP + ci) * ntaps + t) * CoP + co] = lo;
    }
  }
  if (threadIdx.x == 0 && bias_pad != nullptr) bias_pad[co] = (real && bias != nullptr) ? bias[co] : 0.f;
}

// backward of the above: from g_w packed as (ntaps * Ci, Co) rows (tap, ci):
//   g_lognorm[co] = sum g_w w;  g_weight = s g_w - weight * s * (sum g_w weight) / (r (r + eps)),  s = e^lognorm / (r + eps)
__global__ void __launch_bounds__(128) conv_weight_grad_kernel(const float* __restrict__ g_wp, const float* __restrict__ weight,
                                                               const float* __restrict__ lognorm, float* __restrict__ g_weight,
                                                               float* __restrict__ g_lognorm, int Co, int Ci, int ntaps, int groups) {
  __shared__ float s_a[4], s_b[4];
  pdl_launch_dependents();
  pdl_wait();
  const int co = blockIdx.x;
  const int wtaps = groups > 1 ? 1 : ntaps;
  const int n = Ci * wtaps;
  const int my_tap = groups > 1 ? co / (Co / groups) : 0;
  const float* w = weight + static_cast<long long>(co) * n;
  float ss = 0.f, gv = 0.f;
  for (int i = threadIdx.x; i < n; i += 128) {
    const int ci = groups > 1 ? i : i / ntaps, t = groups > 1 ? my_tap : i % ntaps;
    const float g = g_wp[(static_cast<long long>(t) * Ci + ci) * Co + co];
    ss = fmaf(w[i], w[i], ss);
    gv = fmaf(g, w[i], gv);
  }
  ss = warp_sum(ss), gv = warp_sum(gv);
  if ((threadIdx.x & 31) == 0) s_a[threadIdx.x >> 5] = ss, s_b[threadIdx.x >> 5] = gv;
  __syncthreads();
  const float r = sqrtf((s_a[0] + s_a[1]) + (s_a[2] + s_a[3]));
  const float gdotv = (s_b[0] + s_b[1]) + (s_b[2] + s_b[3]);
  const float scale = lognorm != nullptr ? expf(lognorm[co]) / (r + kF32Eps) : 1.f;
  const float back = lognorm != nullptr ? scale * gdotv / (fmaxf(r, 1e-30f) * (r + kF32Eps)) : 0.f;
  for (int i = threadIdx.x; i < n; i += 128) {
    const int ci = groups > 1 ? i : i / ntaps, t = groups > 1 ? my_tap : i % ntaps;
    const float g = g_wp[(static_cast<long long>(t) * Ci + ci) * Co + co];
    g_weight[static_cast<long long>(co) * n + i] = scale * g - back * w[i];
  }
  if (threadIdx.x == 0 && g_lognorm != nullptr) g_lognorm[co] = scale * gdotv;  // sum g_w * w, w = scale * weight
}

// Sum the split partials of conv_wgrad_kernel in split order into the transposed gradient gT[co][R + 1] (row R = bias
// gradient).  A block owns 16 rows x 32 output channels: 128-byte coalesced reads of every partial, eight in flight
// per thread, transposed through shared memory for the store.
__global__ void __launch_bounds__(256) conv_wgrad_reduce_kernel(const float* __restrict__ ws, int splits, float* __restrict__ gT,
                                                                int R1, int Co) {
  __shared__ float s_t[16][33];
  pdl_launch_dependents();
  pdl_wait();
  const int col = threadIdx.x & 31, rl = threadIdx.x >> 5;  // 8 row lanes x 2 rows
  const int co = blockIdx.x * 32 + col;
  const long long split_stride = static_cast<long long>(R1) * Co;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = blockIdx.y * 16 + rl + 8 * h;
    float a = 0.f;
    if (r < R1 && co < Co) {
      const float* src = ws + static_cast<long long>(r) * Co + co;
      int z = 0;
      for (; z + 8 <= splits; z += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (z + u) * split_stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) a += v[u];
      }
      for (; z < splits; ++z) a += __ldg(src + z * split_stride);
    }
    s_t[rl + 8 * h][col] = a;
  }
  __syncthreads();
  // store: thread -> (co_l = t / 8, four rows apart by 8 lanes... ) 16 consecutive r per output channel
  const int co_l = threadIdx.x >> 3, r4 = (threadIdx.x & 7) * 2;
  const int co_o = blockIdx.x * 32 + co_l;
  if (co_o < Co) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int r = blockIdx.y * 16 + r4 + k;
      if (r < R1) gT[static_cast<long long>(co_o) * R1 + r] = s_t[r4 + k][co_l];
    }
  }
}

// weight-normalisation backward from the transposed gradient (rows contiguous), one warp per output channel:
//   g_lognorm[co] = sum g_w w;  g_weight = s g_w - weight s (sum g_w weight) / (r (r + eps)),  s = e^lognorm / (r + eps)
__global__ void __launch_bounds__(256) conv_wgrad_finish_kernel(const float* __restrict__ gT, const float* __restrict__ weight,
                                                                const float* __restrict__ lognorm, float* __restrict__ g_weight,
                                                                float* __restrict__ g_lognorm, float* __restrict__ g_bias, int Co,
                                                                int Ci, int CiP, int ntaps, int groups) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co = blockIdx.x * 8 + warp;
  if (co >= Co) return;       // Co, Ci: the parameter's real channel counts; CiP: padded rows per tap in gT
  const int R = ntaps * CiP;
  const float* g = gT + static_cast<long long>(co) * (R + 1);
  const int wtaps = groups > 1 ? 1 : ntaps;
  const int n = Ci * wtaps;
  const int my_tap = groups > 1 ? co / (Co / groups) : 0;
  const float* w = weight + static_cast<long long>(co) * n;
  float ss = 0.f, gv = 0.f;
  for (int i = lane; i < n; i += 32) {
    const int ci = groups > 1 ? i : i / ntaps, t = groups > 1 ? my_tap : i % ntaps;
    const float wv = __ldg(w + i);
    ss = fmaf(wv, wv, ss);
    gv = fmaf(__ldg(g + t * CiP + ci), wv, gv);
  }
  ss = warp_sum(ss), gv = warp_sum(gv);
  const float r = sqrtf(ss);
  const float scale = lognorm != nullptr ? expf(lognorm[co]) / (r + kF32Eps) : 1.f;
  const float back = lognorm != nullptr ? scale * gv / (fmaxf(r, 1e-30f) * (r + kF32Eps)) : 0.f;
  for (int i = lane; i < n; i += 32) {
    const int ci = groups > 1 ? i : i / ntaps, t = groups > 1 ? my_tap : i % ntaps;
    g_weight[static_cast<long long>(co) * n + i] = scale * __ldg(g + t * CiP + ci) - back * __ldg(w + i);
  }
  if (lane == 0) {
    if (g_lognorm != nullptr) g_lognorm[co] = scale * gv;
    if (g_bias != nullptr) g_bias[co] = g[R];
  }
}

// ---- host side -------------------------------------------------------------------------------------
static int ilog2_ceil(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

// 4-D map over an NHWC tensor (C, W, H, N) with a {32, bw * stride, bh * stride, bi} box and traversal stride
static int make_nhwc_map(CUtensorMap* out, const float* base, int N, int H, int W, int C, int bw, int bh, int bi, int stride,
                         int swizzle) {
  const long long dims[4] = {C, W, H, N};
  const long long strides[3] = {4ll * C, 4ll * C * W, 4ll * C * W * H};
  const int box[4] = {32, bw * stride, bh * stride, bi};
  const int estr[4] = {1, stride, stride, 1};
  return make_tensor_map(out, base, 4, dims, strides, box, estr, swizzle);
}

template <int BN, bool SPLIT, int STAGES>
static int launch_conv(const CUtensorMap& ma, const CUtensorMap& mb, const ConvParams& p, int tiles, cudaStream_t s) {
  constexpr int smem = ConvCfg<BN, SPLIT, STAGES>::kSmem;
  static bool configured = false;
  if (!configured) {
    if (int rc = cuda_ok("cudaFuncSetAttribute", cudaFuncSetAttribute(conv_igemm_kernel<BN, SPLIT, STAGES>,
                                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem)))
      return rc;
    configured = true;
  }
  dim3 grid((p.Co + BN - 1) / BN, tiles, 1);
  return cuda_ok("conv_igemm_kernel", launch_pdl(conv_igemm_kernel<BN, SPLIT, STAGES>, grid, dim3(kGemmThreads), smem, s, ma, mb, p));
}

template <int BN, bool SPLIT, int STAGES>
static int launch_wgrad(const CUtensorMap& mx, const CUtensorMap& mg, const WgradParams& p, int row_blocks, int splits,
                        cudaStream_t s) {
  constexpr int smem = ConvCfg<BN, SPLIT, STAGES>::kSmem;
  static bool configured = false;
  if (!configured) {
    if (int rc = cuda_ok("cudaFuncSetAttribute", cudaFuncSetAttribute(conv_wgrad_kernel<BN, SPLIT, STAGES>,
                                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem)))
      return rc;
    configured = true;
  }
  dim3 grid((p.Co + BN - 1) / BN, row_blocks, splits);
  return cuda_ok("conv_wgrad_kernel", launch_pdl(conv_wgrad_kernel<BN, SPLIT, STAGES>, grid, dim3(kGemmThreads), smem, s, mx, mg, p));
}

// tuning (A/B): 0 = default (narrow tiles: 2 stages, 2 CTAs per SM); 1 = 64-wide tiles with 4 stages / 1 CTA per SM;
// 2 = 32-wide tiles with 4 stages / 1 CTA; 3 = 64-wide tiles also where 128-wide ones would fit
static int g_conv_variant = 0;

static int fill_taps(ConvParams& p, int ntaps, const int* dy, const int* dx, const int* widx) {
  if (ntaps < 0 || ntaps > kMaxTaps) return fail(PVAE_ERR_INVALID, "conv: %d taps not in [0, %d]", ntaps, kMaxTaps);
  p.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) p.taps[t] = ConvTap{static_cast<short>(dy[t]), static_cast<short>(dx[t]), static_cast<short>(widx[t]), 0};
  return PVAE_OK;
}

}  // namespace pvae

using namespace pvae;

extern "C" int pvae_set_conv_variant(int v) {
  if (v < 0 || v > 3) return fail(PVAE_ERR_INVALID, "pvae_set_conv_variant: %d not in [0, 3]", v);
  g_conv_variant = v;
  return PVAE_OK;
}

extern "C" int pvae_conv_weight_prep(const float* weight, const float* lognorm, const float* bias, float* w_fwd, float* w_dgrad,
                                     float* bias_pad, int Co, int Ci, int Co_pad, int Ci_pad, int ntaps, int groups, int with_lo,
                                     void* stream) {
  if (!weight || !w_fwd || Co < 1 || Ci < 1 || Co_pad < Co || Ci_pad < Ci || ntaps < 1 || ntaps > kMaxTaps || groups < 1 ||
      (groups > 1 && (groups != ntaps || Co % groups)))
    return fail(PVAE_ERR_INVALID, "pvae_conv_weight_prep: bad argument");
  return cuda_ok("conv_weight_prep_kernel", launch_pdl(conv_weight_prep_kernel, dim3(Co_pad), dim3(128), 0, static_cast<cudaStream_t>(stream),
                                                       weight, lognorm, bias, w_fwd, w_dgrad, bias_pad, Co, Ci, Co_pad, Ci_pad, ntaps,
                                                       groups, with_lo));
}

extern "C" int pvae_conv_weight_grad(const float* g_wp, const float* weight, const float* lognorm, float* g_weight,
                                     float* g_lognorm, int Co, int Ci, int ntaps, int groups, void* stream) {
  if (!g_wp || !weight || !g_weight || Co < 1 || Ci < 1 || ntaps < 1 || ntaps > kMaxTaps || groups < 1)
    return fail(PVAE_ERR_INVALID, "pvae_conv_weight_grad: bad argument");
  return cuda_ok("conv_weight_grad_kernel", launch_pdl(conv_weight_grad_kernel, dim3(Co), dim3(128), 0, static_cast<cudaStream_t>(stream),
                                                       g_wp, weight, lognorm, g_weight, g_lognorm, Co, Ci, ntaps, groups));
}

extern "C" int pvae_conv_igemm(const float* in, const float* w_packed, float* out, float* out_act, const float* bias,
                               const float* add_pre, const float* dsilu_src, const float* add_post, int NI, int IH, int IW,
                               int Ci, int Co, int OH, int OW, int OHf, int OWf, int out_step, int out_off_h, int out_off_w,
                               int stride, int ntaps, const int* tap_dy, const int* tap_dx, const int* tap_widx,
                               int wtaps, int w_has_lo, int precise, void* stream) {
  if (!in || !w_packed || !out) return fail(PVAE_ERR_INVALID, "pvae_conv_igemm: null pointer");
  if (NI < 1 || IH < 1 || IW < 1 || OH < 1 || OW < 1 || stride < 1 || stride > 2 || out_step < 1)
    return fail(PVAE_ERR_INVALID, "pvae_conv_igemm: bad shape");
  if (Ci % 32 || Co % 32 || Ci < 32 || Co < 32)
    return fail(PVAE_ERR_UNSUPPORTED, "pvae_conv_igemm: Ci=%d Co=%d must be multiples of 32", Ci, Co);
  ConvParams p = {};
  if (int rc = fill_taps(p, ntaps, tap_dy, tap_dx, tap_widx)) return rc;
  p.Ci = Ci, p.Co = Co, p.OH = OH, p.OW = OW, p.NI = NI, p.stride = stride;
  p.lw = std::min(5, ilog2_ceil(OW));
  p.lh = std::min(7 - p.lw, ilog2_ceil(OH));
  const int bw = 1 << p.lw, bh = 1 << p.lh, bi = 128 >> (p.lw + p.lh);
  p.tiles_w = (OW + bw - 1) / bw, p.tiles_h = (OH + bh - 1) / bh;
  const int tiles = p.tiles_w * p.tiles_h * ((NI + bi - 1) / bi);
  p.OHf = OHf, p.OWf = OWf, p.os = out_step, p.ooh = out_off_h, p.oow = out_off_w;
  p.out = out, p.out_act = out_act, p.bias = bias, p.add_pre = add_pre, p.dsilu_src = dsilu_src, p.add_post = add_post;
  CUtensorMap ma, mb;
  if (int rc = make_nhwc_map(&ma, in, NI, IH, IW, Ci, bw, bh, bi, stride, 0)) return rc;
  const int BN = (Co % 128 == 0 && g_conv_variant != 3) ? 128 : (Co % 64 == 0 ? 64 : 32);
  {
    p.w_lo_row = w_has_lo ? Co : 0;
    const long long dims[2] = {static_cast<long long>(wtaps) * Ci, w_has_lo ? 2ll * Co : Co}, strides[1] = {4ll * wtaps * Ci};
    const int box[2] = {32, BN}, estr[2] = {1, 1};
    if (int rc = make_tensor_map(&mb, w_packed, 2, dims, strides, box, estr, 0)) return rc;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (precise) {
    if (BN == 128) return launch_conv<128, true, 3>(ma, mb, p, tiles, s);
    if (BN == 64) return g_conv_variant == 1 ? launch_conv<64, true, 4>(ma, mb, p, tiles, s) : launch_conv<64, true, 2>(ma, mb, p, tiles, s);
    return g_conv_variant == 2 ? launch_conv<32, true, 4>(ma, mb, p, tiles, s) : launch_conv<32, true, 2>(ma, mb, p, tiles, s);
  }
  if (BN == 128) return launch_conv<128, false, 6>(ma, mb, p, tiles, s);
  if (BN == 64) return launch_conv<64, false, 6>(ma, mb, p, tiles, s);
  return launch_conv<32, false, 6>(ma, mb, p, tiles, s);
}

static int wgrad_bn(int Co) { return Co % 128 == 0 ? 128 : (Co % 64 == 0 ? 64 : 32); }
static int wgrad_max_splits(int Ci, int Co, int ntaps) {
  const int ctas = ((ntaps * (Ci / 32) + 1 + 3) / 4) * (Co / wgrad_bn(Co));
  return std::max(1, std::min(64, (2 * 148 + ctas - 1) / ctas));
}

extern "C" int64_t pvae_conv_wgrad_ws_floats(int Ci, int Co, int ntaps) {
  if (Ci < 32 || Co < 32 || ntaps < 1) return 0;  // split partials + the reduced, transposed gradient
  return static_cast<int64_t>(wgrad_max_splits(Ci, Co, ntaps) + 1) * (ntaps * Ci + 1) * Co;
}

extern "C" int pvae_conv_wgrad(const float* in, const float* g_out, const float* weight, const float* lognorm, float* g_weight,
                               float* g_lognorm, float* g_bias, float* ws, int NI, int IH, int IW, int Ci, int Co, int OH,
                               int OW, int stride, int ntaps, const int* tap_dy, const int* tap_dx, const int* tap_widx,
                               int groups, int Ci_real, int Co_real, int precise, void* stream) {
  if (!in || !g_out || !weight || !g_weight || !ws) return fail(PVAE_ERR_INVALID, "pvae_conv_wgrad: null pointer");
  if (Ci_real < 1 || Ci_real > Ci || Co_real < 1 || Co_real > Co) return fail(PVAE_ERR_INVALID, "pvae_conv_wgrad: bad real channel counts");
  if (Ci % 32 || Co % 32 || Ci < 32 || Co < 32)
    return fail(PVAE_ERR_UNSUPPORTED, "pvae_conv_wgrad: Ci=%d Co=%d must be multiples of 32", Ci, Co);
  if (NI < 1 || stride < 1 || stride > 2 || ntaps < 1 || groups < 1 || (groups > 1 && groups != ntaps))
    return fail(PVAE_ERR_INVALID, "pvae_conv_wgrad: bad shape");
  WgradParams p = {};
  {
    ConvParams tmp = {};
    if (int rc = fill_taps(tmp, ntaps, tap_dy, tap_dx, tap_widx)) return rc;
    p.ntaps = ntaps;
    for (int t = 0; t < ntaps; ++t) p.taps[t] = tmp.taps[t];
  }
  p.Ci = Ci, p.Co = Co, p.OH = OH, p.OW = OW, p.NI = NI, p.stride = stride;
  p.lw = std::min(5, ilog2_ceil(OW));
  p.lh = std::min(5 - p.lw, ilog2_ceil(OH));
  const int pw = 1 << p.lw, ph = 1 << p.lh, pi = 32 >> (p.lw + p.lh);
  p.tiles_w = (OW + pw - 1) / pw, p.tiles_h = (OH + ph - 1) / ph;
  p.num_kb = p.tiles_w * p.tiles_h * ((NI + pi - 1) / pi);
  const int BN = wgrad_bn(Co);
  const int row_blocks = (ntaps * (Ci / 32) + 1 + 3) / 4;  // + the all-ones slab (bias gradient)
  int splits = std::min(p.num_kb, wgrad_max_splits(Ci, Co, ntaps));
  p.kb_per_split = (p.num_kb + splits - 1) / splits;
  splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.out = ws;
  CUtensorMap mx, mg;
  if (int rc = make_nhwc_map(&mx, in, NI, IH, IW, Ci, pw, ph, pi, stride, 1)) return rc;
  if (int rc = make_nhwc_map(&mg, g_out, NI, OH, OW, Co, pw, ph, pi, 1, 1)) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc;
  if (precise) {
    rc = BN == 128 ? launch_wgrad<128, true, 3>(mx, mg, p, row_blocks, splits, s)
                   : (BN == 64 ? launch_wgrad<64, true, 4>(mx, mg, p, row_blocks, splits, s)
                               : launch_wgrad<32, true, 4>(mx, mg, p, row_blocks, splits, s));
  } else {
    rc = BN == 128 ? launch_wgrad<128, false, 6>(mx, mg, p, row_blocks, splits, s)
                   : (BN == 64 ? launch_wgrad<64, false, 6>(mx, mg, p, row_blocks, splits, s)
                               : launch_wgrad<32, false, 6>(mx, mg, p, row_blocks, splits, s));
  }
  if (rc) return rc;
  const int R1 = ntaps * Ci + 1;
  float* gT = ws + static_cast<long long>(splits) * R1 * Co;
  if (int rc2 = cuda_ok("conv_wgrad_reduce_kernel", launch_pdl(conv_wgrad_reduce_kernel, dim3(Co / 32, (R1 + 15) / 16), dim3(256), 0, s,
                                                               static_cast<const float*>(ws), splits, gT, R1, Co)))
    return rc2;
  return cuda_ok("conv_wgrad_finish_kernel", launch_pdl(conv_wgrad_finish_kernel, dim3((Co_real + 7) / 8), dim3(256), 0, s,
                                                        static_cast<const float*>(gT), weight, lognorm, g_weight, g_lognorm, g_bias,
                                                        Co_real, Ci_real, Ci, ntaps, groups));
}
