// Dense contractions of the lin|lin P-VAE on the 5th-generation tensor cores (sm_100a):
//   C (M,N) = op(A) * op(B)^T-like products, fp32 in HBM, TF32 tcgen05.mma, fp32 accumulators in TMEM.
//
// One 128 x BN output tile per CTA (optionally one K-split of it), warp-specialised:
//   warp 0      TMA producer     cp.async.bulk.tensor (128B swizzle) into a 4-stage smem ring
//   warp 1      MMA issuer       one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) and
//                                commits to the ring's "empty" barriers / the accumulator-full barrier
//   warps 2-5   epilogue         tcgen05.ld of the accumulator (32 lanes x 32 columns per instruction),
//                                128-byte row segments stored straight to global memory
// Precision.  kind::tf32 reads 10 mantissa bits of each operand and ignores the 13 below (truncation -
// measured, tools/probe_tf32_rounding.py).  In the default "3xTF32" mode the four epilogue warps compute, for
// every landed stage, lo = x - trunc_tf32(x) (exact) into a second plane; the raw tile serves as hi, and the
// issuer accumulates hi*hi + lo*hi + hi*lo: the dropped lo*lo term and the truncation of lo are ~2^-22
// relative, i.e. fp32-grade products with fp32 accumulation in TMEM.
// Both operand majors are supported without any transposed copy in HBM: a K-major operand is a
// (rows, Kd) row-major matrix loaded as one {32 x 128} box per stage; an MN-major operand is a (Kd, rows)
// row-major matrix loaded as BN/32 (or 4) {32 x 32} boxes whose 1024-byte swizzle atoms are walked along
// K by the UMMA descriptor.  That covers forward (x W^T), dgrad (g W) and wgrad (g^T x) of a linear layer.
//
// Replaces: F.linear in Linear.forward base/common.py:409-414 (fc_enc main/vae.py:51, fc_dec :59-60)
// and the three GEMMs autograd derives from each of them.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <array>
#include <map>
#include <mutex>
#include <tuple>

#include "../../include/pvae_b200.h"
#include "pvae_device.cuh"
#include "pvae_host.h"
#include "tc_common.cuh"

namespace pvae {

struct GemmParams {
  float* C;              // (M,N) row-major, or the split-K workspace (splits, M, N)
  int M, N, Kd;
  int k_per_split;       // multiple of kBK
  int a_kmajor, b_kmajor;
  long long split_stride;  // M * N when splits > 1
  // residual epilogue (decoder GEMM of the training step): instead of C = y the kernel writes
  // G = g_scale * (y - X) and row_part[m, n_tile] = sum over the tile's columns of (y - X)^2
  const float* X;
  float* G;
  float* G_lo;  // optional: TF32 low plane of G for the GEMMs that consume it
  float* row_part;
  float g_scale;
  unsigned long long* dbg_ts;  // optional: %globaltimer stamps of the tile (0,0,0) and extremes over all CTAs (pvae_set_gemm_timeline)
};

__device__ __forceinline__ void gstamp(const GemmParams& p, int idx, bool all = false) {
  if (p.dbg_ts == nullptr) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  if (all) {
    if (idx & 1) atomicMax(p.dbg_ts + idx, t); else atomicMin(p.dbg_ts + idx, t);
  } else if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    p.dbg_ts[idx] = t;
  }
}

// SPLIT: 0 = single-pass TF32; 1 = 3xTF32, the low parts computed in the kernel by the epilogue warps (operands as they
// are in HBM); 2 = 3xTF32, the low parts arrive as their own planes by TMA (written once by whoever produced the operand:
// sampler / residual / Adamax epilogues), so the main loop is pure TMA -> tcgen05.mma.  Modes 1 and 2 issue the same MMAs
// in the same order on the same bits: bit-identical results.
// Warps 0 / 1: TMA producer / MMA issuer; warps 2..9: EIGHT epilogue warps - two per TMEM lane quarter, each half of the
// tile's columns (the epilogue of a 128 x 128 tile took 2.2 of the kernel's 8.6 us with four, tools/gemm_timeline.py).
constexpr int kEpiWarps = 8;
constexpr int kDenseThreads = 64 + 32 * kEpiWarps;

template <int BN, int SPLIT, int STAGES>
struct GemmCfg {
  static constexpr uint32_t kABytes = kBM * kBK * 4;  // 16 KB
  static constexpr uint32_t kBBytes = BN * kBK * 4;
  static constexpr uint32_t kTileBytes = kABytes + kBBytes;             // what TMA writes per stage
  static constexpr uint32_t kStageBytes = SPLIT ? 2 * kTileBytes : kTileBytes;  // + the lo planes
  static constexpr int kStages = STAGES;
  static constexpr int kSmem = kStages * kStageBytes + 1024 + 256;
  static constexpr int kCtasPerSm = kSmem <= 110 * 1024 ? 2 : 1;
};

template <int BN, int SPLIT, int STAGES>
__global__ void __launch_bounds__(kDenseThreads, GemmCfg<BN, SPLIT, STAGES>::kCtasPerSm)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_a_lo, const __grid_constant__ CUtensorMap tma_b_lo, const GemmParams p) {
  using Cfg = GemmCfg<BN, SPLIT, STAGES>;
  constexpr uint32_t kABytes = Cfg::kABytes, kTileBytes = Cfg::kTileBytes, kStageBytes = Cfg::kStageBytes;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* conv_bar = empty_bar + kStages;
  uint64_t* accum_bar = conv_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * p.k_per_split;
  const int k_end = min(p.Kd, k_begin + p.k_per_split);
  const int num_kb = (k_end - k_begin + kBK - 1) / kBK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
    if (SPLIT == 2) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a_lo)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b_lo)) : "memory");
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&conv_bar[s], kDenseThreads - 64);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // TMEM: BN fp32 accumulator columns x 128 lanes; 3xTF32 keeps the hi*hi sum and the (2^-10 smaller)
  // cross terms in separate accumulators so that the tensor core's truncating fp32 adds act on each alone
  constexpr uint32_t kTmemCols = SPLIT ? 2 * BN : BN;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlaps the previous kernel's tail
  pdl_launch_dependents();
  if (threadIdx.x == 0) gstamp(p, 0), gstamp(p, 700, true);
  pdl_wait();
  if (threadIdx.x == 0) gstamp(p, 1), gstamp(p, 701, true);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t round = kb / kStages;
        mbar_wait(&empty_bar[s], (round & 1) ^ 1);
        uint8_t* sa = smem + s * kStageBytes;
        uint8_t* sb = sa + kABytes;
        mbar_expect_tx(&full_bar[s], SPLIT == 2 ? 2 * kTileBytes : kTileBytes);
        gstamp(p, 10 + kb);
        const int k0 = k_begin + kb * kBK;
#pragma unroll
        for (int plane = 0; plane < (SPLIT == 2 ? 2 : 1); ++plane) {  // plane 1: the TF32 low parts, kTileBytes above
          const CUtensorMap* ma = plane ? &tma_a_lo : &tma_a;
          const CUtensorMap* mb = plane ? &tma_b_lo : &tma_b;
          uint8_t* da = sa + plane * kTileBytes;
          uint8_t* db = sb + plane * kTileBytes;
          if (p.a_kmajor) {
            tma_load_2d(ma, &full_bar[s], da, k0, m0);  // box {32 k, 128 rows}
          } else {
#pragma unroll
            for (int i = 0; i < kBM / 32; ++i) tma_load_2d(ma, &full_bar[s], da + i * 4096, m0 + 32 * i, k0);  // {32 m, 32 k}
          }
          if (p.b_kmajor) {
            tma_load_2d(mb, &full_bar[s], db, k0, n0);  // box {32 k, BN rows}
          } else {
#pragma unroll
            for (int i = 0; i < BN / 32; ++i) tma_load_2d(mb, &full_bar[s], db + i * 4096, n0 + 32 * i, k0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(kBM, BN, !p.a_kmajor, !p.b_kmajor);
      // K-major (SWIZZLE_128B): 8-row groups 1024 B apart (SBO), K advances 32 B inside the swizzle row.
      // MN-major (SWIZZLE_128B, 32-byte atoms): 32-wide slabs 4096 B apart (LBO), 4-deep K groups 512 B
      // apart (SBO), one MMA (K = 8) spans two of them, K advances 1024 B.
      const uint32_t a_lbo = p.a_kmajor ? 16 : 4096, b_lbo = p.b_kmajor ? 16 : 4096;
      const uint32_t a_sbo = p.a_kmajor ? 1024 : 512, b_sbo = p.b_kmajor ? 1024 : 512;
      const uint32_t a_lay = p.a_kmajor ? 2 : 1, b_lay = p.b_kmajor ? 2 : 1;
      const uint32_t a_step = p.a_kmajor ? kUmmaK * 4 : 1024, b_step = p.b_kmajor ? kUmmaK * 4 : 1024;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t round = kb / kStages;
        mbar_wait(SPLIT == 1 ? &conv_bar[s] : &full_bar[s], round & 1);
        gstamp(p, 30 + kb);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem_u32(smem + s * kStageBytes), sb = sa + kABytes;
#pragma unroll
        for (int k = 0; k < kBK / kUmmaK; ++k) {
          const uint64_t da = make_smem_desc(sa + k * a_step, a_lbo, a_sbo, a_lay);
          const uint64_t db = make_smem_desc(sb + k * b_step, b_lbo, b_sbo, b_lay);
          umma_tf32(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          if (SPLIT) {  // + lo(A) hi(B) + hi(A) lo(B); the lo planes sit kTileBytes above the hi planes
            const uint64_t dal = make_smem_desc(sa + kTileBytes + k * a_step, a_lbo, a_sbo, a_lay);
            const uint64_t dbl = make_smem_desc(sb + kTileBytes + k * b_step, b_lbo, b_sbo, b_lay);
            umma_tf32(tmem_base + BN, dal, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_tf32(tmem_base + BN, da, dbl, idesc, 1u);
          }
        }
        umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs have read it
        gstamp(p, 50 + kb);
      }
      umma_commit(accum_bar);  // accumulator complete
    }
  } else {
    if (SPLIT == 1) {
      // ===== operand splitter: lo = x - trunc_tf32(x) into the second plane =====
      const int ct = threadIdx.x - 64;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        mbar_wait(&full_bar[s], (kb / kStages) & 1);
        uint4* tile = reinterpret_cast<uint4*>(smem + s * kStageBytes);
        uint4* tile_lo = reinterpret_cast<uint4*>(smem + s * kStageBytes + kTileBytes);
#pragma unroll 4
        for (int i = ct; i < static_cast<int>(kTileBytes / 16); i += kDenseThreads - 64) {
          const uint4 v = tile[i];
          uint4 hi, lo;
          hi.x = v.x & 0xFFFFE000u, hi.y = v.y & 0xFFFFE000u, hi.z = v.z & 0xFFFFE000u, hi.w = v.w & 0xFFFFE000u;
          lo.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(hi.x));
          lo.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(hi.y));
          lo.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(hi.z));
          lo.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(hi.w));
          tile_lo[i] = lo;  // the tensor core ignores the 13 low mantissa bits, so the raw tile already is `hi`
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> tensor-core reads
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&conv_bar[s])) : "memory");
      }
    }
    // ===== epilogue: TMEM -> registers -> global =====
    mbar_wait(accum_bar, 0);
    if (threadIdx.x == 64) gstamp(p, 70), gstamp(p, 703, true);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int quarter = warp & 3;  // a warp may only read TMEM lanes [32 * (warp % 4), +32)
    const int chalf = (warp - 2) >> 2;  // which half of the tile's columns
    float* cbase = p.C + static_cast<long long>(blockIdx.z) * p.split_stride;
    const bool residual = p.X != nullptr;
    float rowacc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c0);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      if (SPLIT) {  // + the cross-term accumulator
        uint32_t w[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]),
              "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]), "=r"(w[16]),
              "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]), "=r"(w[24]),
              "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
            : "r"(taddr + BN));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
      } else {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      }
      // transpose through shared memory (the operand ring is idle by now) so that a store instruction
      // writes four full 128-byte row segments instead of 32 scattered 16-byte pieces
      float* stg = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 33);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = __uint_as_float(v[j]);
      __syncwarp();
      const int sub = lane >> 3, col = 4 * (lane & 7);
      const int n = n0 + c0 + col;
      if (!residual) {
#pragma unroll
        for (int rr = 0; rr < 32; rr += 4) {
          const int row = rr + sub;
          const int mm = m0 + quarter * 32 + row;
          const float* src = stg + row * 33 + col;
          if (mm < p.M && n < p.N)  // N is a multiple of 4
            *reinterpret_cast<float4*>(cbase + static_cast<long long>(mm) * p.N + n) = make_float4(src[0], src[1], src[2], src[3]);
        }
      } else {
        float4 xv[8];  // all eight row segments of X in flight before any is used
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int mm = m0 + quarter * 32 + 4 * i + sub;
          xv[i] = (mm < p.M && n < p.N) ? ldg4(p.X + static_cast<long long>(mm) * p.N + n) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = 4 * i + sub;
          const int mm = m0 + quarter * 32 + row;
          const float* src = stg + row * 33 + col;
          float sq = 0.f;
          if (mm < p.M && n < p.N) {
            const float4 r = make_float4(src[0] - xv[i].x, src[1] - xv[i].y, src[2] - xv[i].z, src[3] - xv[i].w);
            const float4 gv = make_float4(p.g_scale * r.x, p.g_scale * r.y, p.g_scale * r.z, p.g_scale * r.w);
            stg4(p.G + static_cast<long long>(mm) * p.N + n, gv);
            if (p.G_lo != nullptr) stg4(p.G_lo + static_cast<long long>(mm) * p.N + n, tf32_lo4f(gv));
            sq = (r.x * r.x + r.y * r.y) + (r.z * r.z + r.w * r.w);
          }
          sq += __shfl_xor_sync(0xffffffffu, sq, 1);  // the 8 lanes that share a row
          sq += __shfl_xor_sync(0xffffffffu, sq, 2);
          sq += __shfl_xor_sync(0xffffffffu, sq, 4);
          rowacc[i] += sq;
        }
      }
    }
    if (residual && (lane & 7) == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int mm = m0 + quarter * 32 + 4 * i + (lane >> 3);
        if (mm < p.M) p.row_part[static_cast<long long>(mm) * (2 * gridDim.x) + 2 * blockIdx.x + chalf] = rowacc[i];
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 64) gstamp(p, 71), gstamp(p, 705, true);
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// fixed-order sum of the split-K partials (loads of eight splits in flight per thread)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, float* __restrict__ C, long long n4,
                                                            int splits) {
  pdl_launch_dependents();
  pdl_wait();
  const float4* w = reinterpret_cast<const float4*>(ws);
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += gridDim.x * 256ll) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = 0;
    for (; s + 8 <= splits; s += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(w + (s + u) * n4 + i);
#pragma unroll
      for (int u = 0; u < 8; ++u) a.x += v[u].x, a.y += v[u].y, a.z += v[u].z, a.w += v[u].w;
    }
    for (; s < splits; ++s) {
      const float4 b = __ldg(w + s * n4 + i);
      a.x += b.x, a.y += b.y, a.z += b.z, a.w += b.w;
    }
    reinterpret_cast<float4*>(C)[i] = a;
  }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &st) == cudaSuccess &&
        st == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

int make_tensor_map(CUtensorMap* out, const float* base, int rank, const long long* dims, const long long* strides_bytes,
                    const int* box, const int* estr, int swizzle) {
  if (rank < 2 || rank > 4) return fail(PVAE_ERR_INVALID, "make_tensor_map: rank %d", rank);
  typedef std::tuple<const void*, int, std::array<long long, 4>, std::array<long long, 3>, std::array<int, 4>, std::array<int, 4>, int> Key;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  std::array<long long, 4> kd = {0, 0, 0, 0};
  std::array<long long, 3> ks = {0, 0, 0};
  std::array<int, 4> kb = {0, 0, 0, 0}, ke = {0, 0, 0, 0};
  for (int i = 0; i < rank; ++i) kd[i] = dims[i], kb[i] = box[i], ke[i] = estr[i];
  for (int i = 0; i + 1 < rank; ++i) ks[i] = strides_bytes[i];
  const Key key(base, rank, kd, ks, kb, ke, swizzle);
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return PVAE_OK;
  }
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(PVAE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdims[4], gstr[3];
  cuuint32_t gbox[4], gest[4];
  for (int i = 0; i < rank; ++i) gdims[i] = static_cast<cuuint64_t>(dims[i]), gbox[i] = static_cast<cuuint32_t>(box[i]), gest[i] = static_cast<cuuint32_t>(estr[i]);
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = static_cast<cuuint64_t>(strides_bytes[i]);
  const CUresult rc = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, static_cast<cuuint32_t>(rank), const_cast<float*>(base), gdims, gstr, gbox,
                         gest, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return fail(PVAE_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(rc));
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return PVAE_OK;
}

// 2-D fp32 row-major (rows, cols) tensor, box {box_cols (inner), box_rows}, 128-byte swizzle with 16-byte
// atoms (K-major operands) or 32-byte atoms (MN-major operands), OOB = 0.
static int make_map(CUtensorMap* out, const float* base, long long rows, long long cols, int box_cols, int box_rows,
                    bool atom32 = false) {
  const long long dims[2] = {cols, rows}, strides[1] = {cols * 4};
  const int box[2] = {box_cols, box_rows}, estr[2] = {1, 1};
  return make_tensor_map(out, base, 2, dims, strides, box, estr, atom32 ? 1 : 0);
}

template <int BN, int SPLIT, int STAGES>
static int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mal, const CUtensorMap& mbl, const GemmParams& p,
                       int splits, cudaStream_t s) {
  constexpr int smem = GemmCfg<BN, SPLIT, STAGES>::kSmem;
  static bool configured = false;
  if (!configured) {
    if (int rc = cuda_ok("cudaFuncSetAttribute", cudaFuncSetAttribute(gemm_tf32_kernel<BN, SPLIT, STAGES>,
                                                                     cudaFuncAttributeMaxDynamicSharedMemorySize, smem)))
      return rc;
    configured = true;
  }
  dim3 grid((p.N + BN - 1) / BN, (p.M + kBM - 1) / kBM, splits);
  return cuda_ok("gemm_tf32_kernel", launch_pdl(gemm_tf32_kernel<BN, SPLIT, STAGES>, grid, dim3(kDenseThreads), smem, s, ma, mb, mal, mbl, p));
}

int launch_splitk_reduce(const float* ws, float* C, long long n, int splits, cudaStream_t s) {
  const long long n4 = n / 4;
  const unsigned grid = static_cast<unsigned>(std::min<long long>((n4 + 255) / 256, 148ll * 16));
  return cuda_ok("splitk_reduce_kernel", launch_pdl(splitk_reduce_kernel, dim3(grid), dim3(256), 0, s, ws, C, n4, splits));
}

static int g_gemm_variant = 0;
static unsigned long long* g_gemm_ts = nullptr;  // tuning: 0 = automatic, 1 = wide tiles / 1 CTA per SM, 2 = narrow tiles / 2 CTAs per SM

}  // namespace pvae

using namespace pvae;

extern "C" int64_t pvae_gemm_ws_floats(int64_t M, int64_t N, int splits) { return splits > 1 ? splits * M * N : 0; }

extern "C" int pvae_gemm_effective_splits(int64_t Kd, int splits) {
  if (splits < 1) splits = 1;
  const long long kblocks = (Kd + kBK - 1) / kBK;
  if (splits > kblocks) splits = static_cast<int>(kblocks);
  const long long per = (kblocks + splits - 1) / splits;  // every split owns at least one K block
  return static_cast<int>((kblocks + per - 1) / per);
}

extern "C" int pvae_set_gemm_timeline(uint64_t* ts_dev) {
  g_gemm_ts = reinterpret_cast<unsigned long long*>(ts_dev);
  return PVAE_OK;
}

extern "C" int pvae_set_gemm_variant(int v) {
  if (v < 0 || v > 2) return fail(PVAE_ERR_INVALID, "pvae_set_gemm_variant: %d not in [0, 2]", v);
  g_gemm_variant = v;
  return PVAE_OK;
}

namespace pvae {
// A_lo / B_lo: the operands' TF32 low planes (x - trunc_tf32(x), same shape and layout as A / B), or both NULL: then the
// kernel computes them itself (precise) - the results are the same bits either way.
int gemm_impl(const float* A, const float* A_lo, const float* B, const float* B_lo, float* C, int64_t M, int64_t N, int64_t Kd, int a_kmajor,
              int b_kmajor, int precise, int splits, float* splitk_ws, const float* X, float* G, float* G_lo, float* row_part,
              float g_scale, int force_bn, void* stream) {
  if (M <= 0 || N <= 0 || Kd <= 0) return fail(PVAE_ERR_INVALID, "pvae_gemm_tf32: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)Kd);
  if (M % 4 || N % 4 || Kd % 4)
    return fail(PVAE_ERR_UNSUPPORTED, "pvae_gemm_tf32: M=%lld N=%lld K=%lld must be multiples of 4 (16-byte TMA strides)",
                (long long)M, (long long)N, (long long)Kd);
  if (!A || !B || (!C && !X && !(splits > 1 && splitk_ws))) return fail(PVAE_ERR_INVALID, "pvae_gemm_tf32: null pointer");
  if ((A_lo == nullptr) != (B_lo == nullptr)) return fail(PVAE_ERR_INVALID, "pvae_gemm_tf32: low planes must be given for both operands or none");
  if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(C) | reinterpret_cast<uintptr_t>(A_lo) |
       reinterpret_cast<uintptr_t>(B_lo) | reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(G) | reinterpret_cast<uintptr_t>(G_lo)) & 15)
    return fail(PVAE_ERR_INVALID, "pvae_gemm_tf32: pointers must be 16-byte aligned");
  if (splits < 1) splits = 1;
  const long long kblocks = (Kd + kBK - 1) / kBK;
  if (splits > kblocks) splits = static_cast<int>(kblocks);
  if (splits > 1 && !splitk_ws) return fail(PVAE_ERR_INVALID, "pvae_gemm_tf32: split-K needs a workspace");
  {  // every split must own at least one K block
    const long long per = (kblocks + splits - 1) / splits;
    splits = static_cast<int>((kblocks + per - 1) / per);
  }
  const bool planes = precise && A_lo != nullptr;
  // 128-wide tiles unless that leaves most of the 148 SMs without a CTA
  int BN = (N <= 64 || ((N + 127) / 128) * ((M + kBM - 1) / kBM) * splits < 96) ? 64 : 128;
  const bool narrow = g_gemm_variant == 2;  // measured slower than wide tiles on these shapes; kept as a knob
  if (narrow) BN = 64;
  if (force_bn) BN = force_bn;
  CUtensorMap ma, mb, mal, mbl;
  // K-major operand: (rows, Kd) row-major, box {32 k, rows_per_tile}; MN-major: (Kd, rows) row-major, box {32 rows, 32 k}
  if (int rc = a_kmajor ? make_map(&ma, A, M, Kd, kBK, kBM) : make_map(&ma, A, Kd, M, 32, kBK, true)) return rc;
  if (int rc = b_kmajor ? make_map(&mb, B, N, Kd, kBK, BN) : make_map(&mb, B, Kd, N, 32, kBK, true)) return rc;
  mal = ma, mbl = mb;
  if (planes) {
    if (int rc = a_kmajor ? make_map(&mal, A_lo, M, Kd, kBK, kBM) : make_map(&mal, A_lo, Kd, M, 32, kBK, true)) return rc;
    if (int rc = b_kmajor ? make_map(&mbl, B_lo, N, Kd, kBK, BN) : make_map(&mbl, B_lo, Kd, N, 32, kBK, true)) return rc;
  }
  GemmParams p;
  p.C = splits > 1 ? splitk_ws : C;
  p.M = static_cast<int>(M), p.N = static_cast<int>(N), p.Kd = static_cast<int>(Kd);
  p.k_per_split = static_cast<int>(((kblocks + splits - 1) / splits) * kBK);
  p.a_kmajor = a_kmajor, p.b_kmajor = b_kmajor;
  p.split_stride = splits > 1 ? M * N : 0;
  p.X = X, p.G = G, p.G_lo = G_lo, p.row_part = row_part, p.g_scale = g_scale;
  p.dbg_ts = g_gemm_ts;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc;
  if (planes) {
    rc = BN == 64 ? launch_gemm<64, 2, 4>(ma, mb, mal, mbl, p, splits, s) : launch_gemm<128, 2, 3>(ma, mb, mal, mbl, p, splits, s);
  } else if (narrow) {  // 2 CTAs per SM: one CTA's TMA latency hides behind the other's split + MMA work
    rc = precise ? launch_gemm<64, 1, 2>(ma, mb, mal, mbl, p, splits, s) : launch_gemm<64, 0, 4>(ma, mb, mal, mbl, p, splits, s);
  } else if (precise) {
    rc = BN == 64 ? launch_gemm<64, 1, 4>(ma, mb, mal, mbl, p, splits, s) : launch_gemm<128, 1, 3>(ma, mb, mal, mbl, p, splits, s);
  } else {
    rc = BN == 64 ? launch_gemm<64, 0, 6>(ma, mb, mal, mbl, p, splits, s) : launch_gemm<128, 0, 6>(ma, mb, mal, mbl, p, splits, s);
  }
  if (rc) return rc;
  if (splits > 1) {
    if (!C) return PVAE_OK;  // partials only: the caller reduces them (pvae_adamax_from_partials)
    return launch_splitk_reduce(splitk_ws, C, M * N, splits, s);
  }
  return PVAE_OK;
}
}  // namespace pvae

extern "C" int pvae_gemm_tf32(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t Kd, int a_kmajor,
                              int b_kmajor, int precise, int splits, float* splitk_ws, void* stream) {
  return gemm_impl(A, nullptr, B, nullptr, C, M, N, Kd, a_kmajor, b_kmajor, precise, splits, splitk_ws, nullptr, nullptr, nullptr, nullptr, 0.f,
                   0, stream);
}

extern "C" int pvae_gemm_tf32_planes(const float* A, const float* A_lo, const float* B, const float* B_lo, float* C, int64_t M, int64_t N,
                                     int64_t Kd, int a_kmajor, int b_kmajor, int splits, float* splitk_ws, void* stream) {
  if (!A_lo || !B_lo) return fail(PVAE_ERR_INVALID, "pvae_gemm_tf32_planes: null low plane");
  return gemm_impl(A, A_lo, B, B_lo, C, M, N, Kd, a_kmajor, b_kmajor, 1, splits, splitk_ws, nullptr, nullptr, nullptr, nullptr, 0.f, 0, stream);
}

// 64-wide tiles: the decoder output is only D = 256 wide, so this keeps 4 * M / 128 CTAs in flight; each tile's two
// epilogue column halves write their own partial squared error
extern "C" int64_t pvae_gemm_residual_parts(int64_t N) { return 2 * ((N + 63) / 64); }

extern "C" int pvae_gemm_tf32_residual(const float* A, const float* B, const float* X, float* G, float* row_part, int64_t M,
                                       int64_t N, int64_t Kd, float g_scale, int precise, void* stream) {
  if (!X || !G || !row_part) return fail(PVAE_ERR_INVALID, "pvae_gemm_tf32_residual: null pointer");
  return gemm_impl(A, nullptr, B, nullptr, nullptr, M, N, Kd, 1, 1, precise, 1, nullptr, X, G, nullptr, row_part, g_scale, 64, stream);
}
