// Fused forward of the lin|lin P-VAE after the encoder GEMM (BASELINE config 2: "fused rsample + KL + decoder GEMM"):
//
//   h (B,K) -> softclamps, rate, KL row sums            main/vae.py:393-415, 446-450
//           -> z = Poisson(rate).rsample(), dz_acc      base/distributions.py:55-81
//           -> y = z W_dec^T                            main/vae.py:59-60   (tcgen05, z read from SHARED memory)
//           -> g_y = 2 (y - x) / B, sum_d (y - x)^2     main/vae.py:68-72
//
// in ONE persistent kernel.  The sampler is instruction-issue bound on the SIMT pipes and leaves the tensor pipe idle, the
// decoder GEMM is the opposite: here they share the SM.  A CTA owns a unit of <= 28 rows (B / #SMs of them, so that all
// 148 SMs carry the same sampler load) and ALL K latents of those rows, walked in batches of 256 latents (8 k-blocks):
//
//   warps 0..20  sampler     seven independent groups of three warps; a group owns 4 rows of the unit and runs, per batch,
//                            the schedule of the stand-alone sampler on them: uniform prologue (rates parked in shared memory;
//                            chains whose rate predicts more than a few events are listed), lane groups of 8 walk the listed
//                            long chains 8 events per step, then a work pool over the other quads (k-block major).  Groups
//                            only synchronise among themselves (named barriers): while one group drains the tail of its pool
//                            the others keep the SM's issue slots busy - heavy-tailed chain lengths make CTA-wide barriers
//                            expensive (measured: 65 vs 39 us for the sampler part).  A finished quad is stored to global
//                            memory (z, its TF32 low plane, dz_acc - the backward needs them) AND into the 128-byte-swizzled
//                            K-major operand tile of its k-block in shared memory (a ring of 8 k-block slots, released by the
//                            tensor core as it consumes them); then the lane arrives on that k-block's mbarrier.
//   warp 22      TMA         one thread streams W_dec through a 3-stage ring: one plane (hi, then lo) of one k-block with all
//                            D rows ({32 k, D rows} box, 32 KB) per stage
//   warp 21      MMA         one thread waits for a k-block of z and issues 3xTF32 tcgen05.mma for y = z W_dec^T: A = the z tile
//                            (M = 128 of which the unit's <= 28 rows are live; the rows above alias whatever follows the 4 KB
//                            tile in shared memory and their accumulator lanes are never read), B = the W_dec tile (N = D),
//                            accumulators hi / cross in TMEM (lane = row) - twelve MMAs per k-block.  (A tcgen05.mma costs
//                            ~0.09 us whatever its shape: the transposed product, N = 32 without padding, needed 24.)
//   warps 0,4,8,12           epilogue of a unit (the warps that can read TMEM lanes 0..31): TMEM -> registers, residual
//                            against x, g_y (+ low plane), squared error per row
//
// Full rows per CTA mean no split-K partials: y is complete in TMEM, so the reconstruction residual is fused too.
// Per-chain arithmetic, Philox counters and exit rules are those of sampler.cu (sampler_device.cuh): a quad finished by
// the pool is bit-identical to pvae_sampler_fwd, one finished by a lane group within a few ulp (prefix-sum order).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pvae_b200.h"
#include "pvae_device.cuh"
#include "pvae_host.h"
#include "sampler_device.cuh"
#include "tc_common.cuh"

namespace pvae {

constexpr int kFsGroups = 7;                       // independent sampler groups
constexpr int kFsGroupWarps = 3;
constexpr int kFsGroupLanes = kFsGroupWarps * 32;  // three warps each (704 threads in all: 80 registers per thread)
constexpr int kFsWarps = kFsGroups * kFsGroupWarps; // sampler warps
constexpr int kFsLanes = kFsWarps * 32;
constexpr int kFsThreads = kFsLanes + 64;          // + the TMA warp and the MMA warp
constexpr int kFsMaxUnitRows = 28;                 // rows per unit: at most 4 per group
constexpr int kFsGroupRows = 4;
constexpr int kFsKb = 8;                           // k-blocks (32 latents) per batch
constexpr int kFsBatchK = kFsKb * 32;              // 256 latents per batch
constexpr int kFsGPool = kFsGroupRows * kFsKb * 8; // 256 pool entries (quads) per group and batch
constexpr int kFsMaxK = 1024;
constexpr int kFsWStages = 3;                        // 32 KB each (one plane of a k-block of W_dec, all D = 256 rows)
constexpr int kFsZSlots = 8;                         // operand-tile ring: k-blocks a group may run ahead of the tensor core
constexpr uint32_t kFsZTile = 32 * 128;            // one plane of a 32-row x 32-latent operand tile
constexpr uint32_t kFsWTile = 256 * 128;           // one plane of a D-row x 32-latent W_dec tile = one ring stage (D <= 256)

constexpr uint32_t kOffZ = 0;                                           // [slot][plane][32 rows x 128 B]
constexpr uint32_t kOffW = kOffZ + kFsZSlots * 2 * kFsZTile;            // 64 KB
constexpr uint32_t kOffRate = kOffW + kFsWStages * kFsWTile;            // + 96 KB
constexpr uint32_t kOffList = kOffRate + kFsGroups * 4 * kFsGPool * 4;  // + 28 KB   [group][4][256]
constexpr uint32_t kOffKl = kOffList + kFsGroups * kFsGPool * 2;        // + 3.75 KB [group][256]: listed from the front, late parks from the back
constexpr uint32_t kOffElr = kOffKl + kFsGroups * 32 * kFsGroupRows * 4;  // + 3.75 KB [group][k-block][row]
constexpr uint32_t kOffRecon = kOffElr + kFsMaxK * 4;                   // + 4 KB
constexpr uint32_t kOffMax = kOffRecon + 2 * 8 * 32 * 4;                // + 2 KB
constexpr uint32_t kOffCnt = kOffMax + 128;                             // [group][2 sets][8]
constexpr uint32_t kOffBar = kOffCnt + kFsGroups * 16 * 4;
constexpr uint32_t kFsSmem = kOffBar + 256 + 1024;                      // barriers, alignment slack
constexpr uint32_t kFsTmemCols = 512;                                   // D columns hi*hi + D columns cross terms (lane = row)
static_assert(kFsSmem <= 232448, "fused forward: shared memory");

struct FusedParams {
  SamplerParams sp;
  const float* x;   // (B,D)
  float* g_y;       // (B,D)  2 (y - x) / B_global
  float* g_y_lo;    // optional TF32 low plane of g_y
  float* recon;     // (B)    sum_d (y - x)^2
  int D;
  int rows_unit;    // rows per unit (<= 30)
  int n_units;
  float g_scale;
  float t_need;     // arrival time beyond which a chain can end; rate * t_need ~ events a chain walks
  float park_events;
  int debug;        // timing experiments: 1 = no W_dec loads / MMAs, 2 = hi x hi product only
  unsigned long long* dbg_ts;  // optional (1024 words): globaltimer stamps of CTA 0's phases (pvae_set_fused_timeline)
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void stamp(const FusedParams& fp, int idx) {
  if (fp.dbg_ts != nullptr && blockIdx.x == 0 && idx < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    fp.dbg_ts[idx] = t;
  }
}
// extreme over all CTAs: idx even = earliest, odd = latest
__device__ __forceinline__ void stamp_all(const FusedParams& fp, int idx) {
  if (fp.dbg_ts != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    if (idx & 1) atomicMax(fp.dbg_ts + idx, t); else atomicMin(fp.dbg_ts + idx, t);
  }
}
__device__ __forceinline__ void sts4(uint32_t addr, const float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// one warp reads 32 TMEM lanes x 16 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

template <int IND>
__global__ void __launch_bounds__(kFsThreads, 1)
fused_fwd_kernel(const __grid_constant__ CUtensorMap tma_w, const __grid_constant__ CUtensorMap tma_w_lo, const __grid_constant__ FusedParams fp) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const SamplerParams& p = fp.sp;
  float* const s_elr = reinterpret_cast<float*>(smem + kOffElr);
  float* const s_recon = reinterpret_cast<float*>(smem + kOffRecon);
  float* const s_max = reinterpret_cast<float*>(smem + kOffMax);
  uint64_t* const z_full = reinterpret_cast<uint64_t*>(smem + kOffBar);  // [kFsZSlots]: all quads of the k-block in the slot have arrived
  uint64_t* const z_empty = z_full + kFsZSlots;                          // [kFsZSlots]: the tensor core has read the slot
  uint64_t* const w_full = z_empty + kFsZSlots;
  uint64_t* const w_empty = w_full + kFsWStages;
  uint64_t* const acc_full = w_empty + kFsWStages;
  uint64_t* const acc_empty = acc_full + 1;
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = static_cast<int>(p.K), Kq = K >> 2, D = fp.D;
  const int nkb = K / 32, nbatch = K / kFsBatchK;
  const int R = fp.rows_unit;
  const int n_ep = 4;  // epilogue warps (the four with warp % 4 == 0 among the first 16)

  if (warp == kFsWarps) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_w)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_w_lo)) : "memory");
      for (int i = 0; i < kFsZSlots; ++i) mbar_init(&z_full[i], static_cast<uint32_t>(R * 8)), mbar_init(&z_empty[i], 1);
      for (int i = 0; i < kFsWStages; ++i) mbar_init(&w_full[i], 1), mbar_init(&w_empty[i], 1);
      mbar_init(acc_full, 1);
      mbar_init(acc_empty, static_cast<uint32_t>(n_ep));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kFsTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == kFsWarps + 1) {
    // ===== TMA producer.  A ring stage holds ONE plane of a W_dec tile; the stages of a (k-block, column half) come hi plane
    // first, then lo plane.  (Issuing the loads from the MMA thread made every iteration wait for the completion of the
    // previous iteration's MMAs, ~1 us: the whole decoder then took 54 us per unit instead of hiding under the sampler.)
    if (lane == 0 && fp.debug != 1) {
      int n_my = 0;
      for (int u = blockIdx.x; u < fp.n_units; u += gridDim.x) ++n_my;
      const uint32_t nplanes = fp.debug == 2 ? 1 : 2;
      const uint32_t per_unit = static_cast<uint32_t>(nkb) * nplanes, total = per_unit * n_my;
      const uint32_t stage_bytes = static_cast<uint32_t>(D) * 128;
      for (uint32_t j = 0; j < total; ++j) {
        const uint32_t s = j % kFsWStages, w = j % per_unit, plane = w % nplanes, kbg = w / nplanes;
        mbar_wait(&w_empty[s], ((j / kFsWStages) & 1) ^ 1);
        mbar_expect_tx(&w_full[s], stage_bytes);
        tma_load_2d(plane ? &tma_w_lo : &tma_w, &w_full[s], smem + kOffW + s * kFsWTile, static_cast<int>(kbg) * 32, 0);
      }
    }
  } else if (warp == kFsWarps) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // y (rows, D) += z (rows, 32) W_dec(D, 32)^T: A = the z tile (M = 128 of which the unit's <= 28 rows are live - the rows
      // above alias whatever follows the 4 KB tile in shared memory, their accumulator lanes are never read; an MMA costs
      // the same ~0.09 us whatever its shape, so padding M is free while N = D = 256 halves the instruction count of the
      // transposed product this kernel used first), B = the W_dec tile (N = D rows).  12 MMAs per k-block.
      const uint32_t idesc = make_idesc(128, D, false, false);
      const bool no_mma = fp.debug == 1, hi_only = fp.debug == 2;
      const uint32_t nplanes = hi_only ? 1 : 2;
      stamp(fp, 1);
      uint32_t it = 0, gk = 0, ui = 0;  // gk: k-blocks consumed so far (slot gk % kFsZSlots, phase gk / kFsZSlots)
      for (int u = blockIdx.x; u < fp.n_units; u += gridDim.x, ++ui) {
        if (ui > 0) {  // the previous unit's epilogue has read the accumulators
          mbar_wait(acc_empty, (ui - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        for (int b = 0; b < nbatch; ++b) {
          for (int kb = 0; kb < kFsKb; ++kb, ++gk) {
            const uint32_t slot = gk % kFsZSlots;
            mbar_wait(&z_full[slot], (gk / kFsZSlots) & 1);
            stamp(fp, 100 + gk);
            // the sampler lanes' generic-proxy stores of this k-block, acquired through the barrier, become visible to the
            // tensor core's (async-proxy) operand reads
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t za = smem_u32(smem + kOffZ + slot * 2 * kFsZTile), zl = za + kFsZTile;
            for (uint32_t plane = 0; plane < nplanes && !no_mma; ++plane, ++it) {
              const uint32_t s = it % kFsWStages;
              mbar_wait(&w_full[s], (it / kFsWStages) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t wb = smem_u32(smem + kOffW + s * kFsWTile);
              const uint32_t d_hi = tmem_base, d_x = tmem_base + D;
#pragma unroll
              for (int k = 0; k < kBK / kUmmaK; ++k) {
                const uint64_t dz = make_smem_desc(za + k * 32, 16, 1024, 2), dzl = make_smem_desc(zl + k * 32, 16, 1024, 2);
                const uint64_t dw = make_smem_desc(wb + k * 32, 16, 1024, 2);
                const uint32_t acc = (b > 0 || kb > 0 || k > 0) ? 1u : 0u;
                if (plane == 0) {
                  umma_tf32(d_hi, dz, dw, idesc, acc);               // hi(z) hi(W)
                  if (!hi_only) umma_tf32(d_x, dzl, dw, idesc, acc);  // lo(z) hi(W)
                } else {
                  umma_tf32(d_x, dz, dw, idesc, 1u);                 // hi(z) lo(W)
                }
              }
              umma_commit(&w_empty[s]);  // frees the ring stage once these MMAs have read it
            }
            umma_commit(&z_empty[slot]);  // the operand tiles of this k-block have been read
            stamp(fp, 200 + gk);
          }
        }
        umma_commit(acc_full);
      }
    }
  } else {
    // ===== sampler warps: group g = warps 3 g .. 3 g + 2 =====
    const int g = warp / kFsGroupWarps, gtid = tid - g * kFsGroupLanes;
    float* const s_rate = reinterpret_cast<float*>(smem + kOffRate) + g * (4 * kFsGPool);  // [4][kFsGPool]
    unsigned short* const s_list = reinterpret_cast<unsigned short*>(smem + kOffList) + g * kFsGPool;
    float* const kl_tab = reinterpret_cast<float*>(smem + kOffKl) + g * (32 * kFsGroupRows);  // [k-block][row of the group]
    int* const s_cnt = reinterpret_cast<int*>(smem + kOffCnt) + g * 16;
    const int bar_id = 1 + g;
    // rows of the unit are dealt out evenly: the first R % 7 groups own one more
    const int rows_g = R / kFsGroups + (g < R % kFsGroups ? 1 : 0);
    const int grow0 = g * (R / kFsGroups) + min(g, R % kFsGroups);  // first row of the group inside the unit
    const int pool_n = rows_g * kFsKb * 8;  // entries of a batch, k-block major: entry = (kb * rows_g + r) * 8 + cq
    const int ekb = rows_g * 8;             // entries per k-block
    const uint32_t rmagic = rows_g > 0 ? (65536u + rows_g - 1) / rows_g : 0u;  // x / rows_g == (x * rmagic) >> 16 for x < 256

    uint64_t off = p.offset;
    if (p.offset_dev != nullptr) off += *p.offset_dev;
    const uint32_t off_lo = static_cast<uint32_t>(off), off_hi = static_cast<uint32_t>(off >> 32);
    const int n_a = min(p.phase_a, p.n_exp);
    const float t_exit = p.t_exit, absorb = p.absorb;
    const unsigned lanes_below = (1u << lane) - 1u;
    const uint32_t zst = smem_u32(smem + kOffZ), zbar0 = smem_u32(z_full), zebar0 = smem_u32(z_empty);
    float rmax = 0.f;

    if (tid == 0) stamp(fp, 0), stamp_all(fp, 700), stamp_all(fp, 701);
    for (int k = tid; k < K; k += kFsLanes) s_elr[k] = expf(__ldg(p.log_r + k));
    if (gtid < 16) s_cnt[gtid] = (gtid & 7) == 0 ? kFsGroupLanes : 0;
    named_bar_sync(15, kFsLanes);  // all sampler warps, once

    uint32_t bi = 0, ui = 0;
    for (int u = blockIdx.x; u < fp.n_units; u += gridDim.x, ++ui) {
      const long long row0 = static_cast<long long>(u) * R;           // first row of the unit
      const int nrows = min(rows_g, static_cast<int>(min(static_cast<long long>(R), p.B - row0)) - grow0);  // valid rows of this group (may be <= 0)
      const long long grow = row0 + grow0;                            // first row of the group
      const uint32_t quad0 = static_cast<uint32_t>(((p.row_offset + grow) * p.K) >> 2);
      float* const z0 = p.z + grow * p.K;
      float* const zlo0 = p.z_lo != nullptr ? p.z_lo + grow * p.K : nullptr;
      float* const dz0 = p.dz_out + grow * p.K;

#pragma unroll 1
      for (int b = 0; b < nbatch; ++b, ++bi) {
        int* const cnt = s_cnt + (bi & 1) * 8;  // [0] pool cursor, [1] listed long chains, [2] its cursor, [3] late parks, [4] its cursor
        const uint32_t gk0 = bi * kFsKb;  // the batch's first k-block in the operand-tile ring
        // ---- stage 1: prologue of the batch (uniform, coalesced: 8 lanes = the 32 latents of one row of one k-block)
#pragma unroll 1
        for (int e = gtid; e < pool_n; e += kFsGroupLanes) {
          const int kb = static_cast<int>(((static_cast<uint32_t>(e) >> 3) * rmagic) >> 16), rem = e - kb * ekb, r = rem >> 3, cq = rem & 7;
          float klsum = 0.f;
          bool pre = false;
          if (r < nrows) {
            const int k = b * kFsBatchK + kb * 32 + cq * 4;
            const float4 lr4 = ldg4(p.log_r + k);
            float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.b_enc != nullptr) bias4 = ldg4(p.b_enc + k);
            const float4 hv4 = ldg4(p.h + (grow + r) * p.K + k);
            const float4 elr4 = *reinterpret_cast<const float4*>(s_elr + k);
            const float hv[4] = {hv4.x, hv4.y, hv4.z, hv4.w}, lr[4] = {lr4.x, lr4.y, lr4.z, lr4.w};
            const float bias[4] = {bias4.x, bias4.y, bias4.z, bias4.w}, elr[4] = {elr4.x, elr4.y, elr4.z, elr4.w};
            float klv[4], qmax = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const Latent q = prologue<false, false>(p, hv[j], lr[j], bias[j]);
              s_rate[j * kFsGPool + e] = q.rate;
              qmax = fmaxf(qmax, q.rate);
              const float d = q.log_dr;
              klv[j] = elr[j] * (1.f + expf(d) * (d - 1.f));  // main/vae.py:448-449
            }
            rmax = fmaxf(rmax, qmax);
            klsum = (klv[0] + klv[1]) + (klv[2] + klv[3]);
            // a chain at this rate walks about rate * t_need events: the long ones are walked by lane groups, first
            pre = qmax * fp.t_need > fp.park_events;
          }
          klsum += __shfl_xor_sync(0xffffffffu, klsum, 1);
          klsum += __shfl_xor_sync(0xffffffffu, klsum, 2);
          klsum += __shfl_xor_sync(0xffffffffu, klsum, 4);
          if (cq == 0 && r < nrows) kl_tab[(b * kFsKb + kb) * kFsGroupRows + r] = klsum;
          const unsigned pm = __ballot_sync(0xffffffffu, pre);
          if (pm != 0u) {
            int pbase = 0;
            if (lane == 0) pbase = atomicAdd(&cnt[1], __popc(pm));
            pbase = __shfl_sync(0xffffffffu, pbase, 0);
            if (pre) s_list[pbase + __popc(pm & lanes_below)] = static_cast<unsigned short>(e);
          }
        }
        named_bar_sync(bar_id, kFsGroupLanes);
        if (gtid == 0) stamp(fp, 300 + (g * 8 + bi) * 4);
        if (gtid < 8) s_cnt[((bi + 1) & 1) * 8 + gtid] = gtid == 0 ? kFsGroupLanes : 0;  // the next batch's counters (idle since batch bi - 1)

        // pass 0: the listed long chains (lane groups).  pass 1: the pool over everything else, then the chains that turned
        // out long on the way (lane groups again).
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
          if (pass == 1) {
            if (gtid == 0) stamp(fp, 300 + (g * 8 + bi) * 4 + 1);
            int cur = gtid;
            bool has = cur < pool_n, live = false, skip = false;
            int n = 0, ebase = 0;
            uint32_t quad = 0, zaddr = 0, zbar = 0;
            float rate[4], inv_rate[4], t[4], acc[4], acc2[4];
            auto fetch = [&]() {
              live = false, skip = false;
              if (has) {
                const int kb = static_cast<int>(((static_cast<uint32_t>(cur) >> 3) * rmagic) >> 16), rem = cur - kb * ekb, r = rem >> 3, cq = rem & 7;
                const uint32_t gk = gk0 + kb, slot = gk % kFsZSlots;
                zbar = zbar0 + slot * 8;
                if (gk >= kFsZSlots) mbar_wait_u32(zebar0 + slot * 8, ((gk / kFsZSlots) - 1) & 1);  // the slot's previous tenant has been consumed
                if (r < nrows) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) rate[j] = s_rate[j * kFsGPool + cur];
                  if (fmaxf(fmaxf(rate[0], rate[1]), fmaxf(rate[2], rate[3])) * fp.t_need > fp.park_events) {
                    skip = true;  // listed by the prologue, walked in pass 0
                  } else {
                    live = true;
                    const int qi = r * Kq + b * (kFsBatchK / 4) + kb * 8 + cq;
                    ebase = qi << 2;
                    quad = quad0 + static_cast<uint32_t>(qi);
                    const int ur = grow0 + r;  // row inside the unit = row of the operand tile
                    zaddr = zst + slot * (2 * kFsZTile) + ur * 128 + ((cq ^ (ur & 7)) << 4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      inv_rate[j] = recip_for_div(rate[j]);
                      t[j] = acc[j] = acc2[j] = 0.f;
                    }
                    n = 0;
                  }
                }
              }
            };
            fetch();
            while (__any_sync(0xffffffffu, has)) {
              bool fin = false, park = false;
              if (has && !live) {
                if (!skip) mbar_arrive_u32(zbar);  // a row beyond the batch: nothing to sample, but the k-block's count expects the entry
                fin = true;
              } else if (live) {
                const uint4 rr = philox4x32_10(quad, static_cast<uint32_t>(n), off_lo, off_hi, p.ks);
                const uint32_t bits[4] = {rr.x, rr.y, rr.z, rr.w};
                float excess = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float e = exponential_from_bits(bits[j]);
                  t[j] = __fadd_rn(t[j], div_by_rate<false>(e, rate[j], inv_rate[j]));
                  float f, df = 0.f;
                  soft_term<IND, true>(p, t[j], f, df);
                  acc[j] += f;
                  excess = fmaxf(excess, fmaf(-absorb, acc[j], f));
                  const float gg = df * t[j];
                  acc2[j] += gg;
                  excess = fmaxf(excess, fmaf(-absorb, acc2[j], gg));
                }
                ++n;
                const float tmin = fminf(fminf(t[0], t[1]), fminf(t[2], t[3]));
                if (tmin > t_exit || (absorb > 0.f && tmin > 1.f && excess <= 0.f) || n >= p.n_exp) {  // chain complete
                  const float4 zv = make_float4(acc[0], acc[1], acc[2], acc[3]);
                  stg4_planes(z0, zlo0, ebase, zv);
                  stg4(dz0 + ebase, make_float4(acc2[0], acc2[1], acc2[2], acc2[3]));
                  sts4(zaddr, zv);
                  sts4(zaddr + kFsZTile, tf32_lo4f(zv));
                  mbar_arrive_u32(zbar);  // release; the issuer's acquire + proxy fence order these stores before the MMA's reads
                  fin = true;
                } else if (n >= n_a) {
                  fin = park = true;
                }
              }
              const unsigned fm = __ballot_sync(0xffffffffu, fin);
              if (fm != 0u) {
                const unsigned pm = __ballot_sync(0xffffffffu, park);
                int first = 0, pbase = 0;
                if (lane == 0) {
                  first = atomicAdd(&cnt[0], __popc(fm));
                  if (pm != 0u) pbase = atomicAdd(&cnt[3], __popc(pm));
                }
                first = __shfl_sync(0xffffffffu, first, 0);
                pbase = __shfl_sync(0xffffffffu, pbase, 0);
                if (park) s_list[kFsGPool - 1 - (pbase + __popc(pm & lanes_below))] = static_cast<unsigned short>(cur);  // from the back
                if (fin) {
                  cur = first + __popc(fm & lanes_below);
                  has = cur < pool_n;
                  fetch();
                }
              }
            }
            named_bar_sync(bar_id, kFsGroupLanes);  // every late long chain of the group is on the list
            if (gtid == 0) stamp(fp, 300 + (g * 8 + bi) * 4 + 2);
          }

          // ---- lane groups of kGroup walk the listed chains from their first event, kGroup events per step
          const int n_parked = cnt[1 + 2 * pass];
          int* const cursor = &cnt[2 + 2 * pass];
          if (n_parked > 0) {
            const int gl = lane % kGroup;
            bool has = false, done = true;
            int ebase = 0, n0 = 0;
            uint32_t quad = 0, zaddr = 0, zbar = 0;
            float rate[4] = {1.f, 1.f, 1.f, 1.f}, inv_rate[4] = {1.f, 1.f, 1.f, 1.f}, t0[4] = {0.f, 0.f, 0.f, 0.f};
            float part[4] = {0.f, 0.f, 0.f, 0.f}, part2[4] = {0.f, 0.f, 0.f, 0.f};
            bool exhausted = false;  // warp-uniform: the list has no more entries
            while (true) {
              if (__any_sync(0xffffffffu, done)) {
                float tot[4], tot2[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float v = part[j], v2 = part2[j];
#pragma unroll
                  for (int o = kGroup / 2; o > 0; o >>= 1) {
                    v += __shfl_xor_sync(0xffffffffu, v, o, kGroup);
                    v2 += __shfl_xor_sync(0xffffffffu, v2, o, kGroup);
                  }
                  tot[j] = v, tot2[j] = v2;
                }
                if (done && gl == 0 && has) {
                  const float4 zv = make_float4(tot[0], tot[1], tot[2], tot[3]);
                  stg4_planes(z0, zlo0, ebase, zv);
                  stg4(dz0 + ebase, make_float4(tot2[0], tot2[1], tot2[2], tot2[3]));
                  sts4(zaddr, zv);
                  sts4(zaddr + kFsZTile, tf32_lo4f(zv));
                  mbar_arrive_u32(zbar);
                }
                int next = 0;
                {
                  const unsigned req = __ballot_sync(0xffffffffu, done && gl == 0);
                  int first = n_parked;
                  if (lane == 0 && !exhausted) first = atomicAdd(cursor, __popc(req));
                  first = __shfl_sync(0xffffffffu, first, 0);
                  next = first + __popc(req & lanes_below);
                  exhausted = first + __popc(req) >= n_parked;
                }
                next = __shfl_sync(0xffffffffu, next, 0, kGroup);
                if (done) {
                  has = next < n_parked;
                  if (has) {
                    const int cur = s_list[pass == 0 ? next : kFsGPool - 1 - next];
                    const int kb = static_cast<int>(((static_cast<uint32_t>(cur) >> 3) * rmagic) >> 16), rem = cur - kb * ekb, r = rem >> 3, cq = rem & 7;
                    const int qi = r * Kq + b * (kFsBatchK / 4) + kb * 8 + cq;
                    ebase = qi << 2;
                    quad = quad0 + static_cast<uint32_t>(qi);
                    const int ur = grow0 + r;
                    const uint32_t gk = gk0 + kb, slot = gk % kFsZSlots;
                    if (gk >= kFsZSlots) mbar_wait_u32(zebar0 + slot * 8, ((gk / kFsZSlots) - 1) & 1);
                    zaddr = zst + slot * (2 * kFsZTile) + ur * 128 + ((cq ^ (ur & 7)) << 4);
                    zbar = zbar0 + slot * 8;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      rate[j] = s_rate[j * kFsGPool + cur];
                      inv_rate[j] = recip_for_div(rate[j]);
                      t0[j] = 0.f, part[j] = 0.f, part2[j] = 0.f;
                    }
                    n0 = 0;
                  }
                  done = false;
                }
                if (__all_sync(0xffffffffu, !has)) break;
              }
              const int n = n0 + gl;
              const bool valid = has && !done && n < p.n_exp;
              const uint4 rr = philox4x32_10(quad, static_cast<uint32_t>(n), off_lo, off_hi, p.ks);
              const uint32_t bits[4] = {rr.x, rr.y, rr.z, rr.w};
              float excess = 0.f;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float x = valid ? div_by_rate<false>(exponential_from_bits(bits[j]), rate[j], inv_rate[j]) : 0.f;
#pragma unroll
                for (int o = 1; o < kGroup; o <<= 1) {  // inclusive prefix sum over the group's lanes
                  const float y = __shfl_up_sync(0xffffffffu, x, o, kGroup);
                  if (gl >= o) x += y;
                }
                const float t = t0[j] + x;
                float f, df = 0.f;
                soft_term<IND, true>(p, t, f, df);
                const float gg = df * t;
                part[j] += valid ? f : 0.f;
                // the group's last lane holds the smallest term of the step, its first lane the largest partial sum (a lower
                // bound of the chain's running total)
                excess = fmaxf(excess, fmaf(-absorb, __shfl_sync(0xffffffffu, part[j], 0, kGroup), f));
                part2[j] += valid ? gg : 0.f;
                excess = fmaxf(excess, fmaf(-absorb, __shfl_sync(0xffffffffu, part2[j], 0, kGroup), gg));
                t0[j] = __shfl_sync(0xffffffffu, t, kGroup - 1, kGroup);
              }
              n0 += kGroup;
              const float tmin = fminf(fminf(t0[0], t0[1]), fminf(t0[2], t0[3]));
              const bool absorbed = absorb > 0.f && __shfl_sync(0xffffffffu, excess <= 0.f ? 1 : 0, kGroup - 1, kGroup) != 0 && tmin > 1.f;
              if (has && (n0 >= p.n_exp || tmin > t_exit || absorbed)) done = true;
            }
          }
        }  // passes
        if (b == nbatch - 1 && p.kl_rowsum != nullptr && gtid < nrows) {  // KL row sums of the group's rows (k-blocks in order)
          float s = 0.f;
          for (int kbg = 0; kbg < nkb; ++kbg) s += kl_tab[kbg * kFsGroupRows + gtid];
          p.kl_rowsum[grow + gtid] = s;
        }
        named_bar_sync(bar_id, kFsGroupLanes);  // the group's tables are reused by its next batch
        if (gtid == 0) {
          stamp(fp, 300 + (g * 8 + bi) * 4 + 3), stamp_all(fp, 703);
          if (fp.dbg_ts != nullptr && blockIdx.x < 148) {  // per CTA: latest end of a batch (any group)
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            atomicMax(fp.dbg_ts + 800 + blockIdx.x, t);
          }
        }
      }  // batches

      // ---- unit epilogue: the reconstruction residual from TMEM.  Accumulator lane = row of the unit, so only the warps with
      // warp % 4 == 0 (TMEM lanes 0..31) can read it: warps 0, 4, 8, 12 take D / 4 columns each
      if ((warp & 3) == 0 && (warp >> 2) < n_ep) {
        const int i = warp >> 2, ncol = D / n_ep;  // output columns [ncol i, ncol (i + 1))
        const int urows = static_cast<int>(min(static_cast<long long>(R), p.B - row0));
        mbar_wait(acc_full, ui & 1);
        if (tid == 0) stamp(fp, 600), stamp_all(fp, 705);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tcol = tmem_base + i * ncol;
        const bool rv = lane < urows;
        const long long gbase = (row0 + lane) * D + ncol * i;
        float sq = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < ncol; c0 += 16) {
          uint32_t v[16], w[16];
          tmem_ld_32x16(tcol + c0, v);
          tmem_ld_32x16(tcol + D + c0, w);
          float4 xv[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) xv[c] = rv ? ldg4(fp.x + gbase + c0 + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
          tmem_ld_wait();
          if (rv) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 r4 = make_float4((__uint_as_float(v[4 * c]) + __uint_as_float(w[4 * c])) - xv[c].x,
                                            (__uint_as_float(v[4 * c + 1]) + __uint_as_float(w[4 * c + 1])) - xv[c].y,
                                            (__uint_as_float(v[4 * c + 2]) + __uint_as_float(w[4 * c + 2])) - xv[c].z,
                                            (__uint_as_float(v[4 * c + 3]) + __uint_as_float(w[4 * c + 3])) - xv[c].w);
              const float4 gv = make_float4(fp.g_scale * r4.x, fp.g_scale * r4.y, fp.g_scale * r4.z, fp.g_scale * r4.w);
              stg4(fp.g_y + gbase + c0 + 4 * c, gv);
              if (fp.g_y_lo != nullptr) stg4(fp.g_y_lo + gbase + c0 + 4 * c, tf32_lo4f(gv));
              sq += (r4.x * r4.x + r4.y * r4.y) + (r4.z * r4.z + r4.w * r4.w);
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_u32(smem_u32(acc_empty));
        float* const rec = s_recon + (ui & 1) * 256;
        rec[i * 32 + lane] = sq;
        named_bar_sync(14, n_ep * 32);
        if (i == 0 && rv) {
          float s = rec[lane];
          for (int j = 1; j < n_ep; ++j) s += rec[j * 32 + lane];
          fp.recon[row0 + lane] = s;
        }
        if (tid == 0) stamp(fp, 601), stamp_all(fp, 707);
      }
    }  // units

    if (p.rate_max != nullptr) {
      const float m = warp_max(rmax);
      if (lane == 0) s_max[warp] = m;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0 && p.rate_max != nullptr) {
    float mm = s_max[0];
#pragma unroll
    for (int w = 1; w < kFsWarps; ++w) mm = fmaxf(mm, s_max[w]);
    atomicMax(reinterpret_cast<int*>(p.rate_max), __float_as_int(mm));  // rates are > 0
  }
  if (warp == kFsWarps) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kFsTmemCols));
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------
static int g_fused_debug = 0;
static unsigned long long* g_fused_ts = nullptr;

static int fused_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
    sms = n;
  }
  return sms;
}

// rows per unit: the batch dealt out evenly over m units per SM, m the smallest count that keeps a unit within 28 rows
static int fused_rows_per_unit(int64_t B) {
  const int64_t sms = fused_sms();
  const int64_t m = (B + sms * kFsMaxUnitRows - 1) / (sms * kFsMaxUnitRows);
  return static_cast<int>((B + sms * m - 1) / (sms * m));
}

int fused_fwd_supported(int64_t B, int64_t K, int64_t D) {
  if (B <= 0 || K <= 0 || K % kFsBatchK || K > kFsMaxK || (D != 128 && D != 256)) return 0;
  return fused_rows_per_unit(B) >= 12;  // below that the pool of a batch is too shallow for 640 lanes
}

template <int IND>
static cudaError_t launch_fused(const CUtensorMap& mw, const CUtensorMap& mwl, const FusedParams& fp, int grid, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(fused_fwd_kernel<IND>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kFsSmem));
    if (e != cudaSuccess) return e;
    configured = true;
  }
  return launch_pdl(fused_fwd_kernel<IND>, dim3(static_cast<unsigned>(grid)), dim3(kFsThreads), kFsSmem, s, mw, mwl, fp);
}

int fused_fwd_impl(const float* h, const float* log_r, const float* b_enc, const float* w_dec, const float* w_dec_lo, const float* x, float* z,
                   float* z_lo, float* dz_acc, float* g_y, float* g_y_lo, float* recon, float* kl_rowsum, float* rate_max, int64_t B, int64_t K,
                   int64_t D, int64_t row_offset, int n_exp, float temp, int indicator, int exc_only, float exit_logit, float g_scale,
                   uint64_t seed, uint64_t offset, const uint64_t* offset_dev, void* stream) {
  if (!fused_fwd_supported(B, K, D))
    return fail(PVAE_ERR_UNSUPPORTED, "pvae_fused_fwd: B=%lld K=%lld D=%lld outside the fused kernel's shapes (K %% 256 == 0, K <= 1024, D in {128, 256}, B >= 12 rows per SM)",
                (long long)B, (long long)K, (long long)D);
  if (!h || !log_r || !w_dec || !w_dec_lo || !x || !z || !dz_acc || !g_y || !recon)
    return fail(PVAE_ERR_INVALID, "pvae_fused_fwd: null pointer");
  if (!(temp > 0.f)) return fail(PVAE_ERR_INVALID, "pvae_fused_fwd: the training forward needs temp > 0, got %g", temp);
  if (n_exp < 1 || indicator < 0 || indicator > 4 || row_offset < 0) return fail(PVAE_ERR_INVALID, "pvae_fused_fwd: bad n_exp / indicator / row_offset");
  if (((row_offset + B) * K) >> 2 > 0xFFFFFFFFll) return fail(PVAE_ERR_UNSUPPORTED, "pvae_fused_fwd: more than 2^34 latents");
  FusedParams fp{};
  SamplerParams& p = fp.sp;
  p.h = h, p.log_r = log_r, p.b_enc = b_enc, p.z = z, p.z_lo = z_lo, p.dz_out = dz_acc, p.kl_rowsum = kl_rowsum, p.rate_max = rate_max;
  p.B = B, p.K = K, p.row_offset = row_offset;
  sampler_set_chain(p, n_exp, temp, indicator, exit_logit);
  p.clamp_dr = 10.f, p.clamp_rate = 5.f, p.exc_only = exc_only;  // main/vae.py:403-409
  p.offset = offset, p.offset_dev = offset_dev;
  philox_key_schedule(seed, p.ks);
  fp.x = x, fp.g_y = g_y, fp.g_y_lo = g_y_lo, fp.recon = recon, fp.D = static_cast<int>(D), fp.g_scale = g_scale;
  fp.rows_unit = fused_rows_per_unit(B);
  fp.n_units = static_cast<int>((B + fp.rows_unit - 1) / fp.rows_unit);
  // a chain ends once its arrival time passes t_need: the absorbed-exit rule needs terms below 2^-34 of the sum, i.e.
  // t - 1 > 23.6 temp (sigmoid); without it the chain runs to t_exit
  fp.t_need = p.absorb > 0.f ? std::min(1.f + 24.f * temp, p.t_exit) : p.t_exit;
  fp.debug = g_fused_debug;
  fp.dbg_ts = g_fused_ts;
  fp.park_events = static_cast<float>((p.phase_a < n_exp ? p.phase_a : n_exp) - 1);
  CUtensorMap mw, mwl;
  {
    const long long dims[2] = {K, D}, strides[1] = {K * 4};
    const int box[2] = {32, static_cast<int>(D)}, estr[2] = {1, 1};  // one plane of a k-block: {32 k, D rows}
    if (int rc = make_tensor_map(&mw, w_dec, 2, dims, strides, box, estr, 0)) return rc;
    if (int rc = make_tensor_map(&mwl, w_dec_lo, 2, dims, strides, box, estr, 0)) return rc;
  }
  const int grid = std::min(fp.n_units, fused_sms());
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  switch (indicator) {
    case kSigmoid: e = launch_fused<kSigmoid>(mw, mwl, fp, grid, s); break;
    case kCosine: e = launch_fused<kCosine>(mw, mwl, fp, grid, s); break;
    case kLinear: e = launch_fused<kLinear>(mw, mwl, fp, grid, s); break;
    case kCubic: e = launch_fused<kCubic>(mw, mwl, fp, grid, s); break;
    default: e = launch_fused<kQuintic>(mw, mwl, fp, grid, s); break;
  }
  return cuda_ok("pvae_fused_fwd", e);
}

}  // namespace pvae

extern "C" int pvae_set_fused_debug(int mode) {
  pvae::g_fused_debug = mode;
  return PVAE_OK;
}

extern "C" int pvae_set_fused_timeline(uint64_t* ts_dev) {
  pvae::g_fused_ts = reinterpret_cast<unsigned long long*>(ts_dev);
  return PVAE_OK;
}

extern "C" int pvae_fused_fwd_supported(int64_t B, int64_t K, int64_t D) { return pvae::fused_fwd_supported(B, K, D); }

extern "C" int pvae_fused_fwd(const float* h, const float* log_r, const float* b_enc, const float* w_dec, const float* w_dec_lo, const float* x,
                              float* z, float* z_lo, float* dz_acc, float* g_y, float* g_y_lo, float* recon, float* kl_rowsum,
                              float* rate_max, int64_t B, int64_t K, int64_t D, int64_t row_offset, int n_exp, float temp, int indicator,
                              int exc_only, float exit_logit, float g_scale, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                              void* stream) {
  return pvae::fused_fwd_impl(h, log_r, b_enc, w_dec, w_dec_lo, x, z, z_lo, dz_acc, g_y, g_y_lo, recon, kl_rowsum, rate_max, B, K, D,
                              row_offset, n_exp, temp, indicator, exc_only, exit_logit, g_scale, seed, offset, offset_dev, stream);
}
