// Fused Poisson reparameterised sampler + KL kernels for sm_100a.
//
// Work decomposition.  A block owns a tile of rows (samples) and all K latents of those rows.
//   Phase A  one thread per quad of four consecutive latents (128-bit loads of the encoder output,
//            128-bit stores of the counts): softclamps, rate and KL in registers, then the four
//            arrival chains in lock step - one Philox4x32-10 block per event feeds the four chains.
//            A chain leaves as soon as every later term is exactly zero (lossless early exit).
//            Quads still alive after `phase_a` events are parked in shared memory.
//   Phase B  parked quads are finished by groups of 8 lanes (four quads per warp at a time), 8 events
//            per step: lane i of a group draws event n0+i, the arrival times come from a prefix sum
//            over the group (shfl.up), so one long chain costs n/8 steps of 8 busy lanes instead of
//            n steps of one busy lane.
// The (n_exp, B, K) tensors of the reference never exist; per-sample KL sums and per-latent
// (batch) gradient sums are produced without atomics, in a fixed order.
//
// Hard counts (temp == 0) run Phase A only: the running sum is then the same sequential fp32 sum
// as torch's CUDA cumsum, which makes the hard counts bit-exact.
//
// Reference semantics: base/distributions.py:5-92 (Poisson), main/vae.py:393-415 (infer),
// main/vae.py:446-450 (loss_kl); closed-form backward SURVEY.md section 8 row a13.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pvae_b200.h"
#include "pvae_device.cuh"
#include "pvae_host.h"
#include "sampler_device.cuh"

namespace pvae {

constexpr int kThreads = 128;               // threads per block
constexpr int kWarps = kThreads / 32;
constexpr int kColsPerPass = 4 * kThreads;  // latents of one row a block covers per pass
constexpr int kMaxRows = 3;                 // rows per block tile
constexpr int kSlots = kMaxRows * kThreads; // parked-quad table: one slot per (row, thread)
constexpr int kSlotWords = 20;              // rate[4], t[4], acc[4], acc2[4], cf[4] (one shared-memory plane per word)

// ---- phase A: four chains in lock step, events [0, n_end) ---------------------------------------
// acc: sum_n f(logit_n) (fwd), hard count (hard) or sum_n f'(logit_n) * t_n (bwd).
// Returns true when the chains are complete (exit reached or all n_exp events walked).
template <int MODE, int IND>
__device__ __forceinline__ bool phase_a(const SamplerParams& p, uint32_t quad, uint32_t off_lo, uint32_t off_hi,
                                        const float (&rate)[4], const float (&inv_rate)[4], float (&t)[4],
                                        float (&acc)[4], float (&acc2)[4], int n_end) {
#pragma unroll
  for (int j = 0; j < 4; ++j) t[j] = acc[j] = acc2[j] = 0.f;
  const float t_exit = p.t_exit, absorb = p.absorb;
  for (int n = 0; n < n_end; ++n) {
    const uint4 r = philox4x32_10(quad, static_cast<uint32_t>(n), off_lo, off_hi, p.ks);
    const uint32_t bits[4] = {r.x, r.y, r.z, r.w};
    float excess = 0.f;  // > 0 while some term still matters to its running sum
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float e = exponential_from_bits(bits[j]);
      t[j] = __fadd_rn(t[j], div_by_rate<MODE == kFwdHard>(e, rate[j], inv_rate[j]));
      if (MODE == kFwdHard) {
        acc[j] += (t[j] <= 1.f) ? 1.f : 0.f;
      } else {
        float f, df = 0.f;
        soft_term<IND, MODE == kBwd || MODE == kFwdBoth>(p, t[j], f, df);
        const float g = df * t[j];
        acc[j] += MODE == kBwd ? g : f;
        excess = fmaxf(excess, fmaf(-absorb, acc[j], MODE == kBwd ? g : f));
        if (MODE == kFwdBoth) {
          acc2[j] += g;
          excess = fmaxf(excess, fmaf(-absorb, acc2[j], g));
        }
      }
    }
    const float tmin = fminf(fminf(t[0], t[1]), fminf(t[2], t[3]));
    if (tmin > t_exit) return true;
    // past the threshold (t > 1) the terms only shrink: once each is below 2^-34 of its running sum, all the
    // remaining ones together (<= n_exp < 2^9 of them) stay below half an ulp of the sum
    if (MODE != kFwdHard && absorb > 0.f && tmin > 1.f && excess <= 0.f) return true;  // absorb == 0: rule off (NEVER walks on)
  }
  return n_end >= p.n_exp;
}

template <int MODE, int IND, bool RAW, bool POOL = false>
__global__ void __launch_bounds__(kThreads) sampler_kernel(const SamplerParams p) {
  __shared__ float s_tab[kSlotWords][kSlots];  // parked quads, one word-plane per field
  __shared__ unsigned short s_list[kWarps][POOL ? kSlots : kMaxRows * 32];  // each warp finishes the quads it parked
  __shared__ int s_next;                                    // POOL: cursor into the block's pool of quads
  __shared__ float s_rowsum[kMaxRows][kWarps];
  __shared__ float s_max[kWarps];

  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  // A block owns `rows_block` consecutive rows and walks them in tiles of up to `rows_per_block` rows (the shared-memory
  // tables hold one tile).  rows_block == rows_per_block except with the optional one-wave grid (launch_pool).
  const long long blk_row0 = static_cast<long long>(blockIdx.x) * p.rows_block;
  const int blk_rows = static_cast<int>(min(static_cast<long long>(p.rows_block), p.B - blk_row0));
  uint64_t off = p.offset;
  if (p.offset_dev != nullptr) off += *p.offset_dev;
  const uint32_t off_lo = static_cast<uint32_t>(off), off_hi = static_cast<uint32_t>(off >> 32);
  const int n_a = min(p.phase_a, p.n_exp);

  constexpr bool kIsBwd = MODE == kBwd || MODE == kBwdSaved;
  constexpr bool kCoef = MODE == kFwdBoth && !RAW;  // coefficient mode exists for the training forward only
  const bool coef = kCoef && p.c_ddr != nullptr;
  float rmax = 0.f;
  const bool want_rowsum = !kIsBwd && !RAW && p.kl_rowsum != nullptr;
  const unsigned lanes_below = (1u << lane) - 1u;
#pragma unroll 1
  for (int tile_row = 0; tile_row < blk_rows; tile_row += p.rows_per_block) {
  const long long row0 = blk_row0 + tile_row;
  const int nrows = min(p.rows_per_block, blk_rows - tile_row);
  if (tid < kMaxRows * kWarps) (&s_rowsum[0][0])[tid] = 0.f;
  __syncthreads();

  // every thread runs every pass so that the block-level steps below see whole warps
  for (long long k0 = 0; k0 < p.K; k0 += kColsPerPass) {
    const long long k = k0 + 4 * tid;
    const bool act = k < p.K;
    int n_parked = 0;  // warp-uniform: quads of this warp waiting for phase B
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
    float col_kl[4] = {0.f, 0.f, 0.f, 0.f};  // coefficient mode: this thread's KL column sums over the tile's rows
    unsigned parked = 0;  // bit r: this thread's quad of row r waits for phase B

    // backward epilogue of one quad: gradients from acc = sum_n f'(l_n) t_n
    auto bwd_finish = [&](long long base, const Latent (&q)[4], const float (&acc)[4]) {
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
        // d z / d lambda = (sum_n f'(l_n) t_n) / temp * exp(lambda) / rate; then the clamp chain
        const float g_u = gz[j] * (acc[j] * p.inv_temp * (q[j].explam / q[j].rate)) * q[j].d_rate;
        if (RAW) {
          gh[j] = g_u;
        } else {
          const float d = q[j].log_dr, ed = expf(d);
          gh[j] = (g_u + p.kl_scale * elr[j] * d * ed + gld[j]) * q[j].d_dr;
          col_lr[j] += g_u + p.kl_scale * (elr[j] * (1.f + ed * (d - 1.f)));
          col_b[j] += gh[j];
        }
      }
      stg4_planes(p.g_h, p.g_h_lo, base, make_float4(gh[0], gh[1], gh[2], gh[3]));
    };

    // ---------------- phase A ----------------
    if (POOL) {
      // Work pool per warp.  Chain lengths are heavy-tailed (most quads need 1-2 events, some 8+): walking one row at
      // a time in lock step leaves half of the lanes idle waiting for the slowest quad of the row.  Instead:
      // stage 1 runs the prologue for all rows (uniform, coalesced) and leaves the rates in the warp's shared-memory
      // slots; stage 2 lets every lane walk one quad and, the moment it ends, take the next one of the warp's
      // nrows x 32 quads (ballot + prefix count, no atomics).  Per-chain arithmetic, Philox counters and exit rules
      // are those of phase_a(): results are bit-identical to the row-by-row schedule.
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
            q[j] = kCoef ? prologue<RAW, true>(p, hv[j], lr[j], bias[j]) : prologue<RAW, false>(p, hv[j], lr[j], bias[j]);
            rate[j] = q[j].rate;
          }
          if (p.rate != nullptr) stg4(p.rate + base, make_float4(rate[0], rate[1], rate[2], rate[3]));
          if (p.rate_max != nullptr) rmax = fmaxf(fmaxf(fmaxf(rate[0], rate[1]), fmaxf(rate[2], rate[3])), rmax);
          const int slot = r * kThreads + tid;
          if (!RAW) {
            if (p.log_dr != nullptr)
              stg4(p.log_dr + base, make_float4(q[0].log_dr, q[1].log_dr, q[2].log_dr, q[3].log_dr));
            if (p.kl != nullptr || p.kl_rowsum != nullptr || coef) {
              float klv[4], ck[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float d = q[j].log_dr, ed = expf(d);
                klv[j] = elr[j] * (1.f + ed * (d - 1.f));  // main/vae.py:448-449
                ck[j] = elr[j] * d * ed;                   // d kl / d delta
              }
              if (p.kl != nullptr) stg4(p.kl + base, make_float4(klv[0], klv[1], klv[2], klv[3]));
              klsum = (klv[0] + klv[1]) + (klv[2] + klv[3]);
              if (kCoef && coef) {
                stg4(p.c_ck + base, make_float4(ck[0], ck[1], ck[2], ck[3]));
                stg4(p.c_ddr + base, make_float4(q[0].d_dr, q[1].d_dr, q[2].d_dr, q[3].d_dr));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  col_kl[j] += klv[j];
                  s_tab[16 + j][slot] = coef_factor(p, q[j]);
                }
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) s_tab[j][slot] = rate[j];
        }
        if (want_rowsum) {
          const float sum = warp_sum(klsum);
          if (lane == 0) s_rowsum[r][warp] += sum;
        }
      }
      // stage 2: the block's nrows x kThreads quads form one pool (entry index == slot index); a shared cursor hands
      // out the next entries, one atomicAdd per warp and refill round
      if (tid == 0) s_next = kThreads;
      __syncthreads();
      const int pool_n = nrows * kThreads;
      const float t_exit = p.t_exit, absorb = p.absorb;
      const int Kq = static_cast<int>(p.K >> 2), k0q = static_cast<int>(k0 >> 2);
      const uint32_t quad0 = static_cast<uint32_t>(((p.row_offset + row0) * p.K) >> 2);
      float* const z0 = p.z + row0 * p.K;
      float* const zlo0 = p.z_lo != nullptr ? p.z_lo + row0 * p.K : nullptr;
      float* const dz0 = MODE == kFwdBoth ? p.dz_out + row0 * p.K : nullptr;
      int cur = tid;  // this lane's pool entry: row cur / kThreads, column quad cur % kThreads
      bool has = cur < pool_n, live = false;
      int n = 0, ebase = 0;
      uint32_t quad = 0;
      float rate[4], inv_rate[4], t[4], acc[4], acc2[4];
      auto fetch = [&]() {
        live = false;
        if (has) {
          const int r = cur / kThreads, cq = k0q + cur % kThreads;  // row, column quad
          if (cq < Kq) {
            live = true;
            ebase = (r * Kq + cq) << 2;
            quad = quad0 + static_cast<uint32_t>(r * Kq + cq);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              rate[j] = s_tab[j][cur];
              inv_rate[j] = recip_for_div(rate[j]);
              t[j] = acc[j] = acc2[j] = 0.f;
            }
            n = 0;
          }
        }
      };
      fetch();
      while (__any_sync(0xffffffffu, has)) {
        bool fin = has && !live;  // an entry beyond the last column: nothing to do
        bool park = false;
        if (live) {
          const uint4 rr = philox4x32_10(quad, static_cast<uint32_t>(n), off_lo, off_hi, p.ks);
          const uint32_t bits[4] = {rr.x, rr.y, rr.z, rr.w};
          float excess = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float e = exponential_from_bits(bits[j]);
            t[j] = __fadd_rn(t[j], div_by_rate<false>(e, rate[j], inv_rate[j]));
            float f, df = 0.f;
            soft_term<IND, MODE == kFwdBoth>(p, t[j], f, df);
            acc[j] += f;
            excess = fmaxf(excess, fmaf(-absorb, acc[j], f));
            if (MODE == kFwdBoth) {
              const float g = df * t[j];
              acc2[j] += g;
              excess = fmaxf(excess, fmaf(-absorb, acc2[j], g));
            }
          }
          ++n;
          const float tmin = fminf(fminf(t[0], t[1]), fminf(t[2], t[3]));
          if (tmin > t_exit || (absorb > 0.f && tmin > 1.f && excess <= 0.f) || n >= p.n_exp) {  // chain complete
            stg4_planes(z0, zlo0, ebase, make_float4(acc[0], acc[1], acc[2], acc[3]));
            if (MODE == kFwdBoth) {
              if (kCoef && coef) {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc2[j] *= s_tab[16 + j][cur];
              }
              stg4(dz0 + ebase, make_float4(acc2[0], acc2[1], acc2[2], acc2[3]));
            }
            fin = true;
          } else if (n >= n_a) {  // a long chain: hand it to phase B
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              s_tab[4 + j][cur] = t[j];
              s_tab[8 + j][cur] = acc[j];
              if (MODE == kFwdBoth) s_tab[12 + j][cur] = acc2[j];
            }
            fin = park = true;
          }
        }
        const unsigned fm = __ballot_sync(0xffffffffu, fin);
        if (fm != 0u) {
          const unsigned pm = __ballot_sync(0xffffffffu, park);
          if (park) s_list[warp][n_parked + __popc(pm & lanes_below)] = static_cast<unsigned short>(cur);
          n_parked += __popc(pm);
          int first = 0;
          if (lane == 0) first = atomicAdd(&s_next, __popc(fm));
          first = __shfl_sync(0xffffffffu, first, 0);
          if (fin) {
            cur = first + __popc(fm & lanes_below);
            has = cur < pool_n;
            fetch();
          }
        }
      }
    } else {
#pragma unroll 1
    for (int r = 0; r < nrows; ++r) {
      float klsum = 0.f;
      bool park = false;
      if (act) {
        const long long row = row0 + r;
        const long long base = row * p.K + k;
        const float4 hv4 = ldg4(p.h + base);
        const float hv[4] = {hv4.x, hv4.y, hv4.z, hv4.w};
        Latent q[4];
        float rate[4], inv_rate[4];
        float cf[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          q[j] = (kIsBwd || kCoef) ? prologue<RAW, true>(p, hv[j], lr[j], bias[j]) : prologue<RAW, false>(p, hv[j], lr[j], bias[j]);
          rate[j] = q[j].rate;
          inv_rate[j] = MODE == kBwdSaved ? 0.f : (MODE == kFwdHard ? 0.f : recip_for_div(rate[j]));
          if (kCoef && coef) cf[j] = coef_factor(p, q[j]);
        }
        if (!kIsBwd) {  // outputs that do not depend on the chain
          if (p.rate != nullptr) stg4(p.rate + base, make_float4(rate[0], rate[1], rate[2], rate[3]));
          if (p.rate_max != nullptr) rmax = fmaxf(fmaxf(fmaxf(rate[0], rate[1]), fmaxf(rate[2], rate[3])), rmax);
          if (!RAW) {
            if (p.log_dr != nullptr)
              stg4(p.log_dr + base, make_float4(q[0].log_dr, q[1].log_dr, q[2].log_dr, q[3].log_dr));
            if (p.kl != nullptr || p.kl_rowsum != nullptr || coef) {
              float klv[4], ck[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float d = q[j].log_dr, ed = expf(d);
                klv[j] = elr[j] * (1.f + ed * (d - 1.f));  // main/vae.py:448-449
                ck[j] = elr[j] * d * ed;
              }
              if (p.kl != nullptr) stg4(p.kl + base, make_float4(klv[0], klv[1], klv[2], klv[3]));
              klsum = (klv[0] + klv[1]) + (klv[2] + klv[3]);
              if (kCoef && coef) {
                stg4(p.c_ck + base, make_float4(ck[0], ck[1], ck[2], ck[3]));
                stg4(p.c_ddr + base, make_float4(q[0].d_dr, q[1].d_dr, q[2].d_dr, q[3].d_dr));
#pragma unroll
                for (int j = 0; j < 4; ++j) col_kl[j] += klv[j];
              }
            }
          }
        }
        const uint32_t quad = static_cast<uint32_t>(((p.row_offset + row) * p.K + k) >> 2);
        float t[4], acc[4], acc2[4];
        bool done = true;
        if (MODE == kBwdSaved) {  // the forward already walked the chain
          const float4 d4 = ldg4(p.dz_in + base);
          acc[0] = d4.x, acc[1] = d4.y, acc[2] = d4.z, acc[3] = d4.w;
        } else {
          done = phase_a<MODE, IND>(p, quad, off_lo, off_hi, rate, inv_rate, t, acc, acc2, n_a);
        }
        if (done) {
          if (kIsBwd) {
            bwd_finish(base, q, acc);
          } else {
            stg4_planes(p.z, p.z_lo, base, make_float4(acc[0], acc[1], acc[2], acc[3]));
            if (MODE == kFwdBoth)
              stg4(p.dz_out + base, make_float4(acc2[0] * cf[0], acc2[1] * cf[1], acc2[2] * cf[2], acc2[3] * cf[3]));
          }
        } else {
          const int slot = r * kThreads + tid;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            s_tab[j][slot] = rate[j];
            s_tab[4 + j][slot] = t[j];
            s_tab[8 + j][slot] = acc[j];
            if (MODE == kFwdBoth) s_tab[12 + j][slot] = acc2[j];
            if (kCoef) s_tab[16 + j][slot] = cf[j];
          }
          park = true;
          parked |= 1u << r;
        }
      }
      __syncwarp();
      {  // append this row's parked quads to the warp's list (ballot + prefix count, no atomics)
        const unsigned m = __ballot_sync(0xffffffffu, park);
        if (park) s_list[warp][n_parked + __popc(m & lanes_below)] = static_cast<unsigned short>(r * kThreads + tid);
        n_parked += __popc(m);
      }
      if (want_rowsum) {
        const float s = warp_sum(klsum);
        if (lane == 0) s_rowsum[r][warp] += s;
      }
    }
    __syncwarp();

    }
    // ---------------- phase B (per warp) ----------------
    if (MODE != kFwdHard && MODE != kBwdSaved && n_parked > 0) {
      const int gl = lane % kGroup;
      int next_free = 0;  // warp-uniform cursor into the warp's list
      bool has = false, done = true;  // done: this group needs a (new) quad
      int slot = 0, n0 = 0;
      uint32_t quad = 0;
      float rate[4] = {1.f, 1.f, 1.f, 1.f}, inv_rate[4] = {1.f, 1.f, 1.f, 1.f}, t0[4] = {0.f, 0.f, 0.f, 0.f};
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, part[4] = {0.f, 0.f, 0.f, 0.f};
      float acc2[4] = {0.f, 0.f, 0.f, 0.f}, part2[4] = {0.f, 0.f, 0.f, 0.f};
      while (true) {
        if (__any_sync(0xffffffffu, done)) {
          // retire finished quads (group-wide sum of the per-lane partial sums), then refill from the queue
          float tot[4], tot2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v = part[j], v2 = part2[j];
#pragma unroll
            for (int o = kGroup / 2; o > 0; o >>= 1) {
              v += __shfl_xor_sync(0xffffffffu, v, o, kGroup);
              if (MODE == kFwdBoth) v2 += __shfl_xor_sync(0xffffffffu, v2, o, kGroup);
            }
            tot[j] = acc[j] + v;
            if (MODE == kFwdBoth) tot2[j] = acc2[j] + v2;
          }
          int next = 0;
          if (done && gl == 0) {
            if (has) {
              if (MODE == kBwd) {
#pragma unroll
                for (int j = 0; j < 4; ++j) s_tab[8 + j][slot] = tot[j];
              } else {
                const int r = slot / kThreads, owner = slot % kThreads;
                const long long base = (row0 + r) * p.K + k0 + 4 * owner;
                stg4_planes(p.z, p.z_lo, base, make_float4(tot[0], tot[1], tot[2], tot[3]));
                if (MODE == kFwdBoth) {
                  if (kCoef && coef) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) tot2[j] *= s_tab[16 + j][slot];
                  }
                  stg4(p.dz_out + base, make_float4(tot2[0], tot2[1], tot2[2], tot2[3]));
                }
              }
            }
          }
          {  // hand the next list entries to the requesting groups, in lane order
            const unsigned req = __ballot_sync(0xffffffffu, done && gl == 0);
            next = next_free + __popc(req & lanes_below);
            next_free += __popc(req);
          }
          next = __shfl_sync(0xffffffffu, next, 0, kGroup);
          if (done) {
            has = next < n_parked;
            if (has) {
              slot = s_list[warp][next];
              const int r = slot / kThreads, owner = slot % kThreads;
              quad = static_cast<uint32_t>(((p.row_offset + row0 + r) * p.K + k0 + 4 * owner) >> 2);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                rate[j] = s_tab[j][slot];
                inv_rate[j] = recip_for_div(rate[j]);
                t0[j] = s_tab[4 + j][slot];
                acc[j] = s_tab[8 + j][slot];
                part[j] = 0.f;
                if (MODE == kFwdBoth) acc2[j] = s_tab[12 + j][slot], part2[j] = 0.f;
              }
              n0 = n_a;
            }
            done = false;  // queue exhausted: the group idles (has == false) without asking again
          }
          if (__all_sync(0xffffffffu, !has)) break;
        }
        // one step: kGroup events of this group's quad
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
          soft_term<IND, MODE == kBwd || MODE == kFwdBoth>(p, t, f, df);
          const float g = df * t;
          const float term = MODE == kBwd ? g : f;
          part[j] += valid ? term : 0.f;
          // the group's last lane holds the smallest term of the step; its own sums bound the totals from below
          excess = fmaxf(excess, fmaf(-p.absorb, acc[j] + part[j], term));
          if (MODE == kFwdBoth) {
            part2[j] += valid ? g : 0.f;
            excess = fmaxf(excess, fmaf(-p.absorb, acc2[j] + part2[j], g));
          }
          t0[j] = __shfl_sync(0xffffffffu, t, kGroup - 1, kGroup);
        }
        n0 += kGroup;
        const float tmin = fminf(fminf(t0[0], t0[1]), fminf(t0[2], t0[3]));
        const bool absorbed = p.absorb > 0.f && __shfl_sync(0xffffffffu, excess <= 0.f ? 1 : 0, kGroup - 1, kGroup) != 0 && tmin > 1.f;
        if (has && (n0 >= p.n_exp || tmin > p.t_exit || absorbed)) done = true;
      }
    }
    if (kIsBwd) {
      __syncwarp();
      // the owner finishes its parked quads (prologue recomputed: cheap next to a parked chain)
      for (int r = 0; r < nrows; ++r) {
        if (!((parked >> r) & 1u)) continue;
        const long long base = (row0 + r) * p.K + k;
        const float4 hv4 = ldg4(p.h + base);
        const float hv[4] = {hv4.x, hv4.y, hv4.z, hv4.w};
        Latent q[4];
        float acc[4];
        const int slot = r * kThreads + tid;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          q[j] = prologue<RAW, true>(p, hv[j], lr[j], bias[j]);
          acc[j] = s_tab[8 + j][slot];
        }
        bwd_finish(base, q, acc);
      }
      if (!RAW && act) {
        const long long pbase = static_cast<long long>(blockIdx.x) * p.K + k;
        stg4(p.part_log_r + pbase, make_float4(col_lr[0], col_lr[1], col_lr[2], col_lr[3]));
        if (p.part_b != nullptr) stg4(p.part_b + pbase, make_float4(col_b[0], col_b[1], col_b[2], col_b[3]));
      }
    }
    if (kCoef && coef && act) {  // this block tile's KL column sums (rows in order: deterministic)
      const long long tile = static_cast<long long>(blockIdx.x) * ((p.rows_block + p.rows_per_block - 1) / p.rows_per_block) + tile_row / p.rows_per_block;
      stg4(p.klcol + tile * p.K + k, make_float4(col_kl[0], col_kl[1], col_kl[2], col_kl[3]));
    }
    if (POOL) __syncthreads(); else __syncwarp();  // the slots are reused by the next pass
  }

  if (!kIsBwd && want_rowsum) {
    __syncthreads();
    if (tid < nrows) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += s_rowsum[tid][w];
      p.kl_rowsum[row0 + tid] = s;
    }
  }
  }  // tiles of this block's rows

  if (!kIsBwd) {
    if (p.rate_max != nullptr) {
      const float m = warp_max(rmax);
      if (lane == 0) s_max[warp] = m;
      __syncthreads();
      if (tid == 0) {
        float mm = s_max[0];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) mm = fmaxf(mm, s_max[w]);
        atomicMax(reinterpret_cast<int*>(p.rate_max), __float_as_int(mm));  // rates are > 0
      }
    }
  }
}

// ---- forward sampler, one resident wave of blocks, one pool per tile ------------------------------------
// The schedule of the training forward (modes kFwdSoft / kFwdBoth).  What the per-row-tile pool above left on the table at
// config 2 (ncu, round 1): 1366 three-row blocks over 888 resident slots = 1.54 waves (the SMs run half empty through the
// second one: sub-partitions active 81 % of the kernel) and, inside a block, 3 quads per lane are too few to even out chains
// of 1..8+ events (23 of 32 lanes active).  Here the grid is ONE wave (occupancy x SM count blocks, rows dealt out evenly:
// B / grid or one more), and a block pools ALL quads of its rows - every column pass of up to kPoolCap / (K / 4) rows - so
// that a lane keeps pulling work until the block's last quad is taken.  Stage 1 (prologue: softclamps, rate, KL) is
// uniform and coalesced and parks the rates in shared memory; stage 2 is the refill loop of the pool path above.  Long
// chains (more than phase_a events) are handed to phase B by entry index only: a group of 8 lanes walks such a chain again
// from its first event, 8 events per step (the few events phase A already walked are recomputed - it keeps the parked
// state out of shared memory: 18 KB per block instead of 49 KB).  Chain arithmetic, Philox counters and exit rules are
// those of sampler_kernel; a quad finished by phase A is bit-identical to it, one finished by phase B is within an ulp
// (prefix-sum arrival times from event 0 instead of from the parking point).
constexpr int kPoolCap = 768;  // pool entries (quads) per tile
constexpr int kPoolRows = kPoolCap / kThreads;  // rows of one column pass that fit the pool (K <= 512: 6 rows)

template <int MODE, int IND, bool RAW>
__global__ void __launch_bounds__(kThreads) sampler_pool_kernel(const SamplerParams p) {
  __shared__ float s_rate[4][kPoolCap];
  __shared__ unsigned short s_list[kWarps][kPoolCap];  // entries each warp handed to phase B
  __shared__ int s_next;
  __shared__ float s_rowsum[kPoolRows][kWarps];
  __shared__ float s_max[kWarps];

  pdl_launch_dependents();
  pdl_wait();
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  // rows dealt out evenly over the grid: the first B % grid blocks own one row more
  const long long base_rows = p.B / gridDim.x, extra = p.B % gridDim.x;
  const long long blk_row0 = blockIdx.x * base_rows + min(static_cast<long long>(blockIdx.x), extra);
  const int blk_rows = static_cast<int>(base_rows + (blockIdx.x < extra ? 1 : 0));
  uint64_t off = p.offset;
  if (p.offset_dev != nullptr) off += *p.offset_dev;
  const uint32_t off_lo = static_cast<uint32_t>(off), off_hi = static_cast<uint32_t>(off >> 32);
  const int n_a = min(p.phase_a, p.n_exp);
  const int Kq = static_cast<int>(p.K >> 2);
  const int npass = (Kq + kThreads - 1) / kThreads;
  float rmax = 0.f;
  const bool want_rowsum = !RAW && p.kl_rowsum != nullptr;
  const unsigned lanes_below = (1u << lane) - 1u;
  const float t_exit = p.t_exit, absorb = p.absorb;

#pragma unroll 1
  for (int tile_row = 0; tile_row < blk_rows; tile_row += p.rows_per_block) {
    const long long row0 = blk_row0 + tile_row;
    const int nrows = min(p.rows_per_block, blk_rows - tile_row);
    if (tid < kPoolRows * kWarps) (&s_rowsum[0][0])[tid] = 0.f;
    if (tid == 0) s_next = kThreads;
    __syncthreads();
    // ---- stage 1: prologue of every (pass, row) of the tile; entry = (pass * nrows + r) * kThreads + tid
#pragma unroll 1
    for (int pass = 0; pass < npass; ++pass) {
      const int cq = pass * kThreads + tid;
      const bool act = cq < Kq;
      const long long k = 4ll * cq;
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
#pragma unroll 1
      for (int r = 0; r < nrows; ++r) {
        float klsum = 0.f;
        if (act) {
          const long long gbase = (row0 + r) * p.K + k;
          const float4 hv4 = ldg4(p.h + gbase);
          const float hv[4] = {hv4.x, hv4.y, hv4.z, hv4.w};
          Latent q[4];
          float rate[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            q[j] = prologue<RAW, false>(p, hv[j], lr[j], bias[j]);
            rate[j] = q[j].rate;
          }
          if (p.rate != nullptr) stg4(p.rate + gbase, make_float4(rate[0], rate[1], rate[2], rate[3]));
          if (p.rate_max != nullptr) rmax = fmaxf(fmaxf(fmaxf(rate[0], rate[1]), fmaxf(rate[2], rate[3])), rmax);
          if (!RAW) {
            if (p.log_dr != nullptr) stg4(p.log_dr + gbase, make_float4(q[0].log_dr, q[1].log_dr, q[2].log_dr, q[3].log_dr));
            if (p.kl != nullptr || p.kl_rowsum != nullptr) {
              float klv[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float d = q[j].log_dr;
                klv[j] = elr[j] * (1.f + expf(d) * (d - 1.f));  // main/vae.py:448-449
              }
              if (p.kl != nullptr) stg4(p.kl + gbase, make_float4(klv[0], klv[1], klv[2], klv[3]));
              klsum = (klv[0] + klv[1]) + (klv[2] + klv[3]);
            }
          }
          const int slot = (pass * nrows + r) * kThreads + tid;
#pragma unroll
          for (int j = 0; j < 4; ++j) s_rate[j][slot] = rate[j];
        }
        if (want_rowsum) {
          const float sum = warp_sum(klsum);
          if (lane == 0) s_rowsum[r][warp] += sum;
        }
      }
    }
    __syncthreads();
    // ---- stage 2: the tile's npass * nrows * kThreads quads form one pool (entry index == slot index); a shared
    // cursor hands out the next entries, one atomicAdd per warp and refill round
    const int pool_n = npass * nrows * kThreads;
    const uint32_t quad0 = static_cast<uint32_t>(((p.row_offset + row0) * p.K) >> 2);
    float* const z0 = p.z + row0 * p.K;
    float* const zlo0 = p.z_lo != nullptr ? p.z_lo + row0 * p.K : nullptr;
    float* const dz0 = MODE == kFwdBoth ? p.dz_out + row0 * p.K : nullptr;
    // A long chain's state at the hand-over (arrival times, both running sums: 48 bytes per quad) is parked in the quad's own
    // OUTPUT slots - z, z_lo, dz_acc, which nothing reads before the chain is complete - when all three exist (the training
    // step with operand planes).  The lane group then continues the chain from event phase_a exactly as sampler_kernel's
    // phase B does: bit-identical to schedules 0 / 1.  Without the three slots the chain is walked again from its first event.
    const bool keep = MODE == kFwdBoth && zlo0 != nullptr;
    int n_parked = 0;  // warp-uniform
    // entry -> quad index relative to the tile's first quad (row r, column quad cq), or -1 for a column beyond K
    auto locate = [&](int e) -> int {
      const int t = e % kThreads;
      int pr = e / kThreads, pass = 0;
      while (pr >= nrows) pr -= nrows, ++pass;
      const int cq = pass * kThreads + t;
      return cq < Kq ? pr * Kq + cq : -1;
    };
    {
      int cur = tid;  // this lane's pool entry
      bool has = cur < pool_n, live = false;
      int n = 0, ebase = 0;
      uint32_t quad = 0;
      float rate[4], inv_rate[4], t[4], acc[4], acc2[4];
      auto fetch = [&]() {
        live = false;
        if (has) {
          const int qi = locate(cur);
          if (qi >= 0) {
            live = true;
            ebase = qi << 2;
            quad = quad0 + static_cast<uint32_t>(qi);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              rate[j] = s_rate[j][cur];
              inv_rate[j] = recip_for_div(rate[j]);
              t[j] = acc[j] = acc2[j] = 0.f;
            }
            n = 0;
          }
        }
      };
      fetch();
      while (__any_sync(0xffffffffu, has)) {
        bool fin = has && !live;  // an entry beyond the last column: nothing to do
        bool park = false;
        if (live) {
          const uint4 rr = philox4x32_10(quad, static_cast<uint32_t>(n), off_lo, off_hi, p.ks);
          const uint32_t bits[4] = {rr.x, rr.y, rr.z, rr.w};
          float excess = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float e = exponential_from_bits(bits[j]);
            t[j] = __fadd_rn(t[j], div_by_rate<false>(e, rate[j], inv_rate[j]));
            float f, df = 0.f;
            soft_term<IND, MODE == kFwdBoth>(p, t[j], f, df);
            acc[j] += f;
            excess = fmaxf(excess, fmaf(-absorb, acc[j], f));
            if (MODE == kFwdBoth) {
              const float g = df * t[j];
              acc2[j] += g;
              excess = fmaxf(excess, fmaf(-absorb, acc2[j], g));
            }
          }
          ++n;
          const float tmin = fminf(fminf(t[0], t[1]), fminf(t[2], t[3]));
          if (tmin > t_exit || (absorb > 0.f && tmin > 1.f && excess <= 0.f) || n >= p.n_exp) {  // chain complete
            stg4_planes(z0, zlo0, ebase, make_float4(acc[0], acc[1], acc[2], acc[3]));
            if (MODE == kFwdBoth) stg4(dz0 + ebase, make_float4(acc2[0], acc2[1], acc2[2], acc2[3]));
            fin = true;
          } else if (n >= n_a) {  // a long chain: a lane group finishes it, 8 events per step
            if (keep) {
              stg4(z0 + ebase, make_float4(t[0], t[1], t[2], t[3]));
              stg4(zlo0 + ebase, make_float4(acc[0], acc[1], acc[2], acc[3]));
              stg4(dz0 + ebase, make_float4(acc2[0], acc2[1], acc2[2], acc2[3]));
            }
            fin = park = true;
          }
        }
        const unsigned fm = __ballot_sync(0xffffffffu, fin);
        if (fm != 0u) {
          const unsigned pm = __ballot_sync(0xffffffffu, park);
          if (park) s_list[warp][n_parked + __popc(pm & lanes_below)] = static_cast<unsigned short>(cur);
          n_parked += __popc(pm);
          int first = 0;
          if (lane == 0) first = atomicAdd(&s_next, __popc(fm));
          first = __shfl_sync(0xffffffffu, first, 0);
          if (fin) {
            cur = first + __popc(fm & lanes_below);
            has = cur < pool_n;
            fetch();
          }
        }
      }
    }
    // ---- phase B (per warp): groups of kGroup lanes, kGroup events per step, chains restarted from their first event
    if (n_parked > 0) {
      __syncwarp();
      const int gl = lane % kGroup;
      int next_free = 0;
      bool has = false, done = true;
      int ebase = 0, n0 = 0;
      uint32_t quad = 0;
      float rate[4] = {1.f, 1.f, 1.f, 1.f}, inv_rate[4] = {1.f, 1.f, 1.f, 1.f}, t0[4] = {0.f, 0.f, 0.f, 0.f};
      float part[4] = {0.f, 0.f, 0.f, 0.f}, part2[4] = {0.f, 0.f, 0.f, 0.f};
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};  // the sums phase A handed over (zero when restarted)
      while (true) {
        if (__any_sync(0xffffffffu, done)) {
          float tot[4], tot2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v = part[j], v2 = part2[j];
#pragma unroll
            for (int o = kGroup / 2; o > 0; o >>= 1) {
              v += __shfl_xor_sync(0xffffffffu, v, o, kGroup);
              if (MODE == kFwdBoth) v2 += __shfl_xor_sync(0xffffffffu, v2, o, kGroup);
            }
            tot[j] = acc[j] + v;
            if (MODE == kFwdBoth) tot2[j] = acc2[j] + v2;
          }
          if (done && gl == 0 && has) {
            stg4_planes(z0, zlo0, ebase, make_float4(tot[0], tot[1], tot[2], tot[3]));
            if (MODE == kFwdBoth) stg4(dz0 + ebase, make_float4(tot2[0], tot2[1], tot2[2], tot2[3]));
          }
          int next = 0;
          {
            const unsigned req = __ballot_sync(0xffffffffu, done && gl == 0);
            next = next_free + __popc(req & lanes_below);
            next_free += __popc(req);
          }
          next = __shfl_sync(0xffffffffu, next, 0, kGroup);
          if (done) {
            has = next < n_parked;
            if (has) {
              const int slot = s_list[warp][next];
              const int qi = locate(slot);
              ebase = qi << 2;
              quad = quad0 + static_cast<uint32_t>(qi);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                rate[j] = s_rate[j][slot];
                inv_rate[j] = recip_for_div(rate[j]);
                t0[j] = 0.f, part[j] = 0.f, part2[j] = 0.f, acc[j] = 0.f, acc2[j] = 0.f;
              }
              n0 = 0;
              if (keep) {  // continue where the single lane stopped (coherent loads: this kernel wrote them)
                const float4 tv = ld4(z0 + ebase), av = ld4(zlo0 + ebase), a2 = ld4(dz0 + ebase);
                t0[0] = tv.x, t0[1] = tv.y, t0[2] = tv.z, t0[3] = tv.w;
                acc[0] = av.x, acc[1] = av.y, acc[2] = av.z, acc[3] = av.w;
                acc2[0] = a2.x, acc2[1] = a2.y, acc2[2] = a2.z, acc2[3] = a2.w;
                n0 = n_a;
              }
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
          soft_term<IND, MODE == kFwdBoth>(p, t, f, df);
          const float g = df * t;
          part[j] += valid ? f : 0.f;
          // the group's last lane holds the smallest term of the step; its own sums bound the totals from below
          excess = fmaxf(excess, fmaf(-absorb, acc[j] + part[j], f));
          if (MODE == kFwdBoth) {
            part2[j] += valid ? g : 0.f;
            excess = fmaxf(excess, fmaf(-absorb, acc2[j] + part2[j], g));
          }
          t0[j] = __shfl_sync(0xffffffffu, t, kGroup - 1, kGroup);
        }
        n0 += kGroup;
        const float tmin = fminf(fminf(t0[0], t0[1]), fminf(t0[2], t0[3]));
        const bool absorbed = absorb > 0.f && __shfl_sync(0xffffffffu, excess <= 0.f ? 1 : 0, kGroup - 1, kGroup) != 0 && tmin > 1.f;
        if (has && (n0 >= p.n_exp || tmin > t_exit || absorbed)) done = true;
      }
    }
    __syncthreads();  // the pool and the row sums are complete; the tables are reused by the next tile
    if (want_rowsum && tid < nrows) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) sum += s_rowsum[tid][w];
      p.kl_rowsum[row0 + tid] = sum;
    }
  }
  if (p.rate_max != nullptr) {
    const float m = warp_max(rmax);
    if (lane == 0) s_max[warp] = m;
    __syncthreads();
    if (tid == 0) {
      float mm = s_max[0];
#pragma unroll
      for (int w = 1; w < kWarps; ++w) mm = fmaxf(mm, s_max[w]);
      atomicMax(reinterpret_cast<int*>(p.rate_max), __float_as_int(mm));  // rates are > 0
    }
  }
}

// ---- backward from a stored dz_acc: one elementwise pass ------------------------------------------
// A block owns kSavedRows rows and all K columns; a thread owns a quad of columns and keeps every row's
// loads in flight before it computes (9..24 independent 128-bit loads per thread), then writes its column
// partial sums.  No chain, no RNG.
constexpr int kSavedRows = 4;

// What the elementwise backward needs of one latent, from the softplus intermediates alone: with y1 = 10 - (h + b),
// w = e^{-|y|}, s = 1 + w:  d delta / d h = sigmoid(y1),  e^{delta} = e^{10} sigmoid(-y1) (no second exponential),
// delta = 10 - softplus(y1) (one lg2);  with y2 = 5 - (log_r + delta):  d lambda / d u = sigmoid(y2),
// e^{lambda} = e^5 sigmoid(-y2) (no exponential, no second lg2),  e^{lambda} / rate by an approximate reciprocal.
// 6 MUFU and ~55 instructions per latent instead of ~95 through prologue<.., true>() + expf + IEEE division; the values
// differ from those by their rounding (~2e-7 relative), the forward is untouched.  exc_only keeps the general path.
struct BwdTerms {
  float delta, ed, d_dr, ratio_drate;  // ratio_drate = e^{lambda} / rate * d lambda / d u
};
__device__ __forceinline__ BwdTerms bwd_terms(const SamplerParams& p, float e_dr, float e_rate, float hv, float lr, float bias) {
  BwdTerms o;
  const float y1 = p.clamp_dr - (hv + bias);
  const float w1 = ex2_approx(-fabsf(y1) * kLog2e), s1 = 1.f + w1, r1 = rcp_approx(s1);
  o.d_dr = (y1 >= 0.f ? 1.f : w1) * r1;
  o.ed = e_dr * ((y1 >= 0.f ? w1 : 1.f) * r1);  // e_dr = e^{clamp_dr}, e_rate = e^{clamp_rate}: once per thread
  o.delta = p.clamp_dr - fmaf(kLn2, lg2_approx(s1), fmaxf(y1, 0.f));
  const float y2 = p.clamp_rate - (lr + o.delta);
  const float w2 = ex2_approx(-fabsf(y2) * kLog2e), r2 = rcp_approx(1.f + w2);
  const float explam = e_rate * ((y2 >= 0.f ? w2 : 1.f) * r2);
  const float d_rate = (y2 >= 0.f ? 1.f : w2) * r2;
  o.ratio_drate = explam * rcp_approx(explam + kF32Eps) * d_rate;
  return o;
}

template <bool RAW>
__global__ void __launch_bounds__(kThreads, 7) sampler_bwd_saved_kernel(const SamplerParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const long long row0 = static_cast<long long>(blockIdx.x) * kSavedRows;
  const int nrows = static_cast<int>(min(static_cast<long long>(kSavedRows), p.B - row0));
  const float e_dr = expf(p.clamp_dr), e_rate = expf(p.clamp_rate);
  for (long long k = 4 * threadIdx.x; k < p.K; k += kColsPerPass) {
    float lr[4] = {0.f, 0.f, 0.f, 0.f}, bias[4] = {0.f, 0.f, 0.f, 0.f}, elr[4] = {0.f, 0.f, 0.f, 0.f};
    if (!RAW) {
      const float4 v = ldg4(p.log_r + k);
      lr[0] = v.x, lr[1] = v.y, lr[2] = v.z, lr[3] = v.w;
      if (p.b_enc != nullptr) {
        const float4 b = ldg4(p.b_enc + k);
        bias[0] = b.x, bias[1] = b.y, bias[2] = b.z, bias[3] = b.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) elr[j] = expf(lr[j]);
    }
    const bool has_gld = !RAW && p.g_log_dr != nullptr;
    // two rows at a time: 6 (8) independent 128-bit loads per thread in flight, at up to 32 resident warps per SM
    float col_lr[4] = {0.f, 0.f, 0.f, 0.f}, col_b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int r0 = 0; r0 < nrows; r0 += 2) {
      float4 hv[2], gz[2], dz[2], gl[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (r0 + u < nrows) {
          const long long base = (row0 + r0 + u) * p.K + k;
          // g_h may be written over g_z and its low plane over dz_acc (same index, this thread): coherent loads, not ld.nc
          hv[u] = ldg4(p.h + base), gz[u] = ld4(p.g_z + base), dz[u] = ld4(p.dz_in + base);
          gl[u] = has_gld ? ldg4(p.g_log_dr + base) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (r0 + u < nrows) {
          const long long base = (row0 + r0 + u) * p.K + k;
          const float h4[4] = {hv[u].x, hv[u].y, hv[u].z, hv[u].w};
          const float g4[4] = {gz[u].x, gz[u].y, gz[u].z, gz[u].w};
          const float d4[4] = {dz[u].x, dz[u].y, dz[u].z, dz[u].w};
          const float l4[4] = {gl[u].x, gl[u].y, gl[u].z, gl[u].w};
          float gh[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!RAW && !p.exc_only) {
              const BwdTerms q = bwd_terms(p, e_dr, e_rate, h4[j], lr[j], bias[j]);
              // d z / d lambda = (sum_n f'(l_n) t_n) / temp * exp(lambda) / rate; then the clamp chain
              const float g_u = g4[j] * (d4[j] * p.inv_temp * q.ratio_drate);
              const float d = q.delta, ed = q.ed;
              gh[j] = (g_u + p.kl_scale * elr[j] * d * ed + l4[j]) * q.d_dr;
              col_lr[j] += g_u + p.kl_scale * (elr[j] * (1.f + ed * (d - 1.f)));
              col_b[j] += gh[j];
              continue;
            }
            const Latent q = prologue<RAW, true>(p, h4[j], lr[j], bias[j]);
            const float g_u = g4[j] * (d4[j] * p.inv_temp * (q.explam / q.rate)) * q.d_rate;
            if (RAW) {
              gh[j] = g_u;
            } else {
              const float d = q.log_dr, ed = expf(d);
              gh[j] = (g_u + p.kl_scale * elr[j] * d * ed + l4[j]) * q.d_dr;
              col_lr[j] += g_u + p.kl_scale * (elr[j] * (1.f + ed * (d - 1.f)));
              col_b[j] += gh[j];
            }
          }
          stg4_planes(p.g_h, p.g_h_lo, base, make_float4(gh[0], gh[1], gh[2], gh[3]));
        }
      }
    }
    if (!RAW) {
      const long long pbase = static_cast<long long>(blockIdx.x) * p.K + k;
      stg4(p.part_log_r + pbase, make_float4(col_lr[0], col_lr[1], col_lr[2], col_lr[3]));
      if (p.part_b != nullptr) stg4(p.part_b + pbase, make_float4(col_b[0], col_b[1], col_b[2], col_b[3]));
    }
  }
}

// ---- backward from the forward's coefficients (coefficient mode): pure streaming ------------------------------
//   g_u = g_z A;  g_h = (g_u + kl_scale c_ck) c_ddr;  column partials: sum_b g_u (-> g_log_r, the KL part comes from the
//   forward's klcol sums) and sum_b g_h (-> g_b_enc).  A block owns kCoefRows rows and all K columns.
constexpr int kCoefRows = 4;
__global__ void __launch_bounds__(kThreads, 8) sampler_bwd_coef_kernel(const SamplerParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const long long row0 = static_cast<long long>(blockIdx.x) * kCoefRows;
  const int nrows = static_cast<int>(min(static_cast<long long>(kCoefRows), p.B - row0));
  for (long long k = 4 * threadIdx.x; k < p.K; k += kColsPerPass) {
    float4 col_lr = make_float4(0.f, 0.f, 0.f, 0.f), col_b = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r0 = 0; r0 < kCoefRows; r0 += 2) {  // two rows (eight 128-bit loads) in flight per thread, 32 warps per SM
      float4 gz[2], a[2], dd[2], ck[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (r0 + u < nrows) {
          const long long base = (row0 + r0 + u) * p.K + k;
          gz[u] = ld4(p.g_z + base), a[u] = ld4(p.c_a + base), dd[u] = ldg4(p.c_ddr + base), ck[u] = ldg4(p.c_ck + base);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (r0 + u < nrows) {
          const float4 gu = make_float4(gz[u].x * a[u].x, gz[u].y * a[u].y, gz[u].z * a[u].z, gz[u].w * a[u].w);
          const float4 gh = make_float4(fmaf(p.kl_scale, ck[u].x, gu.x) * dd[u].x, fmaf(p.kl_scale, ck[u].y, gu.y) * dd[u].y,
                                        fmaf(p.kl_scale, ck[u].z, gu.z) * dd[u].z, fmaf(p.kl_scale, ck[u].w, gu.w) * dd[u].w);
          col_lr.x += gu.x, col_lr.y += gu.y, col_lr.z += gu.z, col_lr.w += gu.w;
          col_b.x += gh.x, col_b.y += gh.y, col_b.z += gh.z, col_b.w += gh.w;
          stg4_planes(p.g_h, p.g_h_lo, (row0 + r0 + u) * p.K + k, gh);
        }
      }
    }
    const long long pbase = static_cast<long long>(blockIdx.x) * p.K + k;
    stg4(p.part_log_r + pbase, col_lr);
    if (p.part_b != nullptr) stg4(p.part_b + pbase, col_b);
  }
}

// ---- fixed-order column reduction of the (P,K) partials ----------------------------------------
// One block per 32 columns; 32 row-lanes each keep four independent running sums (rows ty + 32 * (4i + u)),
// then a fixed-order combine through shared memory.  Deterministic for a given P.
constexpr int kRedX = 32, kRedY = 32;
__global__ void __launch_bounds__(kRedX* kRedY) colsum_finalize_kernel(const float* __restrict__ part, float* __restrict__ out,
                                                                      long long P, long long K, int accumulate, float scale) {
  __shared__ float s[kRedY][kRedX + 1];
  pdl_launch_dependents();
  pdl_wait();
  const int tx = threadIdx.x, ty = threadIdx.y;
  const long long k = static_cast<long long>(blockIdx.x) * kRedX + tx;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (k < K) {
    long long r = ty;
    for (; r + 3 * kRedY < P; r += 4 * kRedY) {
      a0 += __ldg(part + r * K + k);
      a1 += __ldg(part + (r + kRedY) * K + k);
      a2 += __ldg(part + (r + 2 * kRedY) * K + k);
      a3 += __ldg(part + (r + 3 * kRedY) * K + k);
    }
    for (; r < P; r += kRedY) a0 += __ldg(part + r * K + k);
  }
  s[ty][tx] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (ty == 0 && k < K) {
    float v = s[0][tx];
#pragma unroll
    for (int y = 1; y < kRedY; ++y) v += s[y][tx];
    out[k] = accumulate ? fmaf(scale, v, out[k]) : (scale == 1.f ? v : scale * v);
  }
}

// ---- standalone KL ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kl_fwd_kernel(const float* __restrict__ log_dr, const float* __restrict__ log_r,
                                                     float* __restrict__ kl, long long n4, long long K) {
  pdl_launch_dependents();
  pdl_wait();
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

// kl_diag of main/train_vae.py:348 (torch.mean(kl, dim=0)) from the encoder output: per-block column sums of the KL over
// kDiagRows rows; colsum_finalize adds the blocks in order and scales by 1 / B.  Runs at logging steps only.
constexpr int kDiagRows = 32;
__global__ void __launch_bounds__(256) kl_diag_kernel(const SamplerParams p, float* __restrict__ part) {
  pdl_launch_dependents();
  pdl_wait();
  const long long row0 = static_cast<long long>(blockIdx.x) * kDiagRows;
  const int nrows = static_cast<int>(min(static_cast<long long>(kDiagRows), p.B - row0));
  for (long long k = 4ll * threadIdx.x; k < p.K; k += 4 * 256) {
    const float4 lr4 = ldg4(p.log_r + k);
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.b_enc != nullptr) b4 = ldg4(p.b_enc + k);
    const float lr[4] = {lr4.x, lr4.y, lr4.z, lr4.w}, bias[4] = {b4.x, b4.y, b4.z, b4.w};
    float elr[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) elr[j] = expf(lr[j]);
    for (int r = 0; r < nrows; ++r) {
      const float4 hv4 = ldg4(p.h + (row0 + r) * p.K + k);
      const float hv[4] = {hv4.x, hv4.y, hv4.z, hv4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = prologue<false, false>(p, hv[j], lr[j], bias[j]).log_dr;
        acc[j] += elr[j] * (1.f + expf(d) * (d - 1.f));  // main/vae.py:448-449
      }
    }
    stg4(part + static_cast<long long>(blockIdx.x) * p.K + k, make_float4(acc[0], acc[1], acc[2], acc[3]));
  }
}

__global__ void __launch_bounds__(kThreads) kl_bwd_kernel(const float* __restrict__ log_dr, const float* __restrict__ log_r,
                                                          const float* __restrict__ g_kl, float* __restrict__ g_log_dr,
                                                          float* __restrict__ part, long long B, long long K, int rows_per_block) {
  pdl_launch_dependents();
  pdl_wait();
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

// ---- parity harness: dump the draws; count mismatches of neg_log_unit against -logf --------------
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

__global__ void debug_log_check_kernel(unsigned long long* mismatches) {
  unsigned long long bad = 0;
  for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m < (1u << 23); m += gridDim.x * blockDim.x) {
    const float u = uniform_from_bits(m << 9);
    if (__float_as_int(neg_log_unit(u)) != __float_as_int(-logf(u))) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

// ---- host-side launch --------------------------------------------------------------------------
static int g_rows_per_block = 0;  // 0 = automatic

static int rows_per_block_for(long long B) {
  if (g_rows_per_block > 0) return g_rows_per_block;
  // enough blocks for several waves over 148 SMs, at most kMaxRows rows per block
  const long long target_blocks = 148ll * 8;
  long long r = B / target_blocks;
  if (r < 1) r = 1;
  if (r > kMaxRows) r = kMaxRows;
  return static_cast<int>(r);
}

static int g_one_wave = 0;  // forward modes: 1 = the one-wave schedule (sampler_pool_kernel, pvae_set_pool(2)); 0 = the per-row-tile pool of sampler_kernel
float g_wave_factor = 1.f;  // blocks of the one-wave schedule as a multiple of the resident wave (pvae_set_wave_factor, A/B)
static int g_pool_auto = 1;  // default (pvae_set_pool(3)): one wave where it is bit-identical to the tile pool (training forward with operand planes)

// One resident wave: occupancy x SM count blocks (queried once per kernel), rows dealt out evenly in the kernel, tiles
// of as many rows as fit the pool.
template <typename Kern>
static cudaError_t launch_one_wave(Kern kern, SamplerParams p, cudaStream_t s) {
  extern float g_wave_factor;
  static int resident = 0;  // one copy per kernel instantiation
  if (resident == 0) {
    int per_sm = 0, dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0) != cudaSuccess || per_sm < 1 || sms < 1)
      per_sm = 6, sms = 148;
    resident = per_sm * sms;
  }
  const long long grid = std::max<long long>(1, std::min<long long>(static_cast<long long>(resident * g_wave_factor + 0.5f), p.B));
  const int npass = static_cast<int>((p.K / 4 + kThreads - 1) / kThreads);
  const long long rows_block = (p.B + grid - 1) / grid;
  p.rows_per_block = static_cast<int>(std::min<long long>(kPoolRows / npass, rows_block));
  return launch_pdl(kern, dim3(static_cast<unsigned>(grid)), dim3(kThreads), 0, s, p);
}

template <typename Kern>
static cudaError_t launch_pool(Kern kern, SamplerParams p, cudaStream_t s) {
  const dim3 grid(static_cast<unsigned>((p.B + p.rows_block - 1) / p.rows_block));
  return launch_pdl(kern, grid, dim3(kThreads), 0, s, p);
}

template <int MODE, bool RAW>
static cudaError_t launch_ind(const SamplerParams& p, int indicator, dim3 grid, cudaStream_t s) {
  const bool auto_wave = g_pool_auto && MODE == kFwdBoth && p.z_lo != nullptr && p.c_ddr == nullptr;
  if ((MODE == kFwdSoft || MODE == kFwdBoth) && p.pool && (g_one_wave || auto_wave) && p.K / 4 <= kPoolCap && p.K / 4 > 0 &&
      kPoolRows / static_cast<int>((p.K / 4 + kThreads - 1) / kThreads) >= 1) {  // one-wave pool (forward modes)
    constexpr int M = (MODE == kFwdSoft || MODE == kFwdBoth) ? MODE : kFwdSoft;
    switch (indicator) {
      case kSigmoid: return launch_one_wave(sampler_pool_kernel<M, kSigmoid, RAW>, p, s);
      case kCosine: return launch_one_wave(sampler_pool_kernel<M, kCosine, RAW>, p, s);
      case kLinear: return launch_one_wave(sampler_pool_kernel<M, kLinear, RAW>, p, s);
      case kCubic: return launch_one_wave(sampler_pool_kernel<M, kCubic, RAW>, p, s);
      case kQuintic: return launch_one_wave(sampler_pool_kernel<M, kQuintic, RAW>, p, s);
      default: return cudaErrorInvalidValue;
    }
  }
  if ((MODE == kFwdSoft || MODE == kFwdBoth) && p.pool) {  // per-row-tile pool (the previous schedule, A/B)
    constexpr int M = (MODE == kFwdSoft || MODE == kFwdBoth) ? MODE : kFwdSoft;
    switch (indicator) {
      case kSigmoid: return launch_pool(sampler_kernel<M, kSigmoid, RAW, true>, p, s);
      case kCosine: return launch_pool(sampler_kernel<M, kCosine, RAW, true>, p, s);
      case kLinear: return launch_pool(sampler_kernel<M, kLinear, RAW, true>, p, s);
      case kCubic: return launch_pool(sampler_kernel<M, kCubic, RAW, true>, p, s);
      case kQuintic: return launch_pool(sampler_kernel<M, kQuintic, RAW, true>, p, s);
      default: return cudaErrorInvalidValue;
    }
  }
  switch (indicator) {
    case kSigmoid: return launch_pdl(sampler_kernel<MODE, kSigmoid, RAW>, grid, dim3(kThreads), 0, s, p);
    case kCosine: return launch_pdl(sampler_kernel<MODE, kCosine, RAW>, grid, dim3(kThreads), 0, s, p);
    case kLinear: return launch_pdl(sampler_kernel<MODE, kLinear, RAW>, grid, dim3(kThreads), 0, s, p);
    case kCubic: return launch_pdl(sampler_kernel<MODE, kCubic, RAW>, grid, dim3(kThreads), 0, s, p);
    case kQuintic: return launch_pdl(sampler_kernel<MODE, kQuintic, RAW>, grid, dim3(kThreads), 0, s, p);
    default: return cudaErrorInvalidValue;
  }
}

static int g_phase_a_events = 8;
static int g_pool = 1;  // forward modes: per-warp work pool (0: row-by-row lock step, the previous schedule)
static bool g_absorb_exit = true;

// Chain-control fields: temperature constants, early-exit threshold and the phase split.
static void set_chain(SamplerParams& p, int n_exp, float temp, int indicator, float exit_logit);
void sampler_set_chain(SamplerParams& p, int n_exp, float temp, int indicator, float exit_logit) { set_chain(p, n_exp, temp, indicator, exit_logit); }
static void set_chain(SamplerParams& p, int n_exp, float temp, int indicator, float exit_logit) {
  const bool hard = temp == 0.f;
  p.n_exp = n_exp;
  p.inv_temp = hard ? 0.f : 1.f / temp;
  p.sig_a = hard ? 0.f : 1.4426950408889634f / temp;
  p.pool = g_pool && n_exp > 0;
  const int split = g_phase_a_events < n_exp ? g_phase_a_events : n_exp;
  p.absorb = 0.f;
  if (hard) {  // sequential fp32 running sum (bit-exact against torch's CUDA cumsum)
    p.t_exit = exit_logit < 0.f ? __builtin_inff() : 1.f;  // every later arrival is > 1 as well
    p.phase_a = n_exp;
  } else if (exit_logit < 0.f) {  // never: walk all n_exp events (same phase split, hence same bits)
    p.t_exit = __builtin_inff();
    p.phase_a = split;
  } else {
    // fp32-absorbed exit (only with the lossless policy): 2^-34 leaves a factor 2^9 > n_exp of head room
    if (exit_logit == 0.f && g_absorb_exit) p.absorb = 5.8207660913467407e-11f;
    // lossless: the reference's fp32 term is exactly 0 below this logit (sigmoid: exp(-l) overflows for
    // l < -88.7228; finite-support indicators vanish for l <= -1); a hair of slack keeps the test
    // t > t_exit on the safe side of the rounding of 1 + c * temp.
    const float c = exit_logit == 0.f ? (indicator == kSigmoid ? 88.73f : 1.0f) : exit_logit;
    p.t_exit = (1.f + c * temp) * (1.f + 4e-7f);
    p.phase_a = split;
  }
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

static int finalize_cols(const float* part, float* out, long long P, long long K, int accumulate, cudaStream_t s, float scale = 1.f) {
  dim3 grid(static_cast<unsigned>((K + kRedX - 1) / kRedX)), block(kRedX, kRedY);
  return cuda_ok("colsum_finalize", launch_pdl(colsum_finalize_kernel, grid, block, 0, s, part, out, P, K, accumulate, scale));
}

}  // namespace pvae

using namespace pvae;

extern "C" int64_t pvae_num_partials(int64_t B) {
  if (B <= 0) return 0;
  const int r = rows_per_block_for(B);
  return (B + r - 1) / r;
}

extern "C" int pvae_set_pool(int on) {  // 0: row-by-row lock step; 1: tile pool; 2: one-wave pool; 3 (default): one wave where bit-identical to 1
  if (on < 0 || on > 3) return fail(PVAE_ERR_INVALID, "pvae_set_pool: %d not in [0, 3]", on);
  g_pool = on != 0;
  g_one_wave = on == 2;
  g_pool_auto = on == 3;
  return PVAE_OK;
}

extern "C" int pvae_set_wave_factor(float f) {
  if (!(f > 0.f) || f > 16.f) return fail(PVAE_ERR_INVALID, "pvae_set_wave_factor: %g not in (0, 16]", f);
  g_wave_factor = f;
  return PVAE_OK;
}

extern "C" int pvae_set_rows_per_block(int n) {
  if (n < 0 || n > kMaxRows) return fail(PVAE_ERR_INVALID, "pvae_set_rows_per_block: n=%d must be in [0, %d]", n, kMaxRows);
  g_rows_per_block = n;
  return PVAE_OK;
}

extern "C" int pvae_set_absorb_exit(int on) {
  g_absorb_exit = on != 0;
  return PVAE_OK;
}

extern "C" int pvae_set_phase_a_events(int n) {
  if (n < 1) return fail(PVAE_ERR_INVALID, "pvae_set_phase_a_events: n=%d must be >= 1", n);
  g_phase_a_events = n;
  return PVAE_OK;
}

extern "C" int pvae_sampler_fwd(const float* h, const float* log_r, const float* b_enc, float* z, float* log_dr,
                                float* rate, float* kl, float* kl_rowsum, float* rate_max, float* dz_acc, int64_t B,
                                int64_t K, int64_t row_offset, int n_exp, float temp, int indicator, int exc_only,
                                float exit_logit, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                                void* stream) {
  return pvae::sampler_fwd_impl(h, log_r, b_enc, z, nullptr, log_dr, rate, kl, kl_rowsum, rate_max, dz_acc, B, K, row_offset, n_exp, temp,
                                indicator, exc_only, exit_logit, seed, offset, offset_dev, stream);
}

int pvae::sampler_fwd_impl(const float* h, const float* log_r, const float* b_enc, float* z, float* z_lo, float* log_dr, float* rate,
                           float* kl, float* kl_rowsum, float* rate_max, float* dz_acc, int64_t B, int64_t K, int64_t row_offset,
                           int n_exp, float temp, int indicator, int exc_only, float exit_logit, uint64_t seed, uint64_t offset,
                           const uint64_t* offset_dev, void* stream) {
  if (int rc = check_common("pvae_sampler_fwd", B, K, row_offset, n_exp, temp, indicator)) return rc;
  if (B == 0) return PVAE_OK;
  if (!h || !log_r || !z) return fail(PVAE_ERR_INVALID, "pvae_sampler_fwd: null h/log_r/z");
  SamplerParams p{};
  p.h = h, p.log_r = log_r, p.b_enc = b_enc, p.z = z, p.z_lo = z_lo, p.log_dr = log_dr, p.rate = rate, p.kl = kl;
  p.kl_rowsum = kl_rowsum, p.rate_max = rate_max, p.dz_out = dz_acc;
  if (dz_acc && temp == 0.f) return fail(PVAE_ERR_INVALID, "pvae_sampler_fwd: hard counts (temp == 0) have no gradient (dz_acc)");
  p.B = B, p.K = K, p.row_offset = row_offset, p.rows_per_block = p.rows_block = rows_per_block_for(B);
  set_chain(p, n_exp, temp, indicator, exit_logit);
  p.clamp_dr = 10.f, p.clamp_rate = 5.f, p.exc_only = exc_only;  // main/vae.py:403-409
  p.offset = offset, p.offset_dev = offset_dev;
  philox_key_schedule(seed, p.ks);
  dim3 grid(static_cast<unsigned>((B + p.rows_per_block - 1) / p.rows_per_block));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = temp == 0.f ? launch_ind<kFwdHard, false>(p, kSigmoid, grid, s)
                  : dz_acc    ? launch_ind<kFwdBoth, false>(p, indicator, grid, s)
                              : launch_ind<kFwdSoft, false>(p, indicator, grid, s);
  return cuda_ok("pvae_sampler_fwd", e);
}

// Training forward in coefficient mode (see SamplerParams::c_ddr) and the matching streaming backward.
int pvae::sampler_fwd_coef_impl(const float* h, const float* log_r, const float* b_enc, float* z, float* z_lo, float* kl_rowsum,
                                float* rate_max, float* c_a, float* c_ddr, float* c_ck, float* klcol, int64_t B, int64_t K,
                                int64_t row_offset, int n_exp, float temp, int indicator, int exc_only, float exit_logit, uint64_t seed,
                                uint64_t offset, const uint64_t* offset_dev, void* stream) {
  if (int rc = check_common("pvae_sampler_fwd", B, K, row_offset, n_exp, temp, indicator)) return rc;
  if (B == 0) return PVAE_OK;
  if (!h || !log_r || !z || !c_a || !c_ddr || !c_ck || !klcol || temp == 0.f)
    return fail(PVAE_ERR_INVALID, "sampler_fwd_coef: null pointer or temp == 0");
  SamplerParams p{};
  p.h = h, p.log_r = log_r, p.b_enc = b_enc, p.z = z, p.z_lo = z_lo, p.kl_rowsum = kl_rowsum, p.rate_max = rate_max;
  p.dz_out = c_a, p.c_ddr = c_ddr, p.c_ck = c_ck, p.klcol = klcol;
  p.B = B, p.K = K, p.row_offset = row_offset, p.rows_per_block = p.rows_block = rows_per_block_for(B);
  set_chain(p, n_exp, temp, indicator, exit_logit);
  p.clamp_dr = 10.f, p.clamp_rate = 5.f, p.exc_only = exc_only;  // main/vae.py:403-409
  p.offset = offset, p.offset_dev = offset_dev;
  philox_key_schedule(seed, p.ks);
  dim3 grid(static_cast<unsigned>((B + p.rows_per_block - 1) / p.rows_per_block));
  const bool one_wave = g_one_wave;
  g_one_wave = 0;  // the one-wave schedule has no coefficient mode
  const cudaError_t e = launch_ind<kFwdBoth, false>(p, indicator, grid, static_cast<cudaStream_t>(stream));
  g_one_wave = one_wave;
  return cuda_ok("pvae_sampler_fwd", e);
}

int pvae::sampler_bwd_coef_impl(const float* g_z, const float* c_a, const float* c_ddr, const float* c_ck, float kl_scale, float* g_h,
                                float* g_h_lo, float* partial_ws, int want_bias, const float* klcol, int64_t klcol_rows, float* g_log_r,
                                float* g_b_enc, int64_t B, int64_t K, void* stream) {
  if (B <= 0 || K <= 0 || K % 4 || !g_z || !c_a || !c_ddr || !c_ck || !g_h || !partial_ws)
    return fail(PVAE_ERR_INVALID, "sampler_bwd_coef: bad argument");
  SamplerParams p{};
  p.g_z = g_z, p.c_a = c_a, p.c_ddr = const_cast<float*>(c_ddr), p.c_ck = const_cast<float*>(c_ck), p.kl_scale = kl_scale, p.g_h = g_h, p.g_h_lo = g_h_lo;
  p.B = B, p.K = K;
  const long long P = (B + kCoefRows - 1) / kCoefRows;
  p.part_log_r = partial_ws;
  p.part_b = want_bias ? partial_ws + P * K : nullptr;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = cuda_ok("sampler_bwd_coef", launch_pdl(sampler_bwd_coef_kernel, dim3(static_cast<unsigned>(P)), dim3(kThreads), 0, s, p)))
    return rc;
  if (g_log_r == nullptr) return PVAE_OK;  // partials only: the caller reduces them (pvae_adamax_from_partials)
  // g_log_r = sum_b g_u + kl_scale sum_b kl; g_b_enc = sum_b g_h
  if (int rc = finalize_cols(p.part_log_r, g_log_r, P, K, 0, s)) return rc;
  if (int rc = finalize_cols(klcol, g_log_r, klcol_rows, K, 1, s, kl_scale)) return rc;
  if (want_bias && g_b_enc)
    if (int rc = finalize_cols(p.part_b, g_b_enc, P, K, 0, s)) return rc;
  return PVAE_OK;
}
int64_t pvae::sampler_bwd_coef_rows(int64_t B) { return B <= 0 ? 0 : (B + kCoefRows - 1) / kCoefRows; }

extern "C" int pvae_sampler_bwd(const float* h, const float* log_r, const float* b_enc, const float* g_z,
                                const float* g_log_dr, const float* dz_acc, float kl_scale, float* g_h, float* g_log_r,
                                float* g_b_enc,
                                float* partial_ws, int accumulate, int64_t B, int64_t K, int64_t row_offset,
                                int n_exp, float temp, int indicator, int exc_only, float exit_logit, uint64_t seed,
                                uint64_t offset, const uint64_t* offset_dev, void* stream) {
  return pvae::sampler_bwd_impl(h, log_r, b_enc, g_z, g_log_dr, dz_acc, kl_scale, g_h, nullptr, g_log_r, g_b_enc, partial_ws, accumulate, B, K,
                                row_offset, n_exp, temp, indicator, exc_only, exit_logit, seed, offset, offset_dev, stream);
}

int pvae::sampler_bwd_impl(const float* h, const float* log_r, const float* b_enc, const float* g_z, const float* g_log_dr,
                           const float* dz_acc, float kl_scale, float* g_h, float* g_h_lo, float* g_log_r, float* g_b_enc,
                           float* partial_ws, int accumulate, int64_t B, int64_t K, int64_t row_offset, int n_exp, float temp,
                           int indicator, int exc_only, float exit_logit, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                           void* stream) {
  if (int rc = check_common("pvae_sampler_bwd", B, K, row_offset, n_exp, temp, indicator)) return rc;
  if (temp == 0.f) return fail(PVAE_ERR_INVALID, "pvae_sampler_bwd: hard counts (temp == 0) have no gradient");
  if (B == 0) return PVAE_OK;
  if (!h || !log_r || !g_z || !g_h || !g_log_r || !partial_ws)
    return fail(PVAE_ERR_INVALID, "pvae_sampler_bwd: null h/log_r/g_z/g_h/g_log_r/partial_ws");
  SamplerParams p{};
  p.h = h, p.log_r = log_r, p.b_enc = b_enc, p.g_z = g_z, p.g_log_dr = g_log_dr, p.g_h = g_h, p.g_h_lo = g_h_lo, p.dz_in = dz_acc;
  p.B = B, p.K = K, p.row_offset = row_offset, p.rows_per_block = p.rows_block = rows_per_block_for(B);
  const long long P = (B + p.rows_per_block - 1) / p.rows_per_block;
  p.part_log_r = partial_ws;
  p.part_b = g_b_enc ? partial_ws + P * K : nullptr;
  set_chain(p, n_exp, temp, indicator, exit_logit);
  p.clamp_dr = 10.f, p.clamp_rate = 5.f, p.exc_only = exc_only, p.kl_scale = kl_scale;
  p.offset = offset, p.offset_dev = offset_dev;
  philox_key_schedule(seed, p.ks);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  long long Pk = P;  // partial rows actually written
  if (dz_acc) {
    Pk = (B + kSavedRows - 1) / kSavedRows;
    if (int rc = cuda_ok("pvae_sampler_bwd", launch_pdl(sampler_bwd_saved_kernel<false>, dim3(static_cast<unsigned>(Pk)),
                                                       dim3(kThreads), 0, s, p)))
      return rc;
  } else if (int rc = cuda_ok("pvae_sampler_bwd", launch_ind<kBwd, false>(p, indicator, dim3(static_cast<unsigned>(P)), s))) {
    return rc;
  }
  if (accumulate < 0) return PVAE_OK;  // partials only: pvae_sampler_bwd_partial_rows(B) rows of K (log_r), then of K (b_enc)
  if (int rc = finalize_cols(p.part_log_r, g_log_r, Pk, K, accumulate, s)) return rc;
  if (g_b_enc)
    if (int rc = finalize_cols(p.part_b, g_b_enc, Pk, K, accumulate, s)) return rc;
  return PVAE_OK;
}

// rows of column partials the saved-derivative backward writes (dz_acc given), and where the b_enc partials start
extern "C" int64_t pvae_sampler_bwd_partial_rows(int64_t B) { return B <= 0 ? 0 : (B + kSavedRows - 1) / kSavedRows; }
extern "C" int64_t pvae_sampler_bwd_bias_offset_rows(int64_t B) { return pvae_num_partials(B); }

extern "C" int pvae_rsample_fwd(const float* log_rate, float* z, float* rate, float* dz_acc, int64_t B, int64_t K, int64_t row_offset,
                                int n_exp, float temp, int indicator, float clamp, float exit_logit, uint64_t seed,
                                uint64_t offset, const uint64_t* offset_dev, void* stream) {
  if (int rc = check_common("pvae_rsample_fwd", B, K, row_offset, n_exp, temp, indicator)) return rc;
  if (B == 0) return PVAE_OK;
  if (!log_rate || !z) return fail(PVAE_ERR_INVALID, "pvae_rsample_fwd: null log_rate/z");
  SamplerParams p{};
  p.h = log_rate, p.z = z, p.rate = rate, p.dz_out = dz_acc;
  if (dz_acc && temp == 0.f) return fail(PVAE_ERR_INVALID, "pvae_rsample_fwd: hard counts (temp == 0) have no gradient (dz_acc)");
  p.B = B, p.K = K, p.row_offset = row_offset, p.rows_per_block = p.rows_block = rows_per_block_for(B);
  set_chain(p, n_exp, temp, indicator, exit_logit);
  p.clamp_rate = clamp;
  p.offset = offset, p.offset_dev = offset_dev;
  philox_key_schedule(seed, p.ks);
  dim3 grid(static_cast<unsigned>((B + p.rows_per_block - 1) / p.rows_per_block));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = temp == 0.f ? launch_ind<kFwdHard, true>(p, kSigmoid, grid, s)
                  : dz_acc    ? launch_ind<kFwdBoth, true>(p, indicator, grid, s)
                              : launch_ind<kFwdSoft, true>(p, indicator, grid, s);
  return cuda_ok("pvae_rsample_fwd", e);
}

extern "C" int pvae_rsample_bwd(const float* log_rate, const float* g_z, const float* dz_acc, float* g_log_rate, int64_t B, int64_t K,
                                int64_t row_offset, int n_exp, float temp, int indicator, float clamp, float exit_logit,
                                uint64_t seed, uint64_t offset, const uint64_t* offset_dev, void* stream) {
  if (int rc = check_common("pvae_rsample_bwd", B, K, row_offset, n_exp, temp, indicator)) return rc;
  if (temp == 0.f) return fail(PVAE_ERR_INVALID, "pvae_rsample_bwd: hard counts (temp == 0) have no gradient");
  if (B == 0) return PVAE_OK;
  if (!log_rate || !g_z || !g_log_rate) return fail(PVAE_ERR_INVALID, "pvae_rsample_bwd: null pointer");
  SamplerParams p{};
  p.h = log_rate, p.g_z = g_z, p.g_h = g_log_rate, p.dz_in = dz_acc;
  p.B = B, p.K = K, p.row_offset = row_offset, p.rows_per_block = p.rows_block = rows_per_block_for(B);
  set_chain(p, n_exp, temp, indicator, exit_logit);
  p.clamp_rate = clamp;
  p.offset = offset, p.offset_dev = offset_dev;
  philox_key_schedule(seed, p.ks);
  dim3 grid(static_cast<unsigned>((B + p.rows_per_block - 1) / p.rows_per_block));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dz_acc)
    return cuda_ok("pvae_rsample_bwd", launch_pdl(sampler_bwd_saved_kernel<true>,
                                                  dim3(static_cast<unsigned>((B + kSavedRows - 1) / kSavedRows)),
                                                  dim3(kThreads), 0, s, p));
  return cuda_ok("pvae_rsample_bwd", launch_ind<kBwd, true>(p, indicator, grid, s));
}

extern "C" int pvae_kl_fwd(const float* log_dr, const float* log_r, float* kl, int64_t B, int64_t K, void* stream) {
  if (B < 0 || K <= 0 || K % 4 != 0) return fail(PVAE_ERR_INVALID, "pvae_kl_fwd: bad shape B=%lld K=%lld", (long long)B, (long long)K);
  if (B == 0) return PVAE_OK;
  if (!log_dr || !log_r || !kl) return fail(PVAE_ERR_INVALID, "pvae_kl_fwd: null pointer");
  const long long n4 = B * K / 4;
  const unsigned grid = static_cast<unsigned>(std::min<long long>((n4 + 255) / 256, 148ll * 16));
  return cuda_ok("pvae_kl_fwd", launch_pdl(kl_fwd_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), log_dr,
                                           log_r, kl, n4, static_cast<long long>(K)));
}

extern "C" int64_t pvae_kl_diag_ws_floats(int64_t B, int64_t K) { return B <= 0 || K <= 0 ? 0 : ((B + kDiagRows - 1) / kDiagRows) * K; }

extern "C" int pvae_kl_diag(const float* h, const float* log_r, const float* b_enc, float* kl_diag, float* ws, int64_t B, int64_t K,
                            int exc_only, void* stream) {
  if (B <= 0 || K <= 0 || K % 4 != 0) return fail(PVAE_ERR_INVALID, "pvae_kl_diag: bad shape B=%lld K=%lld", (long long)B, (long long)K);
  if (!h || !log_r || !kl_diag || !ws) return fail(PVAE_ERR_INVALID, "pvae_kl_diag: null pointer");
  SamplerParams p{};
  p.h = h, p.log_r = log_r, p.b_enc = b_enc, p.B = B, p.K = K;
  p.clamp_dr = 10.f, p.clamp_rate = 5.f, p.exc_only = exc_only;  // main/vae.py:403-409
  const long long P = (B + kDiagRows - 1) / kDiagRows;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = cuda_ok("pvae_kl_diag", launch_pdl(kl_diag_kernel, dim3(static_cast<unsigned>(P)), dim3(256), 0, s, p, ws))) return rc;
  return finalize_cols(ws, kl_diag, P, K, 0, s, 1.f / static_cast<float>(B));
}

extern "C" int pvae_kl_bwd(const float* log_dr, const float* log_r, const float* g_kl, float* g_log_dr, float* g_log_r,
                           float* partial_ws, int accumulate, int64_t B, int64_t K, void* stream) {
  if (B < 0 || K <= 0 || K % 4 != 0) return fail(PVAE_ERR_INVALID, "pvae_kl_bwd: bad shape B=%lld K=%lld", (long long)B, (long long)K);
  if (B == 0) return PVAE_OK;
  if (!log_dr || !log_r || !g_kl || !g_log_dr || !g_log_r || !partial_ws) return fail(PVAE_ERR_INVALID, "pvae_kl_bwd: null pointer");
  const int rpb = rows_per_block_for(B);
  const long long P = (B + rpb - 1) / rpb;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (int rc = cuda_ok("pvae_kl_bwd", launch_pdl(kl_bwd_kernel, dim3(static_cast<unsigned>(P)), dim3(kThreads), 0, s, log_dr, log_r,
                                                 g_kl, g_log_dr, partial_ws, static_cast<long long>(B),
                                                 static_cast<long long>(K), rpb)))
    return rc;
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

extern "C" int pvae_debug_log_check(uint64_t* mismatches_dev, void* stream) {
  if (!mismatches_dev) return fail(PVAE_ERR_INVALID, "pvae_debug_log_check: null pointer");
  debug_log_check_kernel<<<148 * 4, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<unsigned long long*>(mismatches_dev));
  return cuda_ok("pvae_debug_log_check", cudaGetLastError());
}
