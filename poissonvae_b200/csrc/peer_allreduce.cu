// Data-parallel gradient exchange fused with the optimiser: ONE kernel per step that
//   1. publishes "my gradients are complete" to every peer (a release store of the step token into each
//      peer's flag array, over NVLink),
//   2. waits until every peer has published (acquire loads of its OWN flag array - local memory),
//   3. reads every rank's gradient buffer straight from peer memory (P2P loads through NVSwitch), sums them in
//      rank order (so every rank forms bit-identical sums) and applies the Adamax update in the same pass.
// It replaces NCCL all-reduce -> Adamax (two launches, two passes over the buffer, ring/tree latency) for the
// 1 MB gradient of the lin|lin P-VAE (optimiser base/adamax.py:34-100).
// Gradients are double buffered by step parity, so no second barrier is needed: a rank overwrites buffer
// (step & 1) two steps later, after it has seen every peer's token for step + 1, which a peer only publishes
// after its own kernel for `step` - and with it its reads - has completed.
//
// The buffers are symmetric allocations exchanged by the caller (torch.distributed._symmetric_memory); this file
// only sees raw peer pointers.  Never run several ranks of this kernel on one GPU: they wait on one another.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pvae_b200.h"
#include "pvae_device.cuh"
#include "pvae_host.h"

namespace pvae {

constexpr int kMaxPeers = 16;

struct PeerPtrs {
  const float* grad[kMaxPeers];  // rank r's gradient buffer for this step's parity (n_ext floats)
  unsigned* flag[kMaxPeers];     // rank r's flag array (>= world words)
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {  // bypass L1: the data was written by another GPU
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// Up to three [start, start + len) ranges of the flat buffer (in float4 units) form one exchange "bucket"; `tail4` is the
// float4 index of the trailing statistics in the gradient buffers (or -1).
struct ExchangeRanges {
  long long start4[3], len4[3];
  int n;
  long long tail4;
};

constexpr long long kSpinLimit = 4000000000ll;  // ~2 s of SM clocks: a peer that never publishes is reported, not waited for

__global__ void __launch_bounds__(256) allreduce_adamax_kernel(const PeerPtrs pp, int world, int rank, unsigned token, int flag_base,
                                                               float* __restrict__ p, float* __restrict__ p_lo, float* __restrict__ m,
                                                               float* __restrict__ u, const ExchangeRanges rg,
                                                               float* __restrict__ tail_out, float* __restrict__ g_sum_out,
                                                               float clr, float beta1, float beta2, float eps, float wd) {
  pdl_launch_dependents();
  pdl_wait();  // every kernel that wrote this rank's gradients has completed
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(pp.flag[threadIdx.x] + flag_base + rank, token);
  }
  if (threadIdx.x < world) {
    const unsigned* mine = pp.flag[rank] + flag_base + threadIdx.x;
    const long long t0 = clock64();
    while (static_cast<int>(ld_acquire_sys(mine) - token) < 0) {
      if (clock64() - t0 > kSpinLimit) {  // the host reads word 63 (pvae exchange error: token of the step that timed out)
        atomicMax(pp.flag[rank] + 63, token);
        break;
      }
    }
  }
  __syncthreads();
  long long total4 = rg.tail4 >= 0 ? 1 : 0;
  for (int r = 0; r < rg.n; ++r) total4 += rg.len4[r];
  for (long long j = blockIdx.x * 256ll + threadIdx.x; j < total4; j += gridDim.x * 256ll) {
    long long i = rg.tail4, k = j;  // float4 index into the flat buffers
    bool is_tail = true;
    for (int r = 0; r < rg.n; ++r) {
      if (k < rg.len4[r]) {
        i = rg.start4[r] + k, is_tail = false;
        break;
      }
      k -= rg.len4[r];
    }
    // all peers' loads in flight at once (one NVLink round trip instead of `world` dependent ones), eight at a time,
    // then the sum in rank order - every rank forms the same bits
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r0 = 0; r0 < kMaxPeers; r0 += 8) {
      if (r0 >= world) break;
      float4 v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (r0 + q < world) v[q] = ld_peer4(pp.grad[r0 + q] + 4 * i);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (r0 + q < world) {
          if (r0 + q == 0) g = v[q];
          else g.x += v[q].x, g.y += v[q].y, g.z += v[q].z, g.w += v[q].w;
        }
    }
    if (is_tail) {  // trailing statistics ([loss, mse, kl, -]) ride along
      reinterpret_cast<float4*>(tail_out)[0] = g;
      continue;
    }
    if (g_sum_out != nullptr) reinterpret_cast<float4*>(g_sum_out)[i] = g;
    float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], uv = reinterpret_cast<float4*>(u)[i], lo;
    lo.x = adamax_apply(pv.x, mv.x, uv.x, g.x, clr, beta1, beta2, eps, wd);  // base/adamax.py:79-88
    lo.y = adamax_apply(pv.y, mv.y, uv.y, g.y, clr, beta1, beta2, eps, wd);
    lo.z = adamax_apply(pv.z, mv.z, uv.z, g.z, clr, beta1, beta2, eps, wd);
    lo.w = adamax_apply(pv.w, mv.w, uv.w, g.w, clr, beta1, beta2, eps, wd);
    reinterpret_cast<float4*>(p)[i] = pv, reinterpret_cast<float4*>(m)[i] = mv, reinterpret_cast<float4*>(u)[i] = uv;
    if (p_lo != nullptr) reinterpret_cast<float4*>(p_lo)[i] = lo;
  }
}

}  // namespace pvae

using namespace pvae;

// One bucket of the exchange: ranges (floats, multiples of 4) of the flat buffer + optionally the 4-float statistics tail
// at float offset `tail`; flags of bucket b live at words [16 b, 16 b + world) of every rank's flag array.
int pvae::peer_exchange_impl(float* p, float* p_lo, float* exp_avg, float* exp_inf, const int64_t* range_start, const int64_t* range_len,
                             int n_ranges, int64_t tail, const void* const* peer_grads, void* const* peer_flags, int bucket, int world,
                             int rank, uint32_t token, float* tail_out, float* g_sum_out, float lr, int step, float beta1, float beta2,
                             float eps, float weight_decay, void* stream) {
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(PVAE_ERR_INVALID, "pvae_allreduce_adamax: world=%d rank=%d (at most %d ranks)", world, rank, kMaxPeers);
  if (n_ranges < 1 || n_ranges > 3 || bucket < 0 || bucket > 2 || step < 1 || !p || !exp_avg || !exp_inf || !peer_grads || !peer_flags ||
      (tail >= 0 && (!tail_out || tail % 4)))
    return fail(PVAE_ERR_INVALID, "pvae_allreduce_adamax: bad argument");
  PeerPtrs pp{};
  for (int r = 0; r < world; ++r) {
    if (!peer_grads[r] || !peer_flags[r]) return fail(PVAE_ERR_INVALID, "pvae_allreduce_adamax: null peer pointer %d", r);
    pp.grad[r] = static_cast<const float*>(peer_grads[r]);
    pp.flag[r] = static_cast<unsigned*>(peer_flags[r]);
  }
  ExchangeRanges rg{};
  rg.n = n_ranges, rg.tail4 = tail >= 0 ? tail / 4 : -1;
  long long total4 = tail >= 0 ? 1 : 0;
  for (int r = 0; r < n_ranges; ++r) {
    if (range_start[r] < 0 || range_len[r] <= 0 || range_start[r] % 4 || range_len[r] % 4)
      return fail(PVAE_ERR_INVALID, "pvae_allreduce_adamax: range %d must be 4-float aligned", r);
    rg.start4[r] = range_start[r] / 4, rg.len4[r] = range_len[r] / 4;
    total4 += rg.len4[r];
  }
  const double bias_correction = 1.0 - pow(static_cast<double>(beta1), step);  // base/adamax.py:75-76
  const float clr = static_cast<float>(static_cast<double>(lr) / bias_correction);
  const unsigned grid = static_cast<unsigned>(std::min<long long>((total4 + 255) / 256, 148ll * 2));
  return cuda_ok("pvae_allreduce_adamax",
                 launch_pdl(allreduce_adamax_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), pp, world, rank,
                            static_cast<unsigned>(token), 16 * bucket, p, p_lo, exp_avg, exp_inf, rg, tail_out, g_sum_out, clr, beta1, beta2,
                            eps, weight_decay));
}

extern "C" int pvae_allreduce_adamax(float* p, float* exp_avg, float* exp_inf, int64_t n, int64_t n_ext,
                                     const void* const* peer_grads, void* const* peer_flags, int world, int rank,
                                     uint32_t token, float* tail_out, float* g_sum_out, float lr, int step, float beta1,
                                     float beta2, float eps, float weight_decay, void* stream) {
  if (n <= 0 || n % 4 || n_ext < n || n_ext % 4 || (n_ext > n && n_ext != n + 4))
    return fail(PVAE_ERR_INVALID, "pvae_allreduce_adamax: n must be a multiple of 4 and n_ext == n or n + 4");
  const int64_t start = 0, len = n;
  return peer_exchange_impl(p, nullptr, exp_avg, exp_inf, &start, &len, 1, n_ext > n ? n : -1, peer_grads, peer_flags, 0, world, rank, token,
                            tail_out, g_sum_out, lr, step, beta1, beta2, eps, weight_decay, stream);
}
