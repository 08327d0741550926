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

__global__ void __launch_bounds__(256) allreduce_adamax_kernel(const PeerPtrs pp, int world, int rank, unsigned token,
                                                               float* __restrict__ p, float* __restrict__ m,
                                                               float* __restrict__ u, long long n4, long long n4_ext,
                                                               float* __restrict__ tail_out, float* __restrict__ g_sum_out,
                                                               float clr, float beta1, float beta2, float eps, float wd) {
  pdl_launch_dependents();
  pdl_wait();  // every kernel that wrote this rank's gradients has completed
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(pp.flag[threadIdx.x] + rank, token);
  }
  if (threadIdx.x < world) {
    const unsigned* mine = pp.flag[rank] + threadIdx.x;
    while (static_cast<int>(ld_acquire_sys(mine) - token) < 0) {
    }
  }
  __syncthreads();
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4_ext; i += gridDim.x * 256ll) {
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
    if (i >= n4) {  // trailing statistics ([loss, mse, kl, -]) ride along
      reinterpret_cast<float4*>(tail_out)[i - n4] = g;
      continue;
    }
    if (g_sum_out != nullptr) reinterpret_cast<float4*>(g_sum_out)[i] = g;
    float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], uv = reinterpret_cast<float4*>(u)[i];
    float* gp = &g.x;
    float* pp4 = &pv.x;
    float* mp = &mv.x;
    float* up = &uv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // base/adamax.py:79-88
      float gi = gp[j];
      if (wd != 0.f) gi = fmaf(wd, pp4[j], gi);
      mp[j] = mp[j] * beta1 + (1.f - beta1) * gi;
      up[j] = fmaxf(up[j] * beta2, fabsf(gi) + eps);
      pp4[j] = pp4[j] - clr * (mp[j] / up[j]);
    }
    reinterpret_cast<float4*>(p)[i] = pv, reinterpret_cast<float4*>(m)[i] = mv, reinterpret_cast<float4*>(u)[i] = uv;
  }
}

}  // namespace pvae

using namespace pvae;

extern "C" int pvae_allreduce_adamax(float* p, float* exp_avg, float* exp_inf, int64_t n, int64_t n_ext,
                                     const void* const* peer_grads, void* const* peer_flags, int world, int rank,
                                     uint32_t token, float* tail_out, float* g_sum_out, float lr, int step, float beta1,
                                     float beta2, float eps, float weight_decay, void* stream) {
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
    return fail(PVAE_ERR_INVALID, "pvae_allreduce_adamax: world=%d rank=%d (at most %d ranks)", world, rank, kMaxPeers);
  if (n <= 0 || n % 4 || n_ext < n || n_ext % 4 || step < 1 || !p || !exp_avg || !exp_inf || !peer_grads || !peer_flags ||
      (n_ext > n && !tail_out))
    return fail(PVAE_ERR_INVALID, "pvae_allreduce_adamax: bad argument");
  PeerPtrs pp{};
  for (int r = 0; r < world; ++r) {
    if (!peer_grads[r] || !peer_flags[r]) return fail(PVAE_ERR_INVALID, "pvae_allreduce_adamax: null peer pointer %d", r);
    pp.grad[r] = static_cast<const float*>(peer_grads[r]);
    pp.flag[r] = static_cast<unsigned*>(peer_flags[r]);
  }
  const double bias_correction = 1.0 - pow(static_cast<double>(beta1), step);  // base/adamax.py:75-76
  const float clr = static_cast<float>(static_cast<double>(lr) / bias_correction);
  const long long n4 = n / 4, n4_ext = n_ext / 4;
  const unsigned grid = static_cast<unsigned>(std::min<long long>((n4_ext + 255) / 256, 148ll * 2));
  return cuda_ok("pvae_allreduce_adamax",
                 launch_pdl(allreduce_adamax_kernel, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), pp, world, rank,
                            static_cast<unsigned>(token), p, exp_avg, exp_inf, n4, n4_ext, tail_out, g_sum_out, clr, beta1,
                            beta2, eps, weight_decay));
}
