// Host-side error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <algorithm>

#include "../../include/pvae_b200.h"

namespace pvae {

char* last_error_buffer();  // thread-local, 512 bytes (capi.cu)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

inline int cuda_ok(const char* what, cudaError_t e) {
  if (e == cudaSuccess) return PVAE_OK;
  return fail(PVAE_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream drains;
// it calls pdl_wait() (griddepcontrol.wait) before touching anything the predecessor produced, and
// pdl_launch_dependents() early so that its own successor can do the same.  Hides launch latency and
// kernel prologues behind the previous kernel's tail.  pvae_set_pdl(0) turns the attribute off.
bool pdl_enabled();
void count_launch();  // every kernel launch of this library is counted (pvae_launch_count)

// Every kernel of the library asks for the same L1 / shared-memory split (all shared): an SM that has to re-partition its
// unified L1 between two kernels must drain first, which costs microseconds per kernel boundary and keeps a dependent
// kernel from becoming resident under its predecessor's tail.  The small kernels here stream their data once and do not
// miss the L1.  pvae_set_carveout(0) leaves the driver's per-kernel default (A/B).
bool carveout_enabled();
void prefer_shared_carveout(const void* kernel);  // once per kernel (capi.cu)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  if (carveout_enabled()) prefer_shared_carveout(reinterpret_cast<const void*>(kernel));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  count_launch();
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// sampler.cu: pvae_sampler_fwd / pvae_sampler_bwd with the optional TF32 low planes of z / g_h (what the GEMMs that
// consume them load by TMA instead of splitting the operand in shared memory)
int sampler_fwd_impl(const float* h, const float* log_r, const float* b_enc, float* z, float* z_lo, float* log_dr, float* rate, float* kl,
                     float* kl_rowsum, float* rate_max, float* dz_acc, int64_t B, int64_t K, int64_t row_offset, int n_exp, float temp,
                     int indicator, int exc_only, float exit_logit, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                     void* stream);
int sampler_bwd_impl(const float* h, const float* log_r, const float* b_enc, const float* g_z, const float* g_log_dr, const float* dz_acc,
                     float kl_scale, float* g_h, float* g_h_lo, float* g_log_r, float* g_b_enc, float* partial_ws, int accumulate,
                     int64_t B, int64_t K, int64_t row_offset, int n_exp, float temp, int indicator, int exc_only, float exit_logit,
                     uint64_t seed, uint64_t offset, const uint64_t* offset_dev, void* stream);
int sampler_fwd_coef_impl(const float* h, const float* log_r, const float* b_enc, float* z, float* z_lo, float* kl_rowsum, float* rate_max,
                          float* c_a, float* c_ddr, float* c_ck, float* klcol, int64_t B, int64_t K, int64_t row_offset, int n_exp, float temp,
                          int indicator, int exc_only, float exit_logit, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                          void* stream);
int sampler_bwd_coef_impl(const float* g_z, const float* c_a, const float* c_ddr, const float* c_ck, float kl_scale, float* g_h,
                          float* g_h_lo, float* partial_ws, int want_bias, const float* klcol, int64_t klcol_rows, float* g_log_r,
                          float* g_b_enc, int64_t B, int64_t K, void* stream);
int64_t sampler_bwd_coef_rows(int64_t B);  // rows of column partials sampler_bwd_coef_impl writes
// fused_fwd.cu: sampler + KL + decoder GEMM + reconstruction residual in one kernel (see pvae_fused_fwd)
int fused_fwd_supported(int64_t B, int64_t K, int64_t D);
int fused_fwd_impl(const float* h, const float* log_r, const float* b_enc, const float* w_dec, const float* w_dec_lo, const float* x, float* z,
                   float* z_lo, float* dz_acc, float* g_y, float* g_y_lo, float* recon, float* kl_rowsum, float* rate_max, int64_t B, int64_t K,
                   int64_t D, int64_t row_offset, int n_exp, float temp, int indicator, int exc_only, float exit_logit, float g_scale,
                   uint64_t seed, uint64_t offset, const uint64_t* offset_dev, void* stream);
// step.cu
int recon_parts_impl(const float* y_parts, int n_parts, const float* x, float* g_y, float* g_y_lo, float* recon, int64_t B, int64_t D,
                     float g_scale, void* stream);
struct AdamaxDev {  // step-dependent scalars on the device (CUDA-graph replay): state[0] = steps taken so far
  const float* lr_sched = nullptr;
  int n_sched = 0;
  const int* state = nullptr;
};
int adamax_from_partials_impl(float* p, float* p_lo, float* g_out, float* exp_avg, float* exp_inf, const pvae_grad_segment* segs,
                              int n_segs, float lr, int step, float beta1, float beta2, float eps, float weight_decay, const AdamaxDev& dev,
                              void* stream);
int adamax_step_impl(float* p, float* p_lo, const float* g, float* exp_avg, float* exp_inf, int64_t n, float lr, int step, float beta1,
                     float beta2, float eps, float weight_decay, const float* clip_coef, void* stream);
int adamax_step_dev_impl(float* p, float* p_lo, const float* g, float* exp_avg, float* exp_inf, int64_t n, const AdamaxDev& dev, float beta1,
                         float beta2, float eps, float weight_decay, const float* clip_coef, void* stream);
int advance_state(int* state, uint64_t philox_increment, float* rmax_step, float* rmax_ring, int ring_n, void* stream);

// peer_allreduce.cu: one bucket of the peer-memory gradient exchange fused with Adamax (see pvae_allreduce_adamax)
int peer_exchange_impl(float* p, float* p_lo, float* exp_avg, float* exp_inf, const int64_t* range_start, const int64_t* range_len,
                       int n_ranges, int64_t tail, const void* const* peer_grads, void* const* peer_flags, int bucket, int world, int rank,
                       uint32_t token, float* tail_out, float* g_sum_out, float lr, int step, float beta1, float beta2, float eps,
                       float weight_decay, void* stream);

// gemm_tc.cu: the GEMM behind pvae_gemm_tf32 / _planes / _residual with every option (library-internal callers)
int gemm_impl(const float* A, const float* A_lo, const float* B, const float* B_lo, float* C, int64_t M, int64_t N, int64_t Kd, int a_kmajor,
              int b_kmajor, int precise, int splits, float* splitk_ws, const float* X, float* G, float* G_lo, float* row_part,
              float g_scale, int force_bn, void* stream);

}  // namespace pvae
