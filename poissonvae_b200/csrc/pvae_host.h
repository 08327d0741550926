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

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
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

}  // namespace pvae
