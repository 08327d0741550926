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

}  // namespace pvae
