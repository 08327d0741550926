// Library-level C-ABI entry points (version, error string).
#include <atomic>
#include <mutex>

#include "pvae_host.h"

namespace pvae {
char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
static bool g_pdl = true;
bool pdl_enabled() { return g_pdl; }
void set_pdl(bool on) { g_pdl = on; }
static bool g_carveout = true;
bool carveout_enabled() { return g_carveout; }
void prefer_shared_carveout(const void* kernel) {
  static const void* seen[512];
  static std::atomic<int> n_seen{0};
  const int n = n_seen.load(std::memory_order_acquire);
  for (int i = 0; i < n; ++i)
    if (seen[i] == kernel) return;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  const int m = n_seen.load(std::memory_order_acquire);
  for (int i = 0; i < m; ++i)
    if (seen[i] == kernel) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaGetLastError();  // a hint: failure is not an error
  if (m < 512) {
    seen[m] = kernel;
    n_seen.store(m + 1, std::memory_order_release);
  }
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace pvae

namespace pvae { void set_pdl(bool); }
extern "C" int pvae_set_pdl(int on) {
  pvae::set_pdl(on != 0);
  return PVAE_OK;
}
extern "C" int pvae_set_carveout(int on) {
  pvae::g_carveout = on != 0;
  return PVAE_OK;
}
extern "C" int pvae_version(void) { return 100; }
extern "C" const char* pvae_last_error(void) { return pvae::last_error_buffer(); }
extern "C" int64_t pvae_launch_count(void) { return pvae::g_launches.load(std::memory_order_relaxed); }
