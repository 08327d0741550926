// Library-level C-ABI entry points (version, error string).
#include <atomic>

#include "pvae_host.h"

namespace pvae {
char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
static bool g_pdl = true;
bool pdl_enabled() { return g_pdl; }
void set_pdl(bool on) { g_pdl = on; }
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace pvae

namespace pvae { void set_pdl(bool); }
extern "C" int pvae_set_pdl(int on) {
  pvae::set_pdl(on != 0);
  return PVAE_OK;
}
extern "C" int pvae_version(void) { return 100; }
extern "C" const char* pvae_last_error(void) { return pvae::last_error_buffer(); }
extern "C" int64_t pvae_launch_count(void) { return pvae::g_launches.load(std::memory_order_relaxed); }
