// Library-level C-ABI entry points (version, error string).
#include "pvae_host.h"

namespace pvae {
char* last_error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace pvae

extern "C" int pvae_version(void) { return 100; }
extern "C" const char* pvae_last_error(void) { return pvae::last_error_buffer(); }
