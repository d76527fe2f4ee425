// Error reporting / versioning / launch accounting of libprg.so.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace prg {

static thread_local char t_error[1024] = "";
std::atomic<uint64_t> g_launches{0};
bool g_pdl_enabled = getenv("PRG_NO_PDL") == nullptr;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

}  // namespace prg

extern "C" __attribute__((visibility("default"))) const char* prg_last_error(void) { return prg::t_error; }
extern "C" __attribute__((visibility("default"))) int prg_abi_version(void) { return PRG_ABI_VERSION; }
extern "C" __attribute__((visibility("default"))) uint64_t prg_launch_count(void) {
  return prg::g_launches.load(std::memory_order_relaxed);
}
