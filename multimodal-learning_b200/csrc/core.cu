// Library-wide state of the mml_b200 C-ABI: error string, ABI version, launch counter.
#include <stdarg.h>

#include "common.cuh"

namespace mml {

std::atomic<int64_t> g_launch_count{0};

namespace {
thread_local char t_error[512] = "";
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

const char* get_error() { return t_error; }

}  // namespace mml

extern "C" int mml_abi_version(void) { return MML_ABI_VERSION; }
extern "C" const char* mml_last_error(void) { return mml::get_error(); }
extern "C" int64_t mml_launch_count(void) { return mml::g_launch_count.load(std::memory_order_relaxed); }
