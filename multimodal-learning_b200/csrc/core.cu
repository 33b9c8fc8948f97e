// Library-wide state of the mml_b200 C-ABI: error string, ABI version, launch counter.
#include <stdarg.h>

#include <mutex>

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

// One 32-bit word of sticky error flags per device (MML_DEVERR_*): kernels that index memory with caller-supplied ids OR a
// bit in and clamp / drop the offending id instead of reading or writing out of bounds (the reference's index_select
// raises a device-side assert there).  Allocated on first use -- always before any CUDA-graph capture, whose warm-up runs
// the same kernels eagerly -- and read back with mml_device_error_flags.
uint32_t* device_error_word() {
  static std::mutex mu;
  static uint32_t* words[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (words[dev] == nullptr) {
    uint32_t* w = nullptr;
    if (cudaMalloc(&w, sizeof(uint32_t)) != cudaSuccess) return nullptr;
    if (cudaMemset(w, 0, sizeof(uint32_t)) != cudaSuccess) return nullptr;
    words[dev] = w;
  }
  return words[dev];
}

}  // namespace mml

extern "C" int mml_abi_version(void) { return MML_ABI_VERSION; }
extern "C" const char* mml_last_error(void) { return mml::get_error(); }
extern "C" int64_t mml_launch_count(void) { return mml::g_launch_count.load(std::memory_order_relaxed); }

extern "C" int mml_device_error_flags(uint32_t* flags_host, int32_t reset) {
  MML_REQUIRE(flags_host != nullptr, MML_ERR_INVALID_ARG, "device_error_flags: null output");
  uint32_t* w = mml::device_error_word();
  MML_REQUIRE(w != nullptr, MML_ERR_CUDA, "device_error_flags: no error word on this device");
  MML_CUDA(cudaMemcpy(flags_host, w, sizeof(uint32_t), cudaMemcpyDeviceToHost));      // synchronises with the device
  if (reset) MML_CUDA(cudaMemset(w, 0, sizeof(uint32_t)));
  return MML_OK;
}
