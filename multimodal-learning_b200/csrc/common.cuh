// Shared host/device helpers for the mml_b200 C-ABI library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "mml_b200.h"

namespace mml {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launch_count;
uint32_t* device_error_word();          // core.cu: sticky per-device error flags (MML_DEVERR_*), may be NULL

__device__ __forceinline__ void flag_device_error(uint32_t* word, uint32_t bit) {
  if (word != nullptr) atomicOr(word, bit);
}

inline int check_launch(const char* what) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MML_ERR_CUDA;
  }
  return MML_OK;
}

#define MML_REQUIRE(cond, code, ...)     \
  do {                                   \
    if (!(cond)) {                       \
      ::mml::set_error(__VA_ARGS__);     \
      return (code);                     \
    }                                    \
  } while (0)

#define MML_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) {                                                 \
      ::mml::set_error("%s failed: %s", #call, cudaGetErrorString(e__));      \
      return MML_ERR_CUDA;                                                    \
    }                                                                         \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

constexpr unsigned kFullMask = 0xffffffffu;

// 128-bit read-only streaming load: bank rows are read once per (anchor, column).
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  acc = fmaf(a.w, b.w, acc);
  return acc;
}

__device__ __forceinline__ void axpy4(float4& acc, float g, const float4& r) {
  acc.x = fmaf(g, r.x, acc.x);
  acc.y = fmaf(g, r.y, acc.y);
  acc.z = fmaf(g, r.z, acc.z);
  acc.w = fmaf(g, r.w, acc.w);
}

}  // namespace mml
