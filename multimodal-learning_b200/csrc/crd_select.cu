// Selection variant of the contrast memory (reference: MICCAI-2022/CL_utils/memory_new.py:225-397, ContrastMemory_v3).
//
// Before it scores anything, the variant ranks the K+P sampled columns of every anchor by the gap between two cosine
// "relations" (memory_new.py:288-292):
//     t_relation[b,k] = cos(memory_v1[idx[b,k]], v1[b])        s_relation[b,k] = cos(memory_v2[idx[b,k]], v2[b])
// and keeps P2 positives / K2 negatives by the order of  diff = t_relation - s_relation  (:298-357).  The reference
// materialises both gathered [B, K+P, D] tensors, their normalised copies and two bmm outputs; here ONE pass over the
// rows (same 8-lanes-per-row, 128-bit streaming loads as K4 in crd_gather.cu) emits only diff[B, K+P].
// HBM-bound: algorithmic bytes = 2*B*(K+P)*D*4 (rows) + B*(K+P)*(8 + 4) (indices in, diff out).
#include <math.h>

#include "common.cuh"

namespace mml {
namespace {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;

struct RelArgs {
  const float* bank1;
  const float* bank2;
  const float* v1;
  const float* v2;
  const int64_t* idx64;
  const int32_t* idx32;
  float* diff;
  int64_t cols;
  int64_t n_rows;           // ids outside [0, n_rows) are flagged (MML_DEVERR_CRD_INDEX) and read row 0
  uint32_t* err;
  int32_t D;
  int32_t chunk_cols;
};

template <int VPL, int U>
__global__ void __launch_bounds__(kThreads) crd_relation_kernel(const RelArgs a) {
  constexpr int D = 32 * VPL;
  constexpr int ROWS_PER_IT = 4 * U;
  const int b = blockIdx.y;
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * a.chunk_cols;
  const int64_t c1 = min(c0 + static_cast<int64_t>(a.chunk_cols), a.cols);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = lane & 7, q = lane >> 3;

  float4 fv1[VPL], fv2[VPL];
  float n1 = 0.f, n2 = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    fv1[j] = __ldg(reinterpret_cast<const float4*>(a.v1 + static_cast<int64_t>(b) * D) + j * 8 + s);
    fv2[j] = __ldg(reinterpret_cast<const float4*>(a.v2 + static_cast<int64_t>(b) * D) + j * 8 + s);
    n1 = dot4(fv1[j], fv1[j], n1);
    n2 = dot4(fv2[j], fv2[j], n2);
  }
#pragma unroll
  for (int off = 1; off < 8; off <<= 1) {
    n1 += __shfl_xor_sync(kFullMask, n1, off);
    n2 += __shfl_xor_sync(kFullMask, n2, off);
  }
  const float nv1 = sqrtf(n1), nv2 = sqrtf(n2);          // torch.norm(v, dim=1), :289,292

  const int64_t base = static_cast<int64_t>(b) * a.cols;
  for (int64_t cb = c0 + warp * 32; cb < c1; cb += kWarps * 32) {
    const int64_t mycol = cb + lane;
    int32_t myrow = 0;
    if (mycol < c1) {
      const int64_t r64 = a.idx32 ? static_cast<int64_t>(a.idx32[base + mycol]) : a.idx64[base + mycol];
      if (static_cast<uint64_t>(r64) >= static_cast<uint64_t>(a.n_rows)) flag_device_error(a.err, MML_DEVERR_CRD_INDEX);
      else myrow = static_cast<int32_t>(r64);
    }
    const int n_it = (static_cast<int>(min(static_cast<int64_t>(32), c1 - cb)) + ROWS_PER_IT - 1) / ROWS_PER_IT;
    for (int it = 0; it < n_it; ++it) {
      float4 r1[U][VPL], r2[U][VPL];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int32_t row = __shfl_sync(kFullMask, myrow, it * ROWS_PER_IT + u * 4 + q);
        const float* p1 = a.bank1 + static_cast<int64_t>(row) * D + s * 4;
        const float* p2 = a.bank2 + static_cast<int64_t>(row) * D + s * 4;
#pragma unroll
        for (int j = 0; j < VPL; ++j) r1[u][j] = ldg_stream_f4(p1 + j * 32);
#pragma unroll
        for (int j = 0; j < VPL; ++j) r2[u][j] = ldg_stream_f4(p2 + j * 32);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float t_dot = 0.f, t_sq = 0.f, s_dot = 0.f, s_sq = 0.f;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          t_dot = dot4(r1[u][j], fv1[j], t_dot);       // bank-1 row . v1   (:288)
          t_sq = dot4(r1[u][j], r1[u][j], t_sq);
          s_dot = dot4(r2[u][j], fv2[j], s_dot);       // bank-2 row . v2   (:291)
          s_sq = dot4(r2[u][j], r2[u][j], s_sq);
        }
#pragma unroll
        for (int off = 1; off < 8; off <<= 1) {
          t_dot += __shfl_xor_sync(kFullMask, t_dot, off);
          t_sq += __shfl_xor_sync(kFullMask, t_sq, off);
          s_dot += __shfl_xor_sync(kFullMask, s_dot, off);
          s_sq += __shfl_xor_sync(kFullMask, s_sq, off);
        }
        const int64_t col = cb + it * ROWS_PER_IT + u * 4 + q;
        if (s == 0 && col < c1) a.diff[base + col] = t_dot / (sqrtf(t_sq) * nv1) - s_dot / (sqrtf(s_sq) * nv2);
      }
    }
  }
}

// Any-D path: one warp per column.
__global__ void __launch_bounds__(kThreads) crd_relation_generic_kernel(const RelArgs a) {
  const int D = a.D;
  const int b = blockIdx.y;
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * a.chunk_cols;
  const int64_t c1 = min(c0 + static_cast<int64_t>(a.chunk_cols), a.cols);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* v1 = a.v1 + static_cast<int64_t>(b) * D;
  const float* v2 = a.v2 + static_cast<int64_t>(b) * D;
  float n1 = 0.f, n2 = 0.f;
  for (int i = lane; i < D; i += 32) {
    n1 = fmaf(v1[i], v1[i], n1);
    n2 = fmaf(v2[i], v2[i], n2);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    n1 += __shfl_xor_sync(kFullMask, n1, off);
    n2 += __shfl_xor_sync(kFullMask, n2, off);
  }
  const float nv1 = sqrtf(n1), nv2 = sqrtf(n2);
  const int64_t base = static_cast<int64_t>(b) * a.cols;
  for (int64_t col = c0 + warp; col < c1; col += kWarps) {
    int64_t row = a.idx32 ? static_cast<int64_t>(a.idx32[base + col]) : a.idx64[base + col];
    if (static_cast<uint64_t>(row) >= static_cast<uint64_t>(a.n_rows)) {
      if (lane == 0) flag_device_error(a.err, MML_DEVERR_CRD_INDEX);
      row = 0;
    }
    const float* p1 = a.bank1 + row * D;
    const float* p2 = a.bank2 + row * D;
    float t_dot = 0.f, t_sq = 0.f, s_dot = 0.f, s_sq = 0.f;
    for (int i = lane; i < D; i += 32) {
      const float x1 = __ldg(p1 + i), x2 = __ldg(p2 + i);
      t_dot = fmaf(x1, v1[i], t_dot);
      t_sq = fmaf(x1, x1, t_sq);
      s_dot = fmaf(x2, v2[i], s_dot);
      s_sq = fmaf(x2, x2, s_sq);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      t_dot += __shfl_xor_sync(kFullMask, t_dot, off);
      t_sq += __shfl_xor_sync(kFullMask, t_sq, off);
      s_dot += __shfl_xor_sync(kFullMask, s_dot, off);
      s_sq += __shfl_xor_sync(kFullMask, s_sq, off);
    }
    if (lane == 0) a.diff[base + col] = t_dot / (sqrtf(t_sq) * nv1) - s_dot / (sqrtf(s_sq) * nv2);
  }
}

}  // namespace
}  // namespace mml

using namespace mml;

extern "C" int mml_crd_relation_diff(const float* bank1, const float* bank2, int64_t n_rows, int32_t D, const float* v1,
                                     const float* v2, const void* idx, int32_t idx_bytes, int64_t B, int64_t cols,
                                     float* diff, void* stream) {
  MML_REQUIRE(bank1 && bank2 && v1 && v2 && idx && diff, MML_ERR_INVALID_ARG, "crd_relation_diff: null pointer argument");
  MML_REQUIRE(idx_bytes == 8 || idx_bytes == 4, MML_ERR_INVALID_ARG, "crd_relation_diff: idx_bytes must be 8 or 4");
  MML_REQUIRE(B >= 0 && cols >= 1 && n_rows >= 1 && n_rows < (1LL << 31), MML_ERR_INVALID_ARG, "crd_relation_diff: bad sizes");
  MML_REQUIRE(D >= 1 && D <= 2048, MML_ERR_UNSUPPORTED, "crd_relation_diff: feature dim %d outside [1, 2048]", D);
  MML_REQUIRE(B <= 65535, MML_ERR_UNSUPPORTED, "crd_relation_diff: batch %lld > 65535 anchors per call", (long long)B);
  if (B == 0) return MML_OK;
  const bool fast = (D == 32 || D == 64 || D == 128 || D == 256);
  if (fast)
    MML_REQUIRE(aligned16(bank1) && aligned16(bank2) && aligned16(v1) && aligned16(v2), MML_ERR_INVALID_ARG,
                "crd_relation_diff: banks and v1/v2 must be 16-byte aligned");
  RelArgs a{};
  a.bank1 = bank1; a.bank2 = bank2; a.v1 = v1; a.v2 = v2; a.diff = diff; a.cols = cols; a.D = D;
  a.n_rows = n_rows; a.err = device_error_word();
  a.idx64 = idx_bytes == 8 ? static_cast<const int64_t*>(idx) : nullptr;
  a.idx32 = idx_bytes == 4 ? static_cast<const int32_t*>(idx) : nullptr;
  // >= ~8 waves of 148 SMs x 4 CTAs when the problem allows; chunks of whole 128-column CTA passes
  int64_t chunks = (148LL * 4 * 8 + B - 1) / B;
  const int64_t max_chunks = (cols + 127) / 128;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int64_t cc = ((cols + chunks - 1) / chunks + 127) / 128 * 128;
  chunks = (cols + cc - 1) / cc;
  a.chunk_cols = static_cast<int32_t>(cc);
  const dim3 grid(static_cast<unsigned>(chunks), static_cast<unsigned>(B));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (fast ? D : 0) {
    case 32: crd_relation_kernel<1, 4><<<grid, kThreads, 0, st>>>(a); break;
    case 64: crd_relation_kernel<2, 4><<<grid, kThreads, 0, st>>>(a); break;
    case 128: crd_relation_kernel<4, 2><<<grid, kThreads, 0, st>>>(a); break;
    case 256: crd_relation_kernel<8, 1><<<grid, kThreads, 0, st>>>(a); break;
    default: crd_relation_generic_kernel<<<grid, kThreads, 0, st>>>(a);
  }
  return check_launch("crd_relation_kernel");
}

// =====================================================================================================
// Per-anchor ordering of the relation gaps (memory_new.py:303 `torch.sort(diff_pos, descending=True)`, :342
// `torch.sort(diff_neg)` + `[:, :K2]`): one CTA per anchor sorts n <= 16384 columns of diff[b, col0 .. col0+n) in shared
// memory -- a bitonic network over 64-bit composites (order-preserving image of the float in the high word, column number in
// the low word) -- and writes the first m column numbers.  The composite makes the order TOTAL (equal gaps are ordered by
// column), so the result is deterministic; the library sort the reference calls may order exact ties either way, which is
// why parity is defined on the selected ROWS (equal gaps come from the same bank row sampled twice).
// =====================================================================================================
namespace mml {
namespace {

constexpr int kSortThreads = 1024;

__global__ void __launch_bounds__(kSortThreads) crd_sort_columns_kernel(const float* __restrict__ diff, int64_t ld, int64_t col0,
                                                                        int32_t n, int32_t npow2, int32_t descending, int32_t m,
                                                                        int64_t label0, int64_t* __restrict__ out, int64_t out_ld) {
  extern __shared__ unsigned long long sm_keys[];
  const int64_t b = blockIdx.x;
  const float* src = diff + b * ld + col0;
  for (int i = threadIdx.x; i < npow2; i += kSortThreads) {
    unsigned long long key = ~0ull;                                   // padding sorts last
    if (i < n) {
      uint32_t u = __float_as_uint(src[i]);
      u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;                     // unsigned order == float order (ascending)
      if (descending) u = ~u;
      key = (static_cast<unsigned long long>(u) << 32) | static_cast<uint32_t>(i);
    }
    sm_keys[i] = key;
  }
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (npow2 >> 1); t += kSortThreads) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));         // index with bit j clear
        const int hi = lo | j;
        const bool up = (lo & k) == 0;
        const unsigned long long x = sm_keys[lo], y = sm_keys[hi];
        if ((x > y) == up) {
          sm_keys[lo] = y;
          sm_keys[hi] = x;
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < m; i += kSortThreads)
    out[b * out_ld + i] = label0 + static_cast<int64_t>(static_cast<uint32_t>(sm_keys[i]));
}

// ---- register-blocked variant for 2^LOGN in [2048, 16384] columns ----
// 16 keys per thread.  In "layout s" a thread owns the 16 indices that differ in bits [s, s+3]; the compare-exchange
// stages on those four bits run in registers, and only a change of layout is a trip through shared memory
// (sum over phases of ceil(phase / 4) ~ 30 trips for 16384 keys instead of 105 stage-by-stage passes).  Shared addresses are
// padded by one key per 16 (idx + idx / 16): every layout's 64-bit accesses are bank-conflict free, and because the index
// fields of thread and register are disjoint bits, a key's address is (thread base) + (compile-time offset).
__device__ __forceinline__ constexpr int sort_layout(int phase, int bit) {      // s of the layout that stage (phase, bit) runs in
  const int top = phase - 1 - ((phase - 1 - bit) / 4) * 4;                      // top bit of this stage's group of four
  return top - 3 > 0 ? top - 3 : 0;
}
__device__ __forceinline__ constexpr int sort_index(int s, int t, int r) {
  return ((t >> s) << (s + 4)) | (r << s) | (t & ((1 << s) - 1));
}
__device__ __forceinline__ constexpr int sort_phys(int idx) { return idx + (idx >> 4); }

template <int LOGN>
__global__ void __launch_bounds__(1 << (LOGN - 4)) crd_sort_columns_blocked_kernel(
    const float* __restrict__ diff, int64_t ld, int64_t col0, int32_t n, int32_t descending, int32_t m, int64_t label0,
    int64_t* __restrict__ out, int64_t out_ld) {
  constexpr int N = 1 << LOGN, T = N >> 4;
  extern __shared__ unsigned long long sm_keys[];
  const int64_t b = blockIdx.x;
  const int t = threadIdx.x;
  const float* src = diff + b * ld + col0;
  for (int i = t; i < N; i += T) {
    unsigned long long key = ~0ull;                                   // padding sorts last
    if (i < n) {
      uint32_t u = __float_as_uint(src[i]);
      u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;
      if (descending) u = ~u;
      key = (static_cast<unsigned long long>(u) << 32) | static_cast<uint32_t>(i);
    }
    sm_keys[sort_phys(i)] = key;
  }
  __syncthreads();
  unsigned long long v[16];
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = sm_keys[sort_phys(sort_index(0, t, 0)) + sort_phys(sort_index(0, 0, r))];
#pragma unroll
  for (int phase = 1; phase <= LOGN; ++phase) {
#pragma unroll
    for (int bit = phase - 1; bit >= 0; --bit) {
      const int s = sort_layout(phase, bit);
      const int s_prev = bit == phase - 1 ? 0 : sort_layout(phase, bit + 1);
      if (s != s_prev) {                                              // compile-time after unrolling
        __syncthreads();                                              // everyone has read the previous layout
#pragma unroll
        for (int r = 0; r < 16; ++r) sm_keys[sort_phys(sort_index(s_prev, t, 0)) + sort_phys(sort_index(s_prev, 0, r))] = v[r];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = sm_keys[sort_phys(sort_index(s, t, 0)) + sort_phys(sort_index(s, 0, r))];
      }
      const int lb = bit - s;
      // direction of the pair: bit `phase` of its index -- one of the register bits, or a thread bit, or (last phase) 0
      const bool dir_in_regs = phase >= s && phase <= s + 3;
      const bool up_t = phase >= LOGN ? true : ((t >> (phase - 4 > 0 ? phase - 4 : 0)) & 1) == 0;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        if (r & (1 << lb)) continue;
        const bool up = dir_in_regs ? ((r >> (phase - s)) & 1) == 0 : up_t;
        const unsigned long long x = v[r], y = v[r | (1 << lb)];
        const bool sw = (x > y) == up;
        v[r] = sw ? y : x;
        v[r | (1 << lb)] = sw ? x : y;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 16; ++r) sm_keys[sort_phys(sort_index(0, t, 0)) + sort_phys(sort_index(0, 0, r))] = v[r];
  __syncthreads();
  for (int i = t; i < m; i += T) out[b * out_ld + i] = label0 + static_cast<int64_t>(static_cast<uint32_t>(sm_keys[sort_phys(i)]));
}

template <int LOGN>
int launch_sort_blocked(const float* diff, int64_t B, int64_t ld, int64_t col0, int32_t n, int32_t descending, int32_t m,
                        int64_t label0, int64_t* out, int64_t out_ld, cudaStream_t stream) {
  constexpr size_t smem = ((static_cast<size_t>(1) << LOGN) + (static_cast<size_t>(1) << (LOGN - 4))) * sizeof(unsigned long long);
  if (smem > 48 * 1024)
    MML_CUDA(cudaFuncSetAttribute(crd_sort_columns_blocked_kernel<LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem)));
  crd_sort_columns_blocked_kernel<LOGN><<<static_cast<unsigned>(B), 1 << (LOGN - 4), smem, stream>>>(diff, ld, col0, n, descending,
                                                                                                   m, label0, out, out_ld);
  return check_launch("crd_sort_columns_blocked_kernel");
}

}  // namespace
}  // namespace mml

extern "C" int32_t mml_crd_sort_columns_max(void) { return 16384; }

extern "C" int mml_crd_sort_columns(const float* diff, int64_t B, int64_t ld, int64_t col0, int32_t n, int32_t descending,
                                    int32_t m, int64_t label0, int64_t* out, int64_t out_ld, void* stream) {
  MML_REQUIRE(diff && out, MML_ERR_INVALID_ARG, "crd_sort_columns: null pointer");
  MML_REQUIRE(B >= 0 && n >= 1 && m >= 0 && m <= n && col0 >= 0 && col0 + n <= ld && out_ld >= m, MML_ERR_INVALID_ARG,
              "crd_sort_columns: bad sizes");
  MML_REQUIRE(n <= mml_crd_sort_columns_max(), MML_ERR_UNSUPPORTED, "crd_sort_columns: at most %d columns per anchor (got %d)",
              mml_crd_sort_columns_max(), n);
  if (B == 0 || m == 0) return MML_OK;
  int32_t npow2 = 2, logn = 1;
  while (npow2 < n) { npow2 <<= 1; ++logn; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (logn) {
    case 11: return mml::launch_sort_blocked<11>(diff, B, ld, col0, n, descending, m, label0, out, out_ld, st);
    case 12: return mml::launch_sort_blocked<12>(diff, B, ld, col0, n, descending, m, label0, out, out_ld, st);
    case 13: return mml::launch_sort_blocked<13>(diff, B, ld, col0, n, descending, m, label0, out, out_ld, st);
    case 14: return mml::launch_sort_blocked<14>(diff, B, ld, col0, n, descending, m, label0, out, out_ld, st);
    default: break;
  }
  const size_t smem = static_cast<size_t>(npow2) * sizeof(unsigned long long);
  if (smem > 48 * 1024)
    MML_CUDA(cudaFuncSetAttribute(mml::crd_sort_columns_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  mml::crd_sort_columns_kernel<<<static_cast<unsigned>(B), mml::kSortThreads, smem, st>>>(
      diff, ld, col0, n, npow2, descending, m, label0, out, out_ld);
  return mml::check_launch("crd_sort_columns_kernel");
}
