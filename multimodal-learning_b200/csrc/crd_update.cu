// K5: fused momentum row update + L2 renormalisation of both memory banks.
//
// Replaces the 2 x (index_select, mul_, add_, pow, sum, pow, div, index_copy_) chain of
// CL_utils/CRD_criterion.py:66-79 with one launch: one warp per (anchor, bank); repeated ids: last occurrence wins.
// Arithmetic follows the reference op for op in fp32 (separate multiply and add, no
// FMA contraction) so a row differs from torch's only by the order of the norm's sum.
#include "common.cuh"

namespace mml {
namespace {

constexpr int kUpdThreads = 128;

__global__ void __launch_bounds__(kUpdThreads) crd_update_kernel(
    float* __restrict__ bank1, float* __restrict__ bank2, int32_t D, const float* __restrict__ v1,
    const float* __restrict__ v2, const int64_t* __restrict__ y, int64_t B, float m, float one_minus_m,
    int64_t row_begin, int64_t row_end) {
  const int64_t w = (static_cast<int64_t>(blockIdx.x) * kUpdThreads + threadIdx.x) >> 5;   // warp id = anchor*2 + bank
  const int lane = threadIdx.x & 31;
  if (w >= 2 * B) return;
  const int64_t i = w >> 1;
  const int64_t row = y[i];
  if (row < row_begin || row >= row_end) return;          // not owned by this rank
  // Duplicate ids in y (replacement sampling, DistributedSampler padding, the same id on two ranks of a sharded batch):
  // the reference reads every old row first and then index_copy_s, so the stored row is ONE well-formed candidate.  Here
  // only the LAST occurrence of an id updates its row (what index_copy_ does when it runs in order), the earlier ones
  // step aside -- no two warps ever write the same row.
  {
    bool later = false;
    for (int64_t j = i + 1 + lane; j < B; j += 32) later |= (y[j] == row);
    if (__any_sync(kFullMask, later)) return;
  }
  float* r = ((w & 1) ? bank2 : bank1) + (row - row_begin) * D;
  const float* v = ((w & 1) ? v2 : v1) + i * D;
  float ss = 0.f;
  // pass 1: blended row (kept in registers when D <= 256) and its squared norm
  float keep[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int d = lane + 32 * k;
    if (d < D) {
      const float t = __fadd_rn(__fmul_rn(r[d], m), __fmul_rn(v[d], one_minus_m));   // :68-69
      keep[k] = t;
      ss = __fadd_rn(ss, __fmul_rn(t, t));                                            // pow(2).sum, :70
    }
  }
  for (int d = lane + 256; d < D; d += 32) {
    const float t = __fadd_rn(__fmul_rn(r[d], m), __fmul_rn(v[d], one_minus_m));
    ss = __fadd_rn(ss, __fmul_rn(t, t));
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(kFullMask, ss, off);
  const float nrm = sqrtf(ss);                                                         // pow(0.5), :70
  // pass 2: divide and write back (:71-72)
  for (int d = lane + 256; d < D; d += 32) {
    const float t = __fadd_rn(__fmul_rn(r[d], m), __fmul_rn(v[d], one_minus_m));
    r[d] = __fdiv_rn(t, nrm);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int d = lane + 32 * k;
    if (d < D) r[d] = __fdiv_rn(keep[k], nrm);
  }
}

__global__ void alias_gather_prob_kernel(const float* __restrict__ prob, const int64_t* __restrict__ kk, int64_t N,
                                         float* __restrict__ out) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < N;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = __ldg(prob + kk[i]);
}

__global__ void alias_select_kernel(const int64_t* __restrict__ alias, const int64_t* __restrict__ kk,
                                    const float* __restrict__ b, int64_t N, const int64_t* __restrict__ y,
                                    int64_t cols, int64_t* __restrict__ out) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < N;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t k = kk[i];
    // kk*b + alias*(1-b) with b in {0,1}  (CRD_criterion.py:138-141)
    int64_t v = (static_cast<int64_t>(b[i]) != 0) ? k : __ldg(alias + k);
    if (y != nullptr && (i % cols) == 0) v = y[i / cols];        // idx.select(1,0).copy_(y), :39
    out[i] = v;
  }
}

int grid_for(int64_t n, int threads) {
  int64_t g = (n + threads - 1) / threads;
  const int64_t cap = 148LL * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace
}  // namespace mml

using namespace mml;

extern "C" int mml_crd_memory_update(float* bank1, float* bank2, int32_t D, const float* v1, const float* v2,
                                     const int64_t* y, int64_t B, float momentum, int64_t row_begin, int64_t row_end,
                                     void* stream) {
  MML_REQUIRE(bank1 && bank2 && v1 && v2 && y, MML_ERR_INVALID_ARG, "crd_memory_update: null pointer argument");
  MML_REQUIRE(D >= 1 && B >= 0 && row_end >= row_begin, MML_ERR_INVALID_ARG, "crd_memory_update: bad sizes");
  if (B == 0) return MML_OK;
  // the reference computes (1 - momentum) in double and torch narrows it to fp32 (:69)
  const float one_minus_m = static_cast<float>(1.0 - static_cast<double>(momentum));
  const int64_t warps = 2 * B;
  const int64_t blocks = (warps * 32 + kUpdThreads - 1) / kUpdThreads;
  crd_update_kernel<<<static_cast<unsigned>(blocks), kUpdThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      bank1, bank2, D, v1, v2, y, B, momentum, one_minus_m, row_begin, row_end);
  return check_launch("crd_update_kernel");
}

extern "C" int mml_alias_build_host(const float* probs, int64_t n, float* prob_out, int64_t* alias_out) {
  MML_REQUIRE(probs && prob_out && alias_out && n >= 1, MML_ERR_INVALID_ARG, "alias_build: bad arguments");
  // CRD_criterion.py:97-123.  Two LIFO stacks; `smaller`/`larger` are filled in index order
  // and popped from the back, exactly as the reference's python lists are.
  int64_t* stack_small = new int64_t[n];
  int64_t* stack_large = new int64_t[n];
  int64_t ns = 0, nl = 0;
  const float kf = static_cast<float>(n);
  for (int64_t i = 0; i < n; ++i) {
    alias_out[i] = 0;
    const float p = kf * probs[i];            // fp32 product, :101
    prob_out[i] = p;
    if (p < 1.0f) stack_small[ns++] = i;
    else stack_large[nl++] = i;
  }
  while (ns > 0 && nl > 0) {
    const int64_t s = stack_small[--ns];
    const int64_t l = stack_large[--nl];
    alias_out[s] = l;
    volatile float t = prob_out[l] - 1.0f;    // two separately rounded fp32 ops, :115
    const float q = t + prob_out[s];
    prob_out[l] = q;
    if (q < 1.0f) stack_small[ns++] = l;
    else stack_large[nl++] = l;
  }
  for (int64_t i = 0; i < ns; ++i) prob_out[stack_small[i]] = 1.0f;   // :122-123
  for (int64_t i = 0; i < nl; ++i) prob_out[stack_large[i]] = 1.0f;
  delete[] stack_small;
  delete[] stack_large;
  return MML_OK;
}

extern "C" int mml_alias_gather_prob(const float* prob, const int64_t* kk, int64_t N, float* p_out, void* stream) {
  MML_REQUIRE(prob && kk && p_out && N >= 0, MML_ERR_INVALID_ARG, "alias_gather_prob: bad arguments");
  if (N == 0) return MML_OK;
  alias_gather_prob_kernel<<<grid_for(N, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(prob, kk, N, p_out);
  return check_launch("alias_gather_prob_kernel");
}

extern "C" int mml_alias_select(const int64_t* alias, const int64_t* kk, const float* b, int64_t N, const int64_t* y,
                                int64_t cols, int64_t* out, void* stream) {
  MML_REQUIRE(alias && kk && b && out && N >= 0, MML_ERR_INVALID_ARG, "alias_select: bad arguments");
  MML_REQUIRE(y == nullptr || (cols >= 1 && N % cols == 0), MML_ERR_INVALID_ARG, "alias_select: N must be rows*cols");
  if (N == 0) return MML_OK;
  alias_select_kernel<<<grid_for(N, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(alias, kk, b, N, y,
                                                                                        cols > 0 ? cols : 1, out);
  return check_launch("alias_select_kernel");
}

// ------------------------------------------------------------------ fused L2 normalisation (Normalize, CRD_criterion.py:236-245)
namespace mml {
namespace {

// y = x / sqrt(sum x^2)  -- one warp per row; the reference's pow(2).sum(1).pow(0.5) + div as one kernel
__global__ void __launch_bounds__(128) l2norm_fwd_kernel(const float* __restrict__ x, int64_t B, int32_t D,
                                                         float* __restrict__ y, float* __restrict__ nrm) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * 128 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* xr = x + row * D;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) { const float v = xr[d]; ss = fmaf(v, v, ss); }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(kFullMask, ss, off);
  const float n = sqrtf(ss);
  for (int d = lane; d < D; d += 32) y[row * D + d] = __fdiv_rn(xr[d], n);
  if (lane == 0) nrm[row] = n;
}

// gx = (gy - y * sum(gy * y)) / norm
__global__ void __launch_bounds__(128) l2norm_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ y,
                                                         const float* __restrict__ nrm, int64_t B, int32_t D,
                                                         float* __restrict__ gx) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * 128 + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* g = gy + row * D;
  const float* yr = y + row * D;
  float dot = 0.f;
  for (int d = lane; d < D; d += 32) dot = fmaf(g[d], yr[d], dot);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(kFullMask, dot, off);
  const float inv = 1.0f / nrm[row];
  for (int d = lane; d < D; d += 32) gx[row * D + d] = (g[d] - yr[d] * dot) * inv;
}

}  // namespace
}  // namespace mml

extern "C" int mml_l2norm_fwd(const float* x, int64_t B, int32_t D, float* y, float* norm, void* stream) {
  MML_REQUIRE(x && y && norm && B >= 0 && D >= 1, MML_ERR_INVALID_ARG, "l2norm_fwd: bad arguments");
  if (B == 0) return MML_OK;
  mml::l2norm_fwd_kernel<<<static_cast<unsigned>((B * 32 + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(x, B, D, y, norm);
  return check_launch("l2norm_fwd_kernel");
}

extern "C" int mml_l2norm_bwd(const float* gy, const float* y, const float* norm, int64_t B, int32_t D, float* gx,
                              void* stream) {
  MML_REQUIRE(gy && y && norm && gx && B >= 0 && D >= 1, MML_ERR_INVALID_ARG, "l2norm_bwd: bad arguments");
  if (B == 0) return MML_OK;
  mml::l2norm_bwd_kernel<<<static_cast<unsigned>((B * 32 + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(gy, y, norm, B, D, gx);
  return check_launch("l2norm_bwd_kernel");
}
