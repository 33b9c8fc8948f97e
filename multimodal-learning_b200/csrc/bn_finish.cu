// Finisher of encoder1 = Linear -> BatchNorm1d -> ReLU (fusion.py:29, :60): the Kronecker forward
// kernel (kron_tc.cu) emits per-column partial sums of y and y^2 from its epilogue, this file turns them into the batch
// statistics, updates the running statistics exactly as nn.BatchNorm1d does (momentum 0.1, eps 1e-5 by default -- the
// reference passes none; biased variance for the normalisation, unbiased for running_var) and applies
// normalise + affine + ReLU in one pass over [B, N].  When the forward ran split over K (small batches) its epilogue only
// holds partial sums of y, and the statistics are reduced from y itself (two reads of an L2-resident matrix).
// Sums are accumulated in double and in a fixed order: deterministic, and closer to the exact mean / variance than the
// fp32 reductions of the library kernel the reference calls.
#include "common.cuh"

namespace mml {
namespace {

constexpr int kBnCols = 32;
constexpr int kBnGroups = 8;

// grid = ceil(N / 32); block = 32 x 8.  Column statistics from partials [n_part][2][N] (n_part > 0) or from y [B][N].
__global__ void __launch_bounds__(kBnCols * kBnGroups) bn_stats_kernel(
    const float* __restrict__ y, int64_t B, int32_t N, const float* __restrict__ parts, int32_t n_part, float momentum, float eps,
    float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ save_mean,
    float* __restrict__ save_invstd) {
  __shared__ double sm[2][kBnGroups][kBnCols];
  const int c = threadIdx.x & (kBnCols - 1), g = threadIdx.x / kBnCols;
  const int n = blockIdx.x * kBnCols + c;
  double s = 0.0, q = 0.0;
  if (n < N) {
    if (n_part > 0) {
      for (int p = g; p < n_part; p += kBnGroups) {
        s += static_cast<double>(parts[(static_cast<int64_t>(p) * 2 + 0) * N + n]);
        q += static_cast<double>(parts[(static_cast<int64_t>(p) * 2 + 1) * N + n]);
      }
    } else {
      for (int64_t r = g; r < B; r += kBnGroups) {
        const double v = static_cast<double>(y[r * N + n]);
        s += v;
        q += v * v;
      }
    }
  }
  sm[0][g][c] = s;
  sm[1][g][c] = q;
  __syncthreads();
  if (g == 0 && n < N) {
    double ts = 0.0, tq = 0.0;
#pragma unroll
    for (int i = 0; i < kBnGroups; ++i) {
      ts += sm[0][i][c];
      tq += sm[1][i][c];
    }
    const double mean = ts / static_cast<double>(B);
    double var = tq / static_cast<double>(B) - mean * mean;          // biased (normalisation), torch.var(unbiased=False)
    if (var < 0.0) var = 0.0;
    save_mean[n] = static_cast<float>(mean);
    save_invstd[n] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    if (running_mean != nullptr) {
      const double unbiased = B > 1 ? var * static_cast<double>(B) / static_cast<double>(B - 1) : var;
      running_mean[n] = static_cast<float>((1.0 - momentum) * static_cast<double>(running_mean[n]) + momentum * mean);
      running_var[n] = static_cast<float>((1.0 - momentum) * static_cast<double>(running_var[n]) + momentum * unbiased);
    }
  }
}

__global__ void bn_relu_apply_kernel(const float* __restrict__ y, int64_t total, int32_t N, const float* __restrict__ mean,
                                     const float* __restrict__ invstd, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ out) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i % N);
    const float w = gamma ? gamma[n] : 1.f, b = beta ? beta[n] : 0.f;
    const float v = (y[i] - mean[n]) * invstd[n] * w + b;            // same association as the native kernel
    out[i] = v > 0.f ? v : 0.f;
  }
}

}  // namespace
}  // namespace mml

using namespace mml;

extern "C" int mml_bn_relu_fwd(const float* y, int64_t B, int32_t N, const float* col_stats, int32_t n_part, const float* gamma,
                               const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                               float* out, float* save_mean, float* save_invstd, void* stream) {
  MML_REQUIRE(y && out && save_mean && save_invstd, MML_ERR_INVALID_ARG, "bn_relu_fwd: null pointer");
  MML_REQUIRE(B >= 1 && N >= 1 && n_part >= 0, MML_ERR_INVALID_ARG, "bn_relu_fwd: bad sizes");
  MML_REQUIRE((running_mean == nullptr) == (running_var == nullptr), MML_ERR_INVALID_ARG,
              "bn_relu_fwd: running_mean and running_var must both be set or both be absent");
  MML_REQUIRE(n_part == 0 || col_stats != nullptr, MML_ERR_INVALID_ARG, "bn_relu_fwd: n_part > 0 needs col_stats");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bn_stats_kernel<<<(N + kBnCols - 1) / kBnCols, kBnCols * kBnGroups, 0, st>>>(y, B, N, col_stats, n_part, momentum, eps,
                                                                                running_mean, running_var, save_mean, save_invstd);
  int rc = check_launch("bn_stats_kernel");
  if (rc != MML_OK) return rc;
  const int64_t total = B * N;
  int64_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  bn_relu_apply_kernel<<<static_cast<unsigned>(g), 256, 0, st>>>(y, total, N, save_mean, save_invstd, gamma, beta, out);
  return check_launch("bn_relu_apply_kernel");
}
