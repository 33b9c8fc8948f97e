// Exact-fp32 CUDA-core kernels for the Kronecker-fusion encoder (fusion.py:58-60, :126-129):
//   forward   y[b,n]  = sum_k A[b,k] m[b,k] W[n,k] + bias[n]
//   wgrad     dW[n,k] = sum_b dy[b,n] A[b,k] m[b,k]
//   dgrad     df_x[b,.] from dA[b,k] = m[b,k] sum_n dy[b,n] W[n,k]   (never stored: contracted with the other factors)
// A is generated on the fly from the factors (kron_common.cuh); m is the dropout multiplier.
// These are the bit-faithful fp32 path (rel ~1e-6 vs the oracle): used for backward in round 1,
// for shapes the tcgen05 kernel does not take, and as the on-GPU cross-check of the tensor-core kernel.
#include "kron_common.cuh"

namespace mml {
namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------ forward
__global__ void __launch_bounds__(kThreads) kron_fwd_simt_kernel(
    KronShape s, KronDropout dr, const float* __restrict__ f1, const float* __restrict__ f2,
    const float* __restrict__ f3, int64_t B, const float* __restrict__ W, const float* __restrict__ bias, int32_t N,
    float* __restrict__ y) {
  kron_seed(dr, dr.seed_lo, dr.seed_hi);
  dr.seed_dev = nullptr;
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 1];
  __shared__ float Ws[BK][BN + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t b0 = static_cast<int64_t>(blockIdx.x) * BM;
  const int n0 = blockIdx.y * BN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < s.Kk; k0 += BK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + kThreads * r;
      const int bl = e & 63, kk = e >> 6;
      const int64_t b = b0 + bl;
      const int k = k0 + kk;
      float v = 0.f;
      if (b < B && k < s.Kk) v = kron_element(s, f1, f2, f3, b, k) * kron_keep(dr, b, k);
      As[kk][bl] = v;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + kThreads * r;
      const int kk = e & 15, nl = e >> 4;
      const int n = n0 + nl, k = k0 + kk;
      Ws[kk][nl] = (n < N && k < s.Kk) ? __ldg(W + static_cast<int64_t>(n) * s.Kk + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; w[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t b = b0 + ty * 4 + i;
    if (b >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) y[b * N + n] = acc[i][j] + (bias ? bias[n] : 0.f);
    }
  }
}

// ------------------------------------------------------------------ wgrad
__global__ void __launch_bounds__(kThreads) kron_wgrad_simt_kernel(
    KronShape s, KronDropout dr, const float* __restrict__ f1, const float* __restrict__ f2,
    const float* __restrict__ f3, int64_t B, const float* __restrict__ dy, int32_t N, float* __restrict__ dW,
    int64_t rows_per_split, int use_atomic) {
  kron_seed(dr, dr.seed_lo, dr.seed_hi);
  dr.seed_dev = nullptr;
  constexpr int BM = 64, BN = 64, BK = 16;   // BM: n, BN: k, BK: batch rows per step
  __shared__ float Ds[BK][BM + 1];
  __shared__ float As[BK][BN + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int k0 = blockIdx.x * BN;
  const int n0 = blockIdx.y * BM;
  const int64_t bs = static_cast<int64_t>(blockIdx.z) * rows_per_split;
  const int64_t be = min(B, bs + rows_per_split);
  float acc[4][4] = {};
  for (int64_t bb = bs; bb < be; bb += BK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + kThreads * r;
      const int nl = e & 63, bl = e >> 6;
      const int64_t b = bb + bl;
      const int n = n0 + nl;
      Ds[bl][nl] = (b < be && n < N) ? __ldg(dy + b * N + n) : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + kThreads * r;
      const int kl = e & 63, bl = e >> 6;
      const int64_t b = bb + bl;
      const int k = k0 + kl;
      float v = 0.f;
      if (b < be && k < s.Kk) v = kron_element(s, f1, f2, f3, b, k) * kron_keep(dr, b, k);
      As[bl][kl] = v;
    }
    __syncthreads();
#pragma unroll
    for (int bl = 0; bl < BK; ++bl) {
      float d[4], a[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { d[i] = Ds[bl][ty * 4 + i]; a[i] = As[bl][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(d[i], a[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k >= s.Kk) continue;
      float* dst = dW + static_cast<int64_t>(n) * s.Kk + k;
      if (use_atomic) atomicAdd(dst, acc[i][j]);
      else *dst = acc[i][j];
    }
  }
}

// ------------------------------------------------------------------ dgrad
// CTA = 32 batch rows x all k.  dA tiles (32 x 64) are formed in registers by a smem-tiled
// contraction over n and immediately folded into per-row factor gradients held in shared memory.
__global__ void __launch_bounds__(kThreads) kron_dgrad_simt_kernel(
    KronShape s, KronDropout dr, const float* __restrict__ f1, const float* __restrict__ f2,
    const float* __restrict__ f3, int64_t B, const float* __restrict__ W, const float* __restrict__ dy, int32_t N,
    float* __restrict__ df1, float* __restrict__ df2, float* __restrict__ df3) {
  kron_seed(dr, dr.seed_lo, dr.seed_hi);
  dr.seed_dev = nullptr;
  constexpr int BM = 32, BN = 64, BK = 16;   // BM: batch rows, BN: k, BK: n per step
  extern __shared__ float sm_df[];            // [BM][d1 + d2 + d3]
  __shared__ float Ds[BK][BM + 1];
  __shared__ float Ws[BK][BN + 1];
  __shared__ int sm_ijl[BN][3];
  const int dsum = s.d1 + s.d2 + s.d3;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;     // tx: 4 k each, ty: 2 rows each
  const int64_t b0 = static_cast<int64_t>(blockIdx.x) * BM;
  for (int i = tid; i < BM * dsum; i += kThreads) sm_df[i] = 0.f;
  for (int k0 = 0; k0 < s.Kk; k0 += BN) {
    float acc[2][4] = {};
    if (tid < BN) {
      int i, j, l;
      kron_decode(s, min(k0 + tid, s.Kk - 1), i, j, l);
      sm_ijl[tid][0] = i; sm_ijl[tid][1] = j; sm_ijl[tid][2] = l;
    }
    for (int nn = 0; nn < N; nn += BK) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int e = tid + kThreads * r;
        const int nl = e & 15, bl = e >> 4;
        const int64_t b = b0 + bl;
        const int n = nn + nl;
        Ds[nl][bl] = (b < B && n < N) ? __ldg(dy + b * N + n) : 0.f;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int e = tid + kThreads * r;
        const int kl = e & 63, nl = e >> 6;
        const int n = nn + nl, k = k0 + kl;
        Ws[nl][kl] = (n < N && k < s.Kk) ? __ldg(W + static_cast<int64_t>(n) * s.Kk + k) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int nl = 0; nl < BK; ++nl) {
        float d[2], w[4];
        d[0] = Ds[nl][ty * 2];
        d[1] = Ds[nl][ty * 2 + 1];
#pragma unroll
        for (int j = 0; j < 4; ++j) w[j] = Ws[nl][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(d[i], w[j], acc[i][j]);
      }
      __syncthreads();
    }
    // fold dA into the factor gradients
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int bl = ty * 2 + i;
      const int64_t b = b0 + bl;
      if (b >= B) continue;
      float* row = sm_df + bl * dsum;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kl = tx * 4 + j;
        const int k = k0 + kl;
        if (k >= s.Kk) continue;
        const float dA = acc[i][j] * kron_keep(dr, b, k);
        const int ii = sm_ijl[kl][0], jj = sm_ijl[kl][1], ll = sm_ijl[kl][2];
        const float g1 = kron_factor(f1, b, s.d1, ii);
        const float g2 = kron_factor(f2, b, s.d2, jj);
        const float g3 = (s.nf == 3) ? kron_factor(f3, b, s.d3, ll) : 1.f;
        if (ii < s.d1) atomicAdd(row + ii, dA * g2 * g3);             // the appended 1 gets no gradient
        if (jj < s.d2) atomicAdd(row + s.d1 + jj, dA * g1 * g3);
        if (s.nf == 3 && ll < s.d3) atomicAdd(row + s.d1 + s.d2 + ll, dA * g1 * g2);
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < BM * dsum; e += kThreads) {
    const int bl = e / dsum, x = e % dsum;
    const int64_t b = b0 + bl;
    if (b >= B) continue;
    if (x < s.d1) df1[b * s.d1 + x] = sm_df[e];
    else if (x < s.d1 + s.d2) df2[b * s.d2 + (x - s.d1)] = sm_df[e];
    else df3[b * s.d3 + (x - s.d1 - s.d2)] = sm_df[e];
  }
}

int check_shape(const float* f1, const float* f2, const float* f3, int64_t B, int32_t d1, int32_t d2, int32_t d3,
                int32_t N) {
  MML_REQUIRE(f1 && f2, MML_ERR_INVALID_ARG, "kron: null factor pointer");
  MML_REQUIRE((d3 > 0) == (f3 != nullptr), MML_ERR_INVALID_ARG, "kron: f3 and d3 must both be set or both be absent");
  MML_REQUIRE(B >= 0 && d1 >= 1 && d2 >= 1 && d3 >= 0 && N >= 1, MML_ERR_INVALID_ARG, "kron: bad sizes");
  MML_REQUIRE(static_cast<int64_t>(d1 + 1) * (d2 + 1) * (d3 > 0 ? d3 + 1 : 1) < (1LL << 30), MML_ERR_UNSUPPORTED,
              "kron: Kronecker width too large");
  return MML_OK;
}

}  // namespace
}  // namespace mml

using namespace mml;

extern "C" int mml_kron_linear_fwd_simt(const float* f1, const float* f2, const float* f3, int64_t B, int32_t d1,
                                        int32_t d2, int32_t d3, const float* W, const float* bias, int32_t N,
                                        float drop_p, uint64_t seed, const uint64_t* seed_dev, int32_t training, float* y, void* stream) {
  int rc = check_shape(f1, f2, f3, B, d1, d2, d3, N);
  if (rc != MML_OK) return rc;
  MML_REQUIRE(W && y, MML_ERR_INVALID_ARG, "kron_fwd_simt: null pointer");
  if (B == 0) return MML_OK;
  const KronShape s = make_kron_shape(d1, d2, d3);
  const KronDropout dr = make_kron_dropout(drop_p, seed, training, s.Kk, seed_dev);
  const dim3 grid(static_cast<unsigned>((B + 63) / 64), (N + 63) / 64);
  kron_fwd_simt_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(s, dr, f1, f2, f3, B, W, bias, N, y);
  return check_launch("kron_fwd_simt_kernel");
}

extern "C" int mml_kron_linear_bwd_simt(const float* f1, const float* f2, const float* f3, int64_t B, int32_t d1,
                                        int32_t d2, int32_t d3, const float* W, const float* dy, int32_t N,
                                        float drop_p, uint64_t seed, const uint64_t* seed_dev, int32_t training, float* df1, float* df2,
                                        float* df3, float* dW, void* stream) {
  int rc = check_shape(f1, f2, f3, B, d1, d2, d3, N);
  if (rc != MML_OK) return rc;
  MML_REQUIRE(W && dy, MML_ERR_INVALID_ARG, "kron_bwd_simt: null pointer");
  MML_REQUIRE((df1 != nullptr) == (df2 != nullptr) && (d3 == 0 || (df3 != nullptr) == (df1 != nullptr)),
              MML_ERR_INVALID_ARG, "kron_bwd_simt: factor gradients are all-or-none");
  if (B == 0) return MML_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const KronShape s = make_kron_shape(d1, d2, d3);
  const KronDropout dr = make_kron_dropout(drop_p, seed, training, s.Kk, seed_dev);
  if (dW != nullptr) {
    const int tiles = ((s.Kk + 63) / 64) * ((N + 63) / 64);
    int64_t split = (148 * 2 + tiles - 1) / tiles;
    const int64_t max_split = (B + 255) / 256;
    if (split > max_split) split = max_split;
    if (split < 1) split = 1;
    int64_t rows = (B + split - 1) / split;
    rows = (rows + 15) / 16 * 16;
    split = (B + rows - 1) / rows;
    if (split > 1) MML_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * static_cast<size_t>(N) * s.Kk, st));
    const dim3 grid((s.Kk + 63) / 64, (N + 63) / 64, static_cast<unsigned>(split));
    kron_wgrad_simt_kernel<<<grid, kThreads, 0, st>>>(s, dr, f1, f2, f3, B, dy, N, dW, rows, split > 1 ? 1 : 0);
    rc = check_launch("kron_wgrad_simt_kernel");
    if (rc != MML_OK) return rc;
  }
  if (df1 != nullptr) {
    const size_t smem = sizeof(float) * 32 * static_cast<size_t>(d1 + d2 + d3);
    MML_REQUIRE(smem <= 160 * 1024, MML_ERR_UNSUPPORTED, "kron_dgrad_simt: factor widths too large for shared memory");
    if (smem > 40 * 1024)
      MML_CUDA(cudaFuncSetAttribute(kron_dgrad_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(smem)));
    kron_dgrad_simt_kernel<<<static_cast<unsigned>((B + 31) / 32), kThreads, smem, st>>>(s, dr, f1, f2, f3, B, W, dy, N,
                                                                                         df1, df2, df3);
    rc = check_launch("kron_dgrad_simt_kernel");
  }
  return rc;
}
