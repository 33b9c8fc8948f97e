// N3: class-conditional contrast_idx produced ON THE DEVICE.
//
// The reference builds every anchor's `sample_idx = hstack(pos_idx, neg_idx)` in DataLoader worker processes
// (MICCAI-2022/data_loaders_MT.py:174-205 pools, :222-249 draws: `np.random.choice` over O(n) candidate lists per
// sample) and uploads the [B, P+K] int64 tensor every step (train_test_MT.py:164) -- 134 MB per step at BASELINE
// config 2, which makes the 8-GPU end-to-end step PCIe-bound.  Here one kernel writes sample_idx[B, P+K] straight into
// HBM from the labels: thread (b, col) = one index.
//
// Pools (never materialised): samples are kept class-sorted (`order`, `cls_ptr`); cls_positive[c] is the segment of
// class c (:193-195), cls_negative[c] is `order` with that segment cut out (:197-202, same class order).  For the
// survival task the negative pool is every sample but the anchor (:224-227).
// Draws: with replacement when the request exceeds the pool (`replace = k > len(pool)`, :226,243), else WITHOUT
// replacement like `np.random.choice(..., replace=False)`.  The latter is a keyed bijection of [0, pool) (8-round
// Feistel network on an even number of bits + cycle walking): position j of the draw is perm(j), so K distinct
// members come out in parallel with no shuffle of the pool.  Randomness is counter-based Philox4x32-10 keyed by
// (seed ^ *seed_dev, anchor, column): the numpy mt19937 stream of the reference cannot be reproduced on a GPU, so
// parity is "same algorithm as oracle/sampler_oracle.py, bit for bit" + the reference's distributional contract.
#include "common.cuh"

namespace mml {
namespace {

struct SamplerArgs {
  const int64_t* index;       // [B] anchors (dataset indices)
  const int32_t* labels;      // [n] class of every sample (grad task) or NULL (surv task)
  const int32_t* order;       // [n] sample ids sorted by class (stable)
  const int32_t* cls_ptr;     // [C+1]
  int64_t* out;               // [B, P+K]
  int64_t B, n;
  int32_t P, K, pos_mode;     // pos_mode: 0 exact, 1 relax, 2 multi_pos
  uint32_t seed_lo, seed_hi;
  const unsigned long long* seed_dev;
};

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = static_cast<uint32_t>(p1);
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = static_cast<uint32_t>(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x7feb352du;
  h ^= h >> 15;
  h *= 0x846ca68bu;
  h ^= h >> 16;
  return h;
}

// j-th element of a uniform draw of distinct members of [0, M): keyed Feistel bijection + cycle walking
__device__ __forceinline__ uint32_t perm_element(uint32_t j, uint32_t M, const uint32_t* key) {
  int bits = 32 - __clz(M - 1 > 0 ? M - 1 : 1);
  if (M <= 2) bits = 2;
  bits = (bits + 1) & ~1;                       // even
  const int half = bits >> 1;
  const uint32_t mask = (1u << half) - 1u;
  uint32_t x = j;
  do {
    uint32_t L = x >> half, R = x & mask;
    // 8 rounds (round keys: the four Philox words, then the same words plus the golden-ratio constant).  With 4 rounds the
    // first TWO images of a small domain are visibly not jointly uniform (chi-square 1866 on 131 degrees of freedom for
    // M = 12, 6 rounds 269, 8 rounds 126: tests/test_sampler_cpu.py) -- np.random.choice(replace=False) is uniform over
    // ordered K-subsets, so marginal uniformity of each position is not enough.
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const uint32_t t = L ^ (mix32(R ^ (key[r & 3] + 0x9E3779B9u * static_cast<uint32_t>(r >> 2))) & mask);
      L = R;
      R = t;
    }
    x = (L << half) | R;
  } while (x >= M);
  return x;
}

// uniform member of [0, M) from one 32-bit word (multiply-shift; bias <= M / 2^32)
__device__ __forceinline__ uint32_t bounded(uint32_t r, uint32_t M) {
  return static_cast<uint32_t>((static_cast<uint64_t>(r) * M) >> 32);
}

__global__ void __launch_bounds__(256) instance_sample_kernel(const SamplerArgs a) {
  const int64_t b = blockIdx.y;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int cols = a.P + a.K;
  uint32_t k0 = a.seed_lo, k1 = a.seed_hi;
  if (a.seed_dev != nullptr) {
    const unsigned long long s = __ldg(a.seed_dev);
    k0 ^= static_cast<uint32_t>(s);
    k1 ^= static_cast<uint32_t>(s >> 32);
  }
  const int64_t anchor = a.index[b];
  const uint32_t blo = static_cast<uint32_t>(b), bhi = static_cast<uint32_t>(static_cast<uint64_t>(b) >> 32);
  // the Feistel keys depend on the anchor row only: one Philox evaluation per CTA, not per thread
  __shared__ uint32_t sm_key[3][4];            // [0] survival negatives, [1] multi_pos positives, [2] class negatives
  if (threadIdx.x < 3) {
    const uint32_t c1 = threadIdx.x == 0 ? 1u : (threadIdx.x == 1 ? 3u : 5u);
    const uint32_t tag = threadIdx.x == 0 ? 0x20000000u : (threadIdx.x == 1 ? 0x40000000u : 0x60000000u);
    philox4x32_10(0u, c1, blo, bhi ^ tag, k0, k1, sm_key[threadIdx.x]);
  }
  __syncthreads();
  if (col >= cols) return;
  uint32_t rnd[4];
  const uint32_t* key;
  int64_t result;
  if (a.labels == nullptr) {                                   // survival task (:222-227)
    if (col < a.P) {
      result = anchor;                                           // pos_idx = index (P == 1)
    } else {
      const uint32_t M = static_cast<uint32_t>(a.n - 1);
      const uint32_t j = static_cast<uint32_t>(col - a.P);
      uint32_t e;
      if (static_cast<uint32_t>(a.K) > M) {
        philox4x32_10(j, 0u, blo, bhi ^ 0x10000000u, k0, k1, rnd);
        e = bounded(rnd[0], M);
      } else {
        key = sm_key[0];
        e = perm_element(j, M, key);
      }
      result = e < anchor ? e : e + 1;                           // all_neg_idx.remove(index)
    }
  } else {
    const int c = a.labels[anchor];
    const uint32_t seg0 = static_cast<uint32_t>(a.cls_ptr[c]), seg1 = static_cast<uint32_t>(a.cls_ptr[c + 1]);
    const uint32_t Mp = seg1 - seg0, Mn = static_cast<uint32_t>(a.n) - Mp;
    if (col < a.P) {
      if (a.pos_mode == 0 || (a.pos_mode == 2 && col == 0)) {
        result = anchor;                                         // 'exact' (:229-230); pos_idx[0] = index (:238)
      } else if (a.pos_mode == 1) {
        philox4x32_10(static_cast<uint32_t>(col), 2u, blo, bhi ^ 0x30000000u, k0, k1, rnd);
        result = a.order[seg0 + bounded(rnd[0], Mp)];            // 'relax': one member of the anchor's class (:232)
      } else {
        key = sm_key[1];
        result = a.order[seg0 + perm_element(static_cast<uint32_t>(col), Mp, key)];   // 'multi_pos', replace=False (:237)
      }
    } else {
      const uint32_t j = static_cast<uint32_t>(col - a.P);
      uint32_t e;
      if (static_cast<uint32_t>(a.K) > Mn) {                    // replace = k > len(cls_negative) (:243)
        philox4x32_10(j, 4u, blo, bhi ^ 0x50000000u, k0, k1, rnd);
        e = bounded(rnd[0], Mn);
      } else {
        key = sm_key[2];
        e = perm_element(j, Mn, key);
      }
      result = a.order[e < seg0 ? e : e + Mp];                   // cls_negative[c] = order minus class c's segment
    }
  }
  a.out[b * cols + col] = result;
}

}  // namespace
}  // namespace mml

using namespace mml;

extern "C" int mml_instance_sample(const int64_t* index, int64_t B, const int32_t* labels, const int32_t* order,
                                   const int32_t* cls_ptr, int32_t num_classes, int64_t n, int32_t P, int32_t K,
                                   int32_t pos_mode, uint64_t seed, const uint64_t* seed_dev, int64_t* out, void* stream) {
  MML_REQUIRE(index && out, MML_ERR_INVALID_ARG, "instance_sample: null pointer argument");
  MML_REQUIRE(B >= 0 && B <= 65535 && n >= 2 && n < (1LL << 31), MML_ERR_INVALID_ARG, "instance_sample: bad B / n");
  MML_REQUIRE(P >= 1 && K >= 0 && pos_mode >= 0 && pos_mode <= 2, MML_ERR_INVALID_ARG, "instance_sample: bad P / K / pos_mode");
  MML_REQUIRE(pos_mode == 2 || P == 1, MML_ERR_INVALID_ARG, "instance_sample: 'exact' / 'relax' draw one positive (P = 1)");
  if (labels != nullptr)
    MML_REQUIRE(order && cls_ptr && num_classes >= 2, MML_ERR_INVALID_ARG, "instance_sample: class tables missing");
  else
    MML_REQUIRE(pos_mode == 0, MML_ERR_INVALID_ARG, "instance_sample: the survival task has one exact positive");
  if (B == 0) return MML_OK;
  SamplerArgs a{};
  a.index = index; a.labels = labels; a.order = order; a.cls_ptr = cls_ptr; a.out = out; a.B = B; a.n = n;
  a.P = P; a.K = K; a.pos_mode = pos_mode;
  a.seed_lo = static_cast<uint32_t>(seed); a.seed_hi = static_cast<uint32_t>(seed >> 32);
  a.seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
  const int cols = P + K;
  const dim3 grid((cols + 255) / 256, static_cast<unsigned>(B));
  instance_sample_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("instance_sample_kernel");
}
