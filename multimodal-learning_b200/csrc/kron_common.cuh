// Shared definitions of the never-materialised Kronecker operand
//   A[b, k] = g1[b,i] * g2[b,j] (* g3[b,l]),  g(x) = x < d ? f[b,x] : 1     (fusion.py:56-58, :123-127)
//   k = (i*(d2+1) + j)*(d3+1) + l   (row-major flatten; the constant 1 sits in the LAST slot of each factor)
// and of the counter-based dropout mask that replaces `post_fusion_dropout` (fusion.py:59,128):
// a tensor that no longer exists cannot carry a torch mask, so the mask is a pure function of
// (seed, b, k) evaluated identically in forward, dgrad and wgrad.
#pragma once

#include "common.cuh"

namespace mml {

struct KronShape {
  int32_t nf;       // 2 (bilinear) or 3 (trilinear)
  int32_t d1, d2, d3;   // factor widths without the appended 1 (d3 = 0 when nf == 2)
  int32_t e2, e3;   // d2+1, d3+1 (e3 = 1 when nf == 2)
  int32_t Kk;       // (d1+1)*(d2+1)*(d3+1 | 1)
};

inline KronShape make_kron_shape(int32_t d1, int32_t d2, int32_t d3) {
  KronShape s;
  s.nf = d3 > 0 ? 3 : 2;
  s.d1 = d1; s.d2 = d2; s.d3 = d3;
  s.e2 = d2 + 1;
  s.e3 = d3 > 0 ? d3 + 1 : 1;
  s.Kk = (d1 + 1) * s.e2 * s.e3;
  return s;
}

struct KronDropout {
  uint32_t thresh;    // drop iff r16 < thresh;  thresh = round(p * 65536); 0 = no dropout
  float scale;        // 65536 / (65536 - thresh)
  uint32_t seed_lo, seed_hi;
  int64_t pairs_per_row;   // ceil(Kk / 2): one 32-bit hash serves two neighbouring k
  const unsigned long long* seed_dev;   // optional DEVICE word xor-ed into the seed at run time: lets a captured CUDA graph
                                        // draw a fresh mask on every replay (the host `seed` is frozen into the graph)
};

inline KronDropout make_kron_dropout(float p, uint64_t seed, int training, int32_t Kk, const uint64_t* seed_dev = nullptr) {
  KronDropout d;
  double t = (training && p > 0.f) ? static_cast<double>(p) * 65536.0 + 0.5 : 0.0;
  if (t > 65535.0) t = 65535.0;
  d.thresh = static_cast<uint32_t>(t);
  d.scale = 65536.0f / static_cast<float>(65536u - d.thresh);
  d.seed_lo = static_cast<uint32_t>(seed);
  d.seed_hi = static_cast<uint32_t>(seed >> 32);
  d.pairs_per_row = (static_cast<int64_t>(Kk) + 1) / 2;
  d.seed_dev = (d.thresh != 0u) ? reinterpret_cast<const unsigned long long*>(seed_dev) : nullptr;
  return d;
}

__host__ __device__ __forceinline__ uint32_t kron_hash(uint32_t lo, uint32_t hi, uint32_t seed_lo, uint32_t seed_hi) {
  uint32_t h = lo ^ seed_lo;
  h *= 0x9E3779B1u;
  h ^= hi ^ seed_hi;
  h ^= h >> 16;
  h *= 0x7feb352du;
  h ^= h >> 15;
  h *= 0x846ca68bu;
  h ^= h >> 16;
  return h;
}

// effective 64-bit seed of this launch: host seed xor the optional device word
__device__ __forceinline__ void kron_seed(const KronDropout& dr, uint32_t& lo, uint32_t& hi) {
  lo = dr.seed_lo;
  hi = dr.seed_hi;
  if (dr.seed_dev != nullptr) {
    const unsigned long long s = __ldg(dr.seed_dev);
    lo ^= static_cast<uint32_t>(s);
    hi ^= static_cast<uint32_t>(s >> 32);
  }
}

// multiplier of A[b,k] under dropout: 0 or scale
__device__ __forceinline__ float kron_keep(const KronDropout& dr, int64_t b, int32_t k) {   // dr: kron_seed() already folded in
  if (dr.thresh == 0u) return 1.0f;
  const int64_t c = b * dr.pairs_per_row + (k >> 1);
  const uint32_t h = kron_hash(static_cast<uint32_t>(c), static_cast<uint32_t>(static_cast<uint64_t>(c) >> 32),
                               dr.seed_lo, dr.seed_hi);
  const uint32_t r16 = (k & 1) ? (h >> 16) : (h & 0xffffu);
  return r16 >= dr.thresh ? dr.scale : 0.0f;
}

__device__ __forceinline__ void kron_decode(const KronShape& s, int32_t k, int32_t& i, int32_t& j, int32_t& l) {
  l = k % s.e3;
  const int32_t t = k / s.e3;
  j = t % s.e2;
  i = t / s.e2;
}

__device__ __forceinline__ float kron_factor(const float* __restrict__ f, int64_t b, int32_t d, int32_t x) {
  return x < d ? __ldg(f + b * d + x) : 1.0f;
}

// A[b,k] without dropout
__device__ __forceinline__ float kron_element(const KronShape& s, const float* f1, const float* f2, const float* f3,
                                              int64_t b, int32_t k) {
  int32_t i, j, l;
  kron_decode(s, k, i, j, l);
  float v = kron_factor(f1, b, s.d1, i) * kron_factor(f2, b, s.d2, j);   // o12, :58 / :126
  if (s.nf == 3) v *= kron_factor(f3, b, s.d3, l);                         // o123, :127
  return v;
}

}  // namespace mml
