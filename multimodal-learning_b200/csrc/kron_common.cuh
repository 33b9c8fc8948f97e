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

// post_fusion_dropout on the never-stored operand: the DROP flag of (row b, logical column k) is bit (k & 31) of a 32-bit
// word shared by the 32 columns [32w, 32w + 32) of that row.  A word is the bit-sliced comparison U < thresh of 32
// independent 16-bit uniforms whose bit planes are counter-hash words: going from the lowest SET bit of thresh upwards,
//   lt = thresh_i ? (plane_i | lt) : (plane_i & lt),     plane_i = hash((b * words_per_row + w) * 16 + i, seed)
// (planes below the lowest set bit cannot change the outcome and are never drawn), so a word costs 16 - ctz(thresh) hashes:
// 2 for p = 0.25, 1 for p = 0.5.  K1/K2 (thread = row, 16 consecutive k) need about one new word per chunk, K3
// (thread = k, 16 consecutive rows) shares the words of a warp's rows by shuffle.  Restated in numpy in
// oracle/fusion_oracle.kron_dropout_mask.
struct KronDropout {
  uint32_t thresh;    // drop iff U16 < thresh;  thresh = round(p * 65536); 0 = no dropout
  float scale;        // 65536 / (65536 - thresh)
  uint32_t seed_lo, seed_hi;
  int64_t words_per_row;   // ceil(Kk / 32)
  int32_t plane0;          // ctz(thresh): the first bit plane that matters
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
  d.words_per_row = (static_cast<int64_t>(Kk) + 31) / 32;
  d.plane0 = 0;
  while (d.thresh != 0u && ((d.thresh >> d.plane0) & 1u) == 0u) ++d.plane0;
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

// DROP flags of row b (row_words = b * words_per_row), columns [32w, 32w + 32); seed_lo/hi from kron_seed()
__device__ __forceinline__ uint32_t kron_drop_word(const KronDropout& dr, uint32_t seed_lo, uint32_t seed_hi, int64_t row_words,
                                                   int32_t w) {
  const uint64_t c0 = static_cast<uint64_t>(row_words + w) << 4;
  if (dr.plane0 >= 14) {           // one or two planes (p = 0.5, 0.25, 0.75): straight-line code, the two hashes independent
    const uint64_t ca = c0 + static_cast<uint32_t>(dr.plane0), cb = c0 + 15u;
    const uint32_t ra = kron_hash(static_cast<uint32_t>(ca), static_cast<uint32_t>(ca >> 32), seed_lo, seed_hi);
    if (dr.plane0 == 15) return ra;
    const uint32_t rb = kron_hash(static_cast<uint32_t>(cb), static_cast<uint32_t>(cb >> 32), seed_lo, seed_hi);
    return (dr.thresh & 0x8000u) ? (rb | ra) : (rb & ra);
  }
  uint32_t lt = 0u;
  for (int i = dr.plane0; i < 16; ++i) {
    const uint64_t c = c0 + static_cast<uint32_t>(i);
    const uint32_t r = kron_hash(static_cast<uint32_t>(c), static_cast<uint32_t>(c >> 32), seed_lo, seed_hi);
    lt = ((dr.thresh >> i) & 1u) ? (r | lt) : (r & lt);
  }
  return lt;
}

// multiplier of A[b,k] under dropout: 0 or scale (per-element form for the CUDA-core kernels; dr.seed_* already folded)
__device__ __forceinline__ float kron_keep(const KronDropout& dr, int64_t b, int32_t k) {
  if (dr.thresh == 0u) return 1.0f;
  const uint32_t word = kron_drop_word(dr, dr.seed_lo, dr.seed_hi, b * dr.words_per_row, k >> 5);
  return ((word >> (k & 31)) & 1u) ? 0.0f : dr.scale;
}

// The word a row computed last: consecutive chunks of a run cover consecutive k, so the upper word of one chunk is the
// lower word of the next.
struct KronDropCache {
  int32_t w;
  uint32_t v;
};

// DROP flags (bit u = element u) of the 16 columns k0 + u * kstride of one row.  kstride == 1 (every core chunk): one or
// two words and a funnel shift; other strides (the faces that hold an appended 1): a word per element.  All branches are
// uniform over threads that walk the same chunks.
__device__ __forceinline__ uint32_t kron_drop_bits16(const KronDropout& dr, uint32_t seed_lo, uint32_t seed_hi, int64_t row_words,
                                                     int32_t k0, int32_t kstride, KronDropCache& cache) {
  if (kstride == 1) {
    const int32_t w0 = k0 >> 5, sh = k0 & 31;
    const uint32_t lo = (cache.w == w0) ? cache.v : kron_drop_word(dr, seed_lo, seed_hi, row_words, w0);
    uint32_t hi = 0u;
    if (sh > 16) {
      hi = kron_drop_word(dr, seed_lo, seed_hi, row_words, w0 + 1);
      cache.w = w0 + 1;
      cache.v = hi;
    } else {
      cache.w = w0;
      cache.v = lo;
    }
    return __funnelshift_r(lo, hi, sh) & 0xffffu;
  }
  uint32_t bits = 0u;
#pragma unroll 4
  for (int u = 0; u < 16; ++u) {
    const int32_t k = k0 + u * kstride;
    bits |= ((kron_drop_word(dr, seed_lo, seed_hi, row_words, k >> 5) >> (k & 31)) & 1u) << u;
  }
  return bits;
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
