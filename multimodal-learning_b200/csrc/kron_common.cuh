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

// post_fusion_dropout on the never-stored operand: the DROP flag of (row b, logical column k) is bit (b & 31) of a 32-bit
// word shared by the 32 rows [32g, 32g + 32) of that column (g = b >> 5).  A word is the bit-sliced comparison U < thresh
// of 32 independent 16-bit uniforms whose bit planes are counter-hash words: going from the lowest SET bit of thresh upwards,
//   lt = thresh_i ? (plane_i | lt) : (plane_i & lt),     plane_i = hash(((g * Kk + k) << 4) + i, seed)
// (planes below the lowest set bit cannot change the outcome and are never drawn), so a word costs 16 - ctz(thresh) hashes:
// 2 for p = 0.25, 1 for p = 0.5 -- against 16 for a hash per two elements.  K3 (thread = column, 16 consecutive rows) uses
// half a word directly; in K1/K2 (thread = row, 16 columns) the 32 lanes of a warp are the 32 rows of one group: lane j
// hashes the word of the warp's j-th column, the 16 words cross the warp through shared memory and every lane tests its
// own bit.  Restated in numpy in oracle/fusion_oracle.kron_dropout_mask.
struct KronDropout {
  uint32_t thresh;    // drop iff U16 < thresh;  thresh = round(p * 65536); 0 = no dropout
  float scale;        // 65536 / (65536 - thresh)
  uint32_t seed_lo, seed_hi;
  int64_t Kk;         // columns per row: the counter of (group g, column k) is g * Kk + k
  int32_t plane0;     // ctz(thresh): the first bit plane that matters
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
  d.Kk = Kk;
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

// DROP flags of column k for the 32 rows of group g (bit r = row 32 g + r); seed_lo/hi from kron_seed()
__device__ __forceinline__ uint32_t kron_drop_word(const KronDropout& dr, uint32_t seed_lo, uint32_t seed_hi, int64_t g, int32_t k) {
  const uint64_t c0 = static_cast<uint64_t>(g * dr.Kk + k) << 4;
  if (dr.plane0 >= 14) {           // one or two planes (p = 0.5, 0.25, 0.75): straight-line code, the two hashes independent
    const uint64_t ca = c0 + static_cast<uint32_t>(dr.plane0), cb = c0 + 15u;
    const uint32_t ra = kron_hash(static_cast<uint32_t>(ca), static_cast<uint32_t>(ca >> 32), seed_lo, seed_hi);
    const uint32_t rb = kron_hash(static_cast<uint32_t>(cb), static_cast<uint32_t>(cb >> 32), seed_lo, seed_hi);
    return dr.plane0 == 15 ? ra : ((dr.thresh & 0x8000u) ? (rb | ra) : (rb & ra));
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
  const uint32_t word = kron_drop_word(dr, dr.seed_lo, dr.seed_hi, b >> 5, k);
  return ((word >> (b & 31)) & 1u) ? 0.0f : dr.scale;
}

// K1 / K2: the 16 words of a warp's 16 columns k0 + j * kstride (j = 0..15) for the warp's 32 rows (group g), exchanged
// through the warp's 16-word shared slot `xw`.  Lane l hashes column l & 15; afterwards every lane holds all 16 words and
// tests bit `lane` of word j for its element j.  Must be called by all 32 lanes.
__device__ __forceinline__ void kron_drop_words16(const KronDropout& dr, uint32_t seed_lo, uint32_t seed_hi, int64_t g, int32_t k0,
                                                  int32_t kstride, uint32_t* xw, int lane, uint32_t (&w)[16]) {
  const uint32_t mine = kron_drop_word(dr, seed_lo, seed_hi, g, k0 + (lane & 15) * kstride);
  if (lane < 16) xw[lane] = mine;
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 t = *reinterpret_cast<const uint4*>(xw + 4 * q);
    w[4 * q + 0] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
  }
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
