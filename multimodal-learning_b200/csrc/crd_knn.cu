// Full-bank class-masked cosine KNN: the positives of the reference's CRD_criterion_v10 "neighbors" mode
// (`MIA 2023/stage2_unimodal_student/CL_utils/CRD_criterion_v10.py:69-80, :108-116`):
//     sim = class_mask[batch_label] * sklearn.cosine_similarity(memory[idx[:, 0]], memory)        # [B, n], on the CPU
//     neighbors, neighbor_sim = top num_pos of sort(sim, descending)
// for both banks, every step.  Here: one tcgen05 TF32 GEMM pass over the bank with a fused per-anchor candidate filter,
// then an exact fp32 re-score of the few candidates -- nothing of size [B, n] is ever stored.
//
//   K12a knn_invnorm_kernel   1 / |row| of every bank row                                (one streaming pass, n*D*4 bytes)
//   K12b knn_query_kernel     the anchors' own rows, L2-normalised -> qn [B, D] (+ label per anchor)
//   K12c knn_gemm_kernel      S = qn . bank^T by 128 x 128 tiles: qn tile in TENSOR MEMORY (A operand, TF32 round-to-nearest),
//                             bank rows by TMA (128B swizzle), accumulators double-buffered in TMEM; the epilogue warps
//                             (thread = anchor, half of the tile's columns) scale by 1/|row|, apply the class mask
//                             (other classes score exactly 0, as in the reference) and keep the 8 best (score, row) per
//                             thread in registers.  Grid = anchor tiles x bank slices.
//   K12d knn_merge_kernel     one warp per anchor: every candidate of every list is re-scored EXACTLY in fp32, the best P by
//                             (score descending, row ascending) are written, and the anchor is FLAGGED when the TF32 error
//                             bound cannot prove that no better row was filtered out (P-th exact score <= list minimum + eps)
//   K12e knn_exact_kernel     flagged anchors only: exact fp32 scan of the whole bank (rare: needs > 8 - P rows of one list
//                             within eps of the P-th neighbour)
// Ties (equal scores) resolve to the smaller row index; the reference's CUDA sort leaves them undefined.
#include "common.cuh"
#include "kron_tc_common.cuh"

namespace mml {
namespace {

using namespace tc;

constexpr int kKnnC = 8;                   // candidates kept per (bank slice, column half) list = the largest supported P
constexpr int kKnnTileN = 128;             // bank rows per accumulator tile
constexpr uint32_t kKnnBoxBytes = kKnnTileN * 32 * 4; // [128 rows x 32 d] fp32 = 16 KB
constexpr int kKnnSample = 16;             // the sampling pass visits every 16th bank tile
constexpr float kKnnEps = 2.0e-3f;         // >= |TF32 score - exact score|: A rounded (2^-11), B truncated (2^-10), |q| = |r| = 1

__device__ __forceinline__ uint32_t ord_bits(float x) {          // unsigned order == float order
  const uint32_t u = __float_as_uint(x);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float ord_value(uint32_t o) {
  return __uint_as_float(o ^ ((o >> 31) ? 0x80000000u : 0xFFFFFFFFu));
}
// larger key = better candidate: higher score, then smaller row
__device__ __forceinline__ unsigned long long knn_key(float s, uint32_t row) {
  return (static_cast<unsigned long long>(ord_bits(s)) << 32) | (0xFFFFFFFFu - row);
}

// rows == NULL: every row of the bank; else the `count` listed rows (ids outside [0, n) are skipped)
__global__ void knn_invnorm_kernel(const float* __restrict__ bank, int64_t n, int32_t D, const int64_t* __restrict__ rows,
                                   int64_t count, float* __restrict__ invn) {
  // 8 lanes per row, 128-bit streaming loads (the gather kernels' row layout)
  const int64_t item = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  int64_t row = item;
  if (rows != nullptr) row = item < count ? rows[item] : -1;
  if (row < 0 || row >= n) row = n;                      // nothing to do (whole 8-lane group agrees)
  float acc = 0.f;
  if (row < n) {
    const float* p = bank + row * D;
    for (int c = sub * 4; c < D; c += 32) {
      const float4 v = ldg_stream_f4(p + c);
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  }
  acc += __shfl_xor_sync(kFullMask, acc, 1);
  acc += __shfl_xor_sync(kFullMask, acc, 2);
  acc += __shfl_xor_sync(kFullMask, acc, 4);
  if (row < n && sub == 0) invn[row] = acc > 0.f ? 1.0f / sqrtf(acc) : 0.f;      // sklearn leaves all-zero rows at zero
}

// queries == NULL: the query of anchor b is row rows[b] of the bank (the reference's case); else queries[b] (the sharded bank:
// an anchor's own row may live on another rank)
__global__ void knn_query_kernel(const float* __restrict__ bank, int64_t n, int32_t D, const int64_t* __restrict__ rows,
                                 const float* __restrict__ queries, const int64_t* __restrict__ labels_in, int64_t B, int64_t Bpad, float* __restrict__ qn,
                                 float* __restrict__ qt, int32_t* __restrict__ qlab, uint32_t* err) {
  const int64_t b = blockIdx.x;
  const int lane = threadIdx.x;              // 32 threads
  if (b >= B) {                              // padding rows of the last anchor tile
    for (int c = lane; c < D; c += 32) { qn[b * D + c] = 0.f; qt[b * D + c] = 0.f; }
    if (lane == 0) qlab[b] = -1;
    return;
  }
  const float* src = queries != nullptr ? queries + b * D : nullptr;
  if (src == nullptr) {
    int64_t r = rows[b];
    if (r < 0 || r >= n) {
      if (lane == 0) flag_device_error(err, MML_DEVERR_CRD_INDEX);
      r = r < 0 ? 0 : n - 1;
    }
    src = bank + r * D;
  }
  float acc = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float v = src[c];
    acc += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullMask, acc, o);
  const float norm = sqrtf(acc);
  for (int c = lane; c < D; c += 32) {
    const float v = norm > 0.f ? src[c] / norm : 0.f;
    qn[b * D + c] = v;                                                               // exact: re-scoring
    qt[b * D + c] = __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xFFFFE000u);   // nearest TF32: the descriptor-fed A operand
  }
  if (lane == 0) qlab[b] = static_cast<int32_t>(labels_in[b]);
}

// D[tmem] (+)= A[smem desc] * B[smem desc],  kind::tf32, M = 128, K = 8 (both operands K-major, 128B-swizzled boxes)
__device__ __forceinline__ void tc_mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

struct KnnArgs {
  const float* qn;          // [Bpad][D] normalised queries (zero rows past B)
  const int32_t* qlab;      // [Bpad]
  const float* invn;        // [n]
  const int32_t* labels;    // [n]
  unsigned long long* part; // [slices * 2][Bpad][kKnnC]
  float* rej;               // [slices * 2][Bpad] the final threshold of every list: rows it left out score <= this (TF32)
  const float* thr_init;    // [Bpad] or NULL: a proven lower bound of the anchor's 8-th best score (from the sampling pass)
  int32_t tile_mul;         // this pass visits bank tiles t * tile_mul (sampling pass: every 16th tile)
  int32_t max_only;         // sampling pass: every list keeps just its best score (no insertion path at all)
  int64_t n, Bpad;
  int32_t D, nbox;          // nbox = D / 32
  int32_t tiles_total, tiles_per_slice;
  int32_t stages, tmem_cols;
  uint32_t idesc;
};

// kAT = anchor tiles per CTA.  kAT = 1: the query tile is the A operand in TENSOR MEMORY.  kAT = 2: every bank box feeds TWO
// 128-anchor MMAs (queries in shared memory, both operands by descriptor; all 512 TMEM columns are accumulators), which
// halves the L2 -> SM stream per flop: one anchor tile per CTA pulls 64 KB per 1418 tensor cycles = 46 B/cycle/SM, above
// what L2 delivers to every SM at once (r3c: 48 % tensor pipe).  16 epilogue warps instead of 8.
// kTab: at most 3 classes -> the class mask and 1/|row| become ONE multiplier per (class, bank row), tabulated per tile
// (cm[c][j] = label[j] == c ? 1/|row j| : 0); a thread reads its anchor's class row: one FFMA per score instead of
// multiply + compare + select.
template <int kAT, bool kTab>
__global__ void __launch_bounds__((8 * kAT + 2) * 32, 1) knn_gemm_kernel(const __grid_constant__ CUtensorMap tmap_bank,
                                                                         const __grid_constant__ CUtensorMap tmap_q, const KnnArgs a) {
  constexpr int EW = 8 * kAT;                    // epilogue warps; warp EW = TMA producer, warp EW + 1 = MMA issuer
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sm_a = smem;                                                              // kAT == 2: [2][nbox][16 KB] query boxes
  uint8_t* sm_b = smem + (kAT == 2 ? static_cast<size_t>(2 * a.nbox) * kKnnBoxBytes : 0);      // [stages][16 KB]
  constexpr int kRowS = 68;                      // class-row stride: 64 + 4 floats, so lanes of different classes hit different banks
  constexpr int kSide = kTab ? 3 * kRowS + 4 : 128;   // floats per (warp, buffer): 3 skewed class rows, or 64 x 1/|row| + 64 labels
  float* sm_side = reinterpret_cast<float*>(sm_b + static_cast<size_t>(a.stages) * kKnnBoxBytes);   // [EW warps][2][kSide]
  float* sm_zero = sm_side + EW * 2 * kSide;     // 64 zeros: the class row of an anchor whose label matches no class
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_zero + 64);
  uint64_t* bar_full = bars;                     // [stages] bank box landed
  uint64_t* bar_empty = bars + a.stages;         // [stages] MMAs reading the box done
  uint64_t* bar_acc_full = bars + 2 * a.stages;  // [2] accumulator tile(s) complete
  uint64_t* bar_acc_empty = bar_acc_full + 2;    // [2] epilogue has read the tile(s) (EW warp arrivals)
  uint64_t* bar_a = bar_acc_empty + 2;           // kAT == 2: query boxes landed
  uint32_t* sm_tmem = reinterpret_cast<uint32_t*>(bar_a + 1);

  const int warp = warp_idx_sync();
  const int lane = threadIdx.x & 31;
  const int64_t b0 = static_cast<int64_t>(blockIdx.x) * kTileM * kAT;
  const int t_begin = blockIdx.y * a.tiles_per_slice;
  const int t_end = min(a.tiles_total, t_begin + a.tiles_per_slice);

  if (warp == EW && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_bank)) : "memory");
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_acc_full[i], 1);
      mbar_init(&bar_acc_empty[i], EW);
    }
    mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == EW + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm_tmem)), "r"(a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < 64) sm_zero[threadIdx.x] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm_tmem;
  const uint32_t tmem_acc = tmem_base;                   // [kAT][2] x 128 accumulator columns
  const uint32_t tmem_a = tmem_base + 2 * kKnnTileN;     // kAT == 1: query tile, D columns

  // ---- kAT == 1: A operand = this CTA's 128 normalised queries -> TMEM, once ----
  if (kAT == 1 && warp < EW) {
    const int row = threadIdx.x & (kTileM - 1);
    const int half = threadIdx.x >> 7;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float* q = a.qn + (b0 + row) * a.D;
    const int c_lo = half == 0 ? 0 : (a.D / 32 + 1) / 2 * 32;     // warps 0-3: the first half of the 32-column groups
    const int c_hi = half == 0 ? (a.D / 32 + 1) / 2 * 32 : a.D;
    for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
      uint32_t r[16];
#pragma unroll
      for (int u = 0; u < 16; u += 4) {
        const float4 v = *reinterpret_cast<const float4*>(q + c0 + u);
        r[u + 0] = __float_as_uint(v.x) + 0x1000u;                 // round to nearest onto the TF32 grid
        r[u + 1] = __float_as_uint(v.y) + 0x1000u;
        r[u + 2] = __float_as_uint(v.z) + 0x1000u;
        r[u + 3] = __float_as_uint(v.w) + 0x1000u;
      }
      tc_st_32x32b_x16(tmem_a + lane_base + c0, r);
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == EW) {
    // ===== TMA producer: D/32 boxes of [128 bank rows x 32 d] per tile =====
    if (kAT == 2) {
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(bar_a, static_cast<uint32_t>(2 * a.nbox) * kKnnBoxBytes);
        for (int at = 0; at < 2; ++at)
          for (int j = 0; j < a.nbox; ++j)
            tma_load_2d(sm_a + static_cast<size_t>(at * a.nbox + j) * kKnnBoxBytes, &tmap_q, j * 32,
                        static_cast<int32_t>(b0) + at * kTileM, bar_a);
      }
      __syncwarp();
    }
    int s = 0;
    uint32_t ph = 0;
    for (int t = t_begin; t < t_end; ++t)
      for (int j = 0; j < a.nbox; ++j) {
        mbar_wait(&bar_empty[s], ph ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&bar_full[s], kKnnBoxBytes);
          tma_load_2d(sm_b + static_cast<size_t>(s) * kKnnBoxBytes, &tmap_bank, j * 32, t * a.tile_mul * kKnnTileN, &bar_full[s]);
        }
        __syncwarp();
        if (++s == a.stages) { s = 0; ph ^= 1; }
      }
  } else if (warp == EW + 1) {
    // ===== MMA issuer =====
    int s = 0;
    uint32_t ph = 0;
    const uint32_t b_base = smem_u32(sm_b);
    const uint32_t a_base = smem_u32(sm_a);
    if (kAT == 2) {
      mbar_wait(bar_a, 0);
      tc_fence_after();
    }
    for (int t = t_begin; t < t_end; ++t) {
      const int it = t - t_begin;
      const int buf = it & 1;
      mbar_wait(&bar_acc_empty[buf], ((it >> 1) & 1) ^ 1);            // epilogue has drained this accumulator
      tc_fence_after();
      for (int j = 0; j < a.nbox; ++j) {
        mbar_wait(&bar_full[s], ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t b_desc = umma_desc_k_sw128(b_base + static_cast<uint32_t>(s) * kKnnBoxBytes);
          if (kAT == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc_mma_tf32_ts(tmem_acc + buf * kKnnTileN, tmem_a + j * 32 + k * 8, b_desc + 2 * k, a.idesc, (j > 0 || k > 0) ? 1u : 0u);
          } else {
#pragma unroll
            for (int at = 0; at < 2; ++at) {
              const uint64_t a_desc = umma_desc_k_sw128(a_base + static_cast<uint32_t>(at * a.nbox + j) * kKnnBoxBytes);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                tc_mma_tf32_ss(tmem_acc + (at * 2 + buf) * kKnnTileN, a_desc + 2 * k, b_desc + 2 * k, a.idesc, (j > 0 || k > 0) ? 1u : 0u);
            }
          }
          tc_commit(&bar_empty[s]);
          if (j + 1 == a.nbox) tc_commit(&bar_acc_full[buf]);
        }
        __syncwarp();
        if (++s == a.stages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===== epilogue: thread = (anchor row, 64-column half of every tile) =====
    const int at = kAT == 2 ? (warp >> 2) & 1 : 0;            // which of the CTA's anchor tiles
    const int row = at * kTileM + (warp & 3) * 32 + lane;     // anchor within the CTA
    const int half = kAT == 2 ? warp >> 3 : warp >> 2;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int32_t mylab = a.qlab[b0 + row];
    unsigned long long keys[kKnnC];
#pragma unroll
    for (int i = 0; i < kKnnC; ++i) keys[i] = 0ull;       // 0 = empty slot: below every real key
    float best_only = -INFINITY;                          // max_only: the list's best score
    const float floor_thr = a.thr_init != nullptr ? a.thr_init[b0 + row] : -INFINITY;
    float thr = floor_thr;                                // rows scoring <= thr are left out: the list's worst kept score once it
                                                          // is full, never below the bound the sampling pass proved

    auto insert = [&](float s, uint32_t j) {
      unsigned long long mk = keys[0];
      int mp = 0;
#pragma unroll
      for (int i = 1; i < kKnnC; ++i)
        if (keys[i] < mk) { mk = keys[i]; mp = i; }
      const unsigned long long nk = knn_key(s, j);
      if (nk > mk) {
#pragma unroll
        for (int i = 0; i < kKnnC; ++i)
          if (i == mp) keys[i] = nk;
        mk = keys[0];
#pragma unroll
        for (int i = 1; i < kKnnC; ++i) mk = keys[i] < mk ? keys[i] : mk;
        thr = mk == 0ull ? floor_thr : fmaxf(floor_thr, ord_value(static_cast<uint32_t>(mk >> 32)));
      }
    };

    // side data of a tile (1/|row| and label of the warp's 64 bank rows): loaded one tile ahead, two + two values per lane,
    // into a warp-private shared slot -- no CTA-wide barrier couples the epilogue warps
    float* my_side = sm_side + warp * (2 * kSide);
    auto load_side = [&](int t, uint32_t (&v)[4]) {
      v[0] = v[1] = 0u;
      v[2] = v[3] = 0xFFFFFFFEu;                            // -2: no such row
      if (t >= t_end) return;
      const int64_t j = static_cast<int64_t>(t) * a.tile_mul * kKnnTileN + half * 64 + lane;
      if (j < a.n) { v[0] = __float_as_uint(__ldg(a.invn + j)); v[2] = static_cast<uint32_t>(__ldg(a.labels + j)); }
      if (j + 32 < a.n) { v[1] = __float_as_uint(__ldg(a.invn + j + 32)); v[3] = static_cast<uint32_t>(__ldg(a.labels + j + 32)); }
    };
    uint32_t side[4];
    load_side(t_begin, side);
    for (int t = t_begin; t < t_end; ++t) {
      const int it = t - t_begin;
      const int buf = it & 1;
      if (kTab) {
        float* d = my_side + buf * kSide;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          d[c * kRowS + lane] = static_cast<int32_t>(side[2]) == c ? __uint_as_float(side[0]) : 0.f;
          d[c * kRowS + lane + 32] = static_cast<int32_t>(side[3]) == c ? __uint_as_float(side[1]) : 0.f;
        }
      } else {
        uint32_t* d = reinterpret_cast<uint32_t*>(my_side + buf * kSide);
        d[lane] = side[0]; d[lane + 32] = side[1]; d[64 + lane] = side[2]; d[96 + lane] = side[3];
      }
      load_side(t + 1, side);
      __syncwarp();
      mbar_wait(&bar_acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_acc + lane_base + (at * 2 + buf) * kKnnTileN + half * 64;
      const float* iv = kTab ? ((mylab >= 0 && mylab < 3) ? my_side + buf * kSide + mylab * kRowS : sm_zero) : my_side + buf * kSide;
      const int32_t* lb = reinterpret_cast<const int32_t*>(my_side + buf * kSide + 64);      // !kTab only
      const uint32_t j0 = static_cast<uint32_t>(t) * a.tile_mul * kKnnTileN + half * 64;
      const bool whole = (static_cast<int64_t>(t) * a.tile_mul + 1) * kKnnTileN <= a.n;      // uniform: only the last tile can hold rows >= n
      uint32_t accA[16], accB[16];
      // The filter's slow path (a score above the thread's threshold) exists ONCE per accumulator register set and walks
      // the qualifying elements with a run-time loop: 64 inlined copies of the insertion made the kernel 13.6 K
      // instructions and instruction-cache bound (r2y: 8.9 ms, "no_inst" stalls everywhere).
#define MML_KNN_CHUNK(Q, ACC, NEXT_LD)                                                                   \
      {                                                                                                  \
        tc_wait_ld();                                                                                    \
        NEXT_LD;                                                                                         \
        float sc[16];                                                                                    \
        float m = -INFINITY;                                                                             \
        _Pragma("unroll") for (int u = 0; u < 16; u += 4) {                                              \
          const float4 i4 = *reinterpret_cast<const float4*>(iv + (Q) * 16 + u);                         \
          if (kTab) {                                                                                    \
            sc[u + 0] = fmaf(__uint_as_float(ACC[u + 0]), i4.x, 0.0f);                                   \
            sc[u + 1] = fmaf(__uint_as_float(ACC[u + 1]), i4.y, 0.0f);                                   \
            sc[u + 2] = fmaf(__uint_as_float(ACC[u + 2]), i4.z, 0.0f);                                   \
            sc[u + 3] = fmaf(__uint_as_float(ACC[u + 3]), i4.w, 0.0f);                                   \
            if (!whole) {                                                                                \
              const int64_t jj = static_cast<int64_t>(j0) + (Q) * 16 + u;                                \
              sc[u + 0] = jj + 0 < a.n ? sc[u + 0] : -INFINITY;                                          \
              sc[u + 1] = jj + 1 < a.n ? sc[u + 1] : -INFINITY;                                          \
              sc[u + 2] = jj + 2 < a.n ? sc[u + 2] : -INFINITY;                                          \
              sc[u + 3] = jj + 3 < a.n ? sc[u + 3] : -INFINITY;                                          \
            }                                                                                            \
          } else {                                                                                       \
            const int4 l4 = *reinterpret_cast<const int4*>(lb + (Q) * 16 + u);                           \
            sc[u + 0] = l4.x == mylab ? __uint_as_float(ACC[u + 0]) * i4.x + 0.0f : 0.0f;                \
            sc[u + 1] = l4.y == mylab ? __uint_as_float(ACC[u + 1]) * i4.y + 0.0f : 0.0f;                \
            sc[u + 2] = l4.z == mylab ? __uint_as_float(ACC[u + 2]) * i4.z + 0.0f : 0.0f;                \
            sc[u + 3] = l4.w == mylab ? __uint_as_float(ACC[u + 3]) * i4.w + 0.0f : 0.0f;                \
            if (!whole) {                                                                                \
              sc[u + 0] = l4.x == -2 ? -INFINITY : sc[u + 0];                                            \
              sc[u + 1] = l4.y == -2 ? -INFINITY : sc[u + 1];                                            \
              sc[u + 2] = l4.z == -2 ? -INFINITY : sc[u + 2];                                            \
              sc[u + 3] = l4.w == -2 ? -INFINITY : sc[u + 3];                                            \
            }                                                                                            \
          }                                                                                              \
          m = fmaxf(fmaxf(m, fmaxf(sc[u + 0], sc[u + 1])), fmaxf(sc[u + 2], sc[u + 3]));                 \
        }                                                                                                \
        if (a.max_only) {                                                                                \
          best_only = fmaxf(best_only, m);                                                               \
        } else if (m > thr) {                                                                            \
          uint32_t cand = 0u;                                                                            \
          _Pragma("unroll") for (int u = 0; u < 16; ++u) cand |= sc[u] > thr ? (1u << u) : 0u;           \
          while (cand != 0u) {                                                                           \
            const int u = __ffs(cand) - 1;                                                               \
            cand &= cand - 1u;                                                                           \
            float s = sc[0];                                                                             \
            _Pragma("unroll") for (int w = 1; w < 16; ++w) s = u == w ? sc[w] : s;                       \
            if (s > thr) insert(s, j0 + (Q) * 16 + u);                                                   \
          }                                                                                              \
        }                                                                                                \
      }
      tc_ld_32x32b_x16(t_addr, accA);
#pragma unroll 1
      for (int q2 = 0; q2 < 2; ++q2) {
        MML_KNN_CHUNK(2 * q2, accA, tc_ld_32x32b_x16(t_addr + (2 * q2 + 1) * 16, accB))
        MML_KNN_CHUNK(2 * q2 + 1, accB, if (q2 == 0) tc_ld_32x32b_x16(t_addr + 32, accA))
      }
#undef MML_KNN_CHUNK
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_acc_empty[buf]);
    }
    unsigned long long* dst = a.part + ((static_cast<int64_t>(blockIdx.y) * 2 + half) * a.Bpad + (b0 + row)) * kKnnC;
    if (a.max_only && best_only > -INFINITY)              // one real score per list; the list number keeps the keys distinct
      keys[0] = knn_key(best_only, static_cast<uint32_t>(blockIdx.y * 2 + half));
#pragma unroll
    for (int i = 0; i < kKnnC; ++i) dst[i] = keys[i];
    a.rej[(static_cast<int64_t>(blockIdx.y) * 2 + half) * a.Bpad + (b0 + row)] = thr;
    tc_fence_before();
  }
  __syncthreads();
  if (warp == EW + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

// Sampling pass -> per-anchor floor of the full pass.  With many lists per anchor (the usual case) every list of the sampling
// pass keeps only its MAXIMUM (no insertion path at all): the 8-th largest of those maxima is the score of 8 distinct rows,
// hence a lower bound of the anchor's 8-th best score -- about as tight as merging full lists (0.28 vs 0.30 for 1M random
// rows) at a third of the cost.  The 8-th best TF32 score among every 16th bank tile is a lower bound of
// the 8-th best over the whole bank, so a row scoring below (bound - 3 eps) in TF32 cannot be among the anchor's 8 best in
// exact arithmetic.  Starting every list of the full pass at that floor makes the filter's slow path rare from the first
// tile on (without it each of the ~36 short lists of an anchor climbs from -inf: r2z, 40 % of the warp-chunks took it).
__global__ void __launch_bounds__(128) knn_threshold_kernel(const unsigned long long* __restrict__ part, int32_t nlists, int64_t B,
                                                            int64_t Bpad, float* __restrict__ thr_init) {
  const int lane = threadIdx.x & 31;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);      // one warp per anchor
  if (b >= Bpad) return;
  if (b >= B) {
    if (lane == 0) thr_init[b] = -INFINITY;
    return;
  }
  constexpr int kMaxMine = kKnnC;                         // each lane keeps the 8 best of its share: the union holds the global 8 best
  unsigned long long mine[kMaxMine];
#pragma unroll
  for (int i = 0; i < kMaxMine; ++i) mine[i] = 0ull;
  const int total = nlists * kKnnC;
  for (int e = lane; e < total; e += 32) {
    const unsigned long long k = part[(static_cast<int64_t>(e >> 3) * Bpad + b) * kKnnC + (e & 7)];
    unsigned long long mk = mine[0];
    int mp = 0;
#pragma unroll
    for (int q = 1; q < kMaxMine; ++q)
      if (mine[q] < mk) { mk = mine[q]; mp = q; }
    if (k > mk) {
#pragma unroll
      for (int q = 0; q < kMaxMine; ++q)
        if (q == mp) mine[q] = k;
    }
  }
  unsigned long long top = 0ull;
  for (int r = 0; r < kKnnC; ++r) {                       // the 8-th largest key of the union
    unsigned long long m = mine[0];
#pragma unroll
    for (int i = 1; i < kMaxMine; ++i) m = mine[i] > m ? mine[i] : m;
    top = m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(kFullMask, top, o);
      top = other > top ? other : top;
    }
    if (top == 0ull) break;                               // fewer than 8 rows seen: no bound
#pragma unroll
    for (int i = 0; i < kMaxMine; ++i)
      if (mine[i] == top) mine[i] = 0ull;                 // keys are unique: exactly one lane holds it
  }
  if (lane == 0) {
    const float v = top == 0ull ? -INFINITY : ord_value(static_cast<uint32_t>(top >> 32));
    thr_init[b] = (v == -INFINITY || !(v == v)) ? -INFINITY : v - 3.0f * kKnnEps;
  }
}

constexpr int kMergeWarps = 4;

__global__ void __launch_bounds__(kMergeWarps * 32) knn_merge_kernel(
    const float* __restrict__ bank, int32_t D, const float* __restrict__ invn, const int32_t* __restrict__ labels,
    const float* __restrict__ qn, const int32_t* __restrict__ qlab, const unsigned long long* __restrict__ part,
    const float* __restrict__ rej, int32_t nlists, int64_t B, int64_t Bpad, int32_t P, int64_t* __restrict__ out_idx,
    float* __restrict__ out_sim, int32_t* __restrict__ flags, int32_t force_flag) {
  extern __shared__ float sm_qm[];                       // [kMergeWarps][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t b = static_cast<int64_t>(blockIdx.x) * kMergeWarps + warp;
  if (b >= B) return;
  float* sq = sm_qm + warp * D;
  for (int c = lane; c < D; c += 32) sq[c] = qn[b * D + c];
  __syncwarp();
  const int32_t mylab = qlab[b];
  float mmax = -INFINITY;                                // max over the lists of the list's final threshold
  for (int l = lane; l < nlists; l += 32) mmax = fmaxf(mmax, rej[static_cast<int64_t>(l) * Bpad + b]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mmax = fmaxf(mmax, __shfl_xor_sync(kFullMask, mmax, o));
  // Every candidate is re-scored by the WHOLE warp (one coalesced pass over its row, butterfly sum), eight candidates in
  // flight at a time; the running best-8 is kept identically in all lanes.
  unsigned long long best[kKnnC];
#pragma unroll
  for (int i = 0; i < kKnnC; ++i) best[i] = 0ull;
  unsigned long long worst = 0ull;
  const int total = nlists * kKnnC;
  for (int base = 0; base < total; base += 32) {
    const int e = base + lane;
    const unsigned long long k = e < total ? part[(static_cast<int64_t>(e >> 3) * Bpad + b) * kKnnC + (e & 7)] : 0ull;
    uint32_t mask = __ballot_sync(kFullMask, k != 0ull);
    while (mask != 0u) {
      uint32_t rowj[8];
      float acc[8], inv[8];
      int32_t lab[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        acc[g] = 0.f;
        rowj[g] = 0xFFFFFFFFu;
        if (mask != 0u) {
          const int src = __ffs(mask) - 1;
          mask &= mask - 1u;
          rowj[g] = 0xFFFFFFFFu - static_cast<uint32_t>(__shfl_sync(kFullMask, static_cast<uint32_t>(k), src));
        }
      }
      // all loads of the round are independent of each other: label, 1/|row| and the row itself, eight candidates at once
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const uint32_t r = rowj[g] == 0xFFFFFFFFu ? 0u : rowj[g];
        lab[g] = __ldg(labels + r);
        inv[g] = __ldg(invn + r);
      }
      for (int c = lane * 4; c < D; c += 128) {
        float4 v[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint32_t r = rowj[g] == 0xFFFFFFFFu ? 0u : rowj[g];
          v[g] = *reinterpret_cast<const float4*>(bank + static_cast<int64_t>(r) * D + c);
        }
        const float4 q = *reinterpret_cast<const float4*>(sq + c);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          acc[g] = fmaf(v[g].x, q.x, acc[g]);
          acc[g] = fmaf(v[g].y, q.y, acc[g]);
          acc[g] = fmaf(v[g].z, q.z, acc[g]);
          acc[g] = fmaf(v[g].w, q.w, acc[g]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int g = 0; g < 8; ++g) acc[g] += __shfl_xor_sync(kFullMask, acc[g], o);
      }
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        if (rowj[g] == 0xFFFFFFFFu) continue;             // warp-uniform
        const float sc = lab[g] == mylab ? acc[g] * inv[g] + 0.0f : 0.0f;
        const unsigned long long nk = knn_key(sc, rowj[g]);
        if (nk > worst) {
          unsigned long long mk = best[0];
          int mp = 0;
#pragma unroll
          for (int q = 1; q < kKnnC; ++q)
            if (best[q] < mk) { mk = best[q]; mp = q; }
#pragma unroll
          for (int q = 0; q < kKnnC; ++q)
            if (q == mp) best[q] = nk;
          worst = best[0];
#pragma unroll
          for (int q = 1; q < kKnnC; ++q) worst = best[q] < worst ? best[q] : worst;
        }
      }
    }
  }
  // the best P, in order
  float vP = -INFINITY;
  for (int p = 0; p < P; ++p) {
    unsigned long long top = best[0];
#pragma unroll
    for (int q = 1; q < kKnnC; ++q) top = best[q] > top ? best[q] : top;
#pragma unroll
    for (int q = 0; q < kKnnC; ++q)
      if (best[q] == top) best[q] = 0ull;
    const bool have = top != 0ull;
    vP = have ? ord_value(static_cast<uint32_t>(top >> 32)) : -INFINITY;
    if (lane == 0) {
      out_idx[b * P + p] = have ? static_cast<int64_t>(0xFFFFFFFFu - static_cast<uint32_t>(top)) : -1;
      out_sim[b * P + p] = have ? vP : 0.f;
    }
  }
  if (lane == 0) flags[b] = (force_flag || !(vP > mmax + kKnnEps)) ? 1 : 0;   // not provable: re-do this anchor exactly
}

// ---- exact scan of the flagged anchors, 32 anchors at a time -------------------------------------------------------------
// knn_compact_kernel lists the flagged anchors; knn_exact_group_kernel gives every lane of a warp ONE of 32 flagged anchors
// (its normalised query in registers) and streams bank rows through shared memory: a row is read once per warp as broadcast
// 128-bit loads and costs D FMAs for 32 anchors (the per-anchor scan below reads the whole bank per anchor: 85 ms for 1024
// flagged anchors over 1M rows; this one is bound by the fp32 FMA rate: ~5 ms).  Grid = (anchor groups, bank slices); groups
// past the flagged count exit at once.  knn_pick_kernel merges the slices.
__global__ void __launch_bounds__(1024) knn_compact_kernel(const int32_t* __restrict__ flags, int64_t B, int32_t* __restrict__ list,
                                                           int32_t* __restrict__ count) {
  __shared__ int warp_cnt[32];
  __shared__ int running;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) running = 0;
  __syncthreads();
  for (int64_t start = 0; start < B; start += 1024) {
    const int64_t b = start + tid;
    const bool f = b < B && flags[b] != 0;
    const uint32_t bal = __ballot_sync(kFullMask, f);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = running;
    for (int w = 0; w < warp; ++w) before += warp_cnt[w];
    if (f) list[before + __popc(bal & ((1u << lane) - 1u))] = static_cast<int32_t>(b);
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += warp_cnt[w];
      running += tot;
    }
    __syncthreads();
  }
  if (tid == 0) *count = running;
}

constexpr int kEgWarps = 8;
constexpr int kEgRows = 64;                 // bank rows per shared-memory tile

template <int kD>
__global__ void __launch_bounds__(kEgWarps * 32, 1) knn_exact_group_kernel(
    const float* __restrict__ bank, int64_t n, const float* __restrict__ invn, const int32_t* __restrict__ labels,
    const float* __restrict__ qn, const int32_t* __restrict__ qlab, const int32_t* __restrict__ list,
    const int32_t* __restrict__ count, int32_t nslices, int64_t rows_per_slice, int64_t Bpad, unsigned long long* __restrict__ partE) {
  const int cnt = *count;
  extern __shared__ __align__(16) float sm_eg[];            // [2][kEgRows][kD] rows | [2][kEgRows] 1/|row| | [2][kEgRows] labels
  float* sm_rows = sm_eg;
  float* sm_in = sm_eg + 2 * kEgRows * kD;
  int32_t* sm_lb = reinterpret_cast<int32_t*>(sm_in + 2 * kEgRows);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // work items = (group of 32 flagged anchors, bank slice); no flagged anchor -> no item -> the CTA retires at once
  for (int item = blockIdx.x; item < (cnt + 31) / 32 * nslices; item += gridDim.x) {
  const int group = item / nslices, slice = item % nslices;
  const int slot = group * 32 + lane;                       // position in the flagged list
  const int anchor = slot < cnt ? list[slot] : -1;
  float q[kD];
#pragma unroll
  for (int c = 0; c < kD; c += 4) {
    const float4 v = anchor >= 0 ? *reinterpret_cast<const float4*>(qn + static_cast<int64_t>(anchor) * kD + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    q[c] = v.x; q[c + 1] = v.y; q[c + 2] = v.z; q[c + 3] = v.w;
  }
  const int32_t mylab = anchor >= 0 ? qlab[anchor] : -3;
  unsigned long long best[kKnnC];
#pragma unroll
  for (int i = 0; i < kKnnC; ++i) best[i] = 0ull;
  unsigned long long worst = 0ull;
  const int64_t r_begin = static_cast<int64_t>(slice) * rows_per_slice;
  const int64_t r_end = r_begin + rows_per_slice < n ? r_begin + rows_per_slice : n;
  const int ntiles = r_begin < r_end ? static_cast<int>((r_end - r_begin + kEgRows - 1) / kEgRows) : 0;
  auto issue = [&](int t) {                                 // tile t -> buffer t & 1 (16-byte cp.async, coalesced)
    if (t < ntiles) {
      const int64_t r0 = r_begin + static_cast<int64_t>(t) * kEgRows;
      float* dst = sm_rows + (t & 1) * kEgRows * kD;
      for (int i = tid; i < kEgRows * kD / 4; i += kEgWarps * 32) {
        const int64_t r = r0 + (i * 4) / kD;
        const float* src = bank + (r < r_end ? r : r_end - 1) * kD + (i * 4) % kD;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + i * 4)), "l"(src) : "memory");
      }
      if (tid < kEgRows) {
        const int64_t r = r0 + tid;
        sm_in[(t & 1) * kEgRows + tid] = r < r_end ? __ldg(invn + r) : 0.f;
        sm_lb[(t & 1) * kEgRows + tid] = r < r_end ? __ldg(labels + r) : -2;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0);
  for (int t = 0; t < ntiles; ++t) {
    issue(t + 1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const float* tile = sm_rows + (t & 1) * kEgRows * kD;
    const int64_t r0 = r_begin + static_cast<int64_t>(t) * kEgRows;
#pragma unroll 1
    for (int rr = 0; rr < kEgRows / kEgWarps; ++rr) {
      const int row = warp * (kEgRows / kEgWarps) + rr;
      const int32_t lab = sm_lb[(t & 1) * kEgRows + row];
      if (lab == -2) continue;                              // past the slice (uniform)
      const float* rp = tile + row * kD;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int c = 0; c < kD; c += 4) {
        const float4 v = *reinterpret_cast<const float4*>(rp + c);      // broadcast: every lane reads the same row
        a0 = fmaf(v.x, q[c], a0);
        a1 = fmaf(v.y, q[c + 1], a1);
        a2 = fmaf(v.z, q[c + 2], a2);
        a3 = fmaf(v.w, q[c + 3], a3);
      }
      const float sc = lab == mylab ? ((a0 + a1) + (a2 + a3)) * sm_in[(t & 1) * kEgRows + row] + 0.0f : 0.0f;
      const unsigned long long nk = knn_key(sc, static_cast<uint32_t>(r0 + row));
      if (nk > worst) {
        unsigned long long mk = best[0];
        int mp = 0;
#pragma unroll
        for (int i = 1; i < kKnnC; ++i)
          if (best[i] < mk) { mk = best[i]; mp = i; }
#pragma unroll
        for (int i = 0; i < kKnnC; ++i)
          if (i == mp) best[i] = nk;
        worst = best[0];
#pragma unroll
        for (int i = 1; i < kKnnC; ++i) worst = best[i] < worst ? best[i] : worst;
      }
    }
    __syncthreads();                                        // the buffer is free for tile t + 2
  }
  // merge the 8 warps' lists per anchor (shared memory: the tile buffers are idle now), one thread per anchor
  unsigned long long* sm_k = reinterpret_cast<unsigned long long*>(sm_eg);      // [kEgWarps][32][kKnnC]
#pragma unroll
  for (int i = 0; i < kKnnC; ++i) sm_k[(warp * 32 + lane) * kKnnC + i] = best[i];
  __syncthreads();
  if (warp == 0 && anchor >= 0) {
#pragma unroll
    for (int i = 0; i < kKnnC; ++i) best[i] = 0ull;
    worst = 0ull;
    for (int w = 0; w < kEgWarps; ++w)
      for (int i = 0; i < kKnnC; ++i) {
        const unsigned long long nk = sm_k[(w * 32 + lane) * kKnnC + i];
        if (nk > worst) {
          unsigned long long mk = best[0];
          int mp = 0;
#pragma unroll
          for (int q2 = 1; q2 < kKnnC; ++q2)
            if (best[q2] < mk) { mk = best[q2]; mp = q2; }
#pragma unroll
          for (int q2 = 0; q2 < kKnnC; ++q2)
            if (q2 == mp) best[q2] = nk;
          worst = best[0];
#pragma unroll
          for (int q2 = 1; q2 < kKnnC; ++q2) worst = best[q2] < worst ? best[q2] : worst;
        }
      }
    unsigned long long* dst = partE + (static_cast<int64_t>(slice) * Bpad + slot) * kKnnC;
#pragma unroll
    for (int i = 0; i < kKnnC; ++i) dst[i] = best[i];
  }
  __syncthreads();                                          // shared memory is reused by the next item
  }
}

// one warp per flagged anchor: the best P of its slices' exact lists
__global__ void __launch_bounds__(128) knn_pick_kernel(const unsigned long long* __restrict__ partE, int32_t nslices, int64_t Bpad,
                                                       const int32_t* __restrict__ list, const int32_t* __restrict__ count, int32_t P,
                                                       int64_t* __restrict__ out_idx, float* __restrict__ out_sim) {
  const int lane = threadIdx.x & 31;
  const int slot = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (slot >= *count) return;
  const int64_t b = list[slot];
  unsigned long long mine[kKnnC];
#pragma unroll
  for (int i = 0; i < kKnnC; ++i) mine[i] = 0ull;
  const int total = nslices * kKnnC;
  for (int e = lane; e < total; e += 32) {
    const unsigned long long k = partE[(static_cast<int64_t>(e >> 3) * Bpad + slot) * kKnnC + (e & 7)];
    unsigned long long mk = mine[0];
    int mp = 0;
#pragma unroll
    for (int q = 1; q < kKnnC; ++q)
      if (mine[q] < mk) { mk = mine[q]; mp = q; }
    if (k > mk) {
#pragma unroll
      for (int q = 0; q < kKnnC; ++q)
        if (q == mp) mine[q] = k;
    }
  }
  for (int p = 0; p < P; ++p) {
    unsigned long long top = mine[0];
#pragma unroll
    for (int i = 1; i < kKnnC; ++i) top = mine[i] > top ? mine[i] : top;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(kFullMask, top, o);
      top = other > top ? other : top;
    }
#pragma unroll
    for (int i = 0; i < kKnnC; ++i)
      if (mine[i] == top && top != 0ull) mine[i] = 0ull;
    if (lane == 0) {
      const bool have = top != 0ull;
      out_idx[b * P + p] = have ? static_cast<int64_t>(0xFFFFFFFFu - static_cast<uint32_t>(top)) : -1;
      out_sim[b * P + p] = have ? ord_value(static_cast<uint32_t>(top >> 32)) : 0.f;
    }
  }
}

template <int kD>
int launch_exact_group(const float* bank, int64_t n, const float* invn, const int32_t* labels, const float* qn, const int32_t* qlab,
                       const int32_t* list, const int32_t* count, int32_t groups, int32_t slices, int64_t rows_per_slice, int64_t Bpad,
                       unsigned long long* partE, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(2) * kEgRows * kD * sizeof(float) + 2 * kEgRows * 8;
  const size_t need = smem > kEgWarps * 32 * kKnnC * sizeof(unsigned long long) ? smem : kEgWarps * 32 * kKnnC * sizeof(unsigned long long);
  MML_CUDA(cudaFuncSetAttribute(knn_exact_group_kernel<kD>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(need)));
  const int64_t items = static_cast<int64_t>(groups) * slices;
  knn_exact_group_kernel<kD><<<static_cast<unsigned>(items < 296 ? items : 296), kEgWarps * 32, need, st>>>(
      bank, n, invn, labels, qn, qlab, list, count, slices, rows_per_slice, Bpad, partE);
  return check_launch("knn_exact_group_kernel");
}

constexpr int kExactWarps = 8;

__global__ void __launch_bounds__(kExactWarps * 32) knn_exact_kernel(
    const float* __restrict__ bank, int64_t n, int32_t D, const float* __restrict__ invn, const int32_t* __restrict__ labels,
    const float* __restrict__ qn, const int32_t* __restrict__ qlab, const int32_t* __restrict__ flags, int32_t P,
    int64_t* __restrict__ out_idx, float* __restrict__ out_sim) {
  const int64_t b = blockIdx.x;
  if (flags[b] == 0) return;
  extern __shared__ float sm_qe[];                       // [D] query | [kExactWarps][kKnnC] keys
  unsigned long long* sm_keys = reinterpret_cast<unsigned long long*>(sm_qe + ((D + 3) / 4 * 4));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = threadIdx.x; c < D; c += blockDim.x) sm_qe[c] = qn[b * D + c];
  __syncthreads();
  const int32_t mylab = qlab[b];
  unsigned long long best[kKnnC];                        // identical in all lanes of a warp
#pragma unroll
  for (int i = 0; i < kKnnC; ++i) best[i] = 0ull;
  unsigned long long worst = 0ull;
  // lane l holds columns c = 4 l + 128 k of the query; a row is one coalesced pass of the warp
  for (int64_t j = warp; j < n; j += kExactWarps) {
    float s = 0.0f;
    if (labels[j] == mylab) {
      float acc = 0.f;
      for (int c = lane * 4; c < D; c += 128) {
        const float4 v = ldg_stream_f4(bank + j * D + c);
        const float4 q = *reinterpret_cast<const float4*>(sm_qe + c);
        acc = fmaf(v.x, q.x, acc);
        acc = fmaf(v.y, q.y, acc);
        acc = fmaf(v.z, q.z, acc);
        acc = fmaf(v.w, q.w, acc);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFullMask, acc, o);
      s = acc * invn[j] + 0.0f;
    }
    const unsigned long long nk = knn_key(s, static_cast<uint32_t>(j));
    if (nk > worst) {
      unsigned long long mk = best[0];
      int mp = 0;
#pragma unroll
      for (int q = 1; q < kKnnC; ++q)
        if (best[q] < mk) { mk = best[q]; mp = q; }
#pragma unroll
      for (int q = 0; q < kKnnC; ++q)
        if (q == mp) best[q] = nk;
      worst = best[0];
#pragma unroll
      for (int q = 1; q < kKnnC; ++q) worst = best[q] < worst ? best[q] : worst;
    }
  }
  if (lane == 0)
    for (int i = 0; i < kKnnC; ++i) sm_keys[warp * kKnnC + i] = best[i];
  __syncthreads();
  if (warp == 0) {
    for (int p = 0; p < P; ++p) {
      unsigned long long top = 0ull;
      for (int i = lane; i < kExactWarps * kKnnC; i += 32) top = sm_keys[i] > top ? sm_keys[i] : top;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(kFullMask, top, o);
        top = other > top ? other : top;
      }
      for (int i = lane; i < kExactWarps * kKnnC; i += 32)
        if (sm_keys[i] == top) sm_keys[i] = 0ull;
      __syncwarp();
      if (lane == 0) {
        const bool have = top != 0ull;
        out_idx[b * P + p] = have ? static_cast<int64_t>(0xFFFFFFFFu - static_cast<uint32_t>(top)) : -1;
        out_sim[b * P + p] = have ? ord_value(static_cast<uint32_t>(top >> 32)) : 0.f;
      }
    }
  }
}

struct KnnPlan {
  int64_t Bpad;
  int32_t kat;                                            // anchor tiles per CTA of the GEMM pass (2 when the batch has more than one)
  int32_t atiles, tiles_total, slices, tiles_per_slice, nlists, stages, tmem_cols;
  int32_t tilesA, slicesA, tiles_per_sliceA, nlistsA;      // sampling pass over every kKnnSample-th tile (0 lists: skipped)
  size_t smem;
  bool tensor;              // the tcgen05 pass applies (D a multiple of 32, at most 128; n < 2^31)
  int32_t groupsE, slicesE;                               // exact scan of flagged anchors: 32-anchor groups x bank slices
  int64_t rows_per_sliceE;
  size_t off_invn, off_qn, off_qt, off_qlab, off_part, off_rej, off_partA, off_rejA, off_thr, off_flags, off_list, off_count, off_partE, total;
};

KnnPlan make_knn_plan(int64_t n, int64_t B, int32_t D, int32_t n_classes = 0) {
  KnnPlan p{};
  p.kat = B > kTileM ? 2 : 1;
  p.atiles = static_cast<int32_t>((B + kTileM * p.kat - 1) / (kTileM * p.kat));      // CTAs along the anchors
  if (p.atiles < 1) p.atiles = 1;
  p.Bpad = static_cast<int64_t>(p.atiles) * kTileM * p.kat;
  p.tensor = (D % 32 == 0) && D >= 32 && D <= 128 && n < (static_cast<int64_t>(1) << 31) - kKnnTileN;
  p.tiles_total = static_cast<int32_t>((n + kKnnTileN - 1) / kKnnTileN);
  int32_t slices = 148 / p.atiles;
  if (slices < 1) slices = 1;
  if (slices > p.tiles_total) slices = p.tiles_total;
  p.tiles_per_slice = (p.tiles_total + slices - 1) / slices;
  p.slices = (p.tiles_total + p.tiles_per_slice - 1) / p.tiles_per_slice;
  p.nlists = p.tensor ? p.slices * 2 : 0;
  p.tilesA = p.tiles_total >= 16 * kKnnSample ? (p.tiles_total + kKnnSample - 1) / kKnnSample : 0;
  if (p.tilesA > 0 && p.tensor) {
    int32_t sa = 148 / p.atiles;
    if (sa < 1) sa = 1;
    if (sa > p.tilesA) sa = p.tilesA;
    p.tiles_per_sliceA = (p.tilesA + sa - 1) / sa;
    p.slicesA = (p.tilesA + p.tiles_per_sliceA - 1) / p.tiles_per_sliceA;
    p.nlistsA = p.slicesA * 2;
  } else {
    p.tilesA = 0;
  }
  p.stages = p.kat == 2 ? 4 : 8;
  p.tmem_cols = 512;
  p.smem = (p.kat == 2 ? static_cast<size_t>(2 * (D / 32)) * kKnnBoxBytes : 0) + static_cast<size_t>(p.stages) * kKnnBoxBytes +
           static_cast<size_t>(8 * p.kat) * 2 * ((n_classes >= 1 && n_classes <= 3) ? 208 : 128) * 4 + 256 + (2 * p.stages + 5) * 8 + 16 + 1024;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  p.off_invn = take(static_cast<size_t>(n) * sizeof(float));
  p.off_qn = take(static_cast<size_t>(p.Bpad) * D * sizeof(float));
  p.off_qt = take(static_cast<size_t>(p.Bpad) * D * sizeof(float));
  p.off_qlab = take(static_cast<size_t>(p.Bpad) * sizeof(int32_t));
  p.off_part = take(static_cast<size_t>(p.nlists > 0 ? p.nlists : 1) * p.Bpad * kKnnC * sizeof(unsigned long long));
  p.off_rej = take(static_cast<size_t>(p.nlists > 0 ? p.nlists : 1) * p.Bpad * sizeof(float));
  p.off_partA = take(static_cast<size_t>(p.nlistsA > 0 ? p.nlistsA : 1) * p.Bpad * kKnnC * sizeof(unsigned long long));
  p.off_rejA = take(static_cast<size_t>(p.nlistsA > 0 ? p.nlistsA : 1) * p.Bpad * sizeof(float));
  p.off_thr = take(static_cast<size_t>(p.Bpad) * sizeof(float));
  p.off_flags = take(static_cast<size_t>(p.Bpad) * sizeof(int32_t));
  p.groupsE = static_cast<int32_t>(p.Bpad / 32);
  p.slicesE = 148 * 16 / p.groupsE;
  if (p.slicesE > 148) p.slicesE = 148;
  if (p.slicesE < 1) p.slicesE = 1;
  p.rows_per_sliceE = (n + p.slicesE - 1) / p.slicesE;
  p.rows_per_sliceE = (p.rows_per_sliceE + kEgRows - 1) / kEgRows * kEgRows;
  p.slicesE = static_cast<int32_t>((n + p.rows_per_sliceE - 1) / p.rows_per_sliceE);
  p.off_list = take(static_cast<size_t>(p.Bpad) * sizeof(int32_t));
  p.off_count = take(sizeof(int32_t));
  p.off_partE = take(static_cast<size_t>(p.slicesE) * p.Bpad * kKnnC * sizeof(unsigned long long));
  p.total = off;
  return p;
}

}  // namespace
}  // namespace mml

using namespace mml;

extern "C" int32_t mml_crd_knn_max_positives(void) { return kKnnC; }

extern "C" int64_t mml_crd_knn_workspace_bytes(int64_t n, int64_t B, int32_t D) {
  if (n < 1 || B < 0 || D < 4 || D % 4 != 0) return -1;
  return static_cast<int64_t>(make_knn_plan(n, B, D).total);
}

extern "C" int mml_crd_knn_inv_norms(const float* bank, int64_t n, int32_t D, const int64_t* rows, int64_t count, float* inv_norms,
                                     void* stream) {
  MML_REQUIRE(bank && inv_norms, MML_ERR_INVALID_ARG, "crd_knn_inv_norms: null pointer");
  MML_REQUIRE(n >= 1 && D >= 4 && D % 4 == 0 && count >= 0, MML_ERR_INVALID_ARG, "crd_knn_inv_norms: bad sizes");
  MML_REQUIRE(aligned16(bank), MML_ERR_INVALID_ARG, "crd_knn_inv_norms: bank must be 16-byte aligned");
  const int64_t items = rows != nullptr ? count : n;
  if (items == 0) return MML_OK;
  knn_invnorm_kernel<<<static_cast<unsigned>((items * 8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(bank, n, D, rows, count,
                                                                                                          inv_norms);
  return check_launch("knn_invnorm_kernel");
}

extern "C" int mml_crd_knn_positives(const float* bank, int64_t n, int32_t D, const float* inv_norms, const int32_t* row_labels,
                                     int32_t n_classes, const int64_t* anchor_rows, const float* queries,
                                     const int64_t* anchor_labels, int64_t B,
                                     int32_t P, int32_t exact_only, int64_t* out_idx,
                                     float* out_sim, int32_t* flags_out, void* workspace, size_t workspace_bytes, void* stream) {
  MML_REQUIRE(bank && row_labels && (anchor_rows || queries) && anchor_labels && out_idx && out_sim && workspace,
              MML_ERR_INVALID_ARG, "crd_knn_positives: null pointer");
  MML_REQUIRE(n >= 1 && B >= 0 && D >= 4 && D % 4 == 0 && D <= 1024, MML_ERR_INVALID_ARG, "crd_knn_positives: bad sizes");
  MML_REQUIRE(P >= 1 && P <= kKnnC && P <= n, MML_ERR_UNSUPPORTED, "crd_knn_positives: 1 <= num_pos <= %d supported (got %d)", kKnnC, P);
  MML_REQUIRE(n < (static_cast<int64_t>(1) << 32) - 1, MML_ERR_UNSUPPORTED, "crd_knn_positives: at most 2^32 - 2 bank rows");
  MML_REQUIRE(aligned16(bank) && aligned16(workspace), MML_ERR_INVALID_ARG, "crd_knn_positives: bank / workspace must be 16-byte aligned");
  if (B == 0) return MML_OK;
  const KnnPlan p = make_knn_plan(n, B, D, n_classes);
  MML_REQUIRE(workspace_bytes >= p.total, MML_ERR_INVALID_ARG, "crd_knn_positives: workspace too small (%zu < %zu)", workspace_bytes, p.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const float* invn = inv_norms != nullptr ? inv_norms : reinterpret_cast<float*>(ws + p.off_invn);
  float* qn = reinterpret_cast<float*>(ws + p.off_qn);
  int32_t* qlab = reinterpret_cast<int32_t*>(ws + p.off_qlab);
  unsigned long long* part = reinterpret_cast<unsigned long long*>(ws + p.off_part);
  int32_t* flags = flags_out != nullptr ? flags_out : reinterpret_cast<int32_t*>(ws + p.off_flags);

  if (inv_norms == nullptr) {
    const int64_t threads = n * 8;
    knn_invnorm_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, st>>>(bank, n, D, nullptr, 0,
                                                                                     reinterpret_cast<float*>(ws + p.off_invn));
    const int rc = check_launch("knn_invnorm_kernel");
    if (rc != MML_OK) return rc;
  }
  {
    knn_query_kernel<<<static_cast<unsigned>(p.Bpad), 32, 0, st>>>(bank, n, D, anchor_rows, queries, anchor_labels, B, p.Bpad, qn,
                                                                reinterpret_cast<float*>(ws + p.off_qt), qlab, device_error_word());
    const int rc = check_launch("knn_query_kernel");
    if (rc != MML_OK) return rc;
  }
  const bool tensor = p.tensor && !exact_only;
  float* rej = reinterpret_cast<float*>(ws + p.off_rej);
  if (tensor) {
    CUtensorMap tmap;
    const int rc0 = get_tensor_map_2d(bank, D, n, 32, kKnnTileN, true, &tmap);
    if (rc0 != MML_OK) return rc0;
    KnnArgs a{};
    a.qn = qn; a.qlab = qlab; a.invn = invn; a.labels = row_labels;
    a.n = n; a.Bpad = p.Bpad; a.D = D; a.nbox = D / 32;
    a.stages = p.stages; a.tmem_cols = p.tmem_cols;
    a.idesc = make_idesc_tf32(kTileM, kKnnTileN);
    CUtensorMap tmap_q = tmap;
    if (p.kat == 2) {
      const int rcq = get_tensor_map_2d(reinterpret_cast<float*>(ws + p.off_qt), D, p.Bpad, 32, kTileM, true, &tmap_q);
      if (rcq != MML_OK) return rcq;
    }
    const bool tab = n_classes >= 1 && n_classes <= 3;       // labels in [0, 3): tabulated class multipliers
    const int smem_i = static_cast<int>(p.smem);
    if (p.kat == 2 && tab) MML_CUDA(cudaFuncSetAttribute(knn_gemm_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_i));
    if (p.kat == 2 && !tab) MML_CUDA(cudaFuncSetAttribute(knn_gemm_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_i));
    if (p.kat == 1 && tab) MML_CUDA(cudaFuncSetAttribute(knn_gemm_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_i));
    if (p.kat == 1 && !tab) MML_CUDA(cudaFuncSetAttribute(knn_gemm_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_i));
    auto launch_gemm = [&](int32_t slices) {
      const dim3 grid(static_cast<unsigned>(p.atiles), static_cast<unsigned>(slices));
      if (p.kat == 2 && tab) knn_gemm_kernel<2, true><<<grid, (8 * 2 + 2) * 32, p.smem, st>>>(tmap, tmap_q, a);
      else if (p.kat == 2) knn_gemm_kernel<2, false><<<grid, (8 * 2 + 2) * 32, p.smem, st>>>(tmap, tmap_q, a);
      else if (tab) knn_gemm_kernel<1, true><<<grid, (8 + 2) * 32, p.smem, st>>>(tmap, tmap_q, a);
      else knn_gemm_kernel<1, false><<<grid, (8 + 2) * 32, p.smem, st>>>(tmap, tmap_q, a);
    };
    float* thr_init = nullptr;
    if (p.nlistsA > 0) {             // sampling pass: every 16th tile -> a proven floor for every anchor's lists
      unsigned long long* partA = reinterpret_cast<unsigned long long*>(ws + p.off_partA);
      thr_init = reinterpret_cast<float*>(ws + p.off_thr);
      a.part = partA; a.rej = reinterpret_cast<float*>(ws + p.off_rejA); a.thr_init = nullptr; a.tile_mul = kKnnSample;
      a.max_only = p.nlistsA >= 4 * kKnnC ? 1 : 0;      // enough lists: the 8th best of their maxima is the floor
      a.tiles_total = p.tilesA; a.tiles_per_slice = p.tiles_per_sliceA;
      launch_gemm(p.slicesA);
      int rc = check_launch("knn_gemm_kernel (sampling pass)");
      if (rc != MML_OK) return rc;
      knn_threshold_kernel<<<static_cast<unsigned>((p.Bpad + 3) / 4), 128, 0, st>>>(partA, p.nlistsA, B, p.Bpad, thr_init);
      rc = check_launch("knn_threshold_kernel");
      if (rc != MML_OK) return rc;
    }
    a.part = part; a.rej = rej; a.thr_init = thr_init; a.tile_mul = 1; a.max_only = 0;
    a.tiles_total = p.tiles_total; a.tiles_per_slice = p.tiles_per_slice;
    launch_gemm(p.slices);
    const int rc = check_launch("knn_gemm_kernel");
    if (rc != MML_OK) return rc;
  }
  {
    const size_t smem = static_cast<size_t>(kMergeWarps) * D * sizeof(float);
    knn_merge_kernel<<<static_cast<unsigned>((B + kMergeWarps - 1) / kMergeWarps), kMergeWarps * 32, smem, st>>>(
        bank, D, invn, row_labels, qn, qlab, part, rej, tensor ? p.nlists : 0, B, p.Bpad, P, out_idx, out_sim, flags, tensor ? 0 : 1);
    const int rc = check_launch("knn_merge_kernel");
    if (rc != MML_OK) return rc;
  }
  if (D == 32 || D == 64 || D == 96 || D == 128) {
    int32_t* list = reinterpret_cast<int32_t*>(ws + p.off_list);
    int32_t* count = reinterpret_cast<int32_t*>(ws + p.off_count);
    unsigned long long* partE = reinterpret_cast<unsigned long long*>(ws + p.off_partE);
    knn_compact_kernel<<<1, 1024, 0, st>>>(flags, B, list, count);
    int rc = check_launch("knn_compact_kernel");
    if (rc != MML_OK) return rc;
    switch (D) {
      case 32: rc = launch_exact_group<32>(bank, n, invn, row_labels, qn, qlab, list, count, p.groupsE, p.slicesE, p.rows_per_sliceE, p.Bpad, partE, st); break;
      case 64: rc = launch_exact_group<64>(bank, n, invn, row_labels, qn, qlab, list, count, p.groupsE, p.slicesE, p.rows_per_sliceE, p.Bpad, partE, st); break;
      case 96: rc = launch_exact_group<96>(bank, n, invn, row_labels, qn, qlab, list, count, p.groupsE, p.slicesE, p.rows_per_sliceE, p.Bpad, partE, st); break;
      default: rc = launch_exact_group<128>(bank, n, invn, row_labels, qn, qlab, list, count, p.groupsE, p.slicesE, p.rows_per_sliceE, p.Bpad, partE, st); break;
    }
    if (rc != MML_OK) return rc;
    knn_pick_kernel<<<static_cast<unsigned>((B + 3) / 4), 128, 0, st>>>(partE, p.slicesE, p.Bpad, list, count, P, out_idx, out_sim);
    rc = check_launch("knn_pick_kernel");
    if (rc != MML_OK) return rc;
  } else {                         // other widths: one CTA per flagged anchor
    const size_t smem = static_cast<size_t>((D + 3) / 4 * 4) * sizeof(float) + kExactWarps * kKnnC * sizeof(unsigned long long);
    knn_exact_kernel<<<static_cast<unsigned>(B), kExactWarps * 32, smem, st>>>(bank, n, D, invn, row_labels, qn, qlab, flags, P,
                                                                            out_idx, out_sim);
    const int rc = check_launch("knn_exact_kernel");
    if (rc != MML_OK) return rc;
  }
  return MML_OK;
}
