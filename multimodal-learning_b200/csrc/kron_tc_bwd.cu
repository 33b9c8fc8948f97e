// K3: weight gradient of the Kronecker-fusion encoder on tcgen05 tensor cores.
//
//   dW[n, k] = sum_b dy[b, n] * A[b, k] * m[b, k]                 (autograd of encoder1[0], fusion.py:60 / :129)
//
// Same machinery as the forward (kron_tc.cu) with the roles turned: the 128 TMEM lanes are 128 PACKED k (4 chunks of
// the K permutation), the contraction runs over the batch in stages of 32 rows, the A operand A^T[k, b] is generated
// into tensor memory (thread = (k lane, 16 batch columns)), and the B operand is dy^T [Np, B] streamed by TMA in
// [Np x 32] tiles.  The Kronecker tensor is not materialised here either.  Everything the generators read arrives by
// TMA from ONE transposed copy of the factors, FT [1 + d1 + d2 + d3][Bpad] (row 0 = ones, then the columns of f1, f2,
// f3: row r < n_scal is exactly the per-row scalar R[r]):
//   * the rows of the chunks' scalars R[p], R[q] (one [1 x 32] or [4 x 32] box each when the four chunks share a row or
//     use consecutive rows -- the common case -- else four single rows) and
//   * per distinct vector segment of the tile one [32 rows x 32 b] box (128-byte swizzled),
// so a generator thread does 12 128-bit shared loads, 24 multiplies and one tcgen05.st per stage; there is no global
// load, cp.async or CTA barrier in the loop (mbarriers only).  Lanes past a chunk's length hold finite garbage that the
// unpack kernel never reads.  Operand rounding: dy^T is rounded to nearest onto the TF32 grid; A^T is truncated by the
// tensor core, and the mean of that truncation (-0.5 ulp over log-uniform mantissas) is cancelled by scaling dy^T with
// 1 + 0.69 * 2^-11 before it is rounded -- same RMS error as round-to-nearest, no per-element integer add.
// Output: per-(batch split) partial tiles in the packed layout, reduced and scattered back to the dense [N, Kk] layout
// by kron_unpack_kernel (fixed order -> deterministic).
#include "kron_tc_common.cuh"

namespace mml {
namespace {

using namespace tc;

constexpr int kBlkB = 32;          // batch rows per pipeline stage (= 4 MMAs of K = 8)
constexpr int kWgThreads = kGenThreads + 96;   // 8 generator warps + factor-box TMA warp + MMA warp + dy^T TMA warp
constexpr uint32_t kWgXBytes = 4 * kBlkB * 32 * 4;     // 4 vector-segment slots of [32 rows x 32 b] fp32
constexpr uint32_t kWgSBytes = 8 * kBlkB * 4;          // 8 scalar rows of 32 b

struct WgArgs {
  const int4* table;
  float* part;                // [bsplit][N][Kp]
  int64_t B;
  int32_t d1, d2, d3;
  int32_t N, Np, nchunks, Kp;
  int32_t nblocks, blocks_per_split;
  int32_t stages, xstages, tmem_cols;      // stages: dy^T tiles + TMEM A slots;  xstages: factor-box ring (deeper)
  uint32_t idesc;
  KronDropout dr;
};

template <bool kDropout>
__global__ void __launch_bounds__(kWgThreads, 2)
kron_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dyT, const __grid_constant__ CUtensorMap tmap_x,
                     const __grid_constant__ CUtensorMap tmap_s, const __grid_constant__ CUtensorMap tmap_s4, const WgArgs a) {
  uint32_t seed_lo = 0u, seed_hi = 0u;
  if (kDropout) kron_seed(a.dr, seed_lo, seed_hi);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // Two rings.  D ring (a.stages): dy^T tile [Np x 32] + 32 TMEM columns of A^T per stage, recycled when the stage's MMAs
  // complete.  X ring (a.xstages, deeper): the factor boxes a stage's generators read, recycled as soon as they have read
  // them -- deep enough that the ~1 us TMA round trip of the next boxes hides behind several stages of arithmetic
  // (with one shared ring the generators sat 35 % of the time waiting for their boxes at N = 96).
  const uint32_t dy_bytes = static_cast<uint32_t>(a.Np) * 128u;
  const uint32_t x_bytes = kWgXBytes + 1024u;                        // X slots | S rows (1 KB)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sm_dy = smem;
  uint8_t* sm_x = smem + static_cast<size_t>(a.stages) * dy_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_x + static_cast<size_t>(a.xstages) * x_bytes);
  uint64_t* bar_full = bars;                    // [stages] A^T stored (8 warp arrivals) + dy^T landed (1 arrival + tx)
  uint64_t* bar_empty = bars + a.stages;        // [stages] MMAs of the stage complete
  uint64_t* bar_xfull = bars + 2 * a.stages;    // [xstages] factor boxes landed (1 arrival + tx)
  uint64_t* bar_xempty = bar_xfull + a.xstages; // [xstages] generators have read the boxes (8 warp arrivals)
  uint64_t* bar_acc = bar_xempty + a.xstages;
  uint32_t* sm_tmem = reinterpret_cast<uint32_t*>(bar_acc + 1);

  const int warp = warp_idx_sync();
  const int lane = threadIdx.x & 31;
  const int c_first = blockIdx.x * 4;                                   // this CTA's 4 chunks = 128 packed k
  const int blk_begin = blockIdx.y * a.blocks_per_split;
  const int blk_end = min(a.nblocks, blk_begin + a.blocks_per_split);

  if (warp == kGenWarps && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_dyT)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_s)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_s4)) : "memory");
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&bar_full[s], kGenWarps + 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int s = 0; s < a.xstages; ++s) {
      mbar_init(&bar_xfull[s], 1);
      mbar_init(&bar_xempty[s], kGenWarps);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kGenWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm_tmem)), "r"(a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm_tmem;
  const uint32_t tmem_d = tmem_base;
  const uint32_t tmem_a = tmem_base + static_cast<uint32_t>(a.Np);

  // The tile's chunk descriptors (uniform over the CTA).  Chunks that share a vector segment share a slot.
  int rp[4], rq[4], xrow[4], slot[4];
  bool own[4];
  int nown = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const bool cv = (c_first + c) < a.nchunks;
    const int4 q0 = cv ? __ldg(a.table + 2 * (c_first + c)) : make_int4(0, 0, 0, 0);
    rp[c] = q0.x;
    rq[c] = q0.y;                                  // FT row of a scalar = its index in R = [1, f1, f2]
    xrow[c] = q0.z == 0 ? 0 : (q0.z == 1 ? 1 : (q0.z == 2 ? 1 + a.d1 : 1 + a.d1 + a.d2)) + q0.w;
    slot[c] = c;
    own[c] = true;
#pragma unroll
    for (int p = 0; p < c; ++p)
      if (own[c] && own[p] && xrow[p] == xrow[c]) {
        slot[c] = p;
        own[c] = false;
      }
    nown += own[c] ? 1 : 0;
  }
  // Scalar rows of the four chunks: usually ONE row shared by all four (mode 0: one [32 x 1] box) or four consecutive rows
  // (mode 1: one [32 x 4] box) -- bilinear tiles are (p consecutive, q = 0), trilinear core tiles (p shared, q consecutive) --
  // so a stage costs 4 TMA instructions instead of 10; anything else (mode 2) loads four single rows.
  auto row_mode = [](const int (&r)[4]) {
    if (r[1] == r[0] && r[2] == r[0] && r[3] == r[0]) return 0;
    if (r[1] == r[0] + 1 && r[2] == r[0] + 2 && r[3] == r[0] + 3) return 1;
    return 2;
  };
  const int p_mode = row_mode(rp), q_mode = row_mode(rq);

  if (warp == kGenWarps) {
    // ===== TMA producer of the X ring: the factor boxes the generators read =====
    int k = 0;
    uint32_t ph = 0;
    for (int blk = blk_begin; blk < blk_end; ++blk) {
      mbar_wait(&bar_xempty[k], ph ^ 1);
      if (elect_one_sync()) {
        uint8_t* st = sm_x + static_cast<size_t>(k) * x_bytes;
        const int b0 = blk * kBlkB;
        uint8_t* sp_rows = st + kWgXBytes;                      // R[p] rows at [0, 512), R[q] rows at [512, 1024)
        mbar_arrive_expect_tx(&bar_xfull[k], static_cast<uint32_t>(nown) * (kBlkB * 32 * 4) +
                                                 (p_mode == 0 ? 128u : 512u) + (q_mode == 0 ? 128u : 512u));
        if (p_mode == 0) tma_load_2d(sp_rows, &tmap_s, b0, rp[0], &bar_xfull[k]);
        else if (p_mode == 1) tma_load_2d(sp_rows, &tmap_s4, b0, rp[0], &bar_xfull[k]);
        if (q_mode == 0) tma_load_2d(sp_rows + 512, &tmap_s, b0, rq[0], &bar_xfull[k]);
        else if (q_mode == 1) tma_load_2d(sp_rows + 512, &tmap_s4, b0, rq[0], &bar_xfull[k]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (p_mode == 2) tma_load_2d(sp_rows + c * 128, &tmap_s, b0, rp[c], &bar_xfull[k]);
          if (q_mode == 2) tma_load_2d(sp_rows + 512 + c * 128, &tmap_s, b0, rq[c], &bar_xfull[k]);
          if (own[c]) tma_load_2d(st + c * (kBlkB * 32 * 4), &tmap_x, b0, xrow[c], &bar_xfull[k]);
        }
      }
      __syncwarp();
      if (++k == a.xstages) { k = 0; ph ^= 1; }
    }
  } else if (warp == kGenWarps + 2) {
    // ===== TMA producer of the D ring: dy^T tiles for the MMAs =====
    int s = 0;
    uint32_t ph = 0;
    for (int blk = blk_begin; blk < blk_end; ++blk) {
      mbar_wait(&bar_empty[s], ph ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&bar_full[s], dy_bytes);
        tma_load_2d(sm_dy + static_cast<size_t>(s) * dy_bytes, &tmap_dyT, blk * kBlkB, 0, &bar_full[s]);
      }
      __syncwarp();
      if (++s == a.stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == kGenWarps + 1) {
    // ===== MMA issuer =====
    int s = 0;
    uint32_t ph = 0;
    const uint32_t b_base = smem_u32(sm_dy);
    for (int blk = blk_begin; blk < blk_end; ++blk) {
      mbar_wait(&bar_full[s], ph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t b_desc = umma_desc_k_sw128(b_base + static_cast<uint32_t>(s) * dy_bytes);
        const uint32_t a_col = tmem_a + s * kBlkB;
#pragma unroll
        for (int j = 0; j < kBlkB / 8; ++j)
          tc_mma_tf32_ts(tmem_d, a_col + j * 8, b_desc + 2 * j, a.idesc, (blk > blk_begin || j > 0) ? 1u : 0u);
        tc_commit(&bar_empty[s]);
      }
      __syncwarp();
      if (++s == a.stages) { s = 0; ph ^= 1; }
    }
    if (elect_one_sync()) tc_commit(bar_acc);
    __syncwarp();
  } else {
    // ===== generators: thread = (packed-k lane t = 32 ci + e, 16 batch columns of the stage) =====
    const int gt = threadIdx.x;                        // 0..255
    const int t = gt & (kTileM - 1);
    const int half = gt >> 7;
    const int ci = warp & 3, e = lane;
    const uint32_t lane_base = static_cast<uint32_t>(ci * 32) << 16;
    int my_slot = 0, my_klog = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (c == ci) my_slot = slot[c];
    if (kDropout) {
      const bool cv = (c_first + ci) < a.nchunks;
      const int4 q1 = cv ? __ldg(a.table + 2 * (c_first + ci) + 1) : make_int4(0, 0, 1, 0);
      my_klog = q1.y + (e < q1.x ? e : 0) * q1.z;      // lanes past the chunk's length: any valid counter (never read)
    }
    // byte offsets inside a stage: scalar rows (broadcast reads) and this lane's swizzled row of its vector slot
    const uint32_t off_sp = kWgXBytes + (p_mode == 0 ? 0 : ci) * 128 + half * 64;
    const uint32_t off_sq = kWgXBytes + 512 + (q_mode == 0 ? 0 : ci) * 128 + half * 64;
    const uint32_t off_x = my_slot * (kBlkB * 32 * 4) + e * 128;
    const uint32_t sw = static_cast<uint32_t>(e & 7);
    int s = 0, s_prev = -1, k = 0;
    uint32_t ph = 0, xph = 0;
    for (int blk = blk_begin; blk < blk_end; ++blk) {
      const uint8_t* st = sm_x + static_cast<size_t>(k) * x_bytes;
      // DROP flags of this lane's column for the 16 rows of its half: the block's 32 rows are one mask group
      uint32_t drop = 0u;
      if (kDropout) drop = kron_drop_word(a.dr, seed_lo, seed_hi, blk, my_klog) >> (half * kHalf);
      mbar_wait(&bar_xfull[k], xph);
      uint32_t r[kHalf];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 p4 = *reinterpret_cast<const float4*>(st + off_sp + u * 16);
        const float4 q4 = *reinterpret_cast<const float4*>(st + off_sq + u * 16);
        const float4 x4 = *reinterpret_cast<const float4*>(st + off_x + (((half * 4 + u) ^ sw) << 4));
        float y0 = p4.x * q4.x * x4.x, y1 = p4.y * q4.y * x4.y, y2 = p4.z * q4.z * x4.z, y3 = p4.w * q4.w * x4.w;
        if (kDropout) {
          const uint32_t d4 = drop >> (u * 4);
          y0 = (d4 & 1u) ? 0.f : y0 * a.dr.scale;
          y1 = (d4 & 2u) ? 0.f : y1 * a.dr.scale;
          y2 = (d4 & 4u) ? 0.f : y2 * a.dr.scale;
          y3 = (d4 & 8u) ? 0.f : y3 * a.dr.scale;
        }
        r[u * 4 + 0] = __float_as_uint(y0);
        r[u * 4 + 1] = __float_as_uint(y1);
        r[u * 4 + 2] = __float_as_uint(y2);
        r[u * 4 + 3] = __float_as_uint(y3);
      }
      __syncwarp();                                      // every lane's shared-memory reads of the boxes are done
      if (lane == 0) mbar_arrive(&bar_xempty[k]);        // X slot back to its producer
      if (++k == a.xstages) { k = 0; xph ^= 1; }
      if (s_prev >= 0) {                                 // publish the PREVIOUS stage: its TMEM store had this stage's
        tc_wait_st();                                    // loads and multiplies to complete in
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_full[s_prev]);
      }
      mbar_wait(&bar_empty[s], ph ^ 1);                  // the MMAs that read this TMEM slot last time around are complete
      tc_fence_after();
      tc_st_32x32b_x16(tmem_a + lane_base + s * kBlkB + half * kHalf, r);
      s_prev = s;
      if (++s == a.stages) { s = 0; ph ^= 1; }
    }
    if (s_prev >= 0) {
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[s_prev]);
    }
    // ===== epilogue: D[k lane][n] -> part[split][n][kp] (coalesced over the k lanes) =====
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    const int kp = blockIdx.x * kTileM + t;
    float* dst = a.part + static_cast<int64_t>(blockIdx.y) * a.N * a.Kp + kp;
    const int split_col = ((a.Np / 2 + 15) / 16) * 16;
    const int n_lo = half == 0 ? 0 : split_col;
    const int n_hi = half == 0 ? split_col : a.Np;
    for (int n0 = n_lo; n0 < n_hi; n0 += 16) {
      uint32_t acc[16];
      tc_ld_32x32b_x16(tmem_d + lane_base + n0, acc);
      tc_wait_ld();
      if (kp < a.Kp) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int n = n0 + u;
          if (n < a.N) dst[static_cast<int64_t>(n) * a.Kp] = __uint_as_float(acc[u]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == kGenWarps + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

// dy [B, N] -> dyT [Np, Bpad] (scaled by `comp`, TF32-rounded, zero padded): the K-major B operand of the wgrad MMA
__global__ void kron_transpose_dy_kernel(const float* __restrict__ dy, int64_t B, int32_t N, int32_t Np, int64_t Bpad, float comp,
                                         float* __restrict__ dyT) {
  __shared__ float tile[32][33];
  const int64_t b0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int n0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int64_t b = b0 + r;
    const int n = n0 + threadIdx.x;
    tile[r][threadIdx.x] = (b < B && n < N) ? dy[b * N + n] * comp : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int n = n0 + r;
    const int64_t b = b0 + threadIdx.x;
    if (n < Np && b < Bpad) {
      const uint32_t u = (__float_as_uint(tile[threadIdx.x][r]) + 0x1000u) & 0xFFFFE000u;
      dyT[static_cast<int64_t>(n) * Bpad + b] = __uint_as_float(u);
    }
  }
}

// FT[r][b]: r = 0 -> 1, then the columns of f1, f2 (, f3); zero for b >= B.  Exact fp32 copies.
__global__ void kron_transpose_factors_kernel(const float* __restrict__ f1, const float* __restrict__ f2, const float* __restrict__ f3,
                                              int64_t B, int32_t d1, int32_t d2, int32_t d3, int64_t Bpad, int32_t rows_alloc,
                                              float* __restrict__ FT) {
  __shared__ float tile[32][33];
  const int rows = 1 + d1 + d2 + d3;
  const int64_t b0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int64_t b = b0 + i;
    const int r = r0 + threadIdx.x;
    float x = 0.f;
    if (b < B && r < rows) {
      if (r == 0) x = 1.0f;
      else if (r <= d1) x = f1[b * d1 + (r - 1)];
      else if (r <= d1 + d2) x = f2[b * d2 + (r - 1 - d1)];
      else x = f3[b * d3 + (r - 1 - d1 - d2)];
    }
    tile[i][threadIdx.x] = x;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i;
    const int64_t b = b0 + threadIdx.x;
    if (r < rows_alloc && b < Bpad) FT[static_cast<int64_t>(r) * Bpad + b] = tile[threadIdx.x][i];
  }
}

// dW[n, klog] = sum_z part[z][n][kp]  for every valid packed position kp = (chunk, e)
__global__ void kron_unpack_kernel(const float* __restrict__ part, int32_t bsplit, int32_t N, int32_t Kp, int32_t Kk,
                                   const int4* __restrict__ table, float* __restrict__ dW) {
  const int64_t total = static_cast<int64_t>(N) * Kp;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i / Kp);
    const int kp = static_cast<int>(i % Kp);
    const int4 e1 = __ldg(table + 2 * (kp >> 5) + 1);
    const int e = kp & 31;
    if (e >= e1.x) continue;
    float t = 0.f;
    for (int z = 0; z < bsplit; ++z) t += part[(static_cast<int64_t>(z) * N + n) * Kp + kp];
    dW[static_cast<int64_t>(n) * Kk + e1.y + e * e1.z] = t;
  }
}

struct WgPlan {
  int32_t nchunks, Np, Kp, stages, xstages, tmem_cols, ktiles, nblocks, bsplit, blocks_per_split, ft_rows;
  int64_t Bpad;
  size_t smem, dyT_bytes, ft_bytes, part_bytes;
  bool ok;
};

// Number of batch splits: fill whole waves of CTAs, pay for every split's partial tile (written and re-read once).
inline int32_t pick_split(int64_t units, int64_t tiles, int64_t slots, int64_t min_units, double split_cost, int64_t max_split,
                          double cta_overhead = 4.0) {
  int64_t best = 1;
  double best_cost = 1e300;
  for (int64_t sp = 1; sp <= max_split; ++sp) {
    const int64_t per = (units + sp - 1) / sp;
    if (sp > 1 && per < min_units) break;
    const int64_t waves = (tiles * sp + slots - 1) / slots;
    const double cost = static_cast<double>(waves) * (static_cast<double>(per) + cta_overhead) + split_cost * static_cast<double>(sp - 1);
    if (cost < best_cost * (sp == 1 ? 1.0 : 0.95)) {        // more splits only for a clear (> 5 %) win
      best_cost = cost;
      best = sp;
    }
  }
  return static_cast<int32_t>(best);
}

WgPlan make_wg_plan(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3) {
  WgPlan p{};
  p.nchunks = static_cast<int32_t>(build_chunks(d1, d2, d3).size());
  p.Np = round_np(N);
  p.Kp = p.nchunks * kChunkK;
  p.ktiles = (p.nchunks + 3) / 4;
  p.nblocks = static_cast<int32_t>((B + kBlkB - 1) / kBlkB);
  p.Bpad = static_cast<int64_t>(p.nblocks) * kBlkB;
  p.ft_rows = 1 + d1 + d2 + d3 < 32 ? 32 : 1 + d1 + d2 + d3;       // at least one whole [32 x 32] box (zero rows below the factors)
  // one CTA per SM: D ring of 4-6 stages (dy^T tile + 32 TMEM columns each), the rest of the 227 KB goes to the X ring
  const size_t fixed = 1024 + 1024;
  const size_t dstage = static_cast<size_t>(p.Np) * 128;
  const size_t xstage = kWgXBytes + 1024;
  int stages = p.Np <= 128 ? 6 : 4;
  while (stages > 2 && (p.Np + stages * kBlkB > 512 || fixed + stages * dstage + 3 * xstage > 227 * 1024)) --stages;
  int xstages = static_cast<int>((227 * 1024 - fixed - stages * dstage) / xstage);
  if (xstages > 10) xstages = 10;
  int ctas_per_sm = 1;
  if (p.Np <= 128) {
    // narrow outputs: a stage is only 2*Np <= 256 tensor-pipe cycles, less than one generator pass -- TWO CTAs per SM
    // (<= 113 KB shared memory and <= 256 TMEM columns each) let one CTA's generators run under the other's MMAs
    // (measured: 1 CTA with deep rings 0.207 ms vs 2 CTAs 0.164 ms at 33^3 -> 96, B = 8192)
    int st2 = 5;
    while (st2 > 2 && (p.Np + st2 * kBlkB > 256 || fixed + st2 * dstage + 3 * xstage > 113 * 1024)) --st2;
    const int xs2 = static_cast<int>((113 * 1024 - fixed - st2 * dstage) / xstage);
    if (p.Np + st2 * kBlkB <= 256 && xs2 >= 3) {
      stages = st2;
      xstages = xs2 > 6 ? 6 : xs2;
      ctas_per_sm = 2;
    }
  }
  p.stages = stages;
  p.xstages = xstages;
  p.ok = p.Np <= 256 && stages >= 2 && xstages >= 3;
  p.smem = fixed + static_cast<size_t>(stages) * dstage + static_cast<size_t>(xstages > 0 ? xstages : 0) * xstage;
  int cols = p.Np + stages * kBlkB, pow2 = 32;
  while (pow2 < cols) pow2 <<= 1;
  p.tmem_cols = pow2;
  if (pow2 > 512) p.ok = false;
  // one split's partial tile costs N*Kp*8 bytes of traffic ~ N*Kp*8 / 6.5e12 s; a stage costs ~2*Np cycles at 1.9 GHz per SM slot
  const double stage_s = 2.0 * p.Np / 1.9e9;
  const double split_cost = (static_cast<double>(N) * p.Kp * 8.0 / 6.5e12) / stage_s;
  const int64_t max_split = p.nblocks / 8 > 0 ? (p.nblocks / 8 < 64 ? p.nblocks / 8 : 64) : 1;
  p.bsplit = pick_split(p.nblocks, p.ktiles, 148 * ctas_per_sm, 8, split_cost, max_split);
  p.blocks_per_split = static_cast<int32_t>((p.nblocks + p.bsplit - 1) / p.bsplit);
  p.bsplit = (p.nblocks + p.blocks_per_split - 1) / p.blocks_per_split;
  p.dyT_bytes = (static_cast<size_t>(p.Np) * p.Bpad * sizeof(float) + 1023) / 1024 * 1024;
  p.ft_bytes = (static_cast<size_t>(p.ft_rows) * p.Bpad * sizeof(float) + 1023) / 1024 * 1024;
  p.part_bytes = static_cast<size_t>(p.bsplit) * N * p.Kp * sizeof(float);
  return p;
}

}  // namespace
}  // namespace mml

using namespace mml;

extern "C" int mml_kron_wgrad_supported(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3) {
  if (B < 1 || N < 1 || d1 < 1 || d2 < 1 || d3 < 0) return 0;
  return make_wg_plan(B, N, d1, d2, d3).ok ? 1 : 0;
}

extern "C" size_t mml_kron_wgrad_workspace_bytes(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3) {
  if (B < 1 || N < 1 || d1 < 1 || d2 < 1 || d3 < 0) return 0;
  const WgPlan p = make_wg_plan(B, N, d1, d2, d3);
  return p.dyT_bytes + p.ft_bytes + p.part_bytes + 1024;
}

extern "C" int mml_kron_linear_wgrad(const float* f1, const float* f2, const float* f3, int64_t B, int32_t d1, int32_t d2,
                                     int32_t d3, const int32_t* table, const float* dy, int32_t N, float drop_p,
                                     uint64_t seed, const uint64_t* seed_dev, int32_t training, float* dW, void* workspace, size_t workspace_bytes,
                                     void* stream) {
  MML_REQUIRE(f1 && f2 && table && dy && dW && workspace, MML_ERR_INVALID_ARG, "kron_linear_wgrad: null pointer");
  MML_REQUIRE((d3 > 0) == (f3 != nullptr), MML_ERR_INVALID_ARG, "kron_linear_wgrad: f3 and d3 must both be set or both be absent");
  MML_REQUIRE(B >= 1 && d1 >= 1 && d2 >= 1 && d3 >= 0 && N >= 1, MML_ERR_INVALID_ARG, "kron_linear_wgrad: bad sizes");
  MML_REQUIRE(aligned16(table) && (reinterpret_cast<uintptr_t>(workspace) & 1023u) == 0, MML_ERR_INVALID_ARG,
              "kron_linear_wgrad: table must be 16-byte and workspace 1024-byte aligned");
  const WgPlan p = make_wg_plan(B, N, d1, d2, d3);
  MML_REQUIRE(p.ok, MML_ERR_UNSUPPORTED, "kron_linear_wgrad: N=%d (<=256) exceeds the tile budget", N);
  MML_REQUIRE(workspace_bytes >= p.dyT_bytes + p.ft_bytes + p.part_bytes, MML_ERR_WORKSPACE, "kron_linear_wgrad: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* dyT = static_cast<float*>(workspace);
  float* FT = reinterpret_cast<float*>(static_cast<char*>(workspace) + p.dyT_bytes);
  float* part = reinterpret_cast<float*>(static_cast<char*>(workspace) + p.dyT_bytes + p.ft_bytes);
  const KronShape s = make_kron_shape(d1, d2, d3);
  {
    const dim3 grid(static_cast<unsigned>(p.Bpad / 32), (p.Np + 31) / 32);
    kron_transpose_dy_kernel<<<grid, dim3(32, 8), 0, st>>>(dy, B, N, p.Np, p.Bpad, kTruncComp, dyT);
    int rc = check_launch("kron_transpose_dy_kernel");
    if (rc != MML_OK) return rc;
    const dim3 gridf(static_cast<unsigned>(p.Bpad / 32), (p.ft_rows + 31) / 32);
    kron_transpose_factors_kernel<<<gridf, dim3(32, 8), 0, st>>>(f1, f2, f3, B, d1, d2, d3, p.Bpad, p.ft_rows, FT);
    rc = check_launch("kron_transpose_factors_kernel");
    if (rc != MML_OK) return rc;
  }
  CUtensorMap tmap, tmap_x, tmap_s;
  int rc = get_tensor_map(dyT, p.Np, static_cast<int32_t>(p.Bpad), &tmap);
  if (rc != MML_OK) return rc;
  rc = get_tensor_map_2d(FT, p.Bpad, p.ft_rows, kBlkB, 32, true, &tmap_x);
  if (rc != MML_OK) return rc;
  rc = get_tensor_map_2d(FT, p.Bpad, p.ft_rows, kBlkB, 1, false, &tmap_s);
  if (rc != MML_OK) return rc;
  CUtensorMap tmap_s4;
  rc = get_tensor_map_2d(FT, p.Bpad, p.ft_rows, kBlkB, 4, false, &tmap_s4);
  if (rc != MML_OK) return rc;
  WgArgs a{};
  a.table = reinterpret_cast<const int4*>(table);
  a.part = part;
  a.B = B; a.d1 = d1; a.d2 = d2; a.d3 = d3; a.N = N; a.Np = p.Np; a.nchunks = p.nchunks; a.Kp = p.Kp;
  a.nblocks = p.nblocks; a.blocks_per_split = p.blocks_per_split;
  a.stages = p.stages; a.xstages = p.xstages; a.tmem_cols = p.tmem_cols;
  a.idesc = make_idesc_tf32(kTileM, p.Np);
  a.dr = make_kron_dropout(drop_p, seed, training, s.Kk, seed_dev);
  const dim3 grid(p.ktiles, p.bsplit);
  if (a.dr.thresh != 0u) {
    MML_CUDA(cudaFuncSetAttribute(kron_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem)));
    kron_wgrad_tc_kernel<true><<<grid, kWgThreads, p.smem, st>>>(tmap, tmap_x, tmap_s, tmap_s4, a);
  } else {
    MML_CUDA(cudaFuncSetAttribute(kron_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem)));
    kron_wgrad_tc_kernel<false><<<grid, kWgThreads, p.smem, st>>>(tmap, tmap_x, tmap_s, tmap_s4, a);
  }
  rc = check_launch("kron_wgrad_tc_kernel");
  if (rc != MML_OK) return rc;
  const int64_t total = static_cast<int64_t>(N) * p.Kp;
  int64_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  kron_unpack_kernel<<<static_cast<unsigned>(g), 256, 0, st>>>(part, p.bsplit, N, p.Kp, s.Kk, a.table, dW);
  return check_launch("kron_unpack_kernel");
}

// =====================================================================================================
// K2: factor gradients (dgrad) on tcgen05 tensor cores.
//
//   dA[b, k] = m[b,k] * sum_n dy[b,n] W[n,k]        -- never stored: each [128 rows x 128 packed k] tile is
//   accumulated in TENSOR MEMORY (M = 128 batch lanes, N = 128 packed k, K = n) and folded on the spot into
//   the per-row factor gradients by the epilogue warps:
//     chunk (p, q, vector segment x):  A = R[p] R[q] x[e]   =>   dx[e]   += dA[e] R[p] R[q]
//                                                               dR[p]   += (sum_e dA[e] x[e]) R[q],   dR[q] likewise
//   A operand = the dy tile, written ONCE per CTA into TMEM (tcgen05.st); B operand = [128 k x 32 n] boxes of
//   the transposed packed weight WpT [Kp, Np32] streamed by TMA, 1-3 boxes per pipeline stage; two accumulator
//   tiles alternate so the epilogue of tile t overlaps the MMAs of tile t+1.
//   Epilogue: 8 warps, thread = (batch row, 16-element half of every chunk).  The per-row scalars R[p], R[q] and the
//   vector segments come from the transposed factor copy FT [1 + d1 + d2 + d3][Bpad] (coalesced over the rows of
//   the tile; issued before the accumulator is waited for), dx lives in registers for a whole run of chunks
//   (build_chunks emits each vector segment as one run), dR accumulates in shared memory.  Packed FFMA2 math.
//   Split over k tiles -> per-split partial gradients, summed in split order by kron_dgrad_reduce_kernel.
// =====================================================================================================
namespace mml {
namespace {

constexpr int kDgEpiWarps = 8;            // epilogue warp g: TMEM lanes 32*(g%4).., elements 16*(g/4).. of every chunk
constexpr int kDgEpiThreads = kDgEpiWarps * 32;
constexpr int kDgThreads = kDgEpiThreads + 64;   // + TMA warp + MMA warp
constexpr int kDgTileK = 128;             // packed k per accumulator tile (4 chunks)
constexpr int kDgBoxN = 32;               // n per TMA box
constexpr uint32_t kDgBoxBytes = kDgTileK * kDgBoxN * 4;     // 16 KB
constexpr int kDgDsFloats = 2 * 4 * 2 * kTileM;              // sm_ds [tile parity][chunk][half][row]
constexpr int kDgScFloats = 2 * 8 * kTileM;                  // sm_sc [tile parity][chunk, p|q][row]

struct DgArgs {
  const float* FT;            // [1 + d1 + d2 + d3][Bpad], row 0 = ones
  const float* dy;            // [B, N]
  const int4* table;
  float* part;                // [ksplit][B][dsum]  (zero-initialised)
  int64_t B, Bpad;
  int32_t d1, d2, d3, dsum;
  int32_t N, Np32, nchunks;
  int32_t ktiles, tiles_per_split;
  int32_t n_scal, stages, bps, tmem_cols, table_in_smem, fold_mode;
  uint32_t idesc;
  KronDropout dr;
};

// two fp32 FMAs per issue slot (FFMA2): d = a * b + d on a pair of registers
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

__device__ __forceinline__ int dg_xrow(const int4& e0, int32_t d1, int32_t d2) {      // FT row of a vector segment's first element
  return (e0.z == 0 ? 0 : (e0.z == 1 ? 1 : (e0.z == 2 ? 1 + d1 : 1 + d1 + d2))) + e0.w;
}

template <bool kDropout>
__global__ void __launch_bounds__(kDgThreads, 1) kron_dgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_wT, const DgArgs a) {
  uint32_t seed_lo = 0u, seed_hi = 0u;
  if (kDropout) kron_seed(a.dr, seed_lo, seed_hi);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t stage_bytes = static_cast<uint32_t>(a.bps) * kDgBoxBytes;
  uint8_t* sm_b = smem;                                                              // [stages][bps x 16 KB]
  float* sm_dR = reinterpret_cast<float*>(smem + static_cast<size_t>(a.stages) * stage_bytes);    // dR accum  [n_scal][128]
  float* sm_ds = sm_dR + static_cast<size_t>(a.n_scal) * kTileM;                                  // per-chunk <dA, x> halves
  float* sm_sc = sm_ds + kDgDsFloats;                                                             // R[p], R[q] of the tile's chunks [2][8][128]
  int4* sm_tab = reinterpret_cast<int4*>(sm_sc + kDgScFloats);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_tab + (a.table_in_smem ? 2 * a.nchunks : 0));
  uint64_t* bar_full = bars;                     // [stages] weight boxes landed
  uint64_t* bar_empty = bars + a.stages;         // [stages] MMAs reading the boxes done
  uint64_t* bar_acc_full = bars + 2 * a.stages;  // [2] accumulator tile complete
  uint64_t* bar_acc_empty = bar_acc_full + 2;    // [2] epilogue has read the tile out of TMEM (8 warp arrivals)
  uint32_t* sm_tmem = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);
  uint32_t* sm_drop = sm_tmem + 1 + ((16u - ((smem_u32(sm_tmem) + 4u) & 15u)) & 15u) / 4u;   // 16 B aligned; pointer arithmetic keeps it a shared pointer
                                                 // [8 warps][2][16] dropout words crossing a warp (kron_drop_words16)

  const int warp = warp_idx_sync();
  const int lane = threadIdx.x & 31;
  const int64_t b0 = static_cast<int64_t>(blockIdx.x) * kTileM;
  const int t_begin = blockIdx.y * a.tiles_per_split;
  const int t_end = min(a.ktiles, t_begin + a.tiles_per_split);
  const int nbox = a.Np32 / kDgBoxN;

  if (warp == kDgEpiWarps && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_wT)) : "memory");
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_acc_full[i], 1);
      mbar_init(&bar_acc_empty[i], kDgEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kDgEpiWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm_tmem)), "r"(a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (a.table_in_smem)
    for (int i = t_begin * 8 + threadIdx.x; i < min(a.nchunks, t_end * 4) * 2; i += kDgThreads) sm_tab[i] = __ldg(a.table + i);
  for (int i = threadIdx.x; i < a.n_scal * kTileM; i += kDgThreads) sm_dR[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm_tmem;
  const uint32_t tmem_acc = tmem_base;                 // 2 x 128 accumulator columns
  const uint32_t tmem_a = tmem_base + 2 * kDgTileK;    // dy tile: Np32 columns
  // chunk descriptor i (two int4 per chunk): an explicit shared or read-only global load, never a generic one
  auto tab_at = [&](int i) -> int4 { return a.table_in_smem ? sm_tab[i] : __ldg(a.table + i); };

  // ---- A operand: this CTA's dy rows -> TMEM, once (epilogue warps), then a CTA-wide sync publishes them ----
  if (warp < kDgEpiWarps) {
    const int row = threadIdx.x & (kTileM - 1);
    const int half = threadIdx.x >> 7;
    const int64_t b = b0 + row;
    const bool live = b < a.B;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const bool vec = (a.N % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.dy) & 15u) == 0);
    for (int n0 = half * 16; n0 < a.Np32; n0 += 32) {
      uint32_t r[16];
      if (vec && live && n0 + 16 <= a.N) {
#pragma unroll
        for (int u = 0; u < 16; u += 4) {
          const float4 x4 = __ldg(reinterpret_cast<const float4*>(a.dy + b * a.N + n0 + u));
          r[u] = __float_as_uint(x4.x) + 0x1000u;
          r[u + 1] = __float_as_uint(x4.y) + 0x1000u;
          r[u + 2] = __float_as_uint(x4.z) + 0x1000u;
          r[u + 3] = __float_as_uint(x4.w) + 0x1000u;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const int n = n0 + u;
          const float x = (live && n < a.N) ? __ldg(a.dy + b * a.N + n) : 0.f;
          r[u] = __float_as_uint(x) + 0x1000u;
        }
      }
      tc_st_32x32b_x16(tmem_a + lane_base + n0, r);
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == kDgEpiWarps) {
    // ===== TMA producer: bps boxes of [128 k x 32 n] per stage (whole warp walks the ring, one elected lane issues) =====
    int s = 0;
    uint32_t ph = 0;
    for (int t = t_begin; t < t_end; ++t)
      for (int j = 0; j < nbox; j += a.bps) {
        mbar_wait(&bar_empty[s], ph ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&bar_full[s], stage_bytes);
          for (int i = 0; i < a.bps; ++i)
            tma_load_2d(sm_b + static_cast<size_t>(s) * stage_bytes + i * kDgBoxBytes, &tmap_wT, (j + i) * kDgBoxN, t * kDgTileK,
                        &bar_full[s]);
        }
        __syncwarp();
        if (++s == a.stages) { s = 0; ph ^= 1; }
      }
  } else if (warp == kDgEpiWarps + 1) {
    // ===== MMA issuer =====
    int s = 0;
    uint32_t ph = 0;
    const uint32_t b_base = smem_u32(sm_b);
    for (int t = t_begin; t < t_end; ++t) {
      const int it = t - t_begin;
      const int buf = it & 1;
      mbar_wait(&bar_acc_empty[buf], ((it >> 1) & 1) ^ 1);            // epilogue has drained this accumulator
      tc_fence_after();
      for (int j = 0; j < nbox; j += a.bps) {
        mbar_wait(&bar_full[s], ph);
        tc_fence_after();
        if (elect_one_sync()) {
          for (int i = 0; i < a.bps; ++i) {
            const uint64_t b_desc = umma_desc_k_sw128(b_base + static_cast<uint32_t>(s) * stage_bytes + i * kDgBoxBytes);
            const uint32_t a_col = tmem_a + (j + i) * kDgBoxN;
#pragma unroll
            for (int k = 0; k < kDgBoxN / 8; ++k)
              tc_mma_tf32_ts(tmem_acc + buf * kDgTileK, a_col + k * 8, b_desc + 2 * k, a.idesc, (j + i > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&bar_empty[s]);
          if (j + a.bps >= nbox) tc_commit(&bar_acc_full[buf]);
        }
        __syncwarp();
        if (++s == a.stages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===== epilogue: thread = (batch row, 16-element half of every chunk); fold dA tiles into factor gradients =====
    const int row = threadIdx.x & (kTileM - 1);
    const int half = threadIdx.x >> 7;
    const int eb = half * kHalf;
    const int64_t b = b0 + row;
    const bool live = b < a.B;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    float* my_part = a.part + (static_cast<int64_t>(blockIdx.y) * a.B + b) * a.dsum;
    const float* ft_b = a.FT + b;                       // column b of FT: ft_b[r * Bpad] (b < Bpad always; zero past B)
    float v[kHalf], dv[kHalf];
#pragma unroll
    for (int e = 0; e < kHalf; ++e) { v[e] = 0.f; dv[e] = 0.f; }
    int cur_src = -1, cur_col = -1, cur_len = 0, cur_row = -1;
    int run_p = 0;                 // fold_mode 2: scalar whose gradient is being accumulated in run_acc
    float run_acc = 0.f;
    const int64_t row_group = (b0 >> 5) + (warp & 3);      // the warp's 32 rows are one group of the dropout mask
    const uint32_t lane_bit = 1u << lane;
    // A vector segment is one contiguous run of chunks (build_chunks), so its gradient leaves the registers once per CTA:
    //   * factors that are not among the per-row scalars R (f2 when bilinear, f3 when trilinear): plain stores into the
    //     zero-initialised partial buffer -- no read-modify-write anywhere on the global side;
    //   * factors inside R (f1; f2 when trilinear): added to the dR accumulator in shared memory.  The barrier orders the
    //     add after the other half-thread's fold of the previous tile (flush points are uniform over the 256 threads).
    const bool tri = a.d3 > 0;
    auto flush = [&]() {
      if (cur_src <= 0) return;
      if (cur_src == 1 || (cur_src == 2 && tri)) {
        asm volatile("bar.sync 2, %0;" ::"n"(kDgEpiThreads) : "memory");
        float* acc_r = sm_dR + static_cast<size_t>(cur_row + eb) * kTileM + row;
#pragma unroll
        for (int e = 0; e < kHalf; ++e)
          if (eb + e < cur_len) acc_r[e * kTileM] += dv[e];
      } else if (live) {
        float* dst = my_part + (cur_src == 2 ? a.d1 : a.d1 + a.d2) + cur_col + eb;
#pragma unroll
        for (int e = 0; e < kHalf; ++e)
          if (eb + e < cur_len) dst[e] = dv[e];
      }
    };
    auto new_segment = [&](int cg) {      // rare: once per run of chunks sharing a vector segment
      flush();
      const int4 e0 = tab_at(2 * cg);
      const int4 e1 = tab_at(2 * cg + 1);
      cur_src = e0.z; cur_col = e0.w; cur_len = e1.x;
      cur_row = dg_xrow(e0, a.d1, a.d2);
#pragma unroll
      for (int e = 0; e < kHalf; ++e) {
        float x = 0.f;
        if (cur_src == 0) x = (eb + e == 0) ? 1.0f : 0.f;
        else if (eb + e < cur_len) x = __ldg(ft_b + static_cast<int64_t>(cur_row + eb + e) * a.Bpad);
        v[e] = x;
        dv[e] = 0.f;
      }
    };
    // The per-row scalars R[p], R[q] of a tile's chunks travel FT -> shared memory with cp.async ONE TILE AHEAD (half h of
    // a row fetches chunks 2h, 2h+1): by the time a tile starts they are a shared-memory read away, and the shared array
    // is 8 KB instead of the [n_scal][128] copy of every scalar.
    auto stage_scalars = [&](int tt, int par) {
      if (tt < t_end) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c = 2 * half + cc;
          const int cg = tt * 4 + c;
          const int4 e0 = cg < a.nchunks ? tab_at(2 * cg) : make_int4(0, 0, 0, 0);
          float* dstp = sm_sc + par * (8 * kTileM) + (2 * c) * kTileM + row;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dstp)),
                       "l"(ft_b + static_cast<int64_t>(e0.x) * a.Bpad) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dstp + kTileM)),
                       "l"(ft_b + static_cast<int64_t>(e0.y) * a.Bpad) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage_scalars(t_begin, 0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"n"(kDgEpiThreads) : "memory");
    for (int t = t_begin; t < t_end; ++t) {
      const int it = t - t_begin;
      const int buf = it & 1;
      float* ds_slot = sm_ds + (it & 1) * (4 * 2 * kTileM) + half * kTileM + row;     // + c * 2 * kTileM
      stage_scalars(t + 1, (it + 1) & 1);
      // chunk descriptors and this row's scalars of the tile
      int P[4], Q[4], XR[4], KB[4], KS[4];
      float SP[4], SQ[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int cg = t * 4 + c;
        const bool cvalid = cg < a.nchunks;
        const int4 e0 = cvalid ? tab_at(2 * cg) : make_int4(0, 0, 0, 0);
        P[c] = e0.x;
        Q[c] = e0.y;
        XR[c] = cvalid ? dg_xrow(e0, a.d1, a.d2) : -1;
        if (kDropout) {
          const int4 e1 = cvalid ? tab_at(2 * cg + 1) : make_int4(0, 0, 1, 0);
          KB[c] = e1.y;
          KS[c] = e1.z;
        }
        SP[c] = sm_sc[(it & 1) * (8 * kTileM) + (2 * c) * kTileM + row];
        SQ[c] = sm_sc[(it & 1) * (8 * kTileM) + (2 * c + 1) * kTileM + row];
      }
      mbar_wait(&bar_acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_acc + lane_base + buf * kDgTileK + eb;
      // The tile's four chunks, fully unrolled over two register buffers: the tcgen05.ld of chunk c+1 is in flight while
      // chunk c is folded, and no register is copied (a rolled loop has to move the 16 loaded words every iteration).
      uint32_t accA[kHalf], accB[kHalf];
#define MML_DG_CHUNK(C, ACC, NEXT_LD)                                                                   \
      {                                                                                                 \
        if (XR[C] >= 0 && XR[C] != cur_row) new_segment(t * 4 + (C));                                   \
        float s = SP[C] * SQ[C];                                                                        \
        if (kDropout) s *= a.dr.scale;                                                                  \
        tc_wait_ld();                                                                                   \
        NEXT_LD;                                                                                        \
        float g[kHalf];                                                                                 \
        _Pragma("unroll") for (int e = 0; e < kHalf; ++e) g[e] = __uint_as_float(ACC[e]);               \
        if (kDropout) {                                                                                 \
          uint32_t dw[16];                                                                              \
          kron_drop_words16(a.dr, seed_lo, seed_hi, row_group, KB[C] + eb * KS[C], KS[C],               \
                            sm_drop + (warp * 2 + ((C) & 1)) * 16, lane, dw);                           \
          _Pragma("unroll") for (int e = 0; e < kHalf; ++e) g[e] = (dw[e] & lane_bit) ? 0.f : g[e];     \
        }                                                                                               \
        float ds0 = 0.f, ds1 = 0.f, ds2 = 0.f, ds3 = 0.f;                                               \
        _Pragma("unroll") for (int e = 0; e < kHalf; e += 4) {                                          \
          ffma2(ds0, ds1, g[e], g[e + 1], v[e], v[e + 1]);                                              \
          ffma2(ds2, ds3, g[e + 2], g[e + 3], v[e + 2], v[e + 3]);                                      \
          ffma2(dv[e], dv[e + 1], g[e], g[e + 1], s, s);                                                \
          ffma2(dv[e + 2], dv[e + 3], g[e + 2], g[e + 3], s, s);                                        \
        }                                                                                               \
        ds_slot[(C) * 2 * kTileM] = (ds0 + ds1) + (ds2 + ds3);                                          \
      }
      tc_ld_32x32b_x16(t_addr, accA);
      MML_DG_CHUNK(0, accA, tc_ld_32x32b_x16(t_addr + 32, accB))
      MML_DG_CHUNK(1, accB, tc_ld_32x32b_x16(t_addr + 64, accA))
      MML_DG_CHUNK(2, accA, tc_ld_32x32b_x16(t_addr + 96, accB))
      MML_DG_CHUNK(3, accB, (void)0)
      tc_fence_before();                    // this warp's TMEM reads of the tile are complete (wait::ld above)
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_acc_empty[buf]);
      asm volatile("cp.async.wait_group 0;" ::: "memory");                  // next tile's scalars (issued a tile ago) have landed
      asm volatile("bar.sync 1, %0;" ::"n"(kDgEpiThreads) : "memory");      // both halves of every row have posted <dA, x>
      // dR[p] += <dA, x> R[q],  dR[q] += <dA, x> R[p].  Every dR element is touched by ONE thread per tile (fixed order):
      //   fold_mode 1 (bilinear: q == 0 everywhere, p distinct inside a tile): half h folds chunks 2h, 2h+1;
      //   fold_mode 2 (trilinear): half 0 folds the p side -- p repeats over runs of chunks, so the sum rides in a register
      //     (run_p, run_acc) across tiles and reaches shared memory once per run -- half 1 the q side (distinct inside a tile:
      //     four independent read-modify-writes in flight);
      //   fold_mode 0 (factor widths below 4, where those sets could overlap inside a tile): half 0 folds everything.
      // Row 0 of sm_dR (the constant 1, p or q == 0) is read but never written.
      {
        const float* dsr = sm_ds + (it & 1) * (4 * 2 * kTileM) + row;
        float ds[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          ds[c] = dsr[c * 2 * kTileM] + dsr[c * 2 * kTileM + kTileM];
          if (kDropout) ds[c] *= a.dr.scale;
        }
        if (a.fold_mode == 1) {
          const int pa = half == 0 ? P[0] : P[2], pb = half == 0 ? P[1] : P[3];
          float* r0 = sm_dR + pa * kTileM + row;
          float* r1 = sm_dR + pb * kTileM + row;
          const float o0 = *r0, o1 = *r1;                         // two independent read-modify-writes in flight
          if (pa != 0) *r0 = o0 + (half == 0 ? ds[0] * SQ[0] : ds[2] * SQ[2]);
          if (pb != 0) *r1 = o1 + (half == 0 ? ds[1] * SQ[1] : ds[3] * SQ[3]);
        } else if (a.fold_mode == 2) {
          if (half == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              if (P[c] != run_p) {
                if (run_p != 0) sm_dR[run_p * kTileM + row] += run_acc;
                run_p = P[c];
                run_acc = 0.f;
              }
              run_acc = fmaf(ds[c], SQ[c], run_acc);
            }
          } else {
            float* rq[4];
            float old[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              rq[c] = sm_dR + Q[c] * kTileM + row;
              old[c] = *rq[c];
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (Q[c] != 0) *rq[c] = fmaf(ds[c], SP[c], old[c]);
          }
        } else if (half == 0) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (P[c] != 0) sm_dR[P[c] * kTileM + row] += ds[c] * SQ[c];
            if (Q[c] != 0) sm_dR[Q[c] * kTileM + row] += ds[c] * SP[c];
          }
        }
      }
    }
    if (run_p != 0) sm_dR[run_p * kTileM + row] += run_acc;       // (fold_mode 2, half 0) the last run
    flush();
    asm volatile("bar.sync 1, %0;" ::"n"(kDgEpiThreads) : "memory");      // every fold and shared-memory flush has landed
    if (live) {
      const int ns = a.n_scal - 1;                      // R[1..] = f1 (then f2 when trilinear): same offsets as the output layout
      const int per = (ns + 1) / 2;
      const int lo = half * per, hi = min(ns, lo + per);
      for (int i = lo; i < hi; ++i) my_part[i] = sm_dR[(1 + i) * kTileM + row];
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == kDgEpiWarps + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

// WpT[kp][n] = TF32(W[n][klog(kp)]) (0 for padding), n < Np32
__global__ void kron_pack_t_kernel(const float* __restrict__ W, int32_t N, int32_t Np32, int32_t Kk, const int4* __restrict__ table,
                                   int32_t nchunks, float* __restrict__ WpT) {
  const int64_t total = static_cast<int64_t>(nchunks) * kChunkK * Np32;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(i % Np32);
    const int kp = static_cast<int>(i / Np32);
    const int4 e1 = __ldg(table + 2 * (kp >> 5) + 1);
    const int e = kp & 31;
    float w = 0.f;
    if (n < N && e < e1.x) w = __ldg(W + static_cast<int64_t>(n) * Kk + e1.y + e * e1.z);
    const uint32_t u = (__float_as_uint(w) + 0x1000u) & 0xFFFFE000u;
    WpT[i] = __uint_as_float(u);
  }
}

__global__ void kron_dgrad_reduce_kernel(const float* __restrict__ part, int32_t ksplit, int64_t B, int32_t d1, int32_t d2,
                                         int32_t d3, float* __restrict__ df1, float* __restrict__ df2, float* __restrict__ df3) {
  const int dsum = d1 + d2 + d3;
  const int64_t total = B * dsum;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float t = 0.f;
    for (int z = 0; z < ksplit; ++z) t += part[static_cast<int64_t>(z) * total + i];
    const int64_t b = i / dsum;
    const int x = static_cast<int>(i % dsum);
    if (x < d1) df1[b * d1 + x] = t;
    else if (x < d1 + d2) df2[b * d2 + (x - d1)] = t;
    else df3[b * d3 + (x - d1 - d2)] = t;
  }
}

struct DgPlan {
  int32_t nchunks, Np32, Kp, n_scal, stages, bps, tmem_cols, ktiles, ksplit, tiles_per_split, table_in_smem, dsum, fold_mode;
  int32_t ft_rows;
  int64_t Bpad;
  size_t smem, part_bytes, ft_bytes;
  bool ok;
};

DgPlan make_dg_plan(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3) {
  DgPlan p{};
  p.nchunks = static_cast<int32_t>(build_chunks(d1, d2, d3).size());
  p.Kp = p.nchunks * kChunkK;
  p.Np32 = (N + 31) / 32 * 32;
  p.n_scal = 1 + d1 + (d3 > 0 ? d2 : 0);
  p.dsum = d1 + d2 + d3;
  p.ktiles = (p.nchunks + 3) / 4;
  p.fold_mode = d3 > 0 ? ((d1 >= 4 && d2 >= 4) ? 2 : 0) : (d1 >= 4 ? 1 : 0);
  const int nbox = p.Np32 / kDgBoxN;
  p.bps = nbox % 3 == 0 ? 3 : (nbox % 2 == 0 ? 2 : 1);
  const size_t stage = static_cast<size_t>(p.bps) * kDgBoxBytes;
  const size_t table_bytes = static_cast<size_t>(p.nchunks) * sizeof(Chunk);
  p.table_in_smem = table_bytes <= static_cast<size_t>(kMaxSmemTable) ? 1 : 0;
  const size_t fixed = static_cast<size_t>(p.n_scal) * kTileM * sizeof(float) + (kDgDsFloats + kDgScFloats) * sizeof(float) +
                       (p.table_in_smem ? table_bytes : 0) + 256 + 1024 + 1024;    // + the dropout word exchange
  int stages = fixed + 2 * stage <= 227 * 1024 ? static_cast<int>((227 * 1024 - fixed) / stage) : 0;
  if (stages > 6) stages = 6;
  p.stages = stages;
  p.tmem_cols = (2 * kDgTileK + p.Np32) <= 256 ? 256 : 512;
  p.ok = p.Np32 <= 256 && stages >= 2;
  p.smem = fixed + static_cast<size_t>(stages > 0 ? stages : 0) * stage;
  const int64_t mtiles = (B + kTileM - 1) / kTileM;
  p.Bpad = mtiles * kTileM;
  p.ft_rows = 1 + p.dsum;
  // k split: fill whole waves of CTAs; every split costs one more [B, dsum] partial (memset, written, re-read)
  const double tile_s = 8.0 * p.Np32 / 1.9e9;
  const double split_cost = (static_cast<double>(B) * p.dsum * 12.0 / 6.5e12) / tile_s;
  const int64_t max_ks = p.ktiles / 2 > 0 ? (p.ktiles / 2 < 32 ? p.ktiles / 2 : 32) : 1;
  const int64_t ks = pick_split(p.ktiles, mtiles, 148, 2, split_cost, max_ks, 6e-6 / tile_s);   // ~6 us of per-CTA set-up
  p.tiles_per_split = static_cast<int32_t>((p.ktiles + ks - 1) / ks);
  p.ksplit = (p.ktiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.part_bytes = (static_cast<size_t>(p.ksplit) * B * p.dsum * sizeof(float) + 1023) / 1024 * 1024;
  p.ft_bytes = static_cast<size_t>(p.ft_rows) * p.Bpad * sizeof(float);
  return p;
}

}  // namespace
}  // namespace mml

extern "C" int mml_kron_dgrad_supported(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3) {
  if (B < 1 || N < 1 || d1 < 1 || d2 < 1 || d3 < 0) return 0;
  return make_dg_plan(B, N, d1, d2, d3).ok ? 1 : 0;
}

extern "C" int64_t mml_kron_packed_t_floats(int32_t N, int32_t d1, int32_t d2, int32_t d3) {
  if (N < 1 || d1 < 1 || d2 < 1 || d3 < 0) return -1;
  return static_cast<int64_t>(build_chunks(d1, d2, d3).size()) * kChunkK * ((N + 31) / 32 * 32);
}

extern "C" int mml_kron_pack_weight_t(const float* W, int32_t N, int32_t d1, int32_t d2, int32_t d3, const int32_t* table,
                                      float* WpT, void* stream) {
  MML_REQUIRE(W && table && WpT && N >= 1, MML_ERR_INVALID_ARG, "kron_pack_weight_t: bad arguments");
  MML_REQUIRE(aligned16(table) && (reinterpret_cast<uintptr_t>(WpT) & 127u) == 0, MML_ERR_INVALID_ARG,
              "kron_pack_weight_t: table must be 16-byte and WpT 128-byte aligned");
  const KronShape s = make_kron_shape(d1, d2, d3);
  const int32_t nchunks = static_cast<int32_t>(build_chunks(d1, d2, d3).size());
  const int32_t Np32 = (N + 31) / 32 * 32;
  const int64_t total = static_cast<int64_t>(nchunks) * kChunkK * Np32;
  int64_t grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  kron_pack_t_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      W, N, Np32, s.Kk, reinterpret_cast<const int4*>(table), nchunks, WpT);
  return check_launch("kron_pack_t_kernel");
}

extern "C" size_t mml_kron_dgrad_workspace_bytes(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3) {
  if (B < 1 || N < 1 || d1 < 1 || d2 < 1 || d3 < 0) return 0;
  const DgPlan p = make_dg_plan(B, N, d1, d2, d3);
  return p.part_bytes + p.ft_bytes + 256;
}

extern "C" int mml_kron_linear_dgrad(const float* f1, const float* f2, const float* f3, int64_t B, int32_t d1, int32_t d2,
                                     int32_t d3, const int32_t* table, const float* WpT, const float* dy, int32_t N,
                                     float drop_p, uint64_t seed, const uint64_t* seed_dev, int32_t training, float* df1, float* df2, float* df3,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  MML_REQUIRE(f1 && f2 && table && WpT && dy && df1 && df2 && workspace, MML_ERR_INVALID_ARG, "kron_linear_dgrad: null pointer");
  MML_REQUIRE((d3 > 0) == (f3 != nullptr) && (d3 > 0) == (df3 != nullptr), MML_ERR_INVALID_ARG,
              "kron_linear_dgrad: f3/df3 and d3 must all be set or all be absent");
  MML_REQUIRE(B >= 1 && d1 >= 1 && d2 >= 1 && d3 >= 0 && N >= 1, MML_ERR_INVALID_ARG, "kron_linear_dgrad: bad sizes");
  MML_REQUIRE(aligned16(table) && (reinterpret_cast<uintptr_t>(WpT) & 127u) == 0 && aligned16(workspace), MML_ERR_INVALID_ARG,
              "kron_linear_dgrad: table / workspace must be 16-byte and WpT 128-byte aligned");
  const DgPlan p = make_dg_plan(B, N, d1, d2, d3);
  MML_REQUIRE(p.ok, MML_ERR_UNSUPPORTED, "kron_linear_dgrad: N=%d / factor widths (%d,%d,%d) exceed the tile budget", N, d1, d2, d3);
  MML_REQUIRE(workspace_bytes >= p.part_bytes + p.ft_bytes, MML_ERR_WORKSPACE, "kron_linear_dgrad: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MML_CUDA(cudaMemsetAsync(workspace, 0, p.part_bytes, st));
  float* FT = reinterpret_cast<float*>(static_cast<char*>(workspace) + p.part_bytes);
  {
    const dim3 gridf(static_cast<unsigned>(p.Bpad / 32), (p.ft_rows + 31) / 32);
    kron_transpose_factors_kernel<<<gridf, dim3(32, 8), 0, st>>>(f1, f2, f3, B, d1, d2, d3, p.Bpad, p.ft_rows, FT);
    const int rc = check_launch("kron_transpose_factors_kernel");
    if (rc != MML_OK) return rc;
  }
  CUtensorMap tmap;
  int rc = get_tensor_map_2d(WpT, p.Np32, p.Kp, kDgBoxN, kDgTileK, true, &tmap);
  if (rc != MML_OK) return rc;
  const KronShape s = make_kron_shape(d1, d2, d3);
  DgArgs a{};
  a.FT = FT; a.dy = dy;
  a.table = reinterpret_cast<const int4*>(table);
  a.part = static_cast<float*>(workspace);
  a.B = B; a.Bpad = p.Bpad; a.d1 = d1; a.d2 = d2; a.d3 = d3; a.dsum = p.dsum; a.N = N; a.Np32 = p.Np32; a.nchunks = p.nchunks;
  a.ktiles = p.ktiles; a.tiles_per_split = p.tiles_per_split;
  a.n_scal = p.n_scal; a.stages = p.stages; a.bps = p.bps; a.tmem_cols = p.tmem_cols; a.table_in_smem = p.table_in_smem;
  a.fold_mode = p.fold_mode;
  a.idesc = make_idesc_tf32(kTileM, kDgTileK);
  a.dr = make_kron_dropout(drop_p, seed, training, s.Kk, seed_dev);
  const dim3 grid(static_cast<unsigned>((B + kTileM - 1) / kTileM), p.ksplit);
  if (a.dr.thresh != 0u) {
    MML_CUDA(cudaFuncSetAttribute(kron_dgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem)));
    kron_dgrad_tc_kernel<true><<<grid, kDgThreads, p.smem, st>>>(tmap, a);
  } else {
    MML_CUDA(cudaFuncSetAttribute(kron_dgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem)));
    kron_dgrad_tc_kernel<false><<<grid, kDgThreads, p.smem, st>>>(tmap, a);
  }
  rc = check_launch("kron_dgrad_tc_kernel");
  if (rc != MML_OK) return rc;
  const int64_t total = B * p.dsum;
  int64_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  kron_dgrad_reduce_kernel<<<static_cast<unsigned>(g), 256, 0, st>>>(a.part, p.ksplit, B, d1, d2, d3, df1, df2, df3);
  return check_launch("kron_dgrad_reduce_kernel");
}
