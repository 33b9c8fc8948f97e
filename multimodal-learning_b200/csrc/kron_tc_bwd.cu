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
//   * per chunk two [1 x 32] boxes -- the rows of its scalars R[p], R[q] -- and
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
constexpr uint32_t kWgXBytes = 4 * kBlkB * 32 * 4;     // 4 vector-segment slots of [32 rows x 32 b] fp32
constexpr uint32_t kWgSBytes = 8 * kBlkB * 4;          // 8 scalar rows of 32 b
constexpr float kTruncComp = 1.0f + 0.69f / 2048.0f;   // see header comment

struct WgArgs {
  const int4* table;
  float* part;                // [bsplit][N][Kp]
  int64_t B;
  int32_t d1, d2, d3;
  int32_t N, Np, nchunks, Kp;
  int32_t nblocks, blocks_per_split;
  int32_t stages, tmem_cols;
  uint32_t idesc;
  KronDropout dr;
};

template <bool kDropout>
__global__ void __launch_bounds__(kThreadsTc, 1)
kron_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dyT, const __grid_constant__ CUtensorMap tmap_x,
                     const __grid_constant__ CUtensorMap tmap_s, const WgArgs a) {
  uint32_t seed_lo = 0u, seed_hi = 0u;
  if (kDropout) kron_seed(a.dr, seed_lo, seed_hi);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t dy_bytes = static_cast<uint32_t>(a.Np) * 128u;
  const uint32_t stage_bytes = dy_bytes + kWgXBytes + 1024u;         // dy^T tile | X slots | S rows (1 KB, 1024-aligned)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(a.stages) * stage_bytes);
  uint64_t* bar_full = bars;                    // [stages] A^T stored (8 warp arrivals) + dy^T landed (1 arrival + tx)
  uint64_t* bar_empty = bars + a.stages;        // [stages] MMAs of the stage complete
  uint64_t* bar_xfull = bars + 2 * a.stages;    // [stages] factor boxes landed (1 arrival + tx)
  uint64_t* bar_acc = bars + 3 * a.stages;
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
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&bar_full[s], kGenWarps + 1);
      mbar_init(&bar_empty[s], 1);
      mbar_init(&bar_xfull[s], 1);
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

  if (warp == kGenWarps) {
    // ===== TMA producer: per stage the factor boxes (generators) and the dy^T tile (MMA) =====
    int s = 0;
    uint32_t ph = 0;
    for (int blk = blk_begin; blk < blk_end; ++blk) {
      mbar_wait(&bar_empty[s], ph ^ 1);
      if (elect_one_sync()) {
        uint8_t* st = smem + static_cast<size_t>(s) * stage_bytes;
        const int b0 = blk * kBlkB;
        mbar_arrive_expect_tx(&bar_xfull[s], static_cast<uint32_t>(nown) * (kBlkB * 32 * 4) + kWgSBytes);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tma_load_2d(st + dy_bytes + kWgXBytes + (2 * c) * 128, &tmap_s, b0, rp[c], &bar_xfull[s]);
          tma_load_2d(st + dy_bytes + kWgXBytes + (2 * c + 1) * 128, &tmap_s, b0, rq[c], &bar_xfull[s]);
          if (own[c]) tma_load_2d(st + dy_bytes + c * (kBlkB * 32 * 4), &tmap_x, b0, xrow[c], &bar_xfull[s]);
        }
        mbar_arrive_expect_tx(&bar_full[s], dy_bytes);
        tma_load_2d(st, &tmap_dyT, b0, 0, &bar_full[s]);
      }
      __syncwarp();
      if (++s == a.stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == kGenWarps + 1) {
    // ===== MMA issuer =====
    int s = 0;
    uint32_t ph = 0;
    const uint32_t b_base = smem_u32(smem);
    for (int blk = blk_begin; blk < blk_end; ++blk) {
      mbar_wait(&bar_full[s], ph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t b_desc = umma_desc_k_sw128(b_base + static_cast<uint32_t>(s) * stage_bytes);
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
    const uint32_t off_sp = dy_bytes + kWgXBytes + (2 * ci) * 128 + half * 64;
    const uint32_t off_sq = off_sp + 128;
    const uint32_t off_x = dy_bytes + my_slot * (kBlkB * 32 * 4) + e * 128;
    const uint32_t sw = static_cast<uint32_t>(e & 7);
    int s = 0, s_prev = -1;
    uint32_t ph = 0;
    for (int blk = blk_begin; blk < blk_end; ++blk) {
      const uint8_t* st = smem + static_cast<size_t>(s) * stage_bytes;
      mbar_wait(&bar_xfull[s], ph);                    // implies the stage's previous MMAs are complete (producer waited)
      uint32_t r[kHalf];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 p4 = *reinterpret_cast<const float4*>(st + off_sp + u * 16);
        const float4 q4 = *reinterpret_cast<const float4*>(st + off_sq + u * 16);
        const float4 x4 = *reinterpret_cast<const float4*>(st + off_x + (((half * 4 + u) ^ sw) << 4));
        float y0 = p4.x * q4.x * x4.x, y1 = p4.y * q4.y * x4.y, y2 = p4.z * q4.z * x4.z, y3 = p4.w * q4.w * x4.w;
        if (kDropout) {
          const int64_t bb = static_cast<int64_t>(blk) * kBlkB + half * kHalf + u * 4;
          float* yy[4] = {&y0, &y1, &y2, &y3};
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            const int64_t cc = (bb + w) * a.dr.pairs_per_row + (my_klog >> 1);
            const uint32_t h = kron_hash(static_cast<uint32_t>(cc), static_cast<uint32_t>(static_cast<uint64_t>(cc) >> 32),
                                         seed_lo, seed_hi);
            const uint32_t r16 = (my_klog & 1) ? (h >> 16) : (h & 0xffffu);
            *yy[w] = (r16 >= a.dr.thresh) ? *yy[w] * a.dr.scale : 0.f;
          }
        }
        r[u * 4 + 0] = __float_as_uint(y0);
        r[u * 4 + 1] = __float_as_uint(y1);
        r[u * 4 + 2] = __float_as_uint(y2);
        r[u * 4 + 3] = __float_as_uint(y3);
      }
      if (s_prev >= 0) {                                 // publish the PREVIOUS stage: its TMEM store had this stage's
        tc_wait_st();                                    // loads and multiplies to complete in
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_full[s_prev]);
      }
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
  int32_t nchunks, Np, Kp, stages, tmem_cols, ktiles, nblocks, bsplit, blocks_per_split, ft_rows;
  int64_t Bpad;
  size_t smem, dyT_bytes, ft_bytes, part_bytes;
  bool ok;
};

// Number of batch splits: fill whole waves of CTAs, pay for every split's partial tile (written and re-read once).
inline int32_t pick_split(int64_t units, int64_t tiles, int64_t slots, int64_t min_units, double split_cost, int64_t max_split) {
  int64_t best = 1;
  double best_cost = 1e300;
  for (int64_t sp = 1; sp <= max_split; ++sp) {
    const int64_t per = (units + sp - 1) / sp;
    if (sp > 1 && per < min_units) break;
    const int64_t waves = (tiles * sp + slots - 1) / slots;
    const double cost = static_cast<double>(waves) * (static_cast<double>(per) + 4.0) + split_cost * static_cast<double>(sp - 1);
    if (cost < best_cost - 1e-9) {
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
  const size_t fixed = 512 + 1024;
  const size_t stage = static_cast<size_t>(p.Np) * 128 + kWgXBytes + 1024;
  int stages = static_cast<int>((227 * 1024 - fixed) / stage);
  if (stages > 8) stages = 8;
  while (stages > 2 && p.Np + stages * kBlkB > 512) --stages;
  // two CTAs per SM (their generators and MMAs interleave) when each still gets a deep enough ring
  const size_t half_budget = 113 * 1024;
  if (fixed + 3 * stage <= half_budget && p.Np + 3 * kBlkB <= 256) {
    int st2 = static_cast<int>((half_budget - fixed) / stage);
    while (p.Np + st2 * kBlkB > 256) --st2;
    stages = st2;
  }
  p.stages = stages;
  p.ok = p.Np <= 256 && stages >= 2;
  p.smem = fixed + static_cast<size_t>(stages) * stage;
  int cols = p.Np + stages * kBlkB, pow2 = 32;
  while (pow2 < cols) pow2 <<= 1;
  p.tmem_cols = pow2;
  if (pow2 > 512) p.ok = false;
  const int ctas_per_sm = (p.smem <= half_budget && pow2 <= 256) ? 2 : 1;
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
  WgArgs a{};
  a.table = reinterpret_cast<const int4*>(table);
  a.part = part;
  a.B = B; a.d1 = d1; a.d2 = d2; a.d3 = d3; a.N = N; a.Np = p.Np; a.nchunks = p.nchunks; a.Kp = p.Kp;
  a.nblocks = p.nblocks; a.blocks_per_split = p.blocks_per_split;
  a.stages = p.stages; a.tmem_cols = p.tmem_cols;
  a.idesc = make_idesc_tf32(kTileM, p.Np);
  a.dr = make_kron_dropout(drop_p, seed, training, s.Kk, seed_dev);
  const dim3 grid(p.ktiles, p.bsplit);
  if (a.dr.thresh != 0u) {
    MML_CUDA(cudaFuncSetAttribute(kron_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem)));
    kron_wgrad_tc_kernel<true><<<grid, kThreadsTc, p.smem, st>>>(tmap, tmap_x, tmap_s, a);
  } else {
    MML_CUDA(cudaFuncSetAttribute(kron_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(p.smem)));
    kron_wgrad_tc_kernel<false><<<grid, kThreadsTc, p.smem, st>>>(tmap, tmap_x, tmap_s, a);
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
//   the per-row factor gradients by the epilogue warps (thread = batch row):
//     chunk (p, q, vector segment x):  A = R[p] R[q] x[e]   =>   dx[e]   += dA[e] R[p] R[q]
//                                                               dR[p]   += (sum_e dA[e] x[e]) R[q],   dR[q] likewise
//   A operand = the dy tile, written ONCE per CTA into TMEM (tcgen05.st); B operand = [128 k x 32 n] boxes of
//   the transposed packed weight WpT [Kp, Np32] streamed by TMA; two accumulator tiles alternate so the
//   epilogue of tile t overlaps the MMAs of tile t+1.  Split over k tiles -> per-split partial gradients,
//   summed in split order by kron_dgrad_reduce_kernel.
// =====================================================================================================
namespace mml {
namespace {

constexpr int kDgEpiWarps = 8;            // epilogue warp g: TMEM lanes 32*(g%4).., elements 16*(g/4).. of every chunk
constexpr int kDgEpiThreads = kDgEpiWarps * 32;
constexpr int kDgThreads = kDgEpiThreads + 64;   // + TMA warp + MMA warp
constexpr int kDgTileK = 128;             // packed k per accumulator tile (4 chunks)
constexpr int kDgBoxN = 32;               // n per TMA box / pipeline stage
constexpr uint32_t kDgStageBytes = kDgTileK * kDgBoxN * 4;   // 16 KB
constexpr int kDgDsFloats = 2 * 4 * 2 * kTileM;              // sm_ds [tile parity][chunk][half][row]

struct DgArgs {
  const float* f1;
  const float* f2;
  const float* f3;
  const float* dy;            // [B, N]
  const int4* table;
  float* part;                // [ksplit][B][dsum]  (zero-initialised)
  int64_t B;
  int32_t d1, d2, d3, dsum;
  int32_t N, Np32, nchunks;
  int32_t ktiles, tiles_per_split;
  int32_t n_scal, stages, tmem_cols, table_in_smem;
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

template <bool kDropout>
__global__ void __launch_bounds__(kDgThreads, 1) kron_dgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_wT, const DgArgs a) {
  uint32_t seed_lo = 0u, seed_hi = 0u;
  if (kDropout) kron_seed(a.dr, seed_lo, seed_hi);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sm_b = smem;                                                              // [stages][16 KB]
  float* sm_S = reinterpret_cast<float*>(smem + static_cast<size_t>(a.stages) * kDgStageBytes);   // R values  [n_scal][128]
  float* sm_dR = sm_S + static_cast<size_t>(a.n_scal) * kTileM;                                   // dR accum  [n_scal][128]
  float* sm_ds = sm_dR + static_cast<size_t>(a.n_scal) * kTileM;                                  // per-chunk <dA, x> halves
  int4* sm_tab = reinterpret_cast<int4*>(sm_ds + kDgDsFloats);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_tab + (a.table_in_smem ? 2 * a.nchunks : 0));
  uint64_t* bar_full = bars;                     // [stages] weight box landed
  uint64_t* bar_empty = bars + a.stages;         // [stages] MMAs reading the box done
  uint64_t* bar_acc_full = bars + 2 * a.stages;  // [2] accumulator tile complete
  uint64_t* bar_acc_empty = bar_acc_full + 2;    // [2] epilogue has read the tile out of TMEM (8 warp arrivals)
  uint32_t* sm_tmem = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);

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
  if (warp < kDgEpiWarps) {
    // per-row scalars R = [1, f1, (f2)] transposed into shared memory; row r is served by threads r and r + 128
    const int row = threadIdx.x & (kTileM - 1);
    const int half = threadIdx.x >> 7;
    const int64_t b = b0 + row;
    const bool live = b < a.B;
    const int ns = a.n_scal - 1;
    const int per = (ns + 1) / 2;
    const int lo = half * per, hi = min(ns, lo + per);
    if (half == 0) {
      sm_S[row] = 1.0f;
      sm_dR[row] = 0.f;
    }
    for (int i0 = lo; i0 < hi; i0 += 8) {
      float tmp[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u;
        float x = 0.f;
        if (live && i < hi) x = (i < a.d1) ? __ldg(a.f1 + b * a.d1 + i) : __ldg(a.f2 + b * a.d2 + (i - a.d1));
        tmp[u] = x;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u < hi) {
          sm_S[(1 + i0 + u) * kTileM + row] = tmp[u];
          sm_dR[(1 + i0 + u) * kTileM + row] = 0.f;
        }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm_tmem;
  const uint32_t tmem_acc = tmem_base;                 // 2 x 128 accumulator columns
  const uint32_t tmem_a = tmem_base + 2 * kDgTileK;    // dy tile: Np32 columns
  const int4* tab = a.table_in_smem ? sm_tab : a.table;

  // ---- A operand: this CTA's dy rows -> TMEM, once (epilogue warps), then a CTA-wide sync publishes them ----
  if (warp < kDgEpiWarps) {
    const int row = threadIdx.x & (kTileM - 1);
    const int half = threadIdx.x >> 7;
    const int64_t b = b0 + row;
    const bool live = b < a.B;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    for (int n0 = half * 16; n0 < a.Np32; n0 += 32) {
      uint32_t r[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int n = n0 + u;
        const float x = (live && n < a.N) ? __ldg(a.dy + b * a.N + n) : 0.f;
        r[u] = __float_as_uint(x) + 0x1000u;
      }
      tc_st_32x32b_x16(tmem_a + lane_base + n0, r);
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == kDgEpiWarps) {
    // ===== TMA producer: [128 k x 32 n] boxes of WpT (whole warp walks the ring, one elected lane issues) =====
    int s = 0;
    uint32_t ph = 0;
    for (int t = t_begin; t < t_end; ++t)
      for (int j = 0; j < nbox; ++j) {
        mbar_wait(&bar_empty[s], ph ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&bar_full[s], kDgStageBytes);
          tma_load_2d(sm_b + static_cast<size_t>(s) * kDgStageBytes, &tmap_wT, j * kDgBoxN, t * kDgTileK, &bar_full[s]);
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
      for (int j = 0; j < nbox; ++j) {
        mbar_wait(&bar_full[s], ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t b_desc = umma_desc_k_sw128(b_base + static_cast<uint32_t>(s) * kDgStageBytes);
          const uint32_t a_col = tmem_a + j * kDgBoxN;
#pragma unroll
          for (int i = 0; i < kDgBoxN / 8; ++i)
            tc_mma_tf32_ts(tmem_acc + buf * kDgTileK, a_col + i * 8, b_desc + 2 * i, a.idesc, (j > 0 || i > 0) ? 1u : 0u);
          tc_commit(&bar_empty[s]);
          if (j == nbox - 1) tc_commit(&bar_acc_full[buf]);
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
    float v[kHalf], dv[kHalf];
#pragma unroll
    for (int e = 0; e < kHalf; ++e) { v[e] = 0.f; dv[e] = 0.f; }
    int cur_src = -1, cur_col = -1, cur_len = 0;
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
        float* acc_r = sm_dR + static_cast<size_t>((cur_src == 1 ? 1 : 1 + a.d1) + cur_col + eb) * kTileM + row;
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
    auto new_segment = [&](const int4& e0, const int4& e1) {      // rare: once per run of chunks sharing a vector segment
      flush();
      cur_src = e0.z; cur_col = e0.w; cur_len = e1.x;
      const float* src = cur_src == 1 ? a.f1 : (cur_src == 2 ? a.f2 : a.f3);
      const int d = cur_src == 1 ? a.d1 : (cur_src == 2 ? a.d2 : a.d3);
#pragma unroll
      for (int e = 0; e < kHalf; ++e) {
        float x = 0.f;
        if (cur_src == 0) x = (eb + e == 0) ? 1.0f : 0.f;
        else if (live && eb + e < cur_len) x = __ldg(src + b * d + cur_col + eb + e);
        v[e] = x;
        dv[e] = 0.f;
      }
    };
    for (int t = t_begin; t < t_end; ++t) {
      const int it = t - t_begin;
      const int buf = it & 1;
      float* ds_slot = sm_ds + (it & 1) * (4 * 2 * kTileM) + half * kTileM + row;     // + c * 2 * kTileM
      mbar_wait(&bar_acc_full[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_acc + lane_base + buf * kDgTileK + eb;
      // The tile's four chunks, fully unrolled over two register buffers: the tcgen05.ld of chunk c+1 is in flight while
      // chunk c is folded, and no register is copied (a rolled loop has to move the 16 loaded words every iteration).
      uint32_t accA[kHalf], accB[kHalf];
#define MML_DG_CHUNK(C, ACC, NEXT_LD)                                                                   \
      {                                                                                                 \
        const int cg = t * 4 + (C);                                                                     \
        const bool cvalid = cg < a.nchunks;                                                             \
        const int4 e0 = cvalid ? tab[2 * cg] : make_int4(0, 0, 0, 0);                                   \
        const int4 e1 = cvalid ? tab[2 * cg + 1] : make_int4(0, 0, 1, 0);                               \
        if (cvalid && (e0.z != cur_src || e0.w != cur_col)) new_segment(e0, e1);                        \
        const float sp = sm_S[e0.x * kTileM + row], sq = sm_S[e0.y * kTileM + row];                     \
        float s = sp * sq;                                                                              \
        if (kDropout) s *= a.dr.scale;                                                                  \
        tc_wait_ld();                                                                                   \
        NEXT_LD;                                                                                        \
        float g[kHalf];                                                                                 \
        _Pragma("unroll") for (int e = 0; e < kHalf; ++e) g[e] = __uint_as_float(ACC[e]);               \
        if (kDropout) {                                                                                 \
          _Pragma("unroll") for (int e = 0; e < kHalf; ++e) {                                           \
            const int klog = e1.y + (eb + e) * e1.z;                                                    \
            const int64_t cc = b * a.dr.pairs_per_row + (klog >> 1);                                    \
            const uint32_t h = kron_hash(static_cast<uint32_t>(cc),                                     \
                                         static_cast<uint32_t>(static_cast<uint64_t>(cc) >> 32), seed_lo, seed_hi); \
            const uint32_t r16 = (klog & 1) ? (h >> 16) : (h & 0xffffu);                                \
            g[e] = (r16 >= a.dr.thresh) ? g[e] : 0.f;                                                   \
          }                                                                                             \
        }                                                                                               \
        float ds0 = 0.f, ds1 = 0.f, ds2 = 0.f, ds3 = 0.f;                                               \
        _Pragma("unroll") for (int e = 0; e < kHalf; e += 4) {                                          \
          ffma2(ds0, ds1, g[e], g[e + 1], v[e], v[e + 1]);                                              \
          ffma2(ds2, ds3, g[e + 2], g[e + 3], v[e + 2], v[e + 3]);                                      \
          ffma2(dv[e], dv[e + 1], g[e], g[e + 1], s, s);                                                \
          ffma2(dv[e + 2], dv[e + 3], g[e + 2], g[e + 3], s, s);                                        \
        }                                                                                               \
        ds_slot[(C) * 2 * kTileM] = cvalid ? (ds0 + ds1) + (ds2 + ds3) : 0.f;                           \
      }
      tc_ld_32x32b_x16(t_addr, accA);
      MML_DG_CHUNK(0, accA, tc_ld_32x32b_x16(t_addr + 32, accB))
      MML_DG_CHUNK(1, accB, tc_ld_32x32b_x16(t_addr + 64, accA))
      MML_DG_CHUNK(2, accA, tc_ld_32x32b_x16(t_addr + 96, accB))
      MML_DG_CHUNK(3, accB, (void)0)
      tc_fence_before();                    // this warp's TMEM reads of the tile are complete (wait::ld above)
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_acc_empty[buf]);
      asm volatile("bar.sync 1, %0;" ::"n"(kDgEpiThreads) : "memory");      // both halves of every row have posted <dA, x>
      if (half == 0) {
        // dR[p] += <dA, x> R[q],  dR[q] += <dA, x> R[p]; one thread per row owns the dR columns -> fixed summation order
        const float* dsr = sm_ds + (it & 1) * (4 * 2 * kTileM) + row;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int cg = t * 4 + c;
          if (cg >= a.nchunks) break;
          const int4 e0 = tab[2 * cg];
          if ((e0.x | e0.y) == 0) continue;
          float ds = dsr[c * 2 * kTileM] + dsr[c * 2 * kTileM + kTileM];
          if (kDropout) ds *= a.dr.scale;
          const float sp = sm_S[e0.x * kTileM + row], sq = sm_S[e0.y * kTileM + row];
          if (e0.x != 0) sm_dR[e0.x * kTileM + row] += ds * sq;
          if (e0.y != 0) sm_dR[e0.y * kTileM + row] += ds * sp;
        }
      }
    }
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
  int32_t nchunks, Np32, Kp, n_scal, stages, tmem_cols, ktiles, ksplit, tiles_per_split, table_in_smem, dsum;
  size_t smem, part_bytes;
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
  const size_t table_bytes = static_cast<size_t>(p.nchunks) * sizeof(Chunk);
  p.table_in_smem = table_bytes <= static_cast<size_t>(kMaxSmemTable) ? 1 : 0;
  const size_t fixed = 2 * static_cast<size_t>(p.n_scal) * kTileM * sizeof(float) + kDgDsFloats * sizeof(float) +
                       (p.table_in_smem ? table_bytes : 0) + 256 + 1024;
  int stages = fixed + 2 * kDgStageBytes <= 227 * 1024 ? static_cast<int>((227 * 1024 - fixed) / kDgStageBytes) : 0;
  if (stages > 6) stages = 6;
  p.stages = stages;
  p.tmem_cols = (2 * kDgTileK + p.Np32) <= 256 ? 256 : 512;
  p.ok = p.Np32 <= 256 && stages >= 2;
  p.smem = fixed + static_cast<size_t>(stages > 0 ? stages : 0) * kDgStageBytes;
  const int64_t mtiles = (B + kTileM - 1) / kTileM;
  int64_t ks = mtiles < 148 ? (148 + mtiles - 1) / mtiles : 1;
  const int64_t max_ks = p.ktiles / 2 > 0 ? p.ktiles / 2 : 1;
  if (ks > max_ks) ks = max_ks;
  if (ks > 32) ks = 32;
  p.tiles_per_split = static_cast<int32_t>((p.ktiles + ks - 1) / ks);
  p.ksplit = (p.ktiles + p.tiles_per_split - 1) / p.tiles_per_split;
  p.part_bytes = static_cast<size_t>(p.ksplit) * B * p.dsum * sizeof(float);
  return p;
}

// TMA descriptor of WpT [Kp rows, Np32 cols]: boxes of 32 n x 128 k
int get_tensor_map_wT(const float* WpT, int32_t Np32, int32_t Kp, CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  std::lock_guard<std::mutex> lock(mu);
  const MapKey key{WpT, Np32, Kp};
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return MML_OK; }
  EncodeTiledFn enc = get_encode_fn();
  MML_REQUIRE(enc != nullptr, MML_ERR_CUDA, "kron: cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMap m;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(Np32), static_cast<cuuint64_t>(Kp)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(Np32) * sizeof(float)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kDgBoxN), static_cast<cuuint32_t>(kDgTileK)};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(WpT), gdim, gstride, box, estride,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MML_REQUIRE(r == CUDA_SUCCESS, MML_ERR_CUDA, "kron: cuTensorMapEncodeTiled(WpT) failed (%d)", static_cast<int>(r));
  if (cache.size() > 256) cache.clear();
  cache[key] = m;
  *out = m;
  return MML_OK;
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
  return make_dg_plan(B, N, d1, d2, d3).part_bytes + 256;
}

extern "C" int mml_kron_linear_dgrad(const float* f1, const float* f2, const float* f3, int64_t B, int32_t d1, int32_t d2,
                                     int32_t d3, const int32_t* table, const float* WpT, const float* dy, int32_t N,
                                     float drop_p, uint64_t seed, const uint64_t* seed_dev, int32_t training, float* df1, float* df2, float* df3,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  MML_REQUIRE(f1 && f2 && table && WpT && dy && df1 && df2 && workspace, MML_ERR_INVALID_ARG, "kron_linear_dgrad: null pointer");
  MML_REQUIRE((d3 > 0) == (f3 != nullptr) && (d3 > 0) == (df3 != nullptr), MML_ERR_INVALID_ARG,
              "kron_linear_dgrad: f3/df3 and d3 must all be set or all be absent");
  MML_REQUIRE(B >= 1 && d1 >= 1 && d2 >= 1 && d3 >= 0 && N >= 1, MML_ERR_INVALID_ARG, "kron_linear_dgrad: bad sizes");
  MML_REQUIRE(aligned16(table) && (reinterpret_cast<uintptr_t>(WpT) & 127u) == 0, MML_ERR_INVALID_ARG,
              "kron_linear_dgrad: table must be 16-byte and WpT 128-byte aligned");
  const DgPlan p = make_dg_plan(B, N, d1, d2, d3);
  MML_REQUIRE(p.ok, MML_ERR_UNSUPPORTED, "kron_linear_dgrad: N=%d / factor widths (%d,%d,%d) exceed the tile budget", N, d1, d2, d3);
  MML_REQUIRE(workspace_bytes >= p.part_bytes, MML_ERR_WORKSPACE, "kron_linear_dgrad: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MML_CUDA(cudaMemsetAsync(workspace, 0, p.part_bytes, st));
  CUtensorMap tmap;
  int rc = get_tensor_map_wT(WpT, p.Np32, p.Kp, &tmap);
  if (rc != MML_OK) return rc;
  const KronShape s = make_kron_shape(d1, d2, d3);
  DgArgs a{};
  a.f1 = f1; a.f2 = f2; a.f3 = f3; a.dy = dy;
  a.table = reinterpret_cast<const int4*>(table);
  a.part = static_cast<float*>(workspace);
  a.B = B; a.d1 = d1; a.d2 = d2; a.d3 = d3; a.dsum = p.dsum; a.N = N; a.Np32 = p.Np32; a.nchunks = p.nchunks;
  a.ktiles = p.ktiles; a.tiles_per_split = p.tiles_per_split;
  a.n_scal = p.n_scal; a.stages = p.stages; a.tmem_cols = p.tmem_cols; a.table_in_smem = p.table_in_smem;
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
