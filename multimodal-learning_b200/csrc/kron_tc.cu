// K1: Kronecker-fusion encoder forward on the 5th-gen tensor cores (tcgen05 / TMEM / TMA).
//
//   y[b, n] = sum_k A[b,k] m[b,k] W[n,k] + bias[n]        encoder1[0] of fusion.py:29,60 / :94,129
//
// The (d+1)^2 / (d+1)^3 Kronecker operand A (fusion.py:58, :126-127) and its dropout copy
// (fusion.py:59, :128) are never written to HBM: each CTA owns 128 batch rows, its 128 generator
// threads (one per row = one per TMEM lane) compute 32-wide slices  scalar[b] * vector[b, 0..31]
// of A in registers and store them straight into TENSOR MEMORY (tcgen05.st), where the MMA reads
// its A operand (tcgen05.mma kind::tf32, A from TMEM, B from shared memory).  The weight is
// repacked once per weight version into the same chunk order ([Np, 32*chunks] fp32, TF32-rounded,
// 16-byte row pitch) so TMA can stream [Np x 32] tiles (128B-swizzled, K-major) into a ring.
// The fp32 accumulator [128 x Np] lives in TMEM and is read back once (tcgen05.ld) for the epilogue.
//
// Chunk order ("K permutation"): A's columns are regrouped into chunks of <= 32 logical k that
// share one per-row scalar:  core block  o1[i] (x o2[j]) * o_last[32-wide segment],  then the faces /
// edges that contain an appended 1, then the corner (1*1[*1]).  A chunk is (p, q, vsrc, vcol, vlen,
// kbase, kstride): scalar = R[p]*R[q] with R = [1, f1.., f2..] per row, vector = f_vsrc[vcol + t],
// logical k of element t = kbase + t*kstride (used for weight packing and the dropout counter).
//
// Warp roles (192 threads): warps 0-3 generate A / run the epilogue, warp 4 issues TMA,
// warp 5 allocates TMEM and issues the MMAs.  Operand precision: TF32 (10-bit mantissa, fp32 range,
// both operands rounded to nearest) with fp32 accumulation -> rel error ~3e-4 (tolerance 2e-3).
#include "kron_tc_common.cuh"

namespace mml {
namespace {

using namespace tc;

struct TcArgs {
  const float* f1;
  const float* f2;
  const float* f3;
  const int4* table;        // [nchunks][2]
  const float* bias;        // may be NULL
  float* out;               // y [B,N] (ksplit == 1) or partials [ksplit, B, N]
  float* col_stats;         // optional [gridDim.x][2][N]: per-tile column sums of y and y^2 (ksplit == 1 only)
  int64_t B;
  int32_t d1, d2, d3;
  int32_t N, Np;
  int32_t nchunks, chunks_per_split, ksplit;
  int32_t n_scal;           // 1 + d1 (+ d2 when trilinear)
  int32_t stages;
  int32_t tmem_cols;
  int32_t table_in_smem;
  uint32_t idesc;
  KronDropout dr;
};

// kCps = chunks per pipeline stage (1 or 2).  With narrow outputs (Np <= 128) one chunk is only 2*Np <= 256 tensor-pipe
// cycles, about what one round trip through the stage barriers costs the issuing threads: two chunks per stage halve
// the number of waits / fences / arrivals / commits per MMA.
template <bool kDropout, int kCps>
__global__ void __launch_bounds__(kThreadsTc, 1) kron_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_w, const TcArgs a) {
  uint32_t seed_lo = 0u, seed_hi = 0u;
  if (kDropout) kron_seed(a.dr, seed_lo, seed_hi);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: B stages | scalars (transposed: [n_scal][128]) | chunk table | barriers | tmem base
  const uint32_t tile_bytes = static_cast<uint32_t>(a.Np) * 128u;          // one [Np x 32] weight tile
  const uint32_t stage_bytes = tile_bytes * kCps;
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);     // 1024-B aligned, still a shared pointer
  uint8_t* sm_b = smem;
  float* sm_S = reinterpret_cast<float*>(smem + static_cast<size_t>(a.stages) * stage_bytes);
  int4* sm_tab = reinterpret_cast<int4*>(sm_S + static_cast<size_t>(a.n_scal) * kTileM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm_tab + (a.table_in_smem ? 2 * a.nchunks : 0));
  uint64_t* bar_full = bars;                    // [stages]  A stored (1 arrival per generator warp) + B landed (1 arrival + tx bytes)
  uint64_t* bar_empty = bars + a.stages;        // [stages]  MMAs that read the stage have completed
  uint64_t* bar_acc = bars + 2 * a.stages;      // accumulator complete
  uint32_t* sm_tmem = reinterpret_cast<uint32_t*>(bars + 2 * a.stages + 1);
  uint32_t* sm_drop = sm_tmem + 1 + ((16u - ((smem_u32(sm_tmem) + 4u) & 15u)) & 15u) / 4u;   // 16 B aligned; pointer arithmetic keeps it a shared pointer
                                                // [8 warps][2][16] dropout words crossing a warp (kron_drop_words16)

  const int warp = warp_idx_sync();
  const int lane = threadIdx.x & 31;
  const int64_t b0 = static_cast<int64_t>(blockIdx.x) * kTileM;
  const int c_begin = blockIdx.y * a.chunks_per_split;
  const int c_end = min(a.nchunks, c_begin + a.chunks_per_split);

  if (warp == kGenWarps && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w)) : "memory");
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&bar_full[s], kGenWarps + 1);
      mbar_init(&bar_empty[s], 1);
    }
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kGenWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(sm_tmem)), "r"(a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (a.table_in_smem)
    for (int i = c_begin * 2 + threadIdx.x; i < c_end * 2; i += kThreadsTc) sm_tab[i] = __ldg(a.table + i);
  if (warp < kGenWarps) {
    // per-row scalars R = [1, f1, (f2)] transposed into shared memory.  Row r is served by threads r and r+128,
    // each loading half of the scalars, 8 independent loads at a time.
    const int row = threadIdx.x & (kTileM - 1);
    const int half = threadIdx.x >> 7;
    const int64_t b = b0 + row;
    const bool live = b < a.B;
    const int ns = a.n_scal - 1;
    const int per = (ns + 1) / 2;
    const int lo = half * per, hi = min(ns, lo + per);
    if (half == 0) sm_S[row] = 1.0f;
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
        if (i0 + u < hi) sm_S[(1 + i0 + u) * kTileM + row] = tmp[u];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sm_tmem;
  const uint32_t tmem_d = tmem_base;                                  // accumulator: columns [0, Np)
  const uint32_t tmem_a = tmem_base + static_cast<uint32_t>(a.Np);    // A ring: stages x 32 columns
  auto tab_at = [&](int i) -> int4 { return a.table_in_smem ? sm_tab[i] : __ldg(a.table + i); };   // explicit LDS / LDG

  if (warp == kGenWarps) {
    // ===================== TMA producer: weight tiles [Np x 32] =====================
    int s = 0;
    uint32_t ph = 0;
    for (int c = c_begin; c < c_end; c += kCps) {
      mbar_wait(&bar_empty[s], ph ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&bar_full[s], stage_bytes);
#pragma unroll
        for (int i = 0; i < kCps; ++i)        // a chunk index past the table reads zero weights (TMA out-of-bounds fill)
          tma_load_2d(sm_b + static_cast<size_t>(s) * stage_bytes + i * tile_bytes, &tmap_w, (c + i) * kChunkK, 0, &bar_full[s]);
      }
      __syncwarp();
      if (++s == a.stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == kGenWarps + 1) {
    // ===================== MMA issuer: the whole warp walks the ring, one elected lane issues =====================
    int s = 0;
    uint32_t ph = 0;
    const uint32_t b_base = smem_u32(sm_b);
    for (int c = c_begin; c < c_end; c += kCps) {
      mbar_wait(&bar_full[s], ph);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int i = 0; i < kCps; ++i) {
          const uint64_t b_desc = umma_desc_k_sw128(b_base + static_cast<uint32_t>(s) * stage_bytes + i * tile_bytes);
          const uint32_t a_col = tmem_a + (s * kCps + i) * kChunkK;
#pragma unroll
          for (int j = 0; j < kChunkK / 8; ++j)                // +8 tf32 along K = +32 B inside the swizzle row = +2 in the descriptor
            tc_mma_tf32_ts(tmem_d, a_col + j * 8, b_desc + 2 * j, a.idesc, (c > c_begin || i > 0 || j > 0) ? 1u : 0u);
        }
        tc_commit(&bar_empty[s]);          // frees the A columns and the B tiles of this stage
      }
      __syncwarp();
      if (++s == a.stages) { s = 0; ph ^= 1; }
    }
    if (elect_one_sync()) tc_commit(bar_acc);
    __syncwarp();
  } else {
    // ===================== A generators: thread = (batch row / TMEM lane, 16-column half of the chunk) =====================
    const int row = threadIdx.x & (kTileM - 1);
    const int half = threadIdx.x >> 7;
    const int64_t b = b0 + row;
    const bool live = b < a.B;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int ebase = half * kHalf;
    float v[kHalf];
#pragma unroll
    for (int u = 0; u < kHalf; ++u) v[u] = 0.f;
    int cur_src = -1, cur_col = -1, s_prev = -1;
    int s = 0;
    uint32_t ph = 0;
    const int64_t row_group = (b0 >> 5) + (warp & 3);      // the warp's 32 rows are one group of the dropout mask
    const uint32_t lane_bit = 1u << lane;
    for (int c0 = c_begin; c0 < c_end; c0 += kCps) {
      uint32_t r[kCps][kHalf];
#pragma unroll
      for (int i = 0; i < kCps; ++i) {
        const int c = c0 + i;
        if (kCps > 1 && c >= c_end) {                      // odd tail: the stage's second chunk does not exist
#pragma unroll
          for (int u = 0; u < kHalf; ++u) r[i][u] = 0u;
          continue;
        }
        const int4 e0 = tab_at(2 * c);
        const int4 e1 = tab_at(2 * c + 1);
        if (e0.z != cur_src || e0.w != cur_col) {          // (re)load this row's vector segment (rare: once per d1 or d1*d2 chunks)
          cur_src = e0.z;
          cur_col = e0.w;
          const float* src = cur_src == 1 ? a.f1 : (cur_src == 2 ? a.f2 : a.f3);
          const int d = cur_src == 1 ? a.d1 : (cur_src == 2 ? a.d2 : a.d3);
#pragma unroll
          for (int u = 0; u < kHalf; ++u) {
            const int e = ebase + u;
            float x = 0.f;
            if (cur_src == 0) x = (e == 0) ? 1.0f : 0.f;
            else if (live && e < e1.x) x = __ldg(src + b * d + cur_col + e);
            v[u] = x;
          }
        }
        float sc = sm_S[e0.x * kTileM + row] * sm_S[e0.y * kTileM + row];
        uint32_t dw[16];
        if (kDropout) {
          sc *= a.dr.scale;
          kron_drop_words16(a.dr, seed_lo, seed_hi, row_group, e1.y + ebase * e1.z, e1.z, sm_drop + (warp * 2 + (c & 1)) * 16, lane, dw);
        }
#pragma unroll
        for (int u = 0; u < kHalf; ++u) {
          float x = sc * v[u];
          if (kDropout) x = (dw[u] & lane_bit) ? 0.f : x;
          r[i][u] = __float_as_uint(x) + 0x1000u;          // round-to-nearest onto the TF32 grid (hardware truncates)
        }
      }
      if (s_prev >= 0) {                                  // publish the PREVIOUS stage: its TMEM stores had this stage's
        tc_wait_st();                                     // arithmetic to complete in
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_full[s_prev]);    // one arrival per generator warp
      }
      mbar_wait(&bar_empty[s], ph ^ 1);
      tc_fence_after();
#pragma unroll
      for (int i = 0; i < kCps; ++i) tc_st_32x32b_x16(tmem_a + lane_base + (s * kCps + i) * kChunkK + ebase, r[i]);
      s_prev = s;
      if (++s == a.stages) { s = 0; ph ^= 1; }
    }
    if (s_prev >= 0) {
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[s_prev]);
    }
    // ===================== epilogue: TMEM accumulator -> registers -> y =====================
    mbar_wait(bar_acc, 0);
    tc_fence_after();
    float* dst = a.out + (static_cast<int64_t>(blockIdx.y) * a.B + b) * a.N;
    const bool add_bias = (a.ksplit == 1) && (a.bias != nullptr);
    const bool vec_ok = (a.N % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.out) & 15u) == 0);
    const int split_col = ((a.Np / 2 + 15) / 16) * 16;               // warps 0-3: [0, split_col), warps 4-7: [split_col, Np)
    const int n_lo = half == 0 ? 0 : split_col;
    const int n_hi = half == 0 ? split_col : a.Np;
    // BatchNorm batch statistics (fusion.py:29): per-column sums of y and y^2 over the tile's 128 rows.  Each warp reduces
    // its 32 rows x 16 columns with a transposing butterfly (31 shuffles per quantity: after the steps over lane bits 4..0
    // lane l holds column n0 + (l & 15)), the four row-warps meet in shared memory (the weight ring is idle by now) and
    // the tile's partial goes to col_stats[tile][2][N]; a finisher kernel sums the tiles in a fixed order.
    const bool stats = a.col_stats != nullptr;
    float* sm_stat = reinterpret_cast<float*>(sm_b);                  // [4 row-warps][2][Np]
    for (int n0 = n_lo; n0 < n_hi; n0 += 16) {
      uint32_t acc[16];
      tc_ld_32x32b_x16(tmem_d + lane_base + n0, acc);
      tc_wait_ld();
      if (stats) {
        float sv[16], sq[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int n = n0 + e;
          const float yv = (live && n < a.N) ? __uint_as_float(acc[e]) + (add_bias ? __ldg(a.bias + n) : 0.f) : 0.f;
          sv[e] = yv;
          sq[e] = yv * yv;
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          sv[e] += __shfl_xor_sync(kFullMask, sv[e], 16);
          sq[e] += __shfl_xor_sync(kFullMask, sq[e], 16);
        }
#pragma unroll
        for (int w = 8; w >= 1; w >>= 1) {                         // halve the value set, keep the half this lane's bit selects
          const bool up = (lane & w) != 0;
#pragma unroll
          for (int e = 0; e < w; ++e) {
            const float s_send = up ? sv[e] : sv[e + w], q_send = up ? sq[e] : sq[e + w];
            const float s_keep = up ? sv[e + w] : sv[e], q_keep = up ? sq[e + w] : sq[e];
            sv[e] = s_keep + __shfl_xor_sync(kFullMask, s_send, w);
            sq[e] = q_keep + __shfl_xor_sync(kFullMask, q_send, w);
          }
        }
        if (lane < 16) {
          sm_stat[((warp & 3) * 2 + 0) * a.Np + n0 + lane] = sv[0];
          sm_stat[((warp & 3) * 2 + 1) * a.Np + n0 + lane] = sq[0];
        }
      }
      if (live) {
        if (vec_ok && n0 + 16 <= a.N) {
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            float4 o;
            o.x = __uint_as_float(acc[e]);     o.y = __uint_as_float(acc[e + 1]);
            o.z = __uint_as_float(acc[e + 2]); o.w = __uint_as_float(acc[e + 3]);
            if (add_bias) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + n0 + e));
              o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            }
            *reinterpret_cast<float4*>(dst + n0 + e) = o;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int n = n0 + e;
            if (n < a.N) dst[n] = __uint_as_float(acc[e]) + (add_bias ? __ldg(a.bias + n) : 0.f);
          }
        }
      }
    }
    if (stats) {
      asm volatile("bar.sync 1, %0;" ::"n"(kGenThreads) : "memory");
      for (int i = threadIdx.x; i < 2 * a.N; i += kGenThreads) {
        const int qd = i / a.N, n = i - qd * a.N;
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) t += sm_stat[(w * 2 + qd) * a.Np + n];
        a.col_stats[(static_cast<int64_t>(blockIdx.x) * 2 + qd) * a.N + n] = t;
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

// sum split-K partials in split order and add the bias
__global__ void kron_reduce_kernel(const float* __restrict__ part, int32_t ksplit, int64_t BN, int32_t N,
                                   const float* __restrict__ bias, float* __restrict__ y) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < BN;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float t = 0.f;
    for (int z = 0; z < ksplit; ++z) t += part[z * BN + i];
    y[i] = t + (bias ? bias[i % N] : 0.f);
  }
}

// dense W[N, Kk] fp32 -> packed Wp[Np, 32*nchunks] in chunk order, TF32-rounded (RN), zero padded
__global__ void kron_pack_kernel(const float* __restrict__ W, int32_t N, int32_t Np, int32_t Kk, const int4* __restrict__ table,
                                 int32_t nchunks, float* __restrict__ Wp) {
  const int64_t total = static_cast<int64_t>(Np) * nchunks * kChunkK;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t kp = i % (static_cast<int64_t>(nchunks) * kChunkK);
    const int n = static_cast<int>(i / (static_cast<int64_t>(nchunks) * kChunkK));
    const int c = static_cast<int>(kp / kChunkK), e = static_cast<int>(kp % kChunkK);
    const int4 e1 = __ldg(table + 2 * c + 1);
    float w = 0.f;
    if (n < N && e < e1.x) w = __ldg(W + static_cast<int64_t>(n) * Kk + e1.y + e * e1.z);
    uint32_t u = __float_as_uint(w);
    u = (u + 0x1000u) & 0xFFFFE000u;
    Wp[i] = __uint_as_float(u);
  }
}

struct TcPlan {
  int32_t nchunks, Np, n_scal, stages, tmem_cols, ksplit, chunks_per_split, table_in_smem, cps;
  size_t smem;
  bool ok;
};

TcPlan make_tc_plan(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3) {
  TcPlan p{};
  p.nchunks = static_cast<int32_t>(build_chunks(d1, d2, d3).size());
  p.Np = round_np(N);
  p.n_scal = 1 + d1 + (d3 > 0 ? d2 : 0);
  const size_t table_bytes = static_cast<size_t>(p.nchunks) * sizeof(Chunk);
  p.table_in_smem = table_bytes <= static_cast<size_t>(kMaxSmemTable) ? 1 : 0;
  const size_t fixed = static_cast<size_t>(p.n_scal) * kTileM * sizeof(float) + (p.table_in_smem ? table_bytes : 0) + 256 + 1024 +
                       1024;      // barriers, alignment slack, dropout word exchange
  const size_t budget = 227 * 1024;
  p.cps = 1;      // two chunks per stage measured SLOWER on B200 (r2i: N=128 0.154 -> 0.194 ms; the wider TMEM ring costs the second CTA per SM)
  const size_t stage = static_cast<size_t>(p.Np) * 128 * p.cps;
  const int stage_cols = kChunkK * p.cps;
  int stages = fixed < budget ? static_cast<int>((budget - fixed) / stage) : 0;
  if (stages > 4) stages = 4;
  while (stages > 2 && p.Np + stages * stage_cols > 512) --stages;
  const size_t half_budget = 113 * 1024;
  if (fixed + 3 * stage <= half_budget && p.Np + 3 * stage_cols <= 256) {     // two CTAs per SM
    int st2 = static_cast<int>((half_budget - fixed) / stage);
    if (st2 > 4) st2 = 4;
    while (p.Np + st2 * stage_cols > 256) --st2;
    stages = st2;
  }
  p.stages = stages;
  p.ok = p.Np <= 256 && stages >= 2;
  p.smem = fixed + static_cast<size_t>(stages > 0 ? stages : 0) * stage;
  int cols = p.Np + stages * stage_cols;
  int pow2 = 32;
  while (pow2 < cols) pow2 <<= 1;
  p.tmem_cols = pow2;
  if (pow2 > 512) p.ok = false;
  // split K when the batch alone cannot fill the 148 SMs
  const int64_t tiles = (B + kTileM - 1) / kTileM;
  int64_t ks = 1;
  if (tiles < 148) ks = (148 + tiles - 1) / tiles;
  const int64_t max_ks = p.nchunks / 8 > 0 ? p.nchunks / 8 : 1;      // keep >= 8 chunks per split
  if (ks > max_ks) ks = max_ks;
  if (ks > 32) ks = 32;
  p.chunks_per_split = static_cast<int32_t>((p.nchunks + ks - 1) / ks);
  p.chunks_per_split = (p.chunks_per_split + p.cps - 1) / p.cps * p.cps;     // splits start on a stage boundary
  p.ksplit = (p.nchunks + p.chunks_per_split - 1) / p.chunks_per_split;
  return p;
}

}  // namespace
}  // namespace mml

using namespace mml;

extern "C" int64_t mml_kron_num_chunks(int32_t d1, int32_t d2, int32_t d3) {
  if (d1 < 1 || d2 < 1 || d3 < 0) return -1;
  return static_cast<int64_t>(build_chunks(d1, d2, d3).size());
}

extern "C" int mml_kron_chunk_table_host(int32_t d1, int32_t d2, int32_t d3, int32_t* table_host) {
  MML_REQUIRE(table_host && d1 >= 1 && d2 >= 1 && d3 >= 0, MML_ERR_INVALID_ARG, "kron_chunk_table: bad arguments");
  const std::vector<Chunk> ch = build_chunks(d1, d2, d3);
  memcpy(table_host, ch.data(), ch.size() * sizeof(Chunk));
  return MML_OK;
}

extern "C" int mml_kron_pack_weight(const float* W, int32_t N, int32_t d1, int32_t d2, int32_t d3, const int32_t* table,
                                    float* Wp, void* stream) {
  MML_REQUIRE(W && table && Wp && N >= 1, MML_ERR_INVALID_ARG, "kron_pack_weight: bad arguments");
  MML_REQUIRE(aligned16(table) && aligned16(Wp), MML_ERR_INVALID_ARG, "kron_pack_weight: table / Wp must be 16-byte aligned");
  const KronShape s = make_kron_shape(d1, d2, d3);
  const int32_t nchunks = static_cast<int32_t>(build_chunks(d1, d2, d3).size());
  const int64_t total = static_cast<int64_t>(round_np(N)) * nchunks * kChunkK;
  int64_t grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  kron_pack_kernel<<<static_cast<unsigned>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      W, N, round_np(N), s.Kk, reinterpret_cast<const int4*>(table), nchunks, Wp);
  return check_launch("kron_pack_kernel");
}

extern "C" int64_t mml_kron_packed_floats(int32_t N, int32_t d1, int32_t d2, int32_t d3) {
  if (N < 1 || d1 < 1 || d2 < 1 || d3 < 0) return -1;
  return static_cast<int64_t>(round_np(N)) * static_cast<int64_t>(build_chunks(d1, d2, d3).size()) * kChunkK;
}

extern "C" int mml_kron_fwd_supported(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3) {
  if (B < 0 || N < 1 || d1 < 1 || d2 < 1 || d3 < 0) return 0;
  return make_tc_plan(B, N, d1, d2, d3).ok ? 1 : 0;
}

extern "C" size_t mml_kron_fwd_workspace_bytes(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3) {
  if (B < 0 || N < 1 || d1 < 1 || d2 < 1 || d3 < 0) return 0;
  const TcPlan p = make_tc_plan(B, N, d1, d2, d3);
  return (p.ksplit > 1 ? static_cast<size_t>(p.ksplit) * B * N * sizeof(float) : 0) + 256;
}

extern "C" int64_t mml_kron_fwd_stat_tiles(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3) {
  if (B < 1 || N < 1 || d1 < 1 || d2 < 1 || d3 < 0) return 0;
  const TcPlan p = make_tc_plan(B, N, d1, d2, d3);
  return (p.ok && p.ksplit == 1) ? (B + kTileM - 1) / kTileM : 0;
}

extern "C" int mml_kron_linear_fwd(const float* f1, const float* f2, const float* f3, int64_t B, int32_t d1, int32_t d2,
                                   int32_t d3, const int32_t* table, const float* Wp, const float* bias, int32_t N,
                                   float drop_p, uint64_t seed, const uint64_t* seed_dev, int32_t training, float* y, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  return mml_kron_linear_fwd_stats(f1, f2, f3, B, d1, d2, d3, table, Wp, bias, N, drop_p, seed, seed_dev, training, y, nullptr,
                                   workspace, workspace_bytes, stream);
}

extern "C" int mml_kron_linear_fwd_stats(const float* f1, const float* f2, const float* f3, int64_t B, int32_t d1, int32_t d2,
                                         int32_t d3, const int32_t* table, const float* Wp, const float* bias, int32_t N,
                                         float drop_p, uint64_t seed, const uint64_t* seed_dev, int32_t training, float* y,
                                         float* col_stats, void* workspace, size_t workspace_bytes, void* stream) {
  MML_REQUIRE(f1 && f2 && table && Wp && y, MML_ERR_INVALID_ARG, "kron_linear_fwd: null pointer");
  MML_REQUIRE((d3 > 0) == (f3 != nullptr), MML_ERR_INVALID_ARG, "kron_linear_fwd: f3 and d3 must both be set or both be absent");
  MML_REQUIRE(B >= 0 && d1 >= 1 && d2 >= 1 && d3 >= 0 && N >= 1, MML_ERR_INVALID_ARG, "kron_linear_fwd: bad sizes");
  MML_REQUIRE(aligned16(table) && (reinterpret_cast<uintptr_t>(Wp) & 127u) == 0, MML_ERR_INVALID_ARG,
              "kron_linear_fwd: table must be 16-byte and Wp 128-byte aligned");
  if (B == 0) return MML_OK;
  const TcPlan p = make_tc_plan(B, N, d1, d2, d3);
  MML_REQUIRE(p.ok, MML_ERR_UNSUPPORTED, "kron_linear_fwd: N=%d (<=256) / factor widths (%d,%d,%d) exceed the tile budget", N,
              d1, d2, d3);
  MML_REQUIRE(workspace_bytes >= mml_kron_fwd_workspace_bytes(B, N, d1, d2, d3) && (p.ksplit == 1 || workspace != nullptr),
              MML_ERR_WORKSPACE, "kron_linear_fwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap tmap;
  int rc = get_tensor_map(Wp, p.Np, p.nchunks * kChunkK, &tmap);
  if (rc != MML_OK) return rc;
  const KronShape s = make_kron_shape(d1, d2, d3);
  TcArgs a{};
  a.f1 = f1; a.f2 = f2; a.f3 = f3;
  a.table = reinterpret_cast<const int4*>(table);
  a.bias = bias;
  a.out = p.ksplit == 1 ? y : static_cast<float*>(workspace);
  MML_REQUIRE(col_stats == nullptr || p.ksplit == 1, MML_ERR_UNSUPPORTED,
              "kron_linear_fwd_stats: column statistics need an unsplit forward (mml_kron_fwd_stat_tiles() > 0)");
  a.col_stats = col_stats;
  a.B = B; a.d1 = d1; a.d2 = d2; a.d3 = d3; a.N = N; a.Np = p.Np;
  a.nchunks = p.nchunks; a.chunks_per_split = p.chunks_per_split; a.ksplit = p.ksplit;
  a.n_scal = p.n_scal; a.stages = p.stages; a.tmem_cols = p.tmem_cols; a.table_in_smem = p.table_in_smem;
  a.idesc = make_idesc_tf32(kTileM, p.Np);
  a.dr = make_kron_dropout(drop_p, seed, training, s.Kk, seed_dev);
  const dim3 grid(static_cast<unsigned>((B + kTileM - 1) / kTileM), p.ksplit);
#define MML_LAUNCH_FWD(DROP, CPS)                                                                                        \
  {                                                                                                                      \
    MML_CUDA(cudaFuncSetAttribute(kron_fwd_tc_kernel<DROP, CPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,             \
                                  static_cast<int>(p.smem)));                                                            \
    kron_fwd_tc_kernel<DROP, CPS><<<grid, kThreadsTc, p.smem, st>>>(tmap, a);                                            \
  }
  if (a.dr.thresh != 0u) {
    if (p.cps == 2) MML_LAUNCH_FWD(true, 2) else MML_LAUNCH_FWD(true, 1)
  } else {
    if (p.cps == 2) MML_LAUNCH_FWD(false, 2) else MML_LAUNCH_FWD(false, 1)
  }
#undef MML_LAUNCH_FWD
  rc = check_launch("kron_fwd_tc_kernel");
  if (rc != MML_OK) return rc;
  if (p.ksplit > 1) {
    const int64_t BN = B * N;
    int64_t g = (BN + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    kron_reduce_kernel<<<static_cast<unsigned>(g), 256, 0, st>>>(static_cast<const float*>(workspace), p.ksplit, BN, N, bias, y);
    rc = check_launch("kron_reduce_kernel");
  }
  return rc;
}
