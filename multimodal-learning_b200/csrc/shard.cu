// Index routing for the row-sharded memory bank ("move indices and scalars, not rows"):
// a stable counting sort of contrast_idx[B, K+1] by owning rank.  One warp per anchor; ballots give
// the per-owner counts and the stable positions, so (anchor, column) order is preserved inside each
// owner's segment -- column 0 (the positive) stays first when its owner holds it.
#include "common.cuh"

namespace mml {
namespace {

constexpr int kRouteThreads = 128;
constexpr int kRouteWarps = kRouteThreads / 32;

// One warp per (anchor b, column chunk c).  counts / offsets are laid out [world][B][chunks].
template <bool kScatter>
__global__ void __launch_bounds__(kRouteThreads) shard_route_kernel(
    const int64_t* __restrict__ idx, int64_t B, int64_t cols, int32_t chunk_cols, int32_t chunks, int64_t rows_per_rank,
    int32_t world, int64_t* __restrict__ counts, const int64_t* __restrict__ offsets, int32_t* __restrict__ out,
    uint32_t* __restrict__ err) {
  __shared__ int64_t run[kRouteWarps][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t w = static_cast<int64_t>(blockIdx.x) * kRouteWarps + warp;     // = b * chunks + c
  if (w >= B * chunks) return;
  const int64_t b = w / chunks;
  const int c = static_cast<int>(w % chunks);
  const int64_t slot = b * chunks + c;
  const int64_t plane = B * chunks;
  if (lane < world) run[warp][lane] = kScatter ? offsets[static_cast<int64_t>(lane) * plane + slot] : 0;
  __syncwarp();
  const unsigned lt = (1u << lane) - 1u;
  const int64_t k_begin = static_cast<int64_t>(c) * chunk_cols;
  const int64_t k_end = min(cols, k_begin + chunk_cols);
  for (int64_t k0 = k_begin; k0 < k_end; k0 += 32) {
    const int64_t k = k0 + lane;
    const bool valid = k < k_end;
    const int64_t row = valid ? idx[b * cols + k] : 0;
    int owner = valid ? static_cast<int>(row / rows_per_rank) : -1;
    if (valid && (row < 0 || owner >= world)) {        // id outside the bank: flagged and dropped (no slot, no write)
      flag_device_error(err, MML_DEVERR_SHARD_OWNER);
      owner = -1;
    }
    for (int o = 0; o < world; ++o) {
      const unsigned m = __ballot_sync(kFullMask, owner == o);
      if (m == 0u) continue;
      const int64_t base = run[warp][o];
      if (kScatter && owner == o) out[base + __popc(m & lt)] = static_cast<int32_t>(row - static_cast<int64_t>(o) * rows_per_rank);
      __syncwarp();
      if (lane == 0) run[warp][o] = base + __popc(m);
      __syncwarp();
    }
  }
  if (!kScatter && lane < world) counts[static_cast<int64_t>(lane) * plane + slot] = run[warp][lane];
}

// Single-pass variant with a FIXED-STRIDE output: slot (o, b, c) owns ids_out[((o*B + b)*chunks + c)*chunk_cols ..] and
// uses the first counts[(o*B + b)*chunks + c] entries.  No scan, no second pass; the owner's gather kernel reads the
// slots straight out of this (peer-mapped) buffer, so the all_to_all and its host-side split sizes disappear.
__global__ void __launch_bounds__(kRouteThreads) shard_route_strided_kernel(
    const int64_t* __restrict__ idx, int64_t B, int64_t cols, int32_t chunk_cols, int32_t chunks, int64_t rows_per_rank,
    int32_t world, int32_t* __restrict__ counts, int32_t* __restrict__ out, uint32_t* __restrict__ err) {
  __shared__ int32_t run[kRouteWarps][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t w = static_cast<int64_t>(blockIdx.x) * kRouteWarps + warp;     // = b * chunks + c
  if (w >= B * chunks) return;
  const int64_t b = w / chunks;
  const int c = static_cast<int>(w % chunks);
  const int64_t slot = b * chunks + c;
  const int64_t plane = B * chunks;
  if (lane < world) run[warp][lane] = 0;
  __syncwarp();
  const unsigned lt = (1u << lane) - 1u;
  const int64_t k_begin = static_cast<int64_t>(c) * chunk_cols;
  const int64_t k_end = min(cols, k_begin + chunk_cols);
  // the divide is by a runtime constant: do it in 32 bits (row < 2^31 is checked on the host side)
  const uint32_t rpr = static_cast<uint32_t>(rows_per_rank);
  // 4 independent coalesced index loads in flight per lane before the (serial, shared-memory) ranking of each of them
  constexpr int kPre = 4;
  for (int64_t kg = k_begin; kg < k_end; kg += 32 * kPre) {
    uint32_t rowv[kPre];
#pragma unroll
    for (int u = 0; u < kPre; ++u) {
      const int64_t k = kg + u * 32 + lane;
      int64_t r64 = (k < k_end) ? __ldg(idx + b * cols + k) : 0;
      if (static_cast<uint64_t>(r64) >= static_cast<uint64_t>(rows_per_rank) * static_cast<uint64_t>(world)) {
        flag_device_error(err, MML_DEVERR_SHARD_OWNER);       // id outside the bank: flagged and dropped below
        r64 = -1;
      }
      rowv[u] = static_cast<uint32_t>(r64);                    // 0xffffffff marks a dropped id (global ids fit in 31 bits)
    }
#pragma unroll
    for (int u = 0; u < kPre; ++u) {
      const int64_t k = kg + u * 32 + lane;
      if (kg + u * 32 >= k_end) break;                 // warp-uniform
      const uint32_t row = rowv[u];
      const bool valid = k < k_end && row != 0xffffffffu;
      const int owner = valid ? static_cast<int>(row / rpr) : -1;
      const uint32_t local = row - static_cast<uint32_t>(owner < 0 ? 0 : owner) * rpr;
      // lanes with the same owner form a group; rank inside the group = stable position
      const unsigned peers = __match_any_sync(kFullMask, owner);
      int32_t base = 0;
      if (valid) {
        base = run[warp][owner];
        out[(static_cast<int64_t>(owner) * plane + slot) * chunk_cols + base + __popc(peers & lt)] = static_cast<int32_t>(local);
      }
      __syncwarp();
      if (valid && lane == __ffs(peers) - 1) run[warp][owner] = base + __popc(peers);
      __syncwarp();
    }
  }
  if (lane < world) counts[static_cast<int64_t>(lane) * plane + slot] = run[warp][lane];
}

// Blocked form of the same routing (identical output, slot by slot): one CTA of 128 threads per slot.
//   1. the slot's ids are read COALESCED (all of a thread's loads in flight at once) and split into (owner, local id) in
//      shared memory -- owner by a multiply-high with floor(2^32 / rows_per_rank) and a fix-up, not a division;
//   2. thread t owns the kIpt consecutive entries t*kIpt.. : the stable position of an id inside its owner's run is
//      (ids of that owner in earlier threads) + (earlier ids of that owner in this thread) -- per-thread counts in a
//      shared [owner][thread] table, exclusive-scanned over the threads owner by owner, then used as cursors;
//   3. ids are placed, in that order, into a shared staging buffer laid out run by run, and
//   4. written out COALESCED, one run per owner.
// Measured (scripts/bench_route.py, 1024 x 16385 ids): 105 us at world 2, 116 us at world 8 -- the same as the
// warp-per-slot kernel above (108 us, serial match_any ranking).  Both are INSTRUCTION-bound, not memory-bound: ncu
// counts 54 M warp instructions (92 per id) at 49 % issue utilisation against 31 us of HBM time; a first blocked version
// with per-thread global loads / stores took 179 us, one with register-packed counters 191-241 us.
template <int kIpt>
__global__ void __launch_bounds__(kRouteThreads) shard_route_strided_blocked_kernel(
    const int64_t* __restrict__ idx, int64_t B, int64_t cols, int32_t chunk_cols, int32_t chunks, int64_t rows_per_rank,
    uint32_t rpr_magic, int32_t world, int32_t* __restrict__ counts, int32_t* __restrict__ out, uint32_t* __restrict__ err) {
  constexpr int kN = kRouteThreads * kIpt;         // ids per slot
  __shared__ __align__(16) uint8_t s_own[kN];
  __shared__ int32_t s_loc[kRouteThreads * (kIpt + 1)];      // row stride kIpt + 1: conflict-free per-thread rows
  __shared__ int32_t s_out[kN];
  __shared__ uint16_t s_cnt[32 * kRouteThreads];   // [owner][thread]: counts, then exclusive prefixes (cursors)
  __shared__ int32_t s_base[33];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t slot = blockIdx.x;                 // = b * chunks + c
  const int64_t b = slot / chunks;
  const int c = static_cast<int>(slot % chunks);
  const int64_t plane = B * chunks;
  const int64_t k_begin = static_cast<int64_t>(c) * chunk_cols;
  const int len = static_cast<int>(min(static_cast<int64_t>(chunk_cols), cols - k_begin));
  const uint32_t rpr = static_cast<uint32_t>(rows_per_rank);
  const uint64_t n_rows = static_cast<uint64_t>(rows_per_rank) * static_cast<uint64_t>(world);
  const int64_t* src = idx + b * cols + k_begin;
  for (int i = tid; i < world * kRouteThreads; i += kRouteThreads) s_cnt[i] = 0;
  bool bad = false;
  int64_t raw[kIpt];                               // all of the thread's loads in flight before the first use
#pragma unroll
  for (int j = 0; j < kIpt; ++j) {
    const int i = j * kRouteThreads + tid;
    raw[j] = __ldg(src + (i < len ? i : len - 1));
  }
#pragma unroll
  for (int j = 0; j < kIpt; ++j) {
    const int i = j * kRouteThreads + tid;
    uint8_t o = 255;
    int32_t l = 0;
    if (i < len) {
      const int64_t r64 = raw[j];
      if (static_cast<uint64_t>(r64) >= n_rows) {
        bad = true;                                // id outside the bank: flagged and dropped
      } else {
        const uint32_t row = static_cast<uint32_t>(r64);
        uint32_t ow = __umulhi(row, rpr_magic);    // floor(row / rpr) or one less (magic = floor(2^32 / rpr))
        uint32_t rem = row - ow * rpr;
        if (rem >= rpr) { ++ow; rem -= rpr; }
        if (rem >= rpr) { ++ow; rem -= rpr; }
        o = static_cast<uint8_t>(ow);
        l = static_cast<int32_t>(rem);
      }
    }
    s_own[i] = o;
    s_loc[(i / kIpt) * (kIpt + 1) + (i % kIpt)] = l;
  }
  if (bad) flag_device_error(err, MML_DEVERR_SHARD_OWNER);
  __syncthreads();
  // per-thread counts of its kIpt consecutive entries
  uint8_t own[kIpt];
#pragma unroll
  for (int j = 0; j < kIpt; ++j) own[j] = s_own[tid * kIpt + j];
#pragma unroll
  for (int j = 0; j < kIpt; ++j)
    if (own[j] != 255) s_cnt[own[j] * kRouteThreads + tid] += 1;      // only this thread touches column `tid`
  __syncthreads();
  // exclusive scan over the 128 threads, owner by owner: warp w takes owners w, w + 4, ...; a lane scans 4 threads' counts
  for (int o = warp; o < world; o += kRouteWarps) {
    uint16_t* row = s_cnt + o * kRouteThreads + lane * 4;
    const uint32_t c0 = row[0], c1 = row[1], c2 = row[2], c3 = row[3];
    uint32_t v = c0 + c1 + c2 + c3;
    const uint32_t mine = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(kFullMask, v, d);
      if (lane >= d) v += t;
    }
    const uint32_t ex = v - mine;
    row[0] = static_cast<uint16_t>(ex);
    row[1] = static_cast<uint16_t>(ex + c0);
    row[2] = static_cast<uint16_t>(ex + c0 + c1);
    row[3] = static_cast<uint16_t>(ex + c0 + c1 + c2);
    if (lane == 31) s_base[o + 1] = static_cast<int32_t>(v);          // the owner's total, turned into run starts below
  }
  __syncthreads();
  if (tid == 0) {
    int32_t acc = 0;
    s_base[0] = 0;
    for (int o = 0; o < world; ++o) {
      acc += s_base[o + 1];
      s_base[o + 1] = acc;
    }
  }
  __syncthreads();
  // place the ids, in order, into the staging buffer (run by run)
#pragma unroll
  for (int j = 0; j < kIpt; ++j) {
    if (own[j] != 255) {
      uint16_t* cur = s_cnt + own[j] * kRouteThreads + tid;
      const int pos = s_base[own[j]] + *cur;
      *cur += 1;
      s_out[pos] = s_loc[tid * (kIpt + 1) + j];
    }
  }
  __syncthreads();
  for (int o = 0; o < world; ++o) {                // coalesced, one run per owner
    const int r0 = s_base[o], r1 = s_base[o + 1];
    int32_t* dst = out + (static_cast<int64_t>(o) * plane + slot) * chunk_cols;
    for (int i = r0 + tid; i < r1; i += kRouteThreads) dst[i - r0] = s_out[i];
  }
  if (tid < world) counts[static_cast<int64_t>(tid) * plane + slot] = s_base[tid + 1] - s_base[tid];
}

int check(const int64_t* idx, int64_t B, int64_t cols, int32_t chunk_cols, int64_t rows_per_rank, int32_t world) {
  MML_REQUIRE(idx && B >= 0 && cols >= 1 && rows_per_rank >= 1, MML_ERR_INVALID_ARG, "shard_route: bad arguments");
  MML_REQUIRE(chunk_cols >= 32 && chunk_cols % 32 == 0, MML_ERR_INVALID_ARG, "shard_route: chunk_cols must be a multiple of 32");
  MML_REQUIRE(world >= 1 && world <= 32, MML_ERR_UNSUPPORTED, "shard_route: world size must be in [1, 32]");
  MML_REQUIRE(rows_per_rank < (1LL << 31), MML_ERR_UNSUPPORTED, "shard_route: local row ids must fit in int32");
  return MML_OK;
}

}  // namespace
}  // namespace mml

using namespace mml;

extern "C" int mml_shard_count(const int64_t* idx, int64_t B, int64_t cols, int32_t chunk_cols, int64_t rows_per_rank,
                               int32_t world, int64_t* counts, void* stream) {
  int rc = check(idx, B, cols, chunk_cols, rows_per_rank, world);
  if (rc != MML_OK) return rc;
  MML_REQUIRE(counts, MML_ERR_INVALID_ARG, "shard_count: null counts");
  if (B == 0) return MML_OK;
  const int32_t chunks = static_cast<int32_t>((cols + chunk_cols - 1) / chunk_cols);
  const int64_t warps = B * chunks;
  shard_route_kernel<false><<<static_cast<unsigned>((warps + kRouteWarps - 1) / kRouteWarps), kRouteThreads, 0,
                              static_cast<cudaStream_t>(stream)>>>(idx, B, cols, chunk_cols, chunks, rows_per_rank, world,
                                                                   counts, nullptr, nullptr, device_error_word());
  return check_launch("shard_route_kernel<count>");
}

extern "C" int mml_shard_scatter(const int64_t* idx, int64_t B, int64_t cols, int32_t chunk_cols, int64_t rows_per_rank,
                                 int32_t world, const int64_t* offsets, int32_t* out_local_ids, void* stream) {
  int rc = check(idx, B, cols, chunk_cols, rows_per_rank, world);
  if (rc != MML_OK) return rc;
  MML_REQUIRE(offsets && out_local_ids, MML_ERR_INVALID_ARG, "shard_scatter: null pointer");
  if (B == 0) return MML_OK;
  const int32_t chunks = static_cast<int32_t>((cols + chunk_cols - 1) / chunk_cols);
  const int64_t warps = B * chunks;
  shard_route_kernel<true><<<static_cast<unsigned>((warps + kRouteWarps - 1) / kRouteWarps), kRouteThreads, 0,
                             static_cast<cudaStream_t>(stream)>>>(idx, B, cols, chunk_cols, chunks, rows_per_rank, world,
                                                                  nullptr, offsets, out_local_ids, device_error_word());
  return check_launch("shard_route_kernel<scatter>");
}

extern "C" int mml_shard_route_strided(const int64_t* idx, int64_t B, int64_t cols, int32_t chunk_cols,
                                       int64_t rows_per_rank, int32_t world, int32_t* counts, int32_t* ids_out,
                                       void* stream) {
  int rc = check(idx, B, cols, chunk_cols, rows_per_rank, world);
  if (rc != MML_OK) return rc;
  MML_REQUIRE(counts && ids_out, MML_ERR_INVALID_ARG, "shard_route_strided: null pointer");
  MML_REQUIRE(rows_per_rank * world < (1LL << 31), MML_ERR_UNSUPPORTED, "shard_route_strided: global row ids must fit in int32");
  if (B == 0) return MML_OK;
  const int32_t chunks = static_cast<int32_t>((cols + chunk_cols - 1) / chunk_cols);
  const int64_t warps = B * chunks;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (chunk_cols % kRouteThreads == 0 && chunk_cols / kRouteThreads <= 16 && rows_per_rank >= 2) {
    const unsigned grid = static_cast<unsigned>(warps);
    const uint32_t magic = static_cast<uint32_t>((1ull << 32) / static_cast<uint64_t>(rows_per_rank));
    switch (chunk_cols / kRouteThreads) {
      case 16: shard_route_strided_blocked_kernel<16><<<grid, kRouteThreads, 0, st>>>(idx, B, cols, chunk_cols, chunks, rows_per_rank, magic, world, counts, ids_out, device_error_word()); break;
      case 8: shard_route_strided_blocked_kernel<8><<<grid, kRouteThreads, 0, st>>>(idx, B, cols, chunk_cols, chunks, rows_per_rank, magic, world, counts, ids_out, device_error_word()); break;
      case 4: shard_route_strided_blocked_kernel<4><<<grid, kRouteThreads, 0, st>>>(idx, B, cols, chunk_cols, chunks, rows_per_rank, magic, world, counts, ids_out, device_error_word()); break;
      case 2: shard_route_strided_blocked_kernel<2><<<grid, kRouteThreads, 0, st>>>(idx, B, cols, chunk_cols, chunks, rows_per_rank, magic, world, counts, ids_out, device_error_word()); break;
      case 1: shard_route_strided_blocked_kernel<1><<<grid, kRouteThreads, 0, st>>>(idx, B, cols, chunk_cols, chunks, rows_per_rank, magic, world, counts, ids_out, device_error_word()); break;
      default:
        shard_route_strided_kernel<<<static_cast<unsigned>((warps + kRouteWarps - 1) / kRouteWarps), kRouteThreads, 0, st>>>(
            idx, B, cols, chunk_cols, chunks, rows_per_rank, world, counts, ids_out, device_error_word());
    }
    return check_launch("shard_route_strided_blocked_kernel");
  }
  shard_route_strided_kernel<<<static_cast<unsigned>((warps + kRouteWarps - 1) / kRouteWarps), kRouteThreads, 0, st>>>(
      idx, B, cols, chunk_cols, chunks, rows_per_rank, world, counts, ids_out, device_error_word());
  return check_launch("shard_route_strided_kernel");
}
