// K4: ContrastMemory gather / score / NCE loss / gradient -- HBM-bound.
//
// Replaces, in ONE pass over the gathered memory rows, the reference's
//   index_select x2 + bmm x2 + exp + div            CL_utils/CRD_criterion.py:41-49,62-63
//   ContrastLoss.forward x2                         CL_utils/CRD_criterion.py:199-216
//   and the autograd of both (closed form).
// Nothing of size [B, K+1, D] is ever written: the rows go HBM -> registers -> FMAs.
//
// Work decomposition: grid = (chunks, B).  CTA (b, c) handles columns
// [c*chunk_cols, (c+1)*chunk_cols) of anchor b with 4 warps.  A row of D = 32*VPL
// floats is read by 8 lanes (lane s loads float4 #(j*8+s), j < VPL: every load
// instruction covers 4 rows x 128 contiguous bytes), so one warp has 4*U rows of
// EACH bank in flight.  The 8-lane dot products are finished with 3 xor-shuffles;
// the per-row scalar maths (exp, log, 1/(x+c)) is packed so the 2*U scalars of a
// row-slot are evaluated by different lanes of that slot in a single SIMT pass.
// Per-CTA partial results (gradient [2,D], 4 scalars) go to a workspace and are
// reduced in a fixed order by two small finisher kernels -> deterministic output.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace mml {
namespace {

constexpr int kCtaThreads = 128;
constexpr int kCtaWarps = kCtaThreads / 32;
constexpr int kPeerStageIds = 4096;      // routed ids one peer-mode CTA stages in shared memory (16 KB)
constexpr int kPeerMaxGroup = 32;        // slots per CTA the staging prologue handles (one lane per slot)

enum Mode : int { kFused = 0, kScores = 1, kWeighted = 2 };

struct GatherArgs {
  const float* bank1;
  const float* bank2;
  const float* v1;
  const float* v2;
  const int64_t* idx;       // int64 indices (the reference's LongTensor) ...
  const int32_t* idx32;     // ... or int32 local row ids (row-sharded bank); exactly one is non-NULL
  const int64_t* seg_ptr;   // NULL -> dense [B, cols]
  const uint8_t* pos_flag;  // NULL -> first entry of every segment is the positive
  const float* Z;           // {Z_v1, Z_v2} or NULL (raw scores)
  const float* coef1;       // kWeighted only
  const float* coef2;
  float* out1;              // optional, laid out like idx
  float* out2;
  float* part_grad;         // [B, chunks, 2, D]
  float* part_scal;         // [B, chunks, 4]
  int64_t cols;
  int64_t n_rows;           // rows of each bank: ids outside [0, n_rows) are flagged (MML_DEVERR_CRD_INDEX) and read row 0
  uint32_t* err;
  int32_t D;
  int32_t chunk_cols;
  int32_t chunks;
  float inv_T;
  float inv_TB;             // 1 / (T * batch_norm)
  float nce_c;              // fp32(m*Pn + eps)     CRD_criterion.py:208,212
  float nce_kp;             // fp32(m*Pn)           CRD_criterion.py:212
  int64_t npos;             // leading columns of a segment that are positives (1 = CRD_criterion.py; P2 = CRD_loss.py:221-241)
  float pos_w;              // weight of a positive's loss term: 1 / npos (ContrastLoss_v2 averages over the P positives)
  // peer mode (row-sharded bank, indices pulled over NVLink): anchor b = (source rank s, local anchor bl);
  // CTA (chunk c, b) reads peer_ids[s][(bl*chunks + c)*peer_stride ..] of length peer_cnt[s][bl*chunks + c]
  int32_t peer_world;       // 0 = off
  int32_t peer_B_local;
  int32_t peer_stride;
  int32_t peer_route_chunks;   // slots per (source, anchor) in the routed layout
  int32_t peer_group;          // consecutive slots handled by one CTA (keeps >= ~1K rows per CTA at large world sizes)
  const int32_t* peer_cnt[32];   // per source rank: counts of ITS slots destined to this owner (local copy or peer-mapped)
  const int32_t* peer_ids[32];
};

struct Segment {
  const int64_t* idx64;
  const int32_t* idx32;
  int64_t begin;            // offset of the segment in idx
  int64_t c0, c1;           // column range of this CTA inside the segment
  bool has_pos;             // column 0 of the segment is the positive
  bool active;
};

// number of sub-segments CTA (chunk, b) walks: 1 except in peer mode
__device__ __forceinline__ int cta_subsegments(const GatherArgs& a, int chunk) {
  if (a.peer_world == 0) return 1;
  return min(a.peer_group, a.peer_route_chunks - chunk * a.peer_group);
}

template <bool kPeer>
__device__ __forceinline__ Segment cta_segment(const GatherArgs& a, int b, int chunk, int sub) {
  Segment g;
  if (kPeer) {
    const int s = b / a.peer_B_local;
    const int bl = b - s * a.peer_B_local;
    const int rc = chunk * a.peer_group + sub;                       // routed slot
    g.idx64 = nullptr;
    g.idx32 = a.peer_ids[s];
    g.begin = (static_cast<int64_t>(bl) * a.peer_route_chunks + rc) * a.peer_stride;
    g.c0 = 0;
    g.c1 = a.peer_cnt[s][static_cast<int64_t>(bl) * a.peer_route_chunks + rc];
    g.has_pos = (rc == 0) && (a.pos_flag == nullptr || a.pos_flag[b] != 0);
    g.active = true;                         // empty sub-segments still publish zero partials
    return g;
  }
  int64_t len;
  if (a.seg_ptr != nullptr) {
    g.begin = a.seg_ptr[b];
    len = a.seg_ptr[b + 1] - g.begin;
  } else {
    g.begin = static_cast<int64_t>(b) * a.cols;
    len = a.cols;
  }
  g.idx64 = a.idx;
  g.idx32 = a.idx32;
  g.c0 = static_cast<int64_t>(chunk) * a.chunk_cols;
  g.c1 = min(g.c0 + static_cast<int64_t>(a.chunk_cols), len);
  g.has_pos = (a.pos_flag == nullptr) ? true : (a.pos_flag[b] != 0);
  g.active = g.c0 < len;                     // finishers skip the same chunks
  return g;
}

// Per-row scalar stage shared by the fast and the generic kernel.
//   dot      : <row, v>
//   returns g: coefficient of `row` in dL/dv;  accumulates loss / raw-sum terms.
template <int MODE>
__device__ __forceinline__ float score_row(const GatherArgs& a, float dot, float inv_Z, bool valid,
                                           bool is_pos, float& loss_acc, float& sum_acc, float& x_out) {
  const float e = expf(dot * a.inv_T);
  const float x = e * inv_Z;
  x_out = x;
  float g = 0.f;
  if (MODE == kFused) {
    const float denom = x + a.nce_c;
    const float num_l = is_pos ? x : a.nce_kp;
    const float term = logf(num_l / denom);          // log_D1 / log_D0, :208,212
    const float num_g = is_pos ? -a.nce_c : x;       // dL/dx * x  (times 1/(T*bsz) below)
    const float w = is_pos ? a.pos_w : 1.f;          // exactly 1 for the single-positive criterion
    g = valid ? (num_g / denom) * a.inv_TB * w : 0.f;
    loss_acc += valid ? term * w : 0.f;
  } else {
    sum_acc += valid ? e : 0.f;
  }
  return g;
}

template <int VPL, int U, int MODE, int MINB = 1, bool kPeer = false>
__global__ void __launch_bounds__(kCtaThreads, MINB) crd_gather_kernel(const GatherArgs a) {
  constexpr int D = 32 * VPL;
  constexpr int ROWS_PER_IT = 4 * U;        // rows of each bank per warp iteration
  constexpr int NSCAL = 2 * U;              // scalars per row-slot per iteration
  static_assert(NSCAL <= 8, "scalar packing uses the 8 lanes of a row slot");

  const int b = blockIdx.y;
  const int chunk = blockIdx.x;
  const int nsub = kPeer ? cta_subsegments(a, chunk) : 1;
  if (!kPeer && !cta_segment<false>(a, b, chunk, 0).active) return;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int s = lane & 7;                   // position inside the 8-lane row group
  const int q = lane >> 3;                  // row slot 0..3

  float4 fv1[VPL], fv2[VPL];
  if (MODE != kWeighted) {
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      fv1[j] = __ldg(reinterpret_cast<const float4*>(a.v1 + static_cast<int64_t>(b) * D) + j * 8 + s);
      fv2[j] = __ldg(reinterpret_cast<const float4*>(a.v2 + static_cast<int64_t>(b) * D) + j * 8 + s);
    }
  }
  float4 acc1[VPL], acc2[VPL];              // acc1 -> dL/dv1 (bank-2 rows), acc2 -> dL/dv2 (bank-1 rows)
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    acc1[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    acc2[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float loss_acc = 0.f, sum_acc = 0.f;
  float inv_Z1 = 1.f, inv_Z2 = 1.f;
  if (MODE != kWeighted && a.Z != nullptr) {
    inv_Z1 = 1.f / a.Z[0];
    inv_Z2 = 1.f / a.Z[1];
  }
  // lane s < NSCAL evaluates scalar s of its row slot: u = s>>1, side 1 (bank 2 . v1) if s odd.
  const bool my_side1 = (s & 1) != 0;
  const float my_inv_Z = my_side1 ? inv_Z1 : inv_Z2;

  // Peer mode: the routed ids of this CTA's slots live in the SOURCE rank's arena.  Pull them over NVLink ONCE, all
  // loads in flight together (one remote round trip for the counts, one for the ids), compacted into shared memory;
  // the row loop below then runs over one flat local list.  A per-block remote index load would put an NVLink
  // latency in front of every 32 rows.  Slots too full for the staging buffer keep the direct (sub-segment) walk.
  int nwalk = nsub;
  const int32_t* peer_stage_ptr = nullptr;
  bool staged = false;
  int staged_total = 0;
  if constexpr (kPeer) {
    __shared__ int32_t sm_ids[kPeerStageIds];
    __shared__ int32_t sm_off[kPeerMaxGroup + 1];
    if (nsub <= kPeerMaxGroup) {
      const int src_rank = b / a.peer_B_local;
      const int bl = b - src_rank * a.peer_B_local;
      const int64_t slot0 = static_cast<int64_t>(bl) * a.peer_route_chunks + chunk * a.peer_group;
      if (warp == 0) {
        const int32_t cnt = lane < nsub ? a.peer_cnt[src_rank][slot0 + lane] : 0;
        int32_t incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int32_t t = __shfl_up_sync(kFullMask, incl, off);
          if (lane >= off) incl += t;
        }
        if (lane < nsub) sm_off[lane + 1] = incl;
        if (lane == 0) sm_off[0] = 0;
      }
      __syncthreads();
      staged_total = sm_off[nsub];
      staged = staged_total <= kPeerStageIds;
      if (staged) {
        for (int sub = warp; sub < nsub; sub += kCtaWarps) {
          const int32_t o0 = sm_off[sub], cnt = sm_off[sub + 1] - o0;
          const int4* src = reinterpret_cast<const int4*>(a.peer_ids[src_rank] + (slot0 + sub) * a.peer_stride);
          for (int i0 = 0; i0 < cnt; i0 += 256) {          // two 16-byte loads per lane in flight
            const int ia = i0 + lane * 4, ib = ia + 128;
            int4 va = make_int4(0, 0, 0, 0), vb = va;
            if (ia < cnt) va = src[ia >> 2];
            if (ib < cnt) vb = src[ib >> 2];
            const int32_t ea[4] = {va.x, va.y, va.z, va.w}, eb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (ia + e < cnt) sm_ids[o0 + ia + e] = ea[e];
              if (ib + e < cnt) sm_ids[o0 + ib + e] = eb[e];
            }
          }
        }
        nwalk = 1;
      }
      __syncthreads();
      if (staged) peer_stage_ptr = sm_ids;
    }
  }

  for (int sub = 0; sub < nwalk; ++sub) {
  Segment sg;
  if (kPeer && staged) {
    sg.idx64 = nullptr; sg.idx32 = peer_stage_ptr; sg.begin = 0; sg.c0 = 0; sg.c1 = staged_total;
    sg.has_pos = (chunk == 0) && (a.pos_flag == nullptr || a.pos_flag[b] != 0);
    sg.active = true;
  } else {
    sg = cta_segment<kPeer>(a, b, chunk, sub);
  }
  const int64_t seg_begin = sg.begin, c0 = sg.c0, c1 = sg.c1;
  const bool has_pos = sg.has_pos;
  for (int64_t cb = c0 + warp * 32; cb < c1; cb += kCtaWarps * 32) {
    // one coalesced load of 32 indices per warp
    const int64_t mycol = cb + lane;
    const bool myvalid = mycol < c1;
    int32_t myrow = 0;
    if (myvalid) {
      const int64_t r64 = sg.idx32 ? static_cast<int64_t>(sg.idx32[seg_begin + mycol]) : sg.idx64[seg_begin + mycol];
      if (static_cast<uint64_t>(r64) >= static_cast<uint64_t>(a.n_rows)) flag_device_error(a.err, MML_DEVERR_CRD_INDEX);
      else myrow = static_cast<int32_t>(r64);
    }
    float mycf1 = 0.f, mycf2 = 0.f;
    if (MODE == kWeighted && myvalid) {
      mycf1 = a.coef1[seg_begin + mycol];
      mycf2 = a.coef2[seg_begin + mycol];
    }
    const int n_it = (static_cast<int>(min(static_cast<int64_t>(32), c1 - cb)) + ROWS_PER_IT - 1) / ROWS_PER_IT;
    for (int it = 0; it < n_it; ++it) {
      float4 r1[U][VPL], r2[U][VPL];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int src = it * ROWS_PER_IT + u * 4 + q;
        const int32_t row = __shfl_sync(kFullMask, myrow, src);
        const float* p1 = a.bank1 + static_cast<int64_t>(row) * D + s * 4;
        const float* p2 = a.bank2 + static_cast<int64_t>(row) * D + s * 4;
#pragma unroll
        for (int j = 0; j < VPL; ++j) r1[u][j] = ldg_stream_f4(p1 + j * 32);
#pragma unroll
        for (int j = 0; j < VPL; ++j) r2[u][j] = ldg_stream_f4(p2 + j * 32);
      }
      float gg[NSCAL];
      if (MODE == kWeighted) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int src = it * ROWS_PER_IT + u * 4 + q;
          gg[2 * u] = __shfl_sync(kFullMask, mycf2, src);       // bank-1 rows -> g2
          gg[2 * u + 1] = __shfl_sync(kFullMask, mycf1, src);   // bank-2 rows -> g1
        }
      } else {
        float p[NSCAL];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float d_b1 = 0.f, d_b2 = 0.f;
#pragma unroll
          for (int j = 0; j < VPL; ++j) {
            d_b1 = dot4(r1[u][j], fv2[j], d_b1);   // bank 1 . v2   (:43)
            d_b2 = dot4(r2[u][j], fv1[j], d_b2);   // bank 2 . v1   (:48)
          }
          p[2 * u] = d_b1;
          p[2 * u + 1] = d_b2;
        }
#pragma unroll
        for (int t = 0; t < NSCAL; ++t) {
          p[t] += __shfl_xor_sync(kFullMask, p[t], 1);
          p[t] += __shfl_xor_sync(kFullMask, p[t], 2);
          p[t] += __shfl_xor_sync(kFullMask, p[t], 4);
        }
        float mine = p[0];
#pragma unroll
        for (int t = 1; t < NSCAL; ++t) mine = (s == t) ? p[t] : mine;
        const int64_t colm = cb + it * ROWS_PER_IT + (s >> 1) * 4 + q;
        const bool validm = (s < NSCAL) && (colm < c1);
        const bool is_pos = has_pos && (colm < a.npos);
        float x;
        const float g = score_row<MODE>(a, mine, my_inv_Z, validm, is_pos, loss_acc, sum_acc, x);
        if (validm && a.out1 != nullptr) {
          float* o = my_side1 ? a.out1 : a.out2;
          o[seg_begin + colm] = x;
        }
#pragma unroll
        for (int t = 0; t < NSCAL; ++t) gg[t] = __shfl_sync(kFullMask, g, (lane & ~7) | t);
      }
      if (MODE != kScores) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int j = 0; j < VPL; ++j) {
            axpy4(acc2[j], gg[2 * u], r1[u][j]);
            axpy4(acc1[j], gg[2 * u + 1], r2[u][j]);
          }
        }
      }
    }
  }

  }  // sub-segments

  // ---- CTA epilogue: reduce over row slots (lanes), then over warps (smem) ----
  __shared__ float sm_grad[kCtaWarps][2][D];
  __shared__ float sm_scal[kCtaWarps][4];
  if (MODE != kScores) {
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      float* f1 = reinterpret_cast<float*>(&acc1[j]);
      float* f2 = reinterpret_cast<float*>(&acc2[j]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        f1[e] += __shfl_xor_sync(kFullMask, f1[e], 8);
        f1[e] += __shfl_xor_sync(kFullMask, f1[e], 16);
        f2[e] += __shfl_xor_sync(kFullMask, f2[e], 8);
        f2[e] += __shfl_xor_sync(kFullMask, f2[e], 16);
      }
      if (q == 0) {
        *reinterpret_cast<float4*>(&sm_grad[warp][0][(j * 8 + s) * 4]) = acc1[j];
        *reinterpret_cast<float4*>(&sm_grad[warp][1][(j * 8 + s) * 4]) = acc2[j];
      }
    }
  }
  if (MODE != kWeighted) {
    // lanes with odd s hold side-1 terms, even s side-2 terms; xor 2..16 keeps parity classes apart
#pragma unroll
    for (int off = 2; off < 32; off <<= 1) {
      loss_acc += __shfl_xor_sync(kFullMask, loss_acc, off);
      sum_acc += __shfl_xor_sync(kFullMask, sum_acc, off);
    }
    if (lane == 0) { sm_scal[warp][1] = loss_acc; sm_scal[warp][3] = sum_acc; }   // side 2
    if (lane == 1) { sm_scal[warp][0] = loss_acc; sm_scal[warp][2] = sum_acc; }   // side 1
  }
  __syncthreads();
  const int64_t slot = static_cast<int64_t>(b) * a.chunks + chunk;
  if (MODE != kScores) {
    for (int i = threadIdx.x; i < 2 * D; i += kCtaThreads) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kCtaWarps; ++w) t += (&sm_grad[w][0][0])[i];
      a.part_grad[slot * 2 * D + i] = t;
    }
  }
  if (MODE != kWeighted && threadIdx.x < 4) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kCtaWarps; ++w) t += sm_scal[w][threadIdx.x];
    a.part_scal[slot * 4 + threadIdx.x] = t;
  }
}

// Any-D path (D not in {32,64,128,256}): one row per warp at a time, scalar loads.
// Same partial-result contract as the fast kernel.  Accumulators live in shared memory.
template <int MODE>
__global__ void __launch_bounds__(kCtaThreads) crd_gather_generic_kernel(const GatherArgs a) {
  extern __shared__ float sm_dyn[];          // [warps][2][D] accumulators + [2][D] v
  const int D = a.D;
  const int b = blockIdx.y;
  const int chunk = blockIdx.x;
  const bool peer = a.peer_world > 0;
  const int nsub = cta_subsegments(a, chunk);
  if (!peer && !cta_segment<false>(a, b, chunk, 0).active) return;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  float* acc = sm_dyn + static_cast<size_t>(warp) * 2 * D;
  float* sv = sm_dyn + static_cast<size_t>(kCtaWarps) * 2 * D;
  __shared__ float sm_scal[kCtaWarps][4];
  for (int i = threadIdx.x; i < kCtaWarps * 2 * D; i += kCtaThreads) sm_dyn[i] = 0.f;
  if (MODE != kWeighted) {
    for (int i = threadIdx.x; i < D; i += kCtaThreads) {
      sv[i] = a.v1[static_cast<int64_t>(b) * D + i];
      sv[D + i] = a.v2[static_cast<int64_t>(b) * D + i];
    }
  }
  __syncthreads();
  float inv_Z1 = 1.f, inv_Z2 = 1.f;
  if (MODE != kWeighted && a.Z != nullptr) { inv_Z1 = 1.f / a.Z[0]; inv_Z2 = 1.f / a.Z[1]; }
  float loss1 = 0.f, loss2 = 0.f, sum1 = 0.f, sum2 = 0.f;
  for (int sub = 0; sub < nsub; ++sub) {
  const Segment sg = peer ? cta_segment<true>(a, b, chunk, sub) : cta_segment<false>(a, b, chunk, sub);
  const int64_t seg_begin = sg.begin, c0 = sg.c0, c1 = sg.c1;
  const bool has_pos = sg.has_pos;
  for (int64_t col = c0 + warp; col < c1; col += kCtaWarps) {
    int64_t row = sg.idx32 ? static_cast<int64_t>(sg.idx32[seg_begin + col]) : sg.idx64[seg_begin + col];
    if (static_cast<uint64_t>(row) >= static_cast<uint64_t>(a.n_rows)) {
      if (lane == 0) flag_device_error(a.err, MML_DEVERR_CRD_INDEX);
      row = 0;
    }
    const float* p1 = a.bank1 + row * D;
    const float* p2 = a.bank2 + row * D;
    float g1, g2;
    if (MODE == kWeighted) {
      g1 = a.coef1[seg_begin + col];
      g2 = a.coef2[seg_begin + col];
    } else {
      float d_b1 = 0.f, d_b2 = 0.f;
      for (int i = lane; i < D; i += 32) {
        d_b1 = fmaf(__ldg(p1 + i), sv[D + i], d_b1);
        d_b2 = fmaf(__ldg(p2 + i), sv[i], d_b2);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        d_b1 += __shfl_xor_sync(kFullMask, d_b1, off);
        d_b2 += __shfl_xor_sync(kFullMask, d_b2, off);
      }
      const bool is_pos = has_pos && (col < a.npos);
      float x1, x2;
      g2 = score_row<MODE>(a, d_b1, inv_Z2, true, is_pos, loss2, sum2, x2);
      g1 = score_row<MODE>(a, d_b2, inv_Z1, true, is_pos, loss1, sum1, x1);
      if (lane == 0 && a.out1 != nullptr) {
        a.out1[seg_begin + col] = x1;
        a.out2[seg_begin + col] = x2;
      }
    }
    if (MODE != kScores) {
      for (int i = lane; i < D; i += 32) {
        acc[i] = fmaf(g1, __ldg(p2 + i), acc[i]);           // dL/dv1 <- bank-2 rows
        acc[D + i] = fmaf(g2, __ldg(p1 + i), acc[D + i]);   // dL/dv2 <- bank-1 rows
      }
    }
  }
  }  // sub-segments
  if (lane == 0) {
    sm_scal[warp][0] = loss1; sm_scal[warp][1] = loss2;
    sm_scal[warp][2] = sum1;  sm_scal[warp][3] = sum2;
  }
  __syncthreads();
  const int64_t slot = static_cast<int64_t>(b) * a.chunks + chunk;
  if (MODE != kScores) {
    for (int i = threadIdx.x; i < 2 * D; i += kCtaThreads) {
      float t = 0.f;
      for (int w = 0; w < kCtaWarps; ++w) t += sm_dyn[static_cast<size_t>(w) * 2 * D + i];
      a.part_grad[slot * 2 * D + i] = t;
    }
  }
  if (MODE != kWeighted && threadIdx.x < 4) {
    float t = 0.f;
    for (int w = 0; w < kCtaWarps; ++w) t += sm_scal[w][threadIdx.x];
    a.part_scal[slot * 4 + threadIdx.x] = t;
  }
}

// Finisher 1: per anchor, sum the chunk partials in chunk order.
__global__ void __launch_bounds__(128) crd_finish_anchor_kernel(
    const float* __restrict__ part_grad, const float* __restrict__ part_scal,
    const int64_t* __restrict__ seg_ptr, int64_t cols, int32_t chunk_cols, int32_t chunks, int32_t D,
    float* __restrict__ g1, float* __restrict__ g2, double* __restrict__ anchor_scal, int want_grad,
    int want_scal) {
  const int b = blockIdx.x;
  const int64_t len = seg_ptr ? (seg_ptr[b + 1] - seg_ptr[b]) : cols;
  const int used = static_cast<int>((len + chunk_cols - 1) / chunk_cols);
  if (want_grad) {
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) {
      float t = 0.f;
      for (int c = 0; c < used; ++c) t += part_grad[(static_cast<int64_t>(b) * chunks + c) * 2 * D + i];
      if (i < D) g1[static_cast<int64_t>(b) * D + i] = t;
      else g2[static_cast<int64_t>(b) * D + (i - D)] = t;
    }
  }
  if (want_scal && threadIdx.x < 4) {
    double t = 0.0;
    for (int c = 0; c < used; ++c) t += static_cast<double>(part_scal[(static_cast<int64_t>(b) * chunks + c) * 4 + threadIdx.x]);
    anchor_scal[static_cast<int64_t>(b) * 4 + threadIdx.x] = t;
  }
}

// Finisher 2: one CTA, fixed-shape tree over anchors in double.
__global__ void __launch_bounds__(256) crd_finish_total_kernel(
    const double* __restrict__ anchor_scal, int64_t B, double inv_batch, double mean_scale,
    float* __restrict__ loss, float* __restrict__ sums, float* __restrict__ set_Z) {
  __shared__ double sm[4][256];
  double t[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t b = threadIdx.x; b < B; b += 256)
    for (int k = 0; k < 4; ++k) t[k] += anchor_scal[b * 4 + k];
  for (int k = 0; k < 4; ++k) sm[k][threadIdx.x] = t[k];
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (threadIdx.x < off)
      for (int k = 0; k < 4; ++k) sm[k][threadIdx.x] += sm[k][threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (loss) loss[0] = static_cast<float>(-(sm[0][0] + sm[1][0]) * inv_batch);     // :214, both criteria summed (:187)
    if (sums) for (int k = 0; k < 4; ++k) sums[k] = static_cast<float>(sm[k][0]);
    if (set_Z) {                                                                      // :52-59
      if (set_Z[0] < 0.f) set_Z[0] = static_cast<float>(sm[2][0] * mean_scale);
      if (set_Z[1] < 0.f) set_Z[1] = static_cast<float>(sm[3][0] * mean_scale);
    }
  }
}

struct Plan {
  int32_t chunk_cols;
  int32_t chunks;
};

// Column chunking: enough CTAs for >= ~16 waves of 148 SMs x 3 resident CTAs when
// the problem allows, chunks of >= 128 columns (one 32-column block per warp), <= 256 chunks.
Plan make_plan(int64_t B, int64_t cols) {
  const int64_t target_ctas = 148LL * 3 * 16;
  int64_t chunks = (target_ctas + B - 1) / (B > 0 ? B : 1);
  const int64_t max_chunks_by_cols = (cols + 127) / 128;
  if (chunks > max_chunks_by_cols) chunks = max_chunks_by_cols;
  if (chunks > 256) chunks = 256;
  if (chunks < 1) chunks = 1;
  int64_t cc = (cols + chunks - 1) / chunks;
  cc = (cc + 31) / 32 * 32;
  if (cc < 32) cc = 32;
  chunks = (cols + cc - 1) / cc;
  if (chunks < 1) chunks = 1;
  return Plan{static_cast<int32_t>(cc), static_cast<int32_t>(chunks)};
}

struct Workspace {
  float* part_grad;
  float* part_scal;
  double* anchor_scal;
};

size_t ws_bytes_plan(int64_t B, const Plan& p, int32_t D);

size_t ws_bytes(int64_t B, int64_t cols, int32_t D) { return ws_bytes_plan(B, make_plan(B, cols), D); }

size_t ws_bytes_plan(int64_t B, const Plan& p, int32_t D) {
  size_t n = 0;
  n += static_cast<size_t>(B) * p.chunks * 2 * D * sizeof(float);
  n += static_cast<size_t>(B) * p.chunks * 4 * sizeof(float);
  n = (n + 255) / 256 * 256;
  n += static_cast<size_t>(B) * 4 * sizeof(double);
  return n + 256;
}

Workspace carve(void* ws, int64_t B, int32_t D, const Plan& p) {
  Workspace w;
  char* base = static_cast<char*>(ws);
  w.part_grad = reinterpret_cast<float*>(base);
  size_t off = static_cast<size_t>(B) * p.chunks * 2 * D * sizeof(float);
  w.part_scal = reinterpret_cast<float*>(base + off);
  off += static_cast<size_t>(B) * p.chunks * 4 * sizeof(float);
  off = (off + 255) / 256 * 256;
  w.anchor_scal = reinterpret_cast<double*>(base + off);
  return w;
}

template <int MODE>
int launch_gather(const GatherArgs& a, int64_t B, cudaStream_t st) {
  const dim3 grid(a.chunks, static_cast<unsigned>(B));
  const bool peer = a.peer_world > 0;
  switch (a.D) {
    case 32:
      if (peer) crd_gather_kernel<1, 4, MODE, 1, true><<<grid, kCtaThreads, 0, st>>>(a);
      else crd_gather_kernel<1, 4, MODE><<<grid, kCtaThreads, 0, st>>>(a);
      break;
    case 64:
      if (peer) crd_gather_kernel<2, 4, MODE, 1, true><<<grid, kCtaThreads, 0, st>>>(a);
      else crd_gather_kernel<2, 4, MODE><<<grid, kCtaThreads, 0, st>>>(a);
      break;
    case 128:
      if (peer) crd_gather_kernel<4, 2, MODE, 3, true><<<grid, kCtaThreads, 0, st>>>(a);
      else crd_gather_kernel<4, 2, MODE, 3><<<grid, kCtaThreads, 0, st>>>(a);      // 168 regs -> 3 CTAs/SM (2 CTAs/SM is 14 % slower)
      break;
    case 256:
      if (peer) crd_gather_kernel<8, 1, MODE, 1, true><<<grid, kCtaThreads, 0, st>>>(a);
      else crd_gather_kernel<8, 1, MODE><<<grid, kCtaThreads, 0, st>>>(a);
      break;
    default: {
      const size_t smem = (static_cast<size_t>(kCtaWarps) * 2 + 2) * a.D * sizeof(float);
      if (smem > 48 * 1024)      // D > 1228: above the default dynamic shared-memory limit (D <= 2048 needs 80 KB)
        cudaFuncSetAttribute(crd_gather_generic_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      crd_gather_generic_kernel<MODE><<<grid, kCtaThreads, smem, st>>>(a);
    }
  }
  return check_launch("crd_gather_kernel");
}

void set_idx(GatherArgs& a, const void* idx, int32_t idx_bytes) {
  a.idx = idx_bytes == 8 ? static_cast<const int64_t*>(idx) : nullptr;
  a.idx32 = idx_bytes == 4 ? static_cast<const int32_t*>(idx) : nullptr;
}

int common_checks(const float* bank1, const float* bank2, int64_t n_rows, int32_t D, const void* idx, int32_t idx_bytes,
                  int64_t B, int64_t cols, void* ws, size_t ws_size) {
  MML_REQUIRE(bank1 && bank2 && idx && ws, MML_ERR_INVALID_ARG, "crd: null pointer argument");
  MML_REQUIRE(idx_bytes == 8 || idx_bytes == 4, MML_ERR_INVALID_ARG, "crd: idx_bytes must be 8 (int64) or 4 (int32)");
  MML_REQUIRE(B >= 0 && cols >= 1 && n_rows >= 1, MML_ERR_INVALID_ARG, "crd: bad sizes B=%lld cols=%lld n_rows=%lld",
              (long long)B, (long long)cols, (long long)n_rows);
  MML_REQUIRE(D >= 1 && D <= 2048, MML_ERR_UNSUPPORTED, "crd: feature dim %d outside [1, 2048]", D);
  MML_REQUIRE(B <= 65535, MML_ERR_UNSUPPORTED, "crd: batch %lld > 65535 anchors per call", (long long)B);
  MML_REQUIRE(n_rows < (1LL << 31), MML_ERR_UNSUPPORTED, "crd: n_rows must fit in int32");
  if (D % 32 == 0 && D <= 256)
    MML_REQUIRE(aligned16(bank1) && aligned16(bank2), MML_ERR_INVALID_ARG, "crd: banks must be 16-byte aligned");
  MML_REQUIRE(ws_size >= ws_bytes(B, cols, D), MML_ERR_WORKSPACE, "crd: workspace %zu < required %zu", ws_size,
              ws_bytes(B, cols, D));
  return MML_OK;
}

}  // namespace
}  // namespace mml

using namespace mml;

extern "C" size_t mml_crd_workspace_bytes(int64_t B, int64_t cols, int32_t D) {
  if (B < 0 || cols < 1 || D < 1) return 0;
  return ws_bytes(B, cols, D);
}

static int fused_loss_grad_impl(
    const float* bank1, const float* bank2, int64_t n_rows, int32_t D, const float* v1, const float* v2,
    const void* idx, int32_t idx_bytes, const int64_t* seg_ptr, const uint8_t* pos_flag, int64_t B, int64_t cols, float T,
    const float* Z, int64_t n_data, int64_t nce_k, int64_t n_pos, int64_t batch_norm, float* loss, float* sums, float* grad_v1,
    float* grad_v2, float* out_v1, float* out_v2, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = common_checks(bank1, bank2, n_rows, D, idx, idx_bytes, B, cols, workspace, workspace_bytes);
  if (rc != MML_OK) return rc;
  MML_REQUIRE(v1 && v2 && Z && grad_v1 && grad_v2, MML_ERR_INVALID_ARG, "crd_fused: null pointer argument");
  MML_REQUIRE((out_v1 == nullptr) == (out_v2 == nullptr), MML_ERR_INVALID_ARG, "crd_fused: out_v1/out_v2 both or neither");
  MML_REQUIRE(T > 0.f && n_data > 0 && batch_norm > 0 && nce_k >= 0, MML_ERR_INVALID_ARG, "crd_fused: bad scalars");
  if (D % 32 == 0 && D <= 256)
    MML_REQUIRE(aligned16(v1) && aligned16(v2), MML_ERR_INVALID_ARG, "crd_fused: v1/v2 must be 16-byte aligned");
  if (B == 0) return MML_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Plan p = make_plan(B, cols);
  const Workspace w = carve(workspace, B, D, p);
  GatherArgs a{};
  a.bank1 = bank1; a.bank2 = bank2; a.v1 = v1; a.v2 = v2; set_idx(a, idx, idx_bytes); a.seg_ptr = seg_ptr; a.pos_flag = pos_flag;
  a.Z = Z; a.out1 = out_v1; a.out2 = out_v2; a.part_grad = w.part_grad; a.part_scal = w.part_scal;
  a.cols = cols; a.D = D; a.chunk_cols = p.chunk_cols; a.chunks = p.chunks;
  a.n_rows = n_rows; a.err = device_error_word();
  a.inv_T = 1.0f / T;
  a.inv_TB = 1.0f / (T * static_cast<float>(batch_norm));
  const double Pn = 1.0 / static_cast<double>(n_data);                       // :204
  a.nce_kp = static_cast<float>(static_cast<double>(nce_k) * Pn);            // fill_(m*Pn), :212
  a.nce_c = static_cast<float>(static_cast<double>(nce_k) * Pn + 1e-7);      // add(m*Pn+eps), :208,212
  a.npos = n_pos;
  a.pos_w = n_pos == 1 ? 1.0f : 1.0f / static_cast<float>(n_pos);
  rc = launch_gather<kFused>(a, B, st);
  if (rc != MML_OK) return rc;
  crd_finish_anchor_kernel<<<static_cast<unsigned>(B), 128, 0, st>>>(w.part_grad, w.part_scal, seg_ptr, cols, p.chunk_cols,
                                                                      p.chunks, D, grad_v1, grad_v2, w.anchor_scal, 1, 1);
  rc = check_launch("crd_finish_anchor_kernel");
  if (rc != MML_OK) return rc;
  if (loss || sums) {
    crd_finish_total_kernel<<<1, 256, 0, st>>>(w.anchor_scal, B, 1.0 / static_cast<double>(batch_norm), 0.0, loss, sums,
                                               nullptr);
    rc = check_launch("crd_finish_total_kernel");
  }
  return rc;
}

extern "C" int mml_crd_fused_loss_grad(
    const float* bank1, const float* bank2, int64_t n_rows, int32_t D, const float* v1, const float* v2,
    const void* idx, int32_t idx_bytes, const int64_t* seg_ptr, const uint8_t* pos_flag, int64_t B, int64_t cols, float T,
    const float* Z, int64_t n_data, int64_t nce_k, int64_t batch_norm, float* loss, float* sums, float* grad_v1,
    float* grad_v2, float* out_v1, float* out_v2, void* workspace, size_t workspace_bytes, void* stream) {
  return fused_loss_grad_impl(bank1, bank2, n_rows, D, v1, v2, idx, idx_bytes, seg_ptr, pos_flag, B, cols, T, Z, n_data, nce_k,
                              1, batch_norm, loss, sums, grad_v1, grad_v2, out_v1, out_v2, workspace, workspace_bytes, stream);
}

extern "C" int mml_crd_fused_loss_grad_multipos(
    const float* bank1, const float* bank2, int64_t n_rows, int32_t D, const float* v1, const float* v2,
    const void* idx, int32_t idx_bytes, int64_t B, int64_t cols, int64_t n_pos, float T, const float* Z, int64_t n_data,
    float* loss, float* grad_v1, float* grad_v2, float* out_v1, float* out_v2, void* workspace, size_t workspace_bytes,
    void* stream) {
  MML_REQUIRE(n_pos >= 1 && n_pos <= cols, MML_ERR_INVALID_ARG, "crd_fused_multipos: n_pos %lld outside [1, cols=%lld]",
              (long long)n_pos, (long long)cols);
  // ContrastLoss_v2 (CRD_loss.py:221-241): m = number of negative columns = cols - P
  return fused_loss_grad_impl(bank1, bank2, n_rows, D, v1, v2, idx, idx_bytes, nullptr, nullptr, B, cols, T, Z, n_data,
                              cols - n_pos, n_pos, B, loss, nullptr, grad_v1, grad_v2, out_v1, out_v2, workspace,
                              workspace_bytes, stream);
}

extern "C" int mml_crd_scores(const float* bank1, const float* bank2, int64_t n_rows, int32_t D, const float* v1,
                              const float* v2, const void* idx, int32_t idx_bytes, const int64_t* seg_ptr, int64_t B,
                              int64_t cols, float T, const float* Z, float* sums, float* set_Z, float* out_v1,
                              float* out_v2, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = common_checks(bank1, bank2, n_rows, D, idx, idx_bytes, B, cols, workspace, workspace_bytes);
  if (rc != MML_OK) return rc;
  MML_REQUIRE(v1 && v2, MML_ERR_INVALID_ARG, "crd_scores: null pointer argument");
  MML_REQUIRE((out_v1 == nullptr) == (out_v2 == nullptr), MML_ERR_INVALID_ARG, "crd_scores: out_v1/out_v2 both or neither");
  MML_REQUIRE(T > 0.f, MML_ERR_INVALID_ARG, "crd_scores: T must be positive");
  if (D % 32 == 0 && D <= 256)
    MML_REQUIRE(aligned16(v1) && aligned16(v2), MML_ERR_INVALID_ARG, "crd_scores: v1/v2 must be 16-byte aligned");
  if (B == 0) return MML_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Plan p = make_plan(B, cols);
  const Workspace w = carve(workspace, B, D, p);
  GatherArgs a{};
  a.bank1 = bank1; a.bank2 = bank2; a.v1 = v1; a.v2 = v2; set_idx(a, idx, idx_bytes); a.seg_ptr = seg_ptr;
  a.Z = Z; a.out1 = out_v1; a.out2 = out_v2; a.part_grad = w.part_grad; a.part_scal = w.part_scal;
  a.cols = cols; a.D = D; a.chunk_cols = p.chunk_cols; a.chunks = p.chunks;
  a.n_rows = n_rows; a.err = device_error_word();
  a.inv_T = 1.0f / T;
  rc = launch_gather<kScores>(a, B, st);
  if (rc != MML_OK) return rc;
  if (sums || set_Z) {
    crd_finish_anchor_kernel<<<static_cast<unsigned>(B), 128, 0, st>>>(w.part_grad, w.part_scal, seg_ptr, cols,
                                                                        p.chunk_cols, p.chunks, D, nullptr, nullptr,
                                                                        w.anchor_scal, 0, 1);
    rc = check_launch("crd_finish_anchor_kernel");
    if (rc != MML_OK) return rc;
    // mean over B*cols entries times outputSize (:53,57); dense layout only for set_Z
    const double mean_scale = static_cast<double>(n_rows) / (static_cast<double>(B) * static_cast<double>(cols));
    crd_finish_total_kernel<<<1, 256, 0, st>>>(w.anchor_scal, B, 0.0, mean_scale, nullptr, sums, set_Z);
    rc = check_launch("crd_finish_total_kernel");
  }
  return rc;
}

extern "C" int mml_crd_weighted_rows(const float* bank1, const float* bank2, int64_t n_rows, int32_t D,
                                     const void* idx, int32_t idx_bytes, const int64_t* seg_ptr, const float* coef1,
                                     const float* coef2, int64_t B, int64_t cols, float* g1, float* g2,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  int rc = common_checks(bank1, bank2, n_rows, D, idx, idx_bytes, B, cols, workspace, workspace_bytes);
  if (rc != MML_OK) return rc;
  MML_REQUIRE(coef1 && coef2 && g1 && g2, MML_ERR_INVALID_ARG, "crd_weighted_rows: null pointer argument");
  if (B == 0) return MML_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Plan p = make_plan(B, cols);
  const Workspace w = carve(workspace, B, D, p);
  GatherArgs a{};
  a.bank1 = bank1; a.bank2 = bank2; set_idx(a, idx, idx_bytes); a.seg_ptr = seg_ptr; a.coef1 = coef1; a.coef2 = coef2;
  a.part_grad = w.part_grad; a.part_scal = w.part_scal;
  a.cols = cols; a.D = D; a.chunk_cols = p.chunk_cols; a.chunks = p.chunks;
  a.n_rows = n_rows; a.err = device_error_word();
  rc = launch_gather<kWeighted>(a, B, st);
  if (rc != MML_OK) return rc;
  crd_finish_anchor_kernel<<<static_cast<unsigned>(B), 128, 0, st>>>(w.part_grad, w.part_scal, seg_ptr, cols, p.chunk_cols,
                                                                      p.chunks, D, g1, g2, w.anchor_scal, 1, 0);
  return check_launch("crd_finish_anchor_kernel");
}


// ------------------------------------------------------------------ peer (NVLink pull) variants
namespace mml {
namespace {

int peer_setup(GatherArgs& a, Plan& p, const float* bank1, const float* bank2, int64_t n_rows, int32_t D,
               const int32_t* const* peer_ids_host, const int32_t* const* peer_counts_host, int32_t world, int64_t B_local,
               int32_t route_chunks, int32_t route_stride, void* ws, size_t ws_size) {
  MML_REQUIRE(bank1 && bank2 && peer_ids_host && peer_counts_host && ws, MML_ERR_INVALID_ARG, "crd_peer: null pointer argument");
  MML_REQUIRE(world >= 1 && world <= 32 && B_local >= 1 && route_chunks >= 1 && route_stride >= 32, MML_ERR_INVALID_ARG,
              "crd_peer: bad world / B_local / route layout");
  MML_REQUIRE(D >= 1 && D <= 2048, MML_ERR_UNSUPPORTED, "crd_peer: feature dim %d outside [1, 2048]", D);
  const int64_t B = B_local * world;
  MML_REQUIRE(B <= 65535, MML_ERR_UNSUPPORTED, "crd_peer: global batch %lld > 65535 anchors per call", (long long)B);
  MML_REQUIRE(n_rows >= 1 && n_rows < (1LL << 31), MML_ERR_UNSUPPORTED, "crd_peer: n_rows must fit in int32");
  // a CTA walks `group` consecutive slots so that it still sees ~>= 1K rows when each slot holds stride/world of them
  int group = (1024 * world + route_stride - 1) / route_stride;
  if (group < 1) group = 1;
  if (group > route_chunks) group = route_chunks;
  const int32_t ctas_per_anchor = (route_chunks + group - 1) / group;
  p = Plan{route_stride, ctas_per_anchor};
  MML_REQUIRE(ws_size >= ws_bytes_plan(B, Plan{route_stride, route_chunks}, D), MML_ERR_WORKSPACE,
              "crd_peer: workspace %zu < required %zu", ws_size, ws_bytes_plan(B, Plan{route_stride, route_chunks}, D));
  a.bank1 = bank1; a.bank2 = bank2;
  a.cols = static_cast<int64_t>(ctas_per_anchor) * route_stride;   // finishers: every CTA publishes a partial
  a.D = D; a.chunk_cols = route_stride; a.chunks = ctas_per_anchor;
  a.n_rows = n_rows; a.err = device_error_word();
  a.peer_world = world; a.peer_B_local = static_cast<int32_t>(B_local); a.peer_stride = route_stride;
  a.peer_route_chunks = route_chunks; a.peer_group = group;
  for (int i = 0; i < world; ++i) {
    MML_REQUIRE(peer_ids_host[i] != nullptr && peer_counts_host[i] != nullptr, MML_ERR_INVALID_ARG, "crd_peer: null peer buffer %d", i);
    a.peer_ids[i] = peer_ids_host[i];
    a.peer_cnt[i] = peer_counts_host[i];
  }
  return MML_OK;
}

}  // namespace
}  // namespace mml

extern "C" size_t mml_crd_peer_workspace_bytes(int64_t B_global, int32_t route_chunks, int32_t D) {
  if (B_global < 0 || route_chunks < 1 || D < 1) return 0;
  return ws_bytes_plan(B_global, Plan{32, route_chunks}, D);
}

extern "C" int mml_crd_fused_loss_grad_peer(
    const float* bank1, const float* bank2, int64_t n_rows, int32_t D, const float* v1, const float* v2,
    const int32_t* const* peer_ids_host, const int32_t* const* peer_counts_host, int32_t world, int64_t B_local, int32_t route_chunks,
    int32_t route_stride, const uint8_t* pos_flag, float T, const float* Z, int64_t n_data, int64_t nce_k,
    int64_t batch_norm, float* sums, float* grad_v1, float* grad_v2, void* workspace, size_t workspace_bytes,
    void* stream) {
  GatherArgs a{};
  Plan p{};
  int rc = peer_setup(a, p, bank1, bank2, n_rows, D, peer_ids_host, peer_counts_host, world, B_local, route_chunks, route_stride,
                      workspace, workspace_bytes);
  if (rc != MML_OK) return rc;
  MML_REQUIRE(v1 && v2 && Z && grad_v1 && grad_v2, MML_ERR_INVALID_ARG, "crd_fused_peer: null pointer argument");
  MML_REQUIRE(T > 0.f && n_data > 0 && batch_norm > 0 && nce_k >= 0, MML_ERR_INVALID_ARG, "crd_fused_peer: bad scalars");
  const int64_t B = B_local * world;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Workspace w = carve(workspace, B, D, p);
  a.v1 = v1; a.v2 = v2; a.pos_flag = pos_flag; a.Z = Z; a.part_grad = w.part_grad; a.part_scal = w.part_scal;
  a.inv_T = 1.0f / T;
  a.inv_TB = 1.0f / (T * static_cast<float>(batch_norm));
  const double Pn = 1.0 / static_cast<double>(n_data);
  a.nce_kp = static_cast<float>(static_cast<double>(nce_k) * Pn);
  a.nce_c = static_cast<float>(static_cast<double>(nce_k) * Pn + 1e-7);
  a.npos = 1; a.pos_w = 1.0f;
  rc = launch_gather<kFused>(a, B, st);
  if (rc != MML_OK) return rc;
  crd_finish_anchor_kernel<<<static_cast<unsigned>(B), 128, 0, st>>>(w.part_grad, w.part_scal, nullptr, a.cols, p.chunk_cols,
                                                                      p.chunks, D, grad_v1, grad_v2, w.anchor_scal, 1, 1);
  rc = check_launch("crd_finish_anchor_kernel");
  if (rc != MML_OK) return rc;
  if (sums) {
    crd_finish_total_kernel<<<1, 256, 0, st>>>(w.anchor_scal, B, 0.0, 0.0, nullptr, sums, nullptr);
    rc = check_launch("crd_finish_total_kernel");
  }
  return rc;
}

extern "C" int mml_crd_scores_peer(const float* bank1, const float* bank2, int64_t n_rows, int32_t D, const float* v1,
                                   const float* v2, const int32_t* const* peer_ids_host,
                                   const int32_t* const* peer_counts_host, int32_t world, int64_t B_local,
                                   int32_t route_chunks, int32_t route_stride, float T,
                                   float* sums, void* workspace, size_t workspace_bytes, void* stream) {
  GatherArgs a{};
  Plan p{};
  int rc = peer_setup(a, p, bank1, bank2, n_rows, D, peer_ids_host, peer_counts_host, world, B_local, route_chunks, route_stride,
                      workspace, workspace_bytes);
  if (rc != MML_OK) return rc;
  MML_REQUIRE(v1 && v2 && sums && T > 0.f, MML_ERR_INVALID_ARG, "crd_scores_peer: bad arguments");
  const int64_t B = B_local * world;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Workspace w = carve(workspace, B, D, p);
  a.v1 = v1; a.v2 = v2; a.part_grad = w.part_grad; a.part_scal = w.part_scal;
  a.inv_T = 1.0f / T;
  rc = launch_gather<kScores>(a, B, st);
  if (rc != MML_OK) return rc;
  crd_finish_anchor_kernel<<<static_cast<unsigned>(B), 128, 0, st>>>(w.part_grad, w.part_scal, nullptr, a.cols, p.chunk_cols,
                                                                      p.chunks, D, nullptr, nullptr, w.anchor_scal, 0, 1);
  rc = check_launch("crd_finish_anchor_kernel");
  if (rc != MML_OK) return rc;
  crd_finish_total_kernel<<<1, 256, 0, st>>>(w.anchor_scal, B, 0.0, 0.0, nullptr, sums, nullptr);
  return check_launch("crd_finish_total_kernel");
}
