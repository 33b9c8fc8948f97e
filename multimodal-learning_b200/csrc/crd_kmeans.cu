// K13 -- per-class k-means centres of a memory bank (the extra positives of pos_extra == "centers" with more than two positives).
//
// Replaces, for one bank, the host round trip of the reference (CL_utils/CRD_criterion_v10.py:84-92 and :122-129):
//     for c in classes:  feature = index_select(memory, 0, class_idx[c]).cpu().numpy()
//                        KMeans(n_clusters = num_pos - 1).fit(feature).cluster_centers_
// by Lloyd iterations on the device.  One iteration is two launches:
//
//   kmeans_assign_kernel  one pass over the listed bank rows (every row is read exactly once: 4 D + 8 bytes per row).  A CTA belongs to one
//                         class and holds that class's k centres in shared memory.  Eight lanes share a row (D/8 values each,
//                         128-byte coalesced segments) and every group keeps kR rows in flight, so a centre chunk read from
//                         shared memory is used 4 * kR times; the next iteration's rows (and the row numbers of the one after)
//                         are already in flight while this one is scored.  Scores are sklearn's E-step form |c|^2 - 2 x.c (lowest index wins
//                         ties).  Every lane owns a private shared-memory accumulator per centre for its D/8 values: no atomics,
//                         a fixed row -> lane mapping, fixed-order reductions afterwards -> bit-reproducible sums.
//   kmeans_update_kernel  one CTA per centre adds the per-CTA partial sums in a fixed order, writes the new centre (an empty
//                         cluster keeps its centre), and the last CTA of a class compares the summed squared shift with the
//                         class's tolerance and raises that class's `done` flag -- later iterations of a finished class return
//                         at once, so the host may enqueue iterations in batches and read the flags between batches.
//
// The same assign pass with update = 0 yields each listed row's squared distance to its nearest centre (the D^2 weights of a
// k-means++ initialisation) and the per-centre inertia (the variance that scales sklearn's tolerance, _kmeans.py `_tolerance`).
#include "common.cuh"

namespace mml {
namespace {

constexpr int kKmMaxK = 8;          // centres per class (num_pos - 1)
constexpr int kKmMaxClasses = 32;
constexpr int kKmMaxGrid = 148;     // one CTA per SM: the private accumulators take most of the shared memory

struct KmArgs {
  const float* bank;
  const int64_t* rows;               // concatenated row lists, class after class
  float* centres;                    // [C, k, D]
  const float* tol;                  // [C] or null
  int32_t* done;                     // [C] or null
  float* inertia;                    // [C, k] or null
  int64_t* counts;                   // [C, k] or null
  float* sums;                       // [C, k, D] or null: the rows of every centre, summed (for a reduction across shards)
  float* row_dist;                   // [m] or null
  float* part_sum;                   // [grid, k, D]
  float* part_inertia;               // [grid, k]
  int32_t* part_count;               // [grid, k]
  float* shift;                      // [C, k]
  int32_t* arrive;                   // [C]
  uint32_t* err;
  int64_t n;
  int64_t row_begin[kKmMaxClasses + 1];
  int32_t cta_begin[kKmMaxClasses + 1];
  int32_t D, k, C, update;
};

// Rows in flight per 8-lane group and CTA size.  kV = D / 32 float4 per lane and row; kK = centres per class rounded up to
// 2 / 4 / 8 (compile-time trip counts; the padding centres score +inf and are never chosen).  The lane-private accumulators
// are kK * kV float4 per thread, which bounds the CTA size.
template <int kV, int kK>
struct KmRows {      // kV == 4 with 8 centres runs 256-thread CTAs: registers for 4 rows, and twice the reuse of a centre chunk
  static constexpr int value = kV <= 2 ? 4 : (kV == 4 ? (kK == 8 ? 4 : 2) : 1);
};
__host__ __device__ constexpr int km_threads(int kV, int kK) {
  int t = 512;
  while (t > 64 && t * kK * kV * 16 > 180 * 1024) t /= 2;
  return t;
}

template <int kV, int kK>
__global__ void __launch_bounds__(km_threads(kV, kK)) kmeans_assign_kernel(const KmArgs a) {
  extern __shared__ float4 km_smem[];
  constexpr int D = 32 * kV;
  constexpr int kR = KmRows<kV, kK>::value;
  constexpr int kRowsIter = 4 * kR;
  constexpr int threads = km_threads(kV, kK);
  constexpr int nwarps = threads / 32;
  const int tid = static_cast<int>(threadIdx.x);
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 3, l8 = lane & 7;
  const int cta = static_cast<int>(blockIdx.x);
  const int k = a.k;

  int c = 0;
  while (c + 1 < a.C && cta >= a.cta_begin[c + 1]) ++c;
  if (a.done != nullptr && a.done[c] != 0) return;

  float4* sm_centre = km_smem;                                    // [kK][kV * 8]
  float* sm_cnorm = reinterpret_cast<float*>(sm_centre + kK * kV * 8);   // [8]
  float4* sm_acc = sm_centre + kK * kV * 8 + 2;                   // [kK * kV][threads], lane-private
  const float4* centre_g = reinterpret_cast<const float4*>(a.centres + static_cast<int64_t>(c) * k * D);
  for (int i = tid; i < kK * kV * 8; i += threads) sm_centre[i] = i < k * kV * 8 ? centre_g[i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < kK * kV; ++i) sm_acc[i * threads + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  for (int j = warp; j < kK; j += nwarps) {                       // |c_j|^2, a warp per centre
    float s = 0.f;
    for (int i = lane; i < kV * 8; i += 32) {
      const float4 v = sm_centre[j * kV * 8 + i];
      s = dot4(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFullMask, s, o);
    if (lane == 0) sm_cnorm[j] = j < k ? s : __int_as_float(0x7f800000);
  }
  __syncthreads();
  float cn[kK];
#pragma unroll
  for (int j = 0; j < kK; ++j) cn[j] = sm_cnorm[j];

  // this CTA's contiguous slice of the class's row list, in whole warp-iterations of 4 * kR rows
  const int64_t m0 = a.row_begin[c], m1 = a.row_begin[c + 1];
  const int nctas = a.cta_begin[c + 1] - a.cta_begin[c];
  const int64_t iters = (m1 - m0 + kRowsIter - 1) / kRowsIter;
  const int64_t per = (iters + nctas - 1) / nctas;
  const int64_t it0 = per * (cta - a.cta_begin[c]);
  const int64_t it1 = it0 + per < iters ? it0 + per : iters;

  int32_t cnt[kK];
  float inr[kK];
#pragma unroll
  for (int j = 0; j < kK; ++j) cnt[j] = 0, inr[j] = 0.f;

  // two-deep software pipeline: row numbers two iterations ahead, rows one iteration ahead
  auto row_ids = [&](int64_t it, int64_t (&rid)[kR]) {
    const int64_t p0 = m0 + it * kRowsIter + g * kR;
#pragma unroll
    for (int t = 0; t < kR; ++t) rid[t] = a.rows[p0 + t < m1 ? p0 + t : m0];
  };
  auto fetch = [&](const int64_t (&rid)[kR], float4 (&dst)[kR][kV]) {
#pragma unroll
    for (int t = 0; t < kR; ++t) {
      int64_t r = rid[t];
      if (r < 0 || r >= a.n) {                                    // a bad row number: flag it, read row 0
        flag_device_error(a.err, MML_DEVERR_CRD_INDEX);
        r = 0;
      }
      const float* src = a.bank + r * D + l8 * 4;
#pragma unroll
      for (int i = 0; i < kV; ++i) dst[t][i] = ldg_stream_f4(src + i * 32);
    }
  };
  int64_t rid[kR];
  float4 xn[kR][kV];
  int64_t it = it0 + warp;
  if (it < it1) {
    row_ids(it, rid);
    fetch(rid, xn);
    if (it + nwarps < it1) row_ids(it + nwarps, rid);
  }
  for (; it < it1; it += nwarps) {
    float4 x[kR][kV];
#pragma unroll
    for (int t = 0; t < kR; ++t)
#pragma unroll
      for (int i = 0; i < kV; ++i) x[t][i] = xn[t][i];
    if (it + nwarps < it1) {
      fetch(rid, xn);
      if (it + 2 * nwarps < it1) row_ids(it + 2 * nwarps, rid);
    }
    const int64_t p0 = m0 + it * kRowsIter + g * kR;
    float acc[kR][kK], xx[kR];
#pragma unroll
    for (int t = 0; t < kR; ++t) {
      xx[t] = 0.f;
#pragma unroll
      for (int i = 0; i < kV; ++i) xx[t] = dot4(x[t][i], x[t][i], xx[t]);
#pragma unroll
      for (int j = 0; j < kK; ++j) acc[t][j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < kK; ++j) {
#pragma unroll
      for (int i = 0; i < kV; ++i) {
        const float4 c4 = sm_centre[(j * kV + i) * 8 + l8];
#pragma unroll
        for (int t = 0; t < kR; ++t) acc[t][j] = dot4(x[t][i], c4, acc[t][j]);
      }
    }
    // Sum over the group's 8 lanes.  A plain butterfly would move every value three times; instead the first level(s) hand
    // half of the ROWS to each half of the group (a lane sends the partial sums of the rows it gives up and keeps the
    // others), so afterwards lanes [t * 8 / kR, (t + 1) * 8 / kR) own row t: kR = 2: 5 + 2 * 5 shuffles instead of 30.
    float w[kK + 1];
    int my_t = 0;
    if constexpr (kR == 4) {
      const bool h2 = (l8 & 4) != 0, h1 = (l8 & 2) != 0;
      float u[2][kK + 1];
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
#pragma unroll
        for (int j = 0; j < kK; ++j)
          u[s2][j] = (h2 ? acc[2 + s2][j] : acc[s2][j]) + __shfl_xor_sync(kFullMask, h2 ? acc[s2][j] : acc[2 + s2][j], 4);
        u[s2][kK] = (h2 ? xx[2 + s2] : xx[s2]) + __shfl_xor_sync(kFullMask, h2 ? xx[s2] : xx[2 + s2], 4);
      }
#pragma unroll
      for (int j = 0; j <= kK; ++j) w[j] = (h1 ? u[1][j] : u[0][j]) + __shfl_xor_sync(kFullMask, h1 ? u[0][j] : u[1][j], 2);
#pragma unroll
      for (int j = 0; j <= kK; ++j) w[j] += __shfl_xor_sync(kFullMask, w[j], 1);
      my_t = (h2 ? 2 : 0) + (h1 ? 1 : 0);
    } else if constexpr (kR == 2) {
      const bool h2 = (l8 & 4) != 0;
#pragma unroll
      for (int j = 0; j < kK; ++j) w[j] = (h2 ? acc[1][j] : acc[0][j]) + __shfl_xor_sync(kFullMask, h2 ? acc[0][j] : acc[1][j], 4);
      w[kK] = (h2 ? xx[1] : xx[0]) + __shfl_xor_sync(kFullMask, h2 ? xx[0] : xx[1], 4);
#pragma unroll
      for (int o = 2; o > 0; o >>= 1)
#pragma unroll
        for (int j = 0; j <= kK; ++j) w[j] += __shfl_xor_sync(kFullMask, w[j], o);
      my_t = h2 ? 1 : 0;
    } else {
#pragma unroll
      for (int j = 0; j < kK; ++j) w[j] = acc[0][j];
      w[kK] = xx[0];
#pragma unroll
      for (int o = 4; o > 0; o >>= 1)
#pragma unroll
        for (int j = 0; j <= kK; ++j) w[j] += __shfl_xor_sync(kFullMask, w[j], o);
    }
    // the nearest centre of the row this lane owns ...
    int best_m = 0;
    float bv = fmaf(-2.f, w[0], cn[0]);
#pragma unroll
    for (int j = 1; j < kK; ++j) {
      const float v = fmaf(-2.f, w[j], cn[j]);
      if (v < bv) bv = v, best_m = j;
    }
    if (p0 + my_t < m1) {
      const float d = fmaxf(w[kK] + bv, 0.f);
      if ((l8 & (8 / kR - 1)) == 0 && a.row_dist != nullptr) a.row_dist[p0 + my_t] = d;
#pragma unroll
      for (int j = 0; j < kK; ++j) cnt[j] += j == best_m ? 1 : 0, inr[j] += j == best_m ? d : 0.f;
    }
    // ... and every lane adds its D / 8 values of each of the group's rows to that row's centre
#pragma unroll
    for (int t = 0; t < kR; ++t) {
      const int best = kR == 1 ? best_m : __shfl_sync(kFullMask, best_m, (lane & 24) | (t * (8 / kR)));
      if (p0 + t < m1) {
        float4* slot = sm_acc + best * kV * threads + tid;
#pragma unroll
        for (int i = 0; i < kV; ++i) {
          float4 s = slot[i * threads];
          s.x += x[t][i].x, s.y += x[t][i].y, s.z += x[t][i].z, s.w += x[t][i].w;
          slot[i * threads] = s;
        }
      }
    }
  }
  // counts / inertia were kept by the lanes that own a row: fold the owners of a group into its lane 0 (fixed order)
  if constexpr (kR >= 2) {
#pragma unroll
    for (int j = 0; j < kK; ++j) cnt[j] += __shfl_xor_sync(kFullMask, cnt[j], 4), inr[j] += __shfl_xor_sync(kFullMask, inr[j], 4);
  }
  if constexpr (kR == 4) {
#pragma unroll
    for (int j = 0; j < kK; ++j) cnt[j] += __shfl_xor_sync(kFullMask, cnt[j], 2), inr[j] += __shfl_xor_sync(kFullMask, inr[j], 2);
  }
  __syncthreads();

  // fixed-order reduction of the lane-private accumulators: output float4 (j, i, l8) = sum over warps and row groups
  float4* out = reinterpret_cast<float4*>(a.part_sum + static_cast<int64_t>(cta) * k * D);
  for (int o = tid; o < k * kV * 8; o += threads) {
    const int ji = o >> 3, e8 = o & 7;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int w = 0; w < nwarps * 4; ++w) {
      const float4 v = sm_acc[ji * threads + w * 8 + e8];
      s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
    }
    out[o] = s;
  }
  __syncthreads();
  // counts and inertia: the group leaders' registers through the (now free) accumulator space
  constexpr int groups = threads / 8;
  int32_t* sm_cnt = reinterpret_cast<int32_t*>(sm_acc);           // [kK][groups]
  float* sm_inr = reinterpret_cast<float*>(sm_acc) + kK * groups;
  if (l8 == 0) {
#pragma unroll
    for (int j = 0; j < kK; ++j) sm_cnt[j * groups + (tid >> 3)] = cnt[j], sm_inr[j * groups + (tid >> 3)] = inr[j];
  }
  __syncthreads();
  for (int j = warp; j < k; j += nwarps) {                        // a warp per centre
    int32_t n_j = 0;
    float in_j = 0.f;
    for (int q = lane; q < groups; q += 32) n_j += sm_cnt[j * groups + q], in_j += sm_inr[j * groups + q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_j += __shfl_xor_sync(kFullMask, n_j, o), in_j += __shfl_xor_sync(kFullMask, in_j, o);
    if (lane == 0) a.part_count[cta * k + j] = n_j, a.part_inertia[cta * k + j] = in_j;
  }
}

// One CTA of 256 threads per (class, centre): D / 4 element lanes x 256 / (D / 4) slices of the class's CTAs.
__global__ void __launch_bounds__(256) kmeans_update_kernel(const KmArgs a) {
  const int c = static_cast<int>(blockIdx.x) / a.k, j = static_cast<int>(blockIdx.x) % a.k;
  const int tid = static_cast<int>(threadIdx.x), warp = tid >> 5, lane = tid & 31;
  const int D = a.D, k = a.k;
  if (a.done != nullptr && a.done[c] != 0) return;
  __shared__ float4 sm_slice[256];
  __shared__ float sm_red[8];
  __shared__ int32_t sm_cnt[8];
  __shared__ double sm_inr[8];
  const int b0 = a.cta_begin[c], b1 = a.cta_begin[c + 1];
  {                                                               // rows and inertia of this centre (the grid has <= 256 CTAs)
    int32_t n_t = 0;
    double in_t = 0.0;
    if (b0 + tid < b1) n_t = a.part_count[(b0 + tid) * k + j], in_t = static_cast<double>(a.part_inertia[(b0 + tid) * k + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n_t += __shfl_xor_sync(kFullMask, n_t, o), in_t += __shfl_xor_sync(kFullMask, in_t, o);
    if (lane == 0) sm_cnt[warp] = n_t, sm_inr[warp] = in_t;
  }
  __syncthreads();
  int64_t n_j = 0;
  double in_j = 0.0;
#pragma unroll
  for (int w = 0; w < 8; ++w) n_j += sm_cnt[w], in_j += sm_inr[w];
  if (tid == 0) {
    if (a.counts != nullptr) a.counts[c * k + j] = n_j;
    if (a.inertia != nullptr) a.inertia[c * k + j] = static_cast<float>(in_j);
  }
  if (a.update == 0 && a.sums == nullptr) return;
  const int lanes = D / 4, slices = 256 / lanes;                  // D in 32..512 -> 8..128 lanes, 32..2 slices
  const int e = tid % lanes, sl = tid / lanes;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int b = b0 + sl; b < b1; b += slices) {
    const float4 v = *reinterpret_cast<const float4*>(a.part_sum + (static_cast<int64_t>(b) * k + j) * D + e * 4);
    s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
  }
  sm_slice[tid] = s;
  __syncthreads();
  float d2 = 0.f;
  if (tid < lanes) {
    for (int q = 1; q < slices; ++q) {
      const float4 v = sm_slice[q * lanes + tid];
      s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
    }
    if (a.sums != nullptr) *reinterpret_cast<float4*>(a.sums + (static_cast<int64_t>(c) * k + j) * D + tid * 4) = s;
    if (a.update != 0) {
      float4* dst = reinterpret_cast<float4*>(a.centres + (static_cast<int64_t>(c) * k + j) * D + tid * 4);
      const float4 old = *dst;
      float4 nw = old;
      if (n_j > 0) {                                              // an empty cluster keeps its centre
        const float inv = 1.f / static_cast<float>(n_j);
        nw = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
      }
      const float ex = nw.x - old.x, ey = nw.y - old.y, ez = nw.z - old.z, ew = nw.w - old.w;
      d2 = ex * ex + ey * ey + ez * ez + ew * ew;
      *dst = nw;
    }
  }
  if (a.update == 0) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(kFullMask, d2, o);
  if (lane == 0) sm_red[warp] = d2;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += sm_red[w];
    a.shift[c * k + j] = tot;
    __threadfence();
    if (atomicAdd(a.arrive + c, 1) == k - 1) {                    // the class's last centre: stop rule of _kmeans_single_lloyd
      __threadfence();
      float all = 0.f;
      for (int q = 0; q < k; ++q) all += __ldcg(a.shift + c * k + q);
      if (a.tol != nullptr && a.done != nullptr && all <= a.tol[c]) a.done[c] = 1;
      a.arrive[c] = 0;
    }
  }
}

struct KmPlan {
  size_t off_sum, off_inertia, off_count, off_shift, off_arrive, total;
};

KmPlan make_km_plan(int32_t C, int32_t k, int32_t D) {
  KmPlan p{};
  size_t o = 0;
  auto take = [&](size_t bytes) {
    const size_t at = o;
    o += (bytes + 255) & ~static_cast<size_t>(255);
    return at;
  };
  p.off_sum = take(sizeof(float) * kKmMaxGrid * k * D);
  p.off_inertia = take(sizeof(float) * kKmMaxGrid * k);
  p.off_count = take(sizeof(int32_t) * kKmMaxGrid * k);
  p.off_shift = take(sizeof(float) * C * k);
  p.off_arrive = take(sizeof(int32_t) * C);
  p.total = o;
  return p;
}

template <int kV, int kK>
int launch_assign(const KmArgs& a, int grid, cudaStream_t st) {
  constexpr int threads = km_threads(kV, kK);
  constexpr size_t smem = sizeof(float4) * (kK * kV * 8 + 2 + static_cast<size_t>(kK) * kV * threads);
  if (smem > 48 * 1024)                                           // per device and per call: the attribute is device state
    MML_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel<kV, kK>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kmeans_assign_kernel<kV, kK><<<grid, threads, smem, st>>>(a);
  return check_launch("kmeans_assign_kernel");
}

template <int kV>
int launch_assign_k(const KmArgs& a, int grid, cudaStream_t st) {
  if (a.k <= 2) return launch_assign<kV, 2>(a, grid, st);
  if (a.k <= 4) return launch_assign<kV, 4>(a, grid, st);
  return launch_assign<kV, 8>(a, grid, st);
}

int km_threads_for(int kV, int k) {
  const int kK = k <= 2 ? 2 : (k <= 4 ? 4 : 8);
  return km_threads(kV, kK);
}

int km_rows_iter(int kV, int k) { return 4 * (kV <= 2 ? 4 : (kV == 4 ? (k > 4 ? 4 : 2) : 1)); }

}  // namespace
}  // namespace mml

using namespace mml;

extern "C" int32_t mml_crd_kmeans_max_clusters(void) { return kKmMaxK; }

extern "C" int64_t mml_crd_kmeans_workspace_bytes(int32_t n_classes, int32_t k, int32_t D) {
  if (n_classes < 1 || n_classes > kKmMaxClasses || k < 1 || k > kKmMaxK || D < 32) return 0;
  return static_cast<int64_t>(make_km_plan(n_classes, k, D).total);
}

extern "C" int mml_crd_kmeans_lloyd(const float* bank, int64_t n, int32_t D, const int64_t* rows, const int64_t* class_offsets,
                                    int32_t n_classes, int32_t k, float* centres, const float* tol, int32_t iterations, int32_t update,
                                    int32_t* done, float* inertia, int64_t* counts, float* sums, float* row_dist, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  MML_REQUIRE(bank && rows && class_offsets && centres && workspace, MML_ERR_INVALID_ARG, "crd_kmeans_lloyd: null pointer");
  MML_REQUIRE(n >= 1 && iterations >= 0, MML_ERR_INVALID_ARG, "crd_kmeans_lloyd: bad sizes");
  MML_REQUIRE(D == 32 || D == 64 || D == 128 || D == 256 || D == 512, MML_ERR_UNSUPPORTED,
              "crd_kmeans_lloyd: feature width 32 / 64 / 128 / 256 / 512 supported (got %d)", D);
  MML_REQUIRE(k >= 1 && k <= kKmMaxK, MML_ERR_UNSUPPORTED, "crd_kmeans_lloyd: 1 <= clusters per class <= %d (got %d)", kKmMaxK, k);
  MML_REQUIRE(n_classes >= 1 && n_classes <= kKmMaxClasses, MML_ERR_UNSUPPORTED, "crd_kmeans_lloyd: 1 <= classes <= %d (got %d)",
              kKmMaxClasses, n_classes);
  MML_REQUIRE(aligned16(bank) && aligned16(centres) && aligned16(workspace) && aligned16(sums), MML_ERR_INVALID_ARG,
              "crd_kmeans_lloyd: bank / centres / sums / workspace must be 16-byte aligned");
  const KmPlan p = make_km_plan(n_classes, k, D);
  MML_REQUIRE(workspace_bytes >= p.total, MML_ERR_WORKSPACE, "crd_kmeans_lloyd: workspace too small (%zu < %zu)", workspace_bytes, p.total);
  MML_REQUIRE(class_offsets[0] == 0, MML_ERR_INVALID_ARG, "crd_kmeans_lloyd: class_offsets[0] must be 0");
  for (int c = 0; c < n_classes; ++c)      // a class may be empty HERE (its rows live on other shards), the list as a whole may not
    MML_REQUIRE(class_offsets[c + 1] >= class_offsets[c], MML_ERR_INVALID_ARG, "crd_kmeans_lloyd: class_offsets must not decrease (class %d)", c);
  MML_REQUIRE(class_offsets[n_classes] >= 1, MML_ERR_INVALID_ARG, "crd_kmeans_lloyd: no rows listed");
  if (iterations == 0) return MML_OK;

  const int threads = km_threads_for(D / 32, k);
  const int rows_iter = km_rows_iter(D / 32, k);
  KmArgs a{};
  a.bank = bank, a.rows = rows, a.centres = centres, a.tol = tol, a.done = done, a.inertia = inertia, a.counts = counts, a.sums = sums;
  a.row_dist = row_dist;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  a.part_sum = reinterpret_cast<float*>(ws + p.off_sum);
  a.part_inertia = reinterpret_cast<float*>(ws + p.off_inertia);
  a.part_count = reinterpret_cast<int32_t*>(ws + p.off_count);
  a.shift = reinterpret_cast<float*>(ws + p.off_shift);
  a.arrive = reinterpret_cast<int32_t*>(ws + p.off_arrive);
  a.err = device_error_word();
  a.n = n, a.D = D, a.k = k, a.C = n_classes, a.update = update;

  // CTAs per class in proportion to its rows (at least one, at most one per warp-iteration of work)
  const int64_t m = class_offsets[n_classes];
  const int spare = kKmMaxGrid - n_classes;
  int grid = 0;
  for (int c = 0; c < n_classes; ++c) {
    a.row_begin[c] = class_offsets[c];
    a.cta_begin[c] = grid;
    const int64_t mc = class_offsets[c + 1] - class_offsets[c];
    const int64_t iters = (mc + rows_iter - 1) / rows_iter;
    int64_t want = 1 + static_cast<int64_t>(static_cast<double>(spare) * static_cast<double>(mc) / static_cast<double>(m));
    const int64_t useful = (iters + (threads / 32) - 1) / (threads / 32);
    if (want > useful) want = useful;
    if (want < 1) want = 1;
    grid += static_cast<int>(want);
  }
  a.row_begin[n_classes] = m;
  a.cta_begin[n_classes] = grid;

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MML_CUDA(cudaMemsetAsync(a.arrive, 0, sizeof(int32_t) * n_classes, st));
  for (int it = 0; it < iterations; ++it) {
    int rc;
    switch (D) {
      case 32: rc = launch_assign_k<1>(a, grid, st); break;
      case 64: rc = launch_assign_k<2>(a, grid, st); break;
      case 128: rc = launch_assign_k<4>(a, grid, st); break;
      case 256: rc = launch_assign_k<8>(a, grid, st); break;
      default: rc = launch_assign_k<16>(a, grid, st); break;
    }
    if (rc != MML_OK) return rc;
    kmeans_update_kernel<<<n_classes * k, 256, 0, st>>>(a);
    rc = check_launch("kmeans_update_kernel");
    if (rc != MML_OK) return rc;
  }
  return MML_OK;
}
