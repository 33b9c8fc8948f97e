// K13 -- per-class k-means centres of a memory bank (the extra positives of pos_extra == "centers" with more than two positives).
//
// Replaces, for one bank, the host round trip of the reference (CL_utils/CRD_criterion_v10.py:84-92 and :122-129):
//     for c in classes:  feature = index_select(memory, 0, class_idx[c]).cpu().numpy()
//                        KMeans(n_clusters = num_pos - 1).fit(feature).cluster_centers_
// by Lloyd iterations on the device.  One iteration is two launches:
//
//   kmeans_assign_kernel  one pass over the listed bank rows (HBM-bound: every row is read exactly once).  A CTA belongs to one
//                         class and holds that class's k centres in shared memory.  Eight lanes share a row (D/8 values each,
//                         128-byte coalesced segments) and every group keeps kR rows in flight, so a centre chunk read from
//                         shared memory is used 4 * kR times.  Scores are sklearn's E-step form |c|^2 - 2 x.c (lowest index wins
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

template <int kV, int kR>
__global__ void __launch_bounds__(256) kmeans_assign_kernel(const KmArgs a) {
  extern __shared__ float4 km_smem[];
  constexpr int D = 32 * kV;
  const int threads = static_cast<int>(blockDim.x);
  const int tid = static_cast<int>(threadIdx.x);
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 3, l8 = lane & 7;
  const int cta = static_cast<int>(blockIdx.x);
  const int k = a.k;

  int c = 0;
  while (c + 1 < a.C && cta >= a.cta_begin[c + 1]) ++c;
  if (a.done != nullptr && a.done[c] != 0) return;

  float4* sm_centre = km_smem;                                   // [k][kV * 8]
  float* sm_cnorm = reinterpret_cast<float*>(sm_centre + kKmMaxK * kV * 8);   // [8]
  float4* sm_acc = sm_centre + kKmMaxK * kV * 8 + 2;             // [k * kV][threads], lane-private
  const float4* centre_g = reinterpret_cast<const float4*>(a.centres + static_cast<int64_t>(c) * k * D);
  for (int i = tid; i < k * kV * 8; i += threads) sm_centre[i] = centre_g[i];
  for (int i = 0; i < k * kV; ++i) sm_acc[i * threads + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  for (int j = warp; j < k; j += threads >> 5) {                  // |c_j|^2, a warp per centre
    float s = 0.f;
    for (int i = lane; i < kV * 8; i += 32) {
      const float4 v = sm_centre[j * kV * 8 + i];
      s = dot4(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFullMask, s, o);
    if (lane == 0) sm_cnorm[j] = s;
  }
  __syncthreads();
  float cn[kKmMaxK];
#pragma unroll
  for (int j = 0; j < kKmMaxK; ++j) cn[j] = j < k ? sm_cnorm[j] : 0.f;

  // this CTA's contiguous slice of the class's row list, in whole warp-iterations of 4 * kR rows
  const int64_t m0 = a.row_begin[c], m1 = a.row_begin[c + 1];
  const int nctas = a.cta_begin[c + 1] - a.cta_begin[c];
  constexpr int kRowsIter = 4 * kR;
  const int64_t iters = (m1 - m0 + kRowsIter - 1) / kRowsIter;
  const int64_t per = (iters + nctas - 1) / nctas;
  const int64_t it0 = per * (cta - a.cta_begin[c]);
  const int64_t it1 = it0 + per < iters ? it0 + per : iters;
  const int nwarps = threads >> 5;

  int32_t cnt[kKmMaxK];
  float inr[kKmMaxK];
#pragma unroll
  for (int j = 0; j < kKmMaxK; ++j) cnt[j] = 0, inr[j] = 0.f;

  for (int64_t it = it0 + warp; it < it1; it += nwarps) {
    const int64_t p0 = m0 + it * kRowsIter + g * kR;
    float4 x[kR][kV];
    bool valid[kR];
#pragma unroll
    for (int t = 0; t < kR; ++t) {
      valid[t] = p0 + t < m1;
      int64_t rid = a.rows[valid[t] ? p0 + t : m0];
      if (rid < 0 || rid >= a.n) {                                // a bad row index: flag it, read row 0
        flag_device_error(a.err, MML_DEVERR_CRD_INDEX);
        rid = 0;
      }
      const float* src = a.bank + rid * D + l8 * 4;
#pragma unroll
      for (int i = 0; i < kV; ++i) x[t][i] = ldg_stream_f4(src + i * 32);
    }
    float acc[kR][kKmMaxK], xx[kR];
#pragma unroll
    for (int t = 0; t < kR; ++t) {
      xx[t] = 0.f;
#pragma unroll
      for (int i = 0; i < kV; ++i) xx[t] = dot4(x[t][i], x[t][i], xx[t]);
#pragma unroll
      for (int j = 0; j < kKmMaxK; ++j) acc[t][j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < kKmMaxK; ++j) {
      if (j < k) {
#pragma unroll
        for (int i = 0; i < kV; ++i) {
          const float4 c4 = sm_centre[(j * kV + i) * 8 + l8];
#pragma unroll
          for (int t = 0; t < kR; ++t) acc[t][j] = dot4(x[t][i], c4, acc[t][j]);
        }
      }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
#pragma unroll
      for (int t = 0; t < kR; ++t) {
        xx[t] += __shfl_xor_sync(kFullMask, xx[t], o);
#pragma unroll
        for (int j = 0; j < kKmMaxK; ++j)
          if (j < k) acc[t][j] += __shfl_xor_sync(kFullMask, acc[t][j], o);
      }
    }
#pragma unroll
    for (int t = 0; t < kR; ++t) {
      int best = 0;
      float bv = fmaf(-2.f, acc[t][0], cn[0]);
#pragma unroll
      for (int j = 1; j < kKmMaxK; ++j) {
        const float v = fmaf(-2.f, acc[t][j], cn[j]);
        if (j < k && v < bv) bv = v, best = j;
      }
      if (valid[t]) {
        const float d = fmaxf(xx[t] + bv, 0.f);
        if (l8 == 0) {
          if (a.row_dist != nullptr) a.row_dist[p0 + t] = d;
#pragma unroll
          for (int j = 0; j < kKmMaxK; ++j)
            if (j == best) cnt[j] += 1, inr[j] += d;
        }
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
  __syncthreads();

  // fixed-order reduction of the lane-private accumulators: output float4 (j, i, l8) = sum over warps and row groups
  float4* out = reinterpret_cast<float4*>(a.part_sum + static_cast<int64_t>(cta) * k * D);
  for (int o = tid; o < k * kV * 8; o += threads) {
    const int ji = o >> 3, e8 = o & 7;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int w = 0; w < nwarps; ++w)
#pragma unroll
      for (int gg = 0; gg < 4; ++gg) {
        const float4 v = sm_acc[ji * threads + w * 32 + gg * 8 + e8];
        s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
      }
    out[o] = s;
  }
  __syncthreads();
  // counts and inertia: the group leaders' registers through the (now free) accumulator space
  int32_t* sm_cnt = reinterpret_cast<int32_t*>(sm_acc);           // [k][threads / 8]
  float* sm_inr = reinterpret_cast<float*>(sm_acc) + kKmMaxK * (threads >> 3);
  if (l8 == 0) {
#pragma unroll
    for (int j = 0; j < kKmMaxK; ++j)
      if (j < k) sm_cnt[j * (threads >> 3) + (tid >> 3)] = cnt[j], sm_inr[j * (threads >> 3) + (tid >> 3)] = inr[j];
  }
  __syncthreads();
  if (tid < k) {
    int32_t n_j = 0;
    float in_j = 0.f;
    for (int q = 0; q < (threads >> 3); ++q) n_j += sm_cnt[tid * (threads >> 3) + q], in_j += sm_inr[tid * (threads >> 3) + q];
    a.part_count[cta * k + tid] = n_j;
    a.part_inertia[cta * k + tid] = in_j;
  }
}

// One CTA per (class, centre), D / 4 threads.
__global__ void kmeans_update_kernel(const KmArgs a) {
  const int c = static_cast<int>(blockIdx.x) / a.k, j = static_cast<int>(blockIdx.x) % a.k;
  const int tid = static_cast<int>(threadIdx.x);
  const int D = a.D, k = a.k;
  if (a.done != nullptr && a.done[c] != 0) return;
  __shared__ float sm_red[32];
  __shared__ int64_t sm_count;
  const int b0 = a.cta_begin[c], b1 = a.cta_begin[c + 1];
  if (tid == 0) {
    int64_t n_j = 0;
    double in_j = 0.0;
    for (int b = b0; b < b1; ++b) n_j += a.part_count[b * k + j], in_j += static_cast<double>(a.part_inertia[b * k + j]);
    sm_count = n_j;
    if (a.counts != nullptr) a.counts[c * k + j] = n_j;
    if (a.inertia != nullptr) a.inertia[c * k + j] = static_cast<float>(in_j);
  }
  __syncthreads();
  if (a.update == 0) return;
  const int64_t n_j = sm_count;
  float d2 = 0.f;
  if (tid * 4 < D) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = b0; b < b1; ++b) {
      const float4 v = *reinterpret_cast<const float4*>(a.part_sum + (static_cast<int64_t>(b) * k + j) * D + tid * 4);
      s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
    }
    float4* dst = reinterpret_cast<float4*>(a.centres + (static_cast<int64_t>(c) * k + j) * D + tid * 4);
    const float4 old = *dst;
    float4 nw = old;
    if (n_j > 0) {                                                // an empty cluster keeps its centre
      const float inv = 1.f / static_cast<float>(n_j);
      nw = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
    }
    const float ex = nw.x - old.x, ey = nw.y - old.y, ez = nw.z - old.z, ew = nw.w - old.w;
    d2 = ex * ex + ey * ey + ez * ez + ew * ew;
    *dst = nw;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d2 += __shfl_xor_sync(kFullMask, d2, o);
  if ((tid & 31) == 0) sm_red[tid >> 5] = d2;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int w = 0; w < (static_cast<int>(blockDim.x) + 31) / 32; ++w) tot += sm_red[w];
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

template <int kV, int kR>
int launch_assign(const KmArgs& a, int grid, int threads, cudaStream_t st) {
  const size_t smem = sizeof(float4) * (kKmMaxK * kV * 8 + 2 + static_cast<size_t>(a.k) * kV * threads);
  if (smem > 48 * 1024)                                           // per device and per call: the attribute is device state
    MML_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel<kV, kR>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kmeans_assign_kernel<kV, kR><<<grid, threads, smem, st>>>(a);
  return check_launch("kmeans_assign_kernel");
}

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
                                    int32_t* done, float* inertia, int64_t* counts, float* row_dist, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  MML_REQUIRE(bank && rows && class_offsets && centres && workspace, MML_ERR_INVALID_ARG, "crd_kmeans_lloyd: null pointer");
  MML_REQUIRE(n >= 1 && iterations >= 0, MML_ERR_INVALID_ARG, "crd_kmeans_lloyd: bad sizes");
  MML_REQUIRE(D == 32 || D == 64 || D == 128 || D == 256 || D == 512, MML_ERR_UNSUPPORTED,
              "crd_kmeans_lloyd: feature width 32 / 64 / 128 / 256 / 512 supported (got %d)", D);
  MML_REQUIRE(k >= 1 && k <= kKmMaxK, MML_ERR_UNSUPPORTED, "crd_kmeans_lloyd: 1 <= clusters per class <= %d (got %d)", kKmMaxK, k);
  MML_REQUIRE(n_classes >= 1 && n_classes <= kKmMaxClasses, MML_ERR_UNSUPPORTED, "crd_kmeans_lloyd: 1 <= classes <= %d (got %d)",
              kKmMaxClasses, n_classes);
  MML_REQUIRE(aligned16(bank) && aligned16(centres) && aligned16(workspace), MML_ERR_INVALID_ARG,
              "crd_kmeans_lloyd: bank / centres / workspace must be 16-byte aligned");
  const KmPlan p = make_km_plan(n_classes, k, D);
  MML_REQUIRE(workspace_bytes >= p.total, MML_ERR_WORKSPACE, "crd_kmeans_lloyd: workspace too small (%zu < %zu)", workspace_bytes, p.total);
  MML_REQUIRE(class_offsets[0] == 0, MML_ERR_INVALID_ARG, "crd_kmeans_lloyd: class_offsets[0] must be 0");
  for (int c = 0; c < n_classes; ++c)
    MML_REQUIRE(class_offsets[c + 1] > class_offsets[c], MML_ERR_INVALID_ARG, "crd_kmeans_lloyd: class %d has no rows", c);
  if (iterations == 0) return MML_OK;

  // rows in flight per 8-lane group and CTA size by feature width (the lane-private accumulators are k * D / 2 bytes per thread)
  const int threads = D <= 128 ? 256 : (D == 256 ? 128 : 64);
  const int rows_iter = 4 * (D <= 128 ? 4 : (D == 256 ? 2 : 1));
  KmArgs a{};
  a.bank = bank, a.rows = rows, a.centres = centres, a.tol = tol, a.done = done, a.inertia = inertia, a.counts = counts;
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
      case 32: rc = launch_assign<1, 4>(a, grid, threads, st); break;
      case 64: rc = launch_assign<2, 4>(a, grid, threads, st); break;
      case 128: rc = launch_assign<4, 4>(a, grid, threads, st); break;
      case 256: rc = launch_assign<8, 2>(a, grid, threads, st); break;
      default: rc = launch_assign<16, 1>(a, grid, threads, st); break;
    }
    if (rc != MML_OK) return rc;
    kmeans_update_kernel<<<n_classes * k, D / 4 < 32 ? 32 : D / 4, 0, st>>>(a);
    rc = check_launch("kmeans_update_kernel");
    if (rc != MML_OK) return rc;
  }
  return MML_OK;
}
