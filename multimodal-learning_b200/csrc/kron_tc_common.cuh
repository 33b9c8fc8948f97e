// Shared pieces of the tcgen05 / TMEM / TMA Kronecker kernels (forward: kron_tc.cu, backward: kron_tc_bwd.cu).
#pragma once

#include <cuda.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "kron_common.cuh"

namespace mml {
namespace tc {

constexpr int kTileM = 128;
constexpr int kChunkK = 32;                 // tf32 elements per chunk = one 128-byte swizzle row
constexpr int kGenWarps = 8;                // warp g: TMEM lanes 32*(g%4).., chunk columns 16*(g/4)..
constexpr int kGenThreads = kGenWarps * 32;
constexpr int kThreadsTc = kGenThreads + 64;   // + TMA warp + MMA/alloc warp
constexpr int kHalf = kChunkK / 2;          // columns of a chunk handled by one generator thread
constexpr int kMaxSmemTable = 48 * 1024;    // chunk tables up to this size are staged in shared memory
constexpr uint32_t kSpinLimit = 1u << 28;   // mbarrier spin guard: trap instead of hanging the GPU
// Operand rounding.  The tensor core TRUNCATES fp32 operands to TF32 (10 mantissa bits).  Operands that are packed once
// (weights, dy^T) are rounded to nearest when they are packed, and so is the forward's generated Kronecker operand (one
// integer add per element).  In the weight-gradient kernel the generated operand A^T is summed over thousands of batch
// rows, so there it is left to the truncation, and the mean of the truncation error (-0.5 ulp, i.e. -0.5 * 2^-10 / m relative for mantissa m,
// -0.69 * 2^-11 averaged over log-uniform mantissas) is cancelled by scaling the OTHER, packed operand with this
// constant before it is rounded: the per-element error keeps the RMS of round-to-nearest (0.29-0.31 ulp) and the sum has no
// systematic bias although all Kronecker factors are >= 0 (post-ReLU).
constexpr float kTruncComp = 1.0f + 0.69f / 2048.0f;

struct Chunk {                // 32 bytes, read as two int4
  int32_t p, q, vsrc, vcol;
  int32_t vlen, kbase, kstride, pad;
};

// ---------------------------------------------------------------- chunk table (host)
inline std::vector<Chunk> build_chunks(int d1, int d2, int d3) {
  std::vector<Chunk> out;
  const int e2 = d2 + 1, e3 = d3 > 0 ? d3 + 1 : 1;
  const int P1 = 1, P2 = 1 + d1;                      // positions of f1[0], f2[0] in R = [1, f1, f2]
  auto seg = [](int d, int s) { return d - s < kChunkK ? d - s : kChunkK; };
  auto add = [&](int p, int q, int vsrc, int vcol, int vlen, int kbase, int kstride) {
    out.push_back(Chunk{p, q, vsrc, vcol, vlen, kbase, kstride, 0});
  };
  // Chunks that share a vector segment (vsrc, vcol) are emitted as ONE contiguous run: the dgrad epilogue keeps the
  // segment's gradient in registers and writes it out once per run (a plain store, no read-modify-write).
  if (d3 == 0) {
    for (int js = 0; js < d2; js += kChunkK) {
      for (int i = 0; i < d1; ++i) add(P1 + i, 0, 2, js, seg(d2, js), i * e2 + js, 1);          // core: o1[i] * o2[js..]
      add(0, 0, 2, js, seg(d2, js), d1 * e2 + js, 1);                                            // i = d1 border: 1 * o2[js..]
    }
    for (int is = 0; is < d1; is += kChunkK)          // j = d2 border: o1[is..] * 1
      add(0, 0, 1, is, seg(d1, is), is * e2 + d2, e2);
    add(0, 0, 0, 0, 1, d1 * e2 + d2, 1);              // corner 1*1
  } else {
    for (int ls = 0; ls < d3; ls += kChunkK) {
      for (int i = 0; i < d1; ++i)                    // core: o1[i] o2[j] * o3[ls..]
        for (int j = 0; j < d2; ++j) add(P1 + i, P2 + j, 3, ls, seg(d3, ls), (i * e2 + j) * e3 + ls, 1);
      for (int i = 0; i < d1; ++i)                    // j = d2 face: o1[i] * o3[ls..]
        add(P1 + i, 0, 3, ls, seg(d3, ls), (i * e2 + d2) * e3 + ls, 1);
      for (int j = 0; j < d2; ++j)                    // i = d1 face: o2[j] * o3[ls..]
        add(P2 + j, 0, 3, ls, seg(d3, ls), (d1 * e2 + j) * e3 + ls, 1);
      add(0, 0, 3, ls, seg(d3, ls), (d1 * e2 + d2) * e3 + ls, 1);     // i = d1, j = d2 edge: o3[ls..]
    }
    for (int js = 0; js < d2; js += kChunkK) {
      for (int i = 0; i < d1; ++i)                    // l = d3 face: o1[i] * o2[js..]
        add(P1 + i, 0, 2, js, seg(d2, js), (i * e2 + js) * e3 + d3, e3);
      add(0, 0, 2, js, seg(d2, js), (d1 * e2 + js) * e3 + d3, e3);    // i = d1, l = d3 edge: o2[js..]
    }
    for (int is = 0; is < d1; is += kChunkK)          // j = d2, l = d3 edge: o1[is..]
      add(0, 0, 1, is, seg(d1, is), (is * e2 + d2) * e3 + d3, e2 * e3);
    add(0, 0, 0, 0, 1, (d1 * e2 + d2) * e3 + d3, 1);  // corner
  }
  return out;
}

inline int round_np(int N) { return (N + 15) / 16 * 16; }

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int32_t c0, int32_t c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// One lane of a fully converged warp.  ptxas knows the region guarded by elect.sync runs on a single thread, so the
// uniform-datapath instructions inside it (UTCHMMA / UTCBAR / UTMALDG) are issued directly; guarding them with
// `lane == 0` instead makes the compiler wrap every one of them in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop, which
// costs ~150 issue cycles per MMA and was THE bound of the round-1 kernels (profiles/r2_*).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred)::"memory");
  return pred != 0;
}
// warp index as a provably warp-uniform value (role dispatch without divergence)
__device__ __forceinline__ int warp_idx_sync() { return __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc],  kind::tf32, M = 128, K = 8
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tc_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row swizzle atoms 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);          // start address
  d |= static_cast<uint64_t>(1) << 16;                               // LBO (unused for swizzled K-major) = 1
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                       // SBO = 1024 B between 8-row groups
  d |= static_cast<uint64_t>(1) << 46;                               // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                               // SWIZZLE_128B
  return d;
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  int64_t dim0, dim1;           // elements along the contiguous / the strided dimension
  int32_t box0, box1, swizzle;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && dim0 == o.dim0 && dim1 == o.dim1 && box0 == o.box0 && box1 == o.box1 && swizzle == o.swizzle;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h = h * 1000003u ^ static_cast<size_t>(k.dim0);
    h = h * 1000003u ^ static_cast<size_t>(k.dim1);
    h = h * 1000003u ^ static_cast<size_t>(k.box0 * 4099 + k.box1 * 2 + k.swizzle);
    return h;
  }
};

// TMA descriptor of a row-major fp32 matrix [dim1 rows][dim0 contiguous], boxes of box0 x box1 elements, 128-byte swizzle
// or none.  Cached by pointer + geometry (the only global state of the library).
inline int get_tensor_map_2d(const float* base, int64_t dim0, int64_t dim1, int32_t box0, int32_t box1, bool swizzle128,
                             CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  std::lock_guard<std::mutex> lock(mu);
  const MapKey key{base, dim0, dim1, box0, box1, swizzle128 ? 1 : 0};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return MML_OK;
  }
  EncodeTiledFn enc = get_encode_fn();
  MML_REQUIRE(enc != nullptr, MML_ERR_CUDA, "kron: cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMap m;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(dim0), static_cast<cuuint64_t>(dim1)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(dim0) * sizeof(float)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(box0), static_cast<cuuint32_t>(box1)};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estride,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MML_REQUIRE(r == CUDA_SUCCESS, MML_ERR_CUDA, "kron: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  if (cache.size() > 512) cache.clear();
  cache[key] = m;
  *out = m;
  return MML_OK;
}

// packed operand [Np rows][Kp contiguous]: boxes of [Np x 32] (one 128-byte swizzle row per operand row)
inline int get_tensor_map(const float* Wp, int32_t Np, int32_t Kp, CUtensorMap* out) {
  return get_tensor_map_2d(Wp, Kp, Np, kChunkK, Np, true, out);
}

inline uint32_t make_idesc_tf32(int M, int N) {
  uint32_t d = 0;
  d |= 1u << 4;                                   // D format: F32
  d |= 2u << 7;                                   // A format: TF32
  d |= 2u << 10;                                  // B format: TF32
  d |= static_cast<uint32_t>(N >> 3) << 17;       // N / 8
  d |= static_cast<uint32_t>(M >> 4) << 24;       // M / 16
  return d;                                       // A, B K-major; no negate; dense
}


}  // namespace tc
}  // namespace mml
