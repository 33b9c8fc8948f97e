// Small collectives over peer-mapped (symmetric) memory for the row-sharded CRD step: an all_gather done
// with NVLink STORES (every rank pushes its slice into every peer's buffer) and a reduce_scatter done with
// NVLink LOADS (every rank sums its own slice out of every peer's partial buffer, in rank order).  The
// messages are KBs..MBs, where a hand-rolled one-kernel exchange beats NCCL's launch + protocol latency.
#include "common.cuh"

namespace mml {
namespace {

struct PushArgs {
  char* base[32];
  const char* src[4];
  int64_t off[4];
  int64_t bytes[4];
  int32_t world, nseg;
};

__global__ void __launch_bounds__(256) symm_push_kernel(const PushArgs a) {
  // blockIdx.y = destination rank; the x-grid strides over 16-byte words of all segments
  char* dst = a.base[blockIdx.y];
  for (int sgi = 0; sgi < a.nseg; ++sgi) {
    const int64_t nvec = a.bytes[sgi] >> 4;
    const int4* s = reinterpret_cast<const int4*>(a.src[sgi]);
    int4* d = reinterpret_cast<int4*>(dst + a.off[sgi]);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
      d[i] = s[i];
  }
}

struct PullArgs {
  const char* base[32];
  int64_t part_off, tail_off, rows_total, row_begin, rows;
  int32_t world, D, nt;
  float* out1;
  float* out2;
  float* tail_out;
};

__global__ void __launch_bounds__(256) symm_pull_reduce_kernel(const PullArgs a) {
  const int64_t per_plane = a.rows * a.D;
  const int64_t total = 2 * per_plane;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int plane = i >= per_plane;
    const int64_t j = i - plane * per_plane;                        // = r * D + d
    const int64_t src = (static_cast<int64_t>(plane) * a.rows_total + a.row_begin) * a.D + j;
    float t = 0.f;
    for (int p = 0; p < a.world; ++p) t += reinterpret_cast<const float*>(a.base[p] + a.part_off)[src];
    (plane ? a.out2 : a.out1)[j] = t;
  }
  if (blockIdx.x == 0 && threadIdx.x < a.nt && a.tail_out != nullptr) {
    float t = 0.f;
    for (int p = 0; p < a.world; ++p) t += reinterpret_cast<const float*>(a.base[p] + a.tail_off)[threadIdx.x];
    a.tail_out[threadIdx.x] = t;
  }
}

}  // namespace
}  // namespace mml

using namespace mml;

extern "C" int mml_symm_push(void* const* peer_base_host, int32_t world, const void* const* src, const int64_t* dst_off,
                             const int64_t* bytes, int32_t nseg, void* stream) {
  MML_REQUIRE(peer_base_host && src && dst_off && bytes, MML_ERR_INVALID_ARG, "symm_push: null pointer");
  MML_REQUIRE(world >= 1 && world <= 32 && nseg >= 1 && nseg <= 4, MML_ERR_INVALID_ARG, "symm_push: world <= 32, 1..4 segments");
  PushArgs a{};
  a.world = world; a.nseg = nseg;
  int64_t maxvec = 0;
  for (int i = 0; i < world; ++i) {
    MML_REQUIRE(peer_base_host[i] != nullptr, MML_ERR_INVALID_ARG, "symm_push: null peer base %d", i);
    a.base[i] = static_cast<char*>(peer_base_host[i]);
  }
  for (int i = 0; i < nseg; ++i) {
    MML_REQUIRE(src[i] && bytes[i] >= 0 && (bytes[i] & 15) == 0 && (dst_off[i] & 15) == 0 && aligned16(src[i]),
                MML_ERR_INVALID_ARG, "symm_push: segment %d must be 16-byte aligned/sized", i);
    a.src[i] = static_cast<const char*>(src[i]); a.off[i] = dst_off[i]; a.bytes[i] = bytes[i];
    if ((bytes[i] >> 4) > maxvec) maxvec = bytes[i] >> 4;
  }
  if (maxvec == 0) return MML_OK;
  int64_t gx = (maxvec + 255) / 256;
  if (gx > 64) gx = 64;
  symm_push_kernel<<<dim3(static_cast<unsigned>(gx), world), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("symm_push_kernel");
}

extern "C" int mml_symm_pull_reduce(void* const* peer_base_host, int32_t world, int64_t part_off, int64_t rows_total,
                                    int64_t row_begin, int64_t rows, int32_t D, float* out1, float* out2,
                                    int64_t tail_off, int32_t nt, float* tail_out, void* stream) {
  MML_REQUIRE(peer_base_host && out1 && out2, MML_ERR_INVALID_ARG, "symm_pull_reduce: null pointer");
  MML_REQUIRE(world >= 1 && world <= 32 && rows >= 0 && D >= 1 && nt >= 0 && nt <= 32 && row_begin >= 0 &&
              row_begin + rows <= rows_total, MML_ERR_INVALID_ARG, "symm_pull_reduce: bad sizes");
  PullArgs a{};
  for (int i = 0; i < world; ++i) {
    MML_REQUIRE(peer_base_host[i] != nullptr, MML_ERR_INVALID_ARG, "symm_pull_reduce: null peer base %d", i);
    a.base[i] = static_cast<const char*>(peer_base_host[i]);
  }
  a.part_off = part_off; a.tail_off = tail_off; a.rows_total = rows_total; a.row_begin = row_begin; a.rows = rows;
  a.world = world; a.D = D; a.nt = nt; a.out1 = out1; a.out2 = out2; a.tail_out = tail_out;
  int64_t g = (2 * rows * D + 255) / 256;
  if (g > 148 * 4) g = 148 * 4;
  if (g < 1) g = 1;
  symm_pull_reduce_kernel<<<static_cast<unsigned>(g), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return check_launch("symm_pull_reduce_kernel");
}
