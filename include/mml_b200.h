/* mml_b200.h -- C ABI of the B200-native fusion + CRD distillation hot path.
 *
 * Drop-in boundary for the hot path of CityU-AIM-Group/MultiModal-learning
 * (citations are file:line under MICCAI-2022/ of the reference).  The reference
 * has no FFI: its boundary is Python `nn.Module`s.  The host-side mirror of
 * those modules lives in `multimodal-learning_b200/*.py` and calls ONLY the
 * entry points below (through ctypes).  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch types.  Pointers are DEVICE pointers
 *     unless the name ends in `_host`.  The caller owns every buffer; nothing is
 *     allocated inside (the caller passes a workspace sized by the matching
 *     `*_workspace_bytes` query).
 *   - `stream` is a `cudaStream_t` passed as `void*`; all work is enqueued on it
 *     and nothing synchronises the host.
 *   - return value: 0 = ok; <0 = error (MML_ERR_*); `mml_last_error()` returns a
 *     thread-local description.  There is NO CPU fallback anywhere.
 *   - floats are fp32, indices int64, row-major, as in the reference's tensors.
 */
#ifndef MML_B200_H_
#define MML_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MML_OK               0
#define MML_ERR_INVALID_ARG -1   /* bad shape / null pointer / misalignment        */
#define MML_ERR_CUDA        -2   /* a CUDA runtime / driver call failed            */
#define MML_ERR_UNSUPPORTED -3   /* shape outside what the kernels are built for   */
#define MML_ERR_WORKSPACE   -4   /* workspace smaller than *_workspace_bytes()     */

#define MML_ABI_VERSION 3   /* 2: seed_dev argument of the mml_kron_* entry points; multipos / relation / instance_sample added
                             * 3: mml_device_error_flags, mml_kron_linear_fwd_stats / mml_kron_fwd_stat_tiles / mml_bn_relu_fwd */

int         mml_abi_version(void);
const char* mml_last_error(void);
/* Number of kernels this library has launched in this process (bench `gpu_launches`). */
int64_t     mml_launch_count(void);

/* Sticky per-device error flags.  Kernels that index memory with caller-supplied ids never read or write out of bounds:
 * an id outside the bank is clamped to row 0 (gather / relation kernels) or dropped (routing kernels) and a bit is OR-ed
 * into the current device's flag word -- the reference's index_select raises a device-side assert in the same situation
 * (CRD_criterion.py:41,46).  mml_device_error_flags copies the word to *flags_host (this synchronises with the device)
 * and clears it when `reset` != 0. */
#define MML_DEVERR_CRD_INDEX   1u   /* contrast_idx / idx entry outside [0, n_rows) */
#define MML_DEVERR_SHARD_OWNER 2u   /* routed id outside [0, rows_per_rank * world) */
int         mml_device_error_flags(uint32_t* flags_host, int32_t reset);

/* ------------------------------------------------------------------------- *
 * CRD contrastive memory (CL_utils/CRD_criterion.py)
 *
 * Segment layout shared by the three gather entry points: anchor b owns the
 * index range idx[seg_begin(b) .. seg_begin(b)+seg_len(b)).
 *   dense  (seg_ptr == NULL): seg_begin = b*cols, seg_len = cols  -- the
 *          reference's idx[B, K+1] (CRD_criterion.py:41-49).
 *   ragged (seg_ptr != NULL): seg_begin = seg_ptr[b], seg_len = seg_ptr[b+1] -
 *          seg_ptr[b]; `cols` is then an upper bound on seg_len.  Used by the
 *          row-sharded bank: a rank sees only the indices it owns.
 * `idx` holds int64 (idx_bytes = 8, the reference's LongTensor) or int32 (idx_bytes = 4, local
 * row ids of a bank shard) entries.
 * `pos_flag` (uint8[B], NULL = all ones) says whether the FIRST entry of
 * anchor b's segment is its positive (column 0 of the reference's idx).
 * Bank 1 rows are scored against v2 ("side 2", out_v2, Z_v2); bank 2 rows
 * against v1 ("side 1", out_v1, Z_v1) -- CRD_criterion.py:41-49.
 * ------------------------------------------------------------------------- */

size_t mml_crd_workspace_bytes(int64_t B, int64_t cols, int32_t D);

/* Fused replacement for ContrastMemory.forward's gather/bmm/exp/div
 * (CRD_criterion.py:41-49,62-63) + ContrastLoss.forward (:199-216) + their
 * autograd: one pass over the gathered rows yields the NCE loss AND dL/dv1,
 * dL/dv2 (closed form).  Z must already be set.
 *   Z          float[2] = {Z_v1, Z_v2}                        (params[2:4])
 *   n_data     size of the noise distribution (ContrastLoss.n_data, :197)
 *   nce_k      m of ContrastLoss.forward (:201) = reference cols-1
 *   batch_norm divisor `bsz` of :214 (global batch when sharded)
 *   loss       float[1] or NULL: -(sum of log terms)/batch_norm
 *   sums       float[4] or NULL: {sum log-terms side1, side2, 0, 0} raw totals
 *   grad_v1/2  float[B,D]: dL/dv1, dL/dv2 for THESE segments (partial if sharded)
 *   out_v1/2   optional float[nnz] laid out like idx: exp(dot/T)/Z   (:62-63)   */
int mml_crd_fused_loss_grad(
    const float* bank1, const float* bank2, int64_t n_rows, int32_t D,
    const float* v1, const float* v2,
    const void* idx, int32_t idx_bytes, const int64_t* seg_ptr, const uint8_t* pos_flag,
    int64_t B, int64_t cols,
    float T, const float* Z, int64_t n_data, int64_t nce_k, int64_t batch_norm,
    float* loss, float* sums, float* grad_v1, float* grad_v2,
    float* out_v1, float* out_v2,
    void* workspace, size_t workspace_bytes, void* stream);

/* Multi-positive form of mml_crd_fused_loss_grad for the reference's selection variant (5-arg CRDLoss,
 * CL_utils/CRD_loss.py:153-175 with ContrastLoss_v2 :221-241, sample_KD == "False"): the first `n_pos` columns of
 * every dense idx[B, cols] row are positives, the remaining m = cols - n_pos are negatives;
 *   loss = -(1/B) * [ (1/n_pos) * sum_{b,p<n_pos} log(x/(x+c)) + sum_{b,k>=n_pos} log(m*Pn/(x+c)) ],  c = m*Pn + 1e-7
 * summed over both sides, plus dL/dv1, dL/dv2 in closed form.  n_pos = 1 is bit-identical to
 * mml_crd_fused_loss_grad with nce_k = cols - 1.  Workspace: mml_crd_workspace_bytes(B, cols, D).            */
int mml_crd_fused_loss_grad_multipos(
    const float* bank1, const float* bank2, int64_t n_rows, int32_t D,
    const float* v1, const float* v2,
    const void* idx, int32_t idx_bytes, int64_t B, int64_t cols, int64_t n_pos,
    float T, const float* Z, int64_t n_data,
    float* loss, float* grad_v1, float* grad_v2, float* out_v1, float* out_v2,
    void* workspace, size_t workspace_bytes, void* stream);

/* Relation gap of ContrastMemory_v3 (CL_utils/memory_new.py:288-292,303,342), the quantity its positive / negative
 * selection sorts by:  diff[b,k] = cos(bank1[idx[b,k]], v1[b]) - cos(bank2[idx[b,k]], v2[b])   for dense idx[B, cols].
 * One pass over the rows; nothing of size [B, cols, D] is written.                                           */
int mml_crd_relation_diff(
    const float* bank1, const float* bank2, int64_t n_rows, int32_t D,
    const float* v1, const float* v2,
    const void* idx, int32_t idx_bytes, int64_t B, int64_t cols,
    float* diff, void* stream);

/* Per-anchor ordering of the relation gaps (memory_new.py:303, :342-345): for every anchor b, sort the n columns
 * diff[b*ld + col0 .. + n) -- descending != 0: largest first -- and write the first m column numbers as
 * out[b*out_ld + j] = label0 + (column - col0).  Equal gaps are ordered by column (total, deterministic order).
 * One CTA per anchor, bitonic network in shared memory; n <= mml_crd_sort_columns_max() (16384). */
int32_t mml_crd_sort_columns_max(void);
int     mml_crd_sort_columns(const float* diff, int64_t B, int64_t ld, int64_t col0, int32_t n, int32_t descending,
                             int32_t m, int64_t label0, int64_t* out, int64_t out_ld, void* stream);

/* Full-bank KNN positives of the stage-2 criterion (MIA 2023/stage2_unimodal_student/CL_utils/CRD_criterion_v10.py:69-80,
 * :108-116): for every anchor b, the num_pos rows j of `bank` with the largest
 *     sim[b, j] = (row_labels[j] == anchor_labels[b]) ? cos(bank[anchor_rows[b]], bank[j]) : 0
 * (the reference multiplies sklearn's cosine_similarity by a 0/1 class mask, so rows of other classes score exactly 0),
 * in descending order: out_idx[B, P] (int64 row numbers), out_sim[B, P] (fp32).  Equal scores resolve to the smaller row.
 * One TF32 tcgen05 pass over the bank keeps 8 candidates per (anchor, bank slice, column half); every candidate is then
 * re-scored exactly in fp32, and anchors for which the TF32 error bound cannot prove the result are recomputed by an exact
 * scan (flags_out[B], optional, reports them; exact_only != 0 forces that scan for every anchor).
 * P <= mml_crd_knn_max_positives() (8).  The tcgen05 pass needs D % 32 == 0 and D <= 128; other D use the exact scan.
 * queries: NULL (the query of anchor b is bank row anchor_rows[b], as in the reference), or explicit query vectors [B, D]
 * (row-sharded bank: the anchor's own row may live on another rank; anchor_rows is then ignored).
 * inv_norms: NULL, or 1/|row| of every bank row kept by the caller (mml_crd_knn_inv_norms: all rows when rows == NULL, else
 * the listed ones -- e.g. the batch's rows after the momentum update), which saves the pass over the bank that computes them.
 * n_classes: when 1..3 every label must lie in [0, n_classes) and the class mask is tabulated per bank tile (faster);
 * 0 = labels are arbitrary int32 values, compared one by one. */
int32_t mml_crd_knn_max_positives(void);
int64_t mml_crd_knn_workspace_bytes(int64_t n_rows, int64_t B, int32_t D);
int     mml_crd_knn_inv_norms(const float* bank, int64_t n_rows, int32_t D, const int64_t* rows, int64_t count,
                              float* inv_norms, void* stream);
int     mml_crd_knn_positives(const float* bank, int64_t n_rows, int32_t D, const float* inv_norms,
                              const int32_t* row_labels, int32_t n_classes, const int64_t* anchor_rows, const float* queries, const int64_t* anchor_labels, int64_t B, int32_t P,
                              int32_t exact_only, int64_t* out_idx, float* out_sim, int32_t* flags_out,
                              void* workspace, size_t workspace_bytes, void* stream);

/* Per-class k-means centres of a bank (MIA 2023/stage2_unimodal_student/CL_utils/CRD_criterion_v10.py:84-92, :122-129: the
 * reference copies every class's rows to the host and runs sklearn KMeans(n_clusters = num_pos - 1) on them, every forward).
 * `rows` (device, int64) lists the bank rows class after class; class_offsets (HOST, int64 [n_classes + 1], class_offsets[0]
 * == 0, non-decreasing: a class may have no rows in this list when the bank is sharded) delimits the classes; centres (device, fp32 [n_classes, k, D]) holds the current centres and is updated in place.
 * The call enqueues `iterations` Lloyd iterations (sklearn's `_kmeans_single_lloyd`): every listed row goes to the centre of
 * its class that minimises |c|^2 - 2 x.c (lowest index on ties), each centre becomes the mean of its rows (an empty cluster
 * keeps its centre), and a class whose summed squared centre shift is <= tol[c] sets done[c] = 1 -- further iterations leave a
 * finished class untouched, so iterations may be enqueued in batches and `done` read between batches (tol / done may be NULL:
 * no stopping rule).  update == 0: assignment pass only -- centres stay, and the outputs below describe them.
 * Outputs of the last pass, each optional: inertia [n_classes, k] (sum of squared distances of a centre's rows to it, before
 * the update), counts [n_classes, k] (int64 rows per centre), row_dist [class_offsets[n_classes]] (squared distance of every
 * listed row to its nearest centre -- the D^2 weights of a k-means++ initialisation), sums [n_classes, k, D] (the rows of
 * every centre added up: with counts, what ranks that each hold a shard of the rows all-reduce before they divide).
 * Sums are accumulated in a fixed order: the result is bit-reproducible.  D in {32, 64, 128, 256, 512}, k <=
 * mml_crd_kmeans_max_clusters() (8), n_classes <= 32.  A row index outside [0, n_rows) raises MML_DEVERR_CRD_INDEX. */
int32_t mml_crd_kmeans_max_clusters(void);
int64_t mml_crd_kmeans_workspace_bytes(int32_t n_classes, int32_t k, int32_t D);
int     mml_crd_kmeans_lloyd(const float* bank, int64_t n_rows, int32_t D, const int64_t* rows, const int64_t* class_offsets,
                             int32_t n_classes, int32_t k, float* centres, const float* tol, int32_t iterations,
                             int32_t update, int32_t* done, float* inertia, int64_t* counts, float* sums, float* row_dist,
                             void* workspace, size_t workspace_bytes, void* stream);

/* Scores only (ContrastMemory.forward :41-49 [+ :62-63 when Z != NULL]).
 *   Z == NULL : out = exp(dot/T) (raw);  Z != NULL: out = exp(dot/T)/Z.
 *   sums      float[4] or NULL: {0, 0, sum raw side1, sum raw side2}
 *   set_Z     float[2] or NULL: first-call normaliser of :52-59 -- entries that
 *             are < 0 are replaced by mean(raw)*n_rows (mean over B*cols).
 *   out_v1/2  may be NULL (statistics-only pass).                              */
int mml_crd_scores(
    const float* bank1, const float* bank2, int64_t n_rows, int32_t D,
    const float* v1, const float* v2,
    const void* idx, int32_t idx_bytes, const int64_t* seg_ptr,
    int64_t B, int64_t cols,
    float T, const float* Z, float* sums, float* set_Z,
    float* out_v1, float* out_v2,
    void* workspace, size_t workspace_bytes, void* stream);

/* Weighted gather-sum: g1[b] = sum_k coef1[b,k]*bank2[idx[b,k]],
 * g2[b] = sum_k coef2[b,k]*bank1[idx[b,k]].  Backward of the bmm at
 * CRD_criterion.py:43,48 for callers that use out_v1/out_v2 in their own loss. */
int mml_crd_weighted_rows(
    const float* bank1, const float* bank2, int64_t n_rows, int32_t D,
    const void* idx, int32_t idx_bytes, const int64_t* seg_ptr,
    const float* coef1, const float* coef2,
    int64_t B, int64_t cols,
    float* g1, float* g2,
    void* workspace, size_t workspace_bytes, void* stream);

/* Momentum update + L2 renormalisation of the anchors' rows in both banks
 * (CRD_criterion.py:66-79):  r <- (m*r + (1-m)*v) / ||m*r + (1-m)*v||_2.
 * `y[i]` outside [row_begin, row_end) is skipped (owner-applies rule of the
 * sharded bank; pass 0, n_rows for the whole bank); the row written is
 * bank[y[i]-row_begin].  Duplicate y are order-dependent, as in the reference. */
int mml_crd_memory_update(
    float* bank1, float* bank2, int32_t D,
    const float* v1, const float* v2, const int64_t* y, int64_t B,
    float momentum, int64_t row_begin, int64_t row_end, void* stream);

/* Class-conditional contrast indices on the device: what `Pathomic_InstanceSample.__getitem__`
 * (data_loaders_MT.py:222-249) returns as `sample_idx = hstack(pos_idx, neg_idx)` for every anchor of the batch, written
 * by one kernel instead of numpy draws in loader workers + a [B, P+K] int64 upload per step.
 *   index      int64[B] anchors;  out int64[B, P+K]
 *   labels     int32[n] class per sample ('grad' task) or NULL ('surv' task: negatives = every other sample, :222-227)
 *   order      int32[n] sample ids sorted by class (stable), cls_ptr int32[num_classes+1]: cls_positive[c] =
 *              order[cls_ptr[c]:cls_ptr[c+1]] (:193-195); cls_negative[c] = order without that segment (:197-202)
 *   pos_mode   0 'exact' (pos = anchor), 1 'relax' (one member of the class), 2 'multi_pos' (P distinct members, first
 *              overwritten by the anchor, :229-241)
 *   negatives  K members of the pool, with replacement iff K > |pool| (:243), else distinct
 * Randomness: Philox4x32-10 keyed by (seed ^ *seed_dev, anchor row, column); distinct draws = keyed Feistel bijection
 * with cycle walking.  Restated bit-for-bit in oracle/sampler_oracle.py.  Pool members must be >= 1 where drawn from. */
int mml_instance_sample(const int64_t* index, int64_t B, const int32_t* labels, const int32_t* order,
                        const int32_t* cls_ptr, int32_t num_classes, int64_t n, int32_t P, int32_t K,
                        int32_t pos_mode, uint64_t seed, const uint64_t* seed_dev, int64_t* out, void* stream);

/* Normalize(2) of the Embed heads (CRD_criterion.py:242-245): y = x / ||x||_2 per row (no epsilon, as the
 * reference), norm[b] kept for the backward  gx = (gy - y * <gy, y>) / norm.  One launch each instead of the
 * reference's pow/sum/pow/div chain and its ~8-kernel autograd.                                          */
int mml_l2norm_fwd(const float* x, int64_t B, int32_t D, float* y, float* norm, void* stream);
int mml_l2norm_bwd(const float* gy, const float* y, const float* norm, int64_t B, int32_t D, float* gx,
                   void* stream);

/* Index routing for the row-sharded bank (rank o owns rows [o*rows_per_rank, (o+1)*rows_per_rank)):
 * a stable counting sort of idx[B, cols] by owner, kept in (anchor, column) order inside each owner.
 * Columns are processed in chunks of `chunk_cols` (multiple of 32), chunks = ceil(cols/chunk_cols):
 *   mml_shard_count:   counts[(o*B + b)*chunks + c] = #{k in chunk c : idx[b,k] / rows_per_rank == o}
 *   mml_shard_scatter: out[offsets[(o*B + b)*chunks + c] + j] = (int32) local id of the j-th such
 *                      column, `offsets` = the caller's exclusive scan of `counts` in that order.
 * world <= 32; every idx value must lie in [0, world*rows_per_rank).                         */
int mml_shard_count(const int64_t* idx, int64_t B, int64_t cols, int32_t chunk_cols, int64_t rows_per_rank,
                    int32_t world, int64_t* counts, void* stream);
int mml_shard_scatter(const int64_t* idx, int64_t B, int64_t cols, int32_t chunk_cols, int64_t rows_per_rank,
                      int32_t world, const int64_t* offsets, int32_t* out_local_ids, void* stream);

/* Single-pass, fixed-stride routing for the peer-memory (NVLink pull) exchange: slot (o, b, c) owns
 * ids_out[((o*B + b)*chunks + c)*chunk_cols ..] and uses the first counts[(o*B + b)*chunks + c] (int32)
 * entries.  ids_out: int32[world*B*chunks*chunk_cols] in peer-mapped (symmetric) memory.         */
int mml_shard_route_strided(const int64_t* idx, int64_t B, int64_t cols, int32_t chunk_cols,
                            int64_t rows_per_rank, int32_t world, int32_t* counts, int32_t* ids_out,
                            void* stream);

/* K4 over PEER-resident index slots: the owner's gather kernel reads, for global anchor
 * a = s*B_local + bl and slot c, the ids  peer_ids_host[s][(bl*route_chunks + c)*route_stride ..]
 * (a pointer into rank s's ids_out, already offset to THIS owner's block, valid in this process
 * through CUDA peer mapping) of length peer_counts_host[s][bl*route_chunks + c] (rank s's counts
 * block for this owner: a local copy after an all_to_all, or the peer-mapped original).  The
 * all_to_all of routed indices is thereby fused into the gather kernel as NVLink loads (4 bytes of
 * index per 1 KB of local row traffic).  `peer_ids_host` / `peer_counts_host` are HOST arrays of
 * `world` device pointers; sums as in the non-peer calls.                                       */
size_t mml_crd_peer_workspace_bytes(int64_t B_global, int32_t route_chunks, int32_t D);
int mml_crd_fused_loss_grad_peer(
    const float* bank1, const float* bank2, int64_t n_rows, int32_t D, const float* v1, const float* v2,
    const int32_t* const* peer_ids_host, const int32_t* const* peer_counts_host, int32_t world, int64_t B_local,
    int32_t route_chunks, int32_t route_stride, const uint8_t* pos_flag,
    float T, const float* Z, int64_t n_data, int64_t nce_k, int64_t batch_norm,
    float* sums, float* grad_v1, float* grad_v2,
    void* workspace, size_t workspace_bytes, void* stream);
int mml_crd_scores_peer(
    const float* bank1, const float* bank2, int64_t n_rows, int32_t D, const float* v1, const float* v2,
    const int32_t* const* peer_ids_host, const int32_t* const* peer_counts_host, int32_t world, int64_t B_local,
    int32_t route_chunks, int32_t route_stride, float T, float* sums,
    void* workspace, size_t workspace_bytes, void* stream);

/* Small collectives over peer-mapped (symmetric) memory, for the sharded step's KB..MB messages
 * where NCCL's launch + protocol latency dominates.  `peer_base_host`: HOST array of `world` base
 * pointers of the same symmetric allocation on every rank (index = rank).  The caller separates
 * the calls with cross-rank barriers (torch symmetric-memory `barrier()`).
 *   mml_symm_push:        copy up to 4 local segments (src[i], bytes[i], 16-byte multiples) to byte
 *                         offset dst_off[i] of EVERY rank's allocation (an all_gather by NVLink stores).
 *   mml_symm_pull_reduce: out[r, :] = sum over ranks p (in rank order: deterministic) of the float
 *                         rows [row_begin + r] of the two [rows_total, D] planes at byte offset
 *                         `part_off` of rank p's allocation (a reduce_scatter by NVLink loads), for
 *                         r < rows; tail[k] = sum_p of the nt floats at byte offset `tail_off`.     */
int mml_symm_push(void* const* peer_base_host, int32_t world, const void* const* src, const int64_t* dst_off,
                  const int64_t* bytes, int32_t nseg, void* stream);
int mml_symm_pull_reduce(void* const* peer_base_host, int32_t world, int64_t part_off, int64_t rows_total,
                         int64_t row_begin, int64_t rows, int32_t D, float* out1, float* out2,
                         int64_t tail_off, int32_t nt, float* tail_out, void* stream);

/* ------------------------------------------------------------------------- *
 * AliasMethod (CRD_criterion.py:84-141)
 * ------------------------------------------------------------------------- */

/* HOST function.  Vose tables with the reference's LIFO pairing (:97-123),
 * bit-exact in fp32.  `probs_host` must already be normalised as :90-91 does. */
int mml_alias_build_host(const float* probs_host, int64_t n,
                         float* prob_out_host, int64_t* alias_out_host);

/* draw(), split around the two torch RNG calls it keeps (:133 random_, :137
 * bernoulli) so the same generator state yields the same indices:
 *   mml_alias_gather_prob: p[i] = prob[kk[i]]                           (:134)
 *   mml_alias_select:      out[i] = b[i] ? kk[i] : alias[kk[i]]         (:135-141)
 *     and, when y != NULL, out[r*cols + 0] = y[r] (ContrastMemory.forward :39). */
int mml_alias_gather_prob(const float* prob, const int64_t* kk, int64_t N,
                          float* p_out, void* stream);
int mml_alias_select(const int64_t* alias, const int64_t* kk, const float* b, int64_t N,
                     const int64_t* y, int64_t cols, int64_t* out, void* stream);

/* ------------------------------------------------------------------------- *
 * Gated Kronecker fusion encoder (fusion.py)
 *
 * Replaces, for BilinearFusion (fusion.py:56-60) and TrilinearFusion_A/B
 * (fusion.py:123-129, :192-198), the chain
 *   cat-with-1 -> bmm outer product(s) -> flatten -> post_fusion_dropout -> encoder1[0] (nn.Linear)
 * by   y[b,n] = sum_k A[b,k] m[b,k] W[n,k] + bias[n]
 * where A[b, (i*(d2+1)+j)*(d3+1)+l] = g1[b,i] g2[b,j] g3[b,l], g(x) = f[b,x] for x < d and 1 for
 * x == d, is generated on chip and NEVER written to HBM, forward or backward.
 *   f1,f2,f3   float[B,d1|d2|d3] factors WITHOUT the appended 1 (f3 = NULL, d3 = 0: bilinear)
 *   W          float[N, Kk] = encoder1[0].weight, Kk = (d1+1)(d2+1)(d3+1 | 1), row-major
 *   m          dropout multiplier of `post_fusion_dropout`: 0 or 1/(1-p'), a pure function of
 *              (seed, b, k) (counter-based hash; p' = round(p*65536)/65536); identity when
 *              training == 0 or drop_p == 0.  The effective seed is `seed ^ *seed_dev` when `seed_dev`
 *              (a DEVICE uint64, may be NULL) is given: a launch captured in a CUDA graph freezes `seed`
 *              but re-reads `*seed_dev` on every replay, so each replay can draw a fresh mask.  Forward,
 *              dgrad and wgrad of one step must see the same pair.
 * ------------------------------------------------------------------------- */

/* Tensor-core path (tcgen05 kind::tf32, A generated into TMEM, W streamed by TMA).
 * The weight must first be repacked into chunk order:
 *   n = mml_kron_num_chunks(d1,d2,d3);  table: int32[n*8] built on the HOST by
 *   mml_kron_chunk_table_host and copied to the device by the caller (16-byte aligned);
 *   Wp: float[mml_kron_packed_floats(N,d1,d2,d3)] (128-byte aligned), filled by mml_kron_pack_weight
 *   whenever W changes.  mml_kron_linear_fwd needs N <= 256.                               */
int64_t mml_kron_num_chunks(int32_t d1, int32_t d2, int32_t d3);
int     mml_kron_chunk_table_host(int32_t d1, int32_t d2, int32_t d3, int32_t* table_host);
int64_t mml_kron_packed_floats(int32_t N, int32_t d1, int32_t d2, int32_t d3);
int     mml_kron_pack_weight(const float* W, int32_t N, int32_t d1, int32_t d2, int32_t d3,
                             const int32_t* table, float* Wp, void* stream);
int     mml_kron_fwd_supported(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3);   /* 1 / 0 */
size_t  mml_kron_fwd_workspace_bytes(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3);
int     mml_kron_linear_fwd(const float* f1, const float* f2, const float* f3, int64_t B,
                            int32_t d1, int32_t d2, int32_t d3, const int32_t* table, const float* Wp,
                            const float* bias, int32_t N, float drop_p, uint64_t seed, const uint64_t* seed_dev, int32_t training,
                            float* y, void* workspace, size_t workspace_bytes, void* stream);

/* encoder1 = Linear -> BatchNorm1d -> ReLU (fusion.py:29,60): the forward can emit the per-column sums of y and y^2 of
 * every 128-row tile from its epilogue (col_stats float[tiles][2][N], tiles = mml_kron_fwd_stat_tiles(...) > 0, i.e. the
 * forward is not split over K), and mml_bn_relu_fwd finishes: batch statistics (from those partials when n_part > 0,
 * else from y itself), the nn.BatchNorm1d running-statistics update (running_* may be NULL; momentum, eps as the module's;
 * biased variance for normalisation, unbiased for running_var), then out = relu((y - mean) * invstd * gamma + beta).
 * save_mean / save_invstd [N] are what the BatchNorm backward needs. */
int64_t mml_kron_fwd_stat_tiles(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3);
int     mml_kron_linear_fwd_stats(const float* f1, const float* f2, const float* f3, int64_t B,
                                  int32_t d1, int32_t d2, int32_t d3, const int32_t* table, const float* Wp,
                                  const float* bias, int32_t N, float drop_p, uint64_t seed, const uint64_t* seed_dev,
                                  int32_t training, float* y, float* col_stats, void* workspace, size_t workspace_bytes,
                                  void* stream);
int     mml_bn_relu_fwd(const float* y, int64_t B, int32_t N, const float* col_stats, int32_t n_part,
                        const float* gamma, const float* beta, float* running_mean, float* running_var,
                        float momentum, float eps, float* out, float* save_mean, float* save_invstd, void* stream);

/* Weight gradient on the tensor cores (same K permutation / chunk table as the forward):
 *   dW[n,k] = sum_b dy[b,n] A[b,k] m[b,k]   -> dense float[N, Kk] (every entry written).
 * Lanes = 128 packed k, contraction over the batch, A^T generated into TMEM, dy^T streamed by TMA.
 * workspace (1024-byte aligned) holds dy^T and the per-batch-split partial tiles.              */
int     mml_kron_wgrad_supported(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3);
size_t  mml_kron_wgrad_workspace_bytes(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3);
int     mml_kron_linear_wgrad(const float* f1, const float* f2, const float* f3, int64_t B,
                              int32_t d1, int32_t d2, int32_t d3, const int32_t* table, const float* dy,
                              int32_t N, float drop_p, uint64_t seed, const uint64_t* seed_dev, int32_t training, float* dW,
                              void* workspace, size_t workspace_bytes, void* stream);

/* Factor gradients on the tensor cores.  dA = m * (dy W) is accumulated tile by tile in TMEM
 * (M = 128 batch rows, N = 128 packed k, K = n) and folded into df1/df2/df3 by the epilogue; it is
 * never stored.  Needs the TRANSPOSED packed weight WpT [32*chunks, roundup(N,32)] (128-byte aligned,
 * mml_kron_packed_t_floats floats, refreshed by mml_kron_pack_weight_t whenever W changes).
 * The appended 1 of each factor receives no gradient.                                          */
int     mml_kron_dgrad_supported(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3);
int64_t mml_kron_packed_t_floats(int32_t N, int32_t d1, int32_t d2, int32_t d3);
int     mml_kron_pack_weight_t(const float* W, int32_t N, int32_t d1, int32_t d2, int32_t d3,
                               const int32_t* table, float* WpT, void* stream);
size_t  mml_kron_dgrad_workspace_bytes(int64_t B, int32_t N, int32_t d1, int32_t d2, int32_t d3);
int     mml_kron_linear_dgrad(const float* f1, const float* f2, const float* f3, int64_t B,
                              int32_t d1, int32_t d2, int32_t d3, const int32_t* table, const float* WpT,
                              const float* dy, int32_t N, float drop_p, uint64_t seed, const uint64_t* seed_dev, int32_t training,
                              float* df1, float* df2, float* df3,
                              void* workspace, size_t workspace_bytes, void* stream);

/* Exact-fp32 CUDA-core path on the dense weight (any N, any widths): forward, and backward
 *   dW[n,k] = sum_b dy[b,n] A[b,k] m[b,k]            (NULL to skip)
 *   df_x    = factor gradients through dA = m * (dy W), contracted on chip (all NULL to skip);
 *             the appended 1 receives no gradient.                                         */
int mml_kron_linear_fwd_simt(const float* f1, const float* f2, const float* f3, int64_t B,
                             int32_t d1, int32_t d2, int32_t d3, const float* W, const float* bias,
                             int32_t N, float drop_p, uint64_t seed, const uint64_t* seed_dev, int32_t training, float* y,
                             void* stream);
int mml_kron_linear_bwd_simt(const float* f1, const float* f2, const float* f3, int64_t B,
                             int32_t d1, int32_t d2, int32_t d3, const float* W, const float* dy,
                             int32_t N, float drop_p, uint64_t seed, const uint64_t* seed_dev, int32_t training,
                             float* df1, float* df2, float* df3, float* dW, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MML_B200_H_ */
