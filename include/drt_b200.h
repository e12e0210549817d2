/*
 * drt_b200.h — C ABI of the B200-native exact-MIPS hot path for DenseRetrievalToolkits.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes / a CUDA stream handle passed
 * as `void*` (a `cudaStream_t`; NULL = the legacy default stream), returns an `int` status
 * (0 = ok, < 0 = DRT_E_*), and never lets a C++ exception cross the boundary.  The thread-local
 * message of the last failure is returned by drt_last_error().  There is NO CPU fallback: on a
 * machine without an sm_100 device every compute call fails with DRT_E_NO_DEVICE.
 *
 * Each function cites the interface of the reference it stands in for.  Citations are
 * `path:line` into yhao-wang/DenseRetrievalToolkits (the reference is pure Python; the
 * arithmetic it delegates to faiss / torch is what these functions replace).
 */
#ifndef DRT_B200_H_
#define DRT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRT_B200_ABI_VERSION 2

/* status codes */
#define DRT_OK                0
#define DRT_E_INVALID        -1   /* bad argument (shape, k, dim, null pointer)            */
#define DRT_E_CUDA           -2   /* a CUDA runtime / driver call failed                   */
#define DRT_E_NO_DEVICE      -3   /* no CUDA device, or the device is not sm_100 (B200)     */
#define DRT_E_OOM            -4   /* device allocation failed                              */
#define DRT_E_UNSUPPORTED    -5   /* e.g. dim > 8192, k > DRT_MAX_K, non-"Flat" factory      */
#define DRT_E_INTERNAL       -6   /* kernel-side watchdog / invariant violation             */

#define DRT_MAX_K          2048   /* same ceiling faiss-gpu uses for k-selection            */

/* drt_search flags */
#define DRT_SEARCH_DEFAULT        0u
#define DRT_SEARCH_NO_RESCORE     1u  /* debug: return the bf16 first-pass scores/order       */
#define DRT_SEARCH_FORCE_1CTA     2u  /* use the 1-CTA (M=128) tile variant of the MMA kernel */
#define DRT_SEARCH_FORCE_2CTA     4u  /* use the CTA-pair (M=256, cta_group::2) variant       */
#define DRT_SEARCH_TIME_KERNELS    8u  /* bracket every MMA-filter launch with CUDA events     */

/* opaque: one device-resident corpus shard.  A store owns its search workspace, so add / search /
 * reset / reconstruct on ONE store are serialised by an internal lock (host threads may share a
 * store; different stores run concurrently). */
typedef struct drt_store drt_store;

/* ---- library ------------------------------------------------------------------------------ */

int         drt_abi_version(void);
const char* drt_last_error(void);
/* Number of CUDA devices that are sm_100; < 0 on error.  Does not create a context. */
int         drt_device_count(void);

/* ---- corpus-embedding store (one shard) ----------------------------------------------------
 * Replaces faiss.IndexFlatIP as used by BaseFaissIPRetriever (DRT/evaluator/index.py:16-28) and
 * the .npy/.json + faiss.write_index/read_index round trip of Trainer._encoding_corpus /
 * _index_corpus / _load_index (DRT/trainer/trainer.py:191-262).                               */

/* faiss.IndexFlatIP(d) (index.py:19,23).  `seg_rows` = rows per device segment (0 = default
 * 1<<20); rows live in fixed-size segments so `add` never reallocates or moves data.
 * Any 1 <= dim <= 8192 (the reference's `projection_out_dim`, arguments.py:46, is free): rows
 * are stored zero-padded to the next multiple of 64 elements, which changes no inner product;
 * every pointer the caller passes or receives uses its own `dim` as the row pitch. */
int drt_store_create(drt_store** out, int dim, int device, int64_t seg_rows);
int drt_store_destroy(drt_store* s);

/* index.add(x) (index.py:28; trainer.py:235): append n fp32 rows, ids = insertion order.
 * `rows` is a host pointer (rows_on_device = 0) or a device pointer on the store's device
 * (rows_on_device = 1, the zero-copy path for encoder outputs, trainer.py:204).
 * Stream-ordered on `stream`; with a host source the call returns after the copy completed. */
int drt_store_add(drt_store* s, const float* rows, int64_t n, int rows_on_device, void* stream);

/* Declares that the LAST `tail_dims` dims of every row and query are exactly representable in
 * bf16 (e.g. the norm-augmentation columns of the squared-L2 form of faiss.index_factory(d,
 * "Flat"), index.py:50): the exactness certificate then bounds the rounding error of the head
 * and tail dims separately, which keeps it tight when the tail carries most of a row's norm.
 * Purely an accuracy hint for the bound — results are exact either way (rows whose tail is not
 * exact are still covered: the residual norm is always taken over all dims).  Call before the
 * first add. */
int drt_store_set_exact_tail(drt_store* s, int tail_dims);

int64_t drt_store_ntotal(const drt_store* s);   /* index.ntotal */
int     drt_store_dim(const drt_store* s);      /* index.d      */
int     drt_store_device(const drt_store* s);
int     drt_store_reset(drt_store* s);          /* index.reset(): drop rows, keep segments  */

/* index.reconstruct_n(row0, n): copy fp32 rows back (host or device destination).  Used by
 * the faiss.write_index replacement (trainer.py:245). */
int drt_store_reconstruct(const drt_store* s, int64_t row0, int64_t n, float* out,
                          int out_on_device, void* stream);

/* ---- search --------------------------------------------------------------------------------
 * index.search(x, k) -> (D float32[nq,k], I int64[nq,k]) (index.py:32): exact inner product of
 * every (query,row) pair, per-query top-k sorted by (score desc, id asc); when k > ntotal the
 * tail is filled with score -FLT_MAX and id -1; NaN scores never enter a result.
 * ids are `id_offset + row` (id_offset = this shard's first global row id).
 * io_on_device = 0: q / out_* are host pointers, H2D + D2H copies happen inside the call and it
 *                   returns when the results are in host memory (the drop-in faiss contract);
 * io_on_device = 1: all three are device pointers; the call is stream-ordered and asynchronous
 *                   except for a few bounded host syncs between corpus chunks.                 */
int drt_search(drt_store* s, const float* q, int64_t nq, int k,
               float* out_scores, int64_t* out_ids,
               int io_on_device, int64_t id_offset, uint32_t flags, void* stream);

/* The same search with NO host round trip: device pointers only, fully stream-ordered, the first
 * pass only.  Instead of retrying / refining on the host, the call publishes ONE device byte:
 * *status_out = 1 when the result is not final — a candidate buffer overflowed, or the exactness
 * certificate flagged a query — and the caller must then repeat the search with drt_search.
 * Made for the multi-GPU path (store.py): the shard search, the exchange + merge kernel and the
 * barriers are enqueued back to back, the merge kernel ORs all ranks' status bytes
 * (drt_merge_topk_peers2), and the host reads one word at the end of the step.
 * 1 <= nq <= 16384, dim a multiple of 64, q 16-byte aligned, non-empty store. */
int drt_search_async(drt_store* s, const float* q, int64_t nq, int k,
                     float* out_scores, int64_t* out_ids, int64_t id_offset, uint32_t flags,
                     uint8_t* status_out, void* stream);

/* Counters of the last drt_search on this store (for tests and bench.py):
 *  [0] kernel launches  [1] mma-filter launches  [2] candidate-buffer overflow retries
 *  [3] first-pass candidates per query (k')      [4] queries whose exactness check flagged
 *  [5] ctas per tile (1|2)                        [6] corpus chunks
 *  [7] summed device time of the MMA-filter launches in ns (DRT_SEARCH_TIME_KERNELS only)
 *  [8] queries that needed the exact fp32 first pass (last-resort refinement)
 *  [9] queries whose certificate failed after the FIRST pass (searched again with a larger k';
 *      [4] counts what is still uncertified after the whole ladder)
 *  [10] candidate rows whose fp32 row was actually read by the rescoring kernel   [11] reserved */
/* After drt_search_async this call waits for that search to finish (its counters are collected
 * lazily, so that the search itself needs no host round trip). */
int drt_search_stats(const drt_store* s, int64_t out[12]);

/* Host-side planning, exposed for tests (no device needed).  drt_plan_params: k' (first-pass
 * candidates kept per query) and the candidate-buffer capacity for retry level `attempt`.
 * drt_plan_chunks: the corpus chunk schedule of one search as triples (segment, row0, row1)
 * relative to the segment, written to out[3*i..]; returns the number of chunks (or < 0). */
int drt_plan_params(int k, int attempt, int* kprime, int* cap);
int drt_plan_chunks(int64_t ntotal, int64_t seg_rows, int k, int attempt, int64_t* out, int max_chunks);

/* ---- cross-shard merge ---------------------------------------------------------------------
 * merge_retrieval_results_by_score (DRT/model/utils.py:215-229): union of G per-shard result
 * lists per query (first occurrence of an id wins), sort by score desc (ties: id asc), keep
 * k_out.  Inputs are device arrays laid out [G][nq][k_in]; entries with id < 0 are padding.
 * flags = DRT_MERGE_SORTED_UNIQUE promises that every list is already ordered (score desc,
 * id asc, padding last) and that no id occurs in two lists (row-sharded stores): the merge then
 * ranks entries by binary search instead of sorting. */
#define DRT_MERGE_DEFAULT        0u
#define DRT_MERGE_SORTED_UNIQUE  1u
int drt_merge_topk(int n_lists, const float* scores, const int64_t* ids, int64_t nq, int k_in,
                   int k_out, float* out_scores, int64_t* out_ids, uint32_t flags, int device,
                   void* stream);

/* The candidate exchange and the merge as ONE kernel over peer-mapped memory (row-sharded store,
 * one rank per GPU on NVLink / NVSwitch): list g is read from rank g's buffer, laid out
 * [nq][k_in] (ordered, id-disjoint lists as for DRT_MERGE_SORTED_UNIQUE); the merged rows of
 * queries [q_begin, q_begin + q_count) and a per-query flag "some full list's last score is not
 * strictly below the merged k_out-th" (the exactness check of a reduced per-shard depth) are
 * stored into EVERY rank's out_scores[g] / out_ids[g] ([nq][k_out]) / truncated[g] ([nq]).
 * out_scores[g] / out_ids[g] may both be NULL: rank g then receives only the flags (rank-local
 * results: every rank keeps just the rows of the queries it merged, the reference's per-rank
 * evaluation, DRT/trainer/trainer.py:287-297).
 * The pointer tables are host arrays of n_lists (<= 16) device pointers valid on `device`.
 * The caller provides the cross-rank barriers before (lists complete) and after (results
 * complete).  Replaces all-gather + drt_merge_topk of store.py's NCCL path. */
int drt_merge_topk_peers(int n_lists, const float* const* scores, const int64_t* const* ids,
                         int64_t q_begin, int64_t q_count, int k_in, int k_out,
                         float* const* out_scores, int64_t* const* out_ids,
                         uint8_t* const* truncated, int device, void* stream);

/* ---- in-batch-negative loss ----------------------------------------------------------------
 * SimpleContrastiveLoss.forward (DRT/trainer/losses.py:11-17) and the loss block of
 * DRModel.forward (DRT/model/biencoder.py:107-116): logits = x·yᵀ (fp32), cross entropy against
 * target[i].  All pointers are device pointers; x is [B,d], y is [P,d], row-major fp32.
 *   target      int64[B] or NULL (NULL => target[i] = i * (P / B), losses.py:12-15)
 *   logits_out  float[B,P] or NULL (DROutput.scores, biencoder.py:122; NULL = never stored)
 *   lse_out     float[B]  (saved for backward)
 *   loss_rows   float[B]  per-row loss (reduction='none')
 *   loss_out    float[1]  sum_i loss_rows[i] * loss_scale  (mean: loss_scale = 1/B)          */
int drt_inbatch_ce_fwd(const float* x, const float* y, int64_t B, int64_t P, int dim,
                       const int64_t* target, float loss_scale,
                       float* logits_out, float* lse_out, float* loss_rows, float* loss_out,
                       int device, void* stream);

/* Backward of the above: dlogits[i,j] = g_i * (softmax(logits)[i,j] - [j == target[i]]) with
 * g_i = grad_scale * grad_rows[i * grad_stride]: grad_stride = 1 for a per-row upstream gradient
 * (reduction='none'), 0 when grad_rows points at ONE device float (reduction='mean'/'sum': the
 * scalar upstream gradient, grad_scale = 1/B or 1).  dx = dlogits·y, dy = dlogitsᵀ·x.  `work` is
 * a device scratch of B*P floats (holds dlogits; see drt_inbatch_ce_bwd_needs_work).  `logits` =
 * the forward's logits_out if it was kept (dlogits is then one elementwise pass), or NULL to
 * recompute the logits tile by tile.  Neither `logits` nor any other input is modified, so the
 * backward may run any number of times over the same saved tensors (retain_graph).  dx / dy may
 * be NULL to skip that gradient. */
/* 1 if drt_inbatch_ce_bwd needs the `work` scratch for this shape, 0 if it may be NULL (the
 * small-shape tensor-core path forms dlogits directly as split bf16 operands). Host-only. */
int drt_inbatch_ce_bwd_needs_work(int64_t B, int64_t P, int dim, int have_logits);

int drt_inbatch_ce_bwd(const float* x, const float* y, int64_t B, int64_t P, int dim,
                       const int64_t* target, const float* lse, const float* logits,
                       const float* grad_rows, int grad_stride, float grad_scale, float* work,
                       float* dx, float* dy, int device, void* stream);

/* ---- mining filter -------------------------------------------------------------------------
 * process_sample (DRT/trainer/sampler.py:69-80): walk each query's retrieved ids in rank order,
 * skip ids inside the query's own positive range [pos_begin[i], pos_end[i]), keep the first
 * `num_negative`; unfilled slots get -1.  Device pointers. */
int drt_filter_negatives(const int64_t* ids, int64_t nq, int k, const int64_t* pos_begin,
                         const int64_t* pos_end, int num_negative, int64_t* out_ids,
                         int device, void* stream);

/* drt_merge_topk_peers plus the status exchange of drt_search_async: `status` = one device byte
 * per rank (peer-mapped), `redo` = this rank's byte that receives their OR (same value on every
 * rank, since every rank reads the same bytes between the two barriers). */
int drt_merge_topk_peers2(int n_lists, const float* const* scores, const int64_t* const* ids,
                          int64_t q_begin, int64_t q_count, int k_in, int k_out,
                          float* const* out_scores, int64_t* const* out_ids,
                          uint8_t* const* truncated, const uint8_t* const* status, uint8_t* redo,
                          int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRT_B200_H_ */
