/*
 * b200vs.h -- C-ABI of the B200-native exact vector-search engine (libb200vs.so).
 *
 * The reference (Theseus-AT/mlx-vector-db) has NO FFI for this path: its boundary is the
 * Python class `MLXVectorStore` (service/optimized_vector_store.py:59-246) and the module
 * functions of performance/mlx_optimized.py:26-287, whose arithmetic is delegated to
 * `mlx.core` (mx.matmul / mx.argsort / mx.linalg.norm / mx.concatenate).  Each entry point
 * below replaces one of those delegation sites; the Python layer in
 * `mlx-vector-db_b200/b200vs/` binds them with ctypes and re-creates the reference's
 * class / function surface on top.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - every function returns an int status (VS_OK == 0) unless stated; on failure
 *     `vs_last_error()` returns a thread-local message;
 *   - "device pointer" = CUDA device memory on the store's device; `stream` is a
 *     `cudaStream_t` passed as void* (NULL = the legacy default stream);
 *   - ids are int32 row numbers in insertion order (the reference surfaces uint32 from
 *     mx.argsort as Python ints); unused output slots hold id -1 and score 0;
 *   - result order: cosine / dot_product descending score, euclidean ascending distance,
 *     equal scores -> lower id first (stable argsort of the negated scores,
 *     service/optimized_vector_store.py:176-181);
 *   - there is no CPU fallback: without a usable CUDA device every compute entry point
 *     fails with VS_ERR_CUDA.
 */
#ifndef B200VS_H
#define B200VS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VS_API __attribute__((visibility("default")))

/* status codes -> Python exceptions (see b200vs/_cabi.py) */
#define VS_OK            0
#define VS_ERR_INVALID   1   /* bad argument            -> ValueError   */
#define VS_ERR_CUDA      2   /* CUDA runtime failure    -> RuntimeError */
#define VS_ERR_OOM       3   /* device memory exhausted -> MemoryError  */
#define VS_ERR_STATE     4   /* e.g. metric unsupported -> RuntimeError */

/* metric: service/optimized_vector_store.py:211-213, service/models.py:23-27 */
#define VS_METRIC_COSINE     0
#define VS_METRIC_EUCLIDEAN  1
#define VS_METRIC_DOT        2

/* shadow copy of the database scanned by the low-precision candidate kernels */
#define VS_SHADOW_NONE  0
#define VS_SHADOW_BF16  1   /* 16-bit copy: fp16 of x/||x|| for cosine, bf16 of x otherwise            */
#define VS_SHADOW_FP8   2   /* additional e4m3 copy (cosine only) for VS_SEARCH_GEMM_FP8; may be OR-ed */

/* search flags */
#define VS_SEARCH_AUTO        0   /* K3 (certified exact) when the store has the 16-bit shadow,
                                     k <= 128, >= 65536 rows and a row mask (if any) keeps at
                                     least a fifth of them -- euclidean: for two or more
                                     queries --, else the K2 fp32 scan                        */
#define VS_SEARCH_SCAN_FP32   1   /* K2: fp32 streaming scan (exact)                        */
#define VS_SEARCH_SCAN_BF16   2   /* K2 over the 16-bit shadow + K5 fp32 rescoring (recall) */
#define VS_SEARCH_GEMM        3   /* K3: tcgen05 16-bit GEMM candidates + K5 rescoring,
                                     certified exact; uncertified queries are retried with 4x
                                     the candidates, then re-run through the exact scan      */
#define VS_SEARCH_GEMM_NOCERT 4   /* K3 + K5 without certification (recall reported)        */
#define VS_SEARCH_GEMM_FP8    5   /* K3 over the e4m3 shadow (kind::f8f6f4) + K5 rescoring; no
                                     certification: recall-reported variant                  */
#define VS_SEARCH_MODE_MASK   0xff
#define VS_SEARCH_TMA         0x100  /* K2 variant: cp.async.bulk (TMA) staged tiles        */
#define VS_SEARCH_LDG         0x200  /* K2 variant: direct 128-bit global loads             */

typedef struct vs_store vs_store;

/* Library / device ---------------------------------------------------------------- */

/* Thread-local message for the last failing call on this thread. Never NULL. */
VS_API const char* vs_last_error(void);
/* Library version string. */
VS_API const char* vs_version(void);
/* Number of kernels this library has launched since load (for bench.py's gpu_launches). */
VS_API int64_t vs_launch_count(void);

/* Bracket the dominant kernels (kind 0 = K2 scan, 1 = K3 GEMM) with CUDA events on their
 * launch stream while enabled; vs_profile_read waits for them, returns the summed kernel time
 * and launch count since the last read, and clears the record.  Measurement aid for bench.py. */
VS_API int vs_profile(int enable);
VS_API int vs_profile_read(int kind, double* total_ms, int64_t* launches);

/* Store lifetime -- replaces MLXVectorStore.__init__/_create_empty_store
 * (service/optimized_vector_store.py:60-94): one flat (N, D) fp32 array, here resident in HBM
 * in a growable virtual-memory arena with pre-computed row norms.
 *   max_rows: upper bound used to reserve address space (0 = derive from device memory). */
VS_API int vs_create(int device, int dim, int metric, int shadow, int64_t max_rows,
                     vs_store** out);
VS_API int vs_destroy(vs_store* s);

/* K1 append_norm -- replaces `mx.concatenate([self._vectors, new])` + `mx.eval`
 * (service/optimized_vector_store.py:98-106) and the per-query re-normalisation of the whole
 * database (:34-38).  Copies `m` rows of `dim` fp32 (row-major, contiguous) to the end of the
 * store, computes max(||x||, 1e-8) and ||x||^2 per row and the optional shadow copies.
 *   rows_on_device: 0 = host pointer (copied by the library), 1 = device pointer.
 * Rows become visible to searches enqueued after this call returns. */
VS_API int vs_append(vs_store* s, const float* rows, int64_t m, int rows_on_device,
                     void* stream);

/* Same for one shard of a row-sharded store: the m rows get the global ids
 * first_global_id .. first_global_id + m - 1, and searches report global ids.  Global ids
 * must increase with the local row number so that the tie order (lower id first) of a shard
 * agrees with the global one.  A store uses either vs_append (ids = local row numbers) or
 * vs_append_ids throughout. */
VS_API int vs_append_ids(vs_store* s, const float* rows, int64_t m, int rows_on_device,
                         int64_t first_global_id, void* stream);

/* self._vector_count (service/optimized_vector_store.py:105) */
VS_API int64_t vs_count(const vs_store* s);
/* in-memory part of MLXVectorStore.clear() (:198-209): forget all rows, keep the arena */
VS_API int vs_reset(vs_store* s);
/* bytes of device memory currently mapped for this store (get_stats()['memory_usage_mb']) */
VS_API int64_t vs_memory_bytes(const vs_store* s);

/* Copy rows [first, first+m) of the fp32 master copy to `out` (device or host pointer);
 * serves persistence (`_save_store`, :218-223) and the filtered-subset gather (:167). */
VS_API int vs_read_rows(vs_store* s, int64_t first, int64_t m, float* out, int out_on_device,
                        void* stream);

/* K2/K3 (+K4 merge, +K5 rescoring) -- replaces
 *   _compiled_cosine_similarity / _compiled_euclidean_distance + mx.argsort(...)[:k] + gather
 *   (service/optimized_vector_store.py:31-48,176-184) for B == 1, and
 *   compute_cosine_similarity_batch + mx.argsort(axis=1)[:, :k]
 *   (performance/mlx_optimized.py:59-88,217-248) for B > 1.
 * q: (B, dim) fp32 row-major DEVICE pointer.  out_scores / out_ids: (B, k) DEVICE pointers.
 * Row b holds min(k, count) results, best first; remaining slots id -1.
 * `row_mask` (nullable): device bitmap, bit i of word i/32 set = row i takes part
 * (the metadata filter of :159-167 pushed into the scan / the GEMM epilogue); `mask_live` = its
 * number of set bits (-1 = unknown: the search then takes the masked scan; ignored without a mask).
 * The work is enqueued on `stream`; a certified GEMM search (VS_SEARCH_GEMM, and AUTO when it
 * picks K3) additionally WAITS on the host for its certification count (one int per 2048
 * queries) and, when a query could not be certified, re-runs it before returning.  Use
 * vs_search_submit / vs_search_complete to keep that wait off the critical path. */
VS_API int vs_search(vs_store* s, const float* q, int B, int k, int flags,
                     const uint32_t* row_mask, int64_t mask_live, float* out_scores,
                     int32_t* out_ids, void* stream);

/* The same search split in two so that a caller can keep several batches in flight:
 *   vs_search_submit    enqueues every kernel of the search on `stream` and returns without
 *                       waiting for the GPU (the certification count travels to pinned host
 *                       memory behind the last kernel);
 *   vs_search_complete  waits for THAT search's count only, re-runs the queries that could not
 *                       be certified (enqueued on the same stream, results land in the same
 *                       output rows) and frees the ticket.  In the common case (count 0) it
 *                       enqueues nothing.
 * `q`, `row_mask` and the output buffers must stay valid and unmodified until
 * vs_search_complete returns; the results are final (in stream order) only after it.
 * New in this engine (the reference is synchronous, service/optimized_vector_store.py:116-192). */
typedef struct vs_ticket vs_ticket;
VS_API int vs_search_submit(vs_store* s, const float* q, int B, int k, int flags,
                            const uint32_t* row_mask, int64_t mask_live, float* out_scores,
                            int32_t* out_ids, void* stream, vs_ticket** ticket_out);
VS_API int vs_search_complete(vs_store* s, vs_ticket* ticket);
/* vs_search_submit on another stream than the caller's: `search_stream` first waits for everything
 * enqueued on `cur_stream` so far (the queries, an earlier reader of the output buffers), and an event
 * is recorded behind the search for vs_exchange_result.  One call instead of the caller's own event
 * plumbing: a sharded server is host-bound at batch 1 otherwise. */
VS_API int vs_search_submit_on(vs_store* s, const float* q, int B, int k, int flags,
                               const uint32_t* row_mask, int64_t mask_live, float* out_scores,
                               int32_t* out_ids, void* cur_stream, void* search_stream,
                               vs_ticket** ticket_out);

/* Same with HOST buffers: H2D of the queries, search, D2H of the results, stream
 * synchronised on return.  This is the call MLXVectorStore.query()/batch_query() make. */
VS_API int vs_search_host(vs_store* s, const float* q_host, int B, int k, int flags,
                          const uint32_t* row_mask_dev, int64_t mask_live,
                          float* out_scores_host, int32_t* out_ids_host);

/* Per-query certification flags of the last VS_SEARCH_GEMM call on this store/stream are
 * folded into the result (uncertified queries are re-run exactly); this returns how many
 * queries took the exact fallback in total since creation (diagnostic). */
VS_API int64_t vs_fallback_count(const vs_store* s);
/* Queries the GEMM path could not certify with its default candidate count and retried once
 * with four times as many before (if still uncertified) falling back (diagnostic). */
VS_API int64_t vs_retry_count(const vs_store* s);

/* K4 merge_topk -- new (the reference is single-device): merge G candidate lists of k
 * entries per query into (B, k).  Group g's (B, k) block starts at g * group_stride elements
 * from both base pointers (0 = dense (G, B, k)).  Entries with id < 0 are ignored; scores
 * are in the reference convention (euclidean: distances).  Used after the NCCL all-gather
 * of per-GPU local results. */
VS_API int vs_merge(int device, int metric, const float* cand_scores, const int32_t* cand_ids,
                    int G, int B, int k, int64_t group_stride, float* out_scores,
                    int32_t* out_ids, void* stream);

/* Candidate exchange over NVLink peer memory -- new (the reference is single-device): the
 * fused replacement of "all-gather the per-GPU (B, k) results, then merge".  Every rank owns an
 * exchange buffer all ranks of the box have mapped (the host side gets the peer pointers from
 * torch.distributed's symmetric-memory rendezvous).
 *   vs_exchange_push  copies `bytes` (a multiple of 16) from `src` (device) to peer_dst[g] for
 *                     every rank g with direct stores over NVLink, then stores `step` with release
 *                     semantics into peer_flag[g] (this rank's flag word in rank g's buffer).
 *                     peer_dst / peer_flag are HOST arrays of G device pointers; `counter` is a
 *                     zero-initialised device uint32 owned by the caller (block-completion count).
 *   vs_exchange_wait  enqueues a one-warp kernel that returns once the G flag words at `flags`
 *                     (this rank's own memory) are all >= step (traps after ~10 s: a dead rank
 *                     must fail loudly).  vs_merge over the local buffer follows on the same stream. */
VS_API int vs_exchange_push(int device, const void* src, int64_t bytes, void* const* peer_dst,
                            void* const* peer_flag, int G, uint32_t step, void* counter, void* stream);
VS_API int vs_exchange_wait(int device, const void* flags, int G, uint32_t step, void* stream);
/* vs_exchange_wait + the merge of the G blocks (each (2, B, k) int32: fp32 score bits, then ids; block g
 * starts g * block_words words after `blocks`) in ONE kernel: one warp per query, no shared memory, so
 * it runs next to the following search's GEMM instead of waiting for a gap.  G * k <= 256; order and
 * conventions of vs_merge. */
VS_API int vs_exchange_wait_merge(int device, int metric, const void* flags, int G, uint32_t step,
                                  const void* blocks, int64_t block_words, int B, int k,
                                  float* out_scores, int32_t* out_ids, void* stream);

/* One call per sharded search result: vs_search_complete(ticket) (the host wait for the certification
 * count), then on `xs_stream`, ordered behind the search: vs_exchange_push of `src_block`, the wait for all
 * G ranks' blocks and the merge into (B, k) `out_*` (vs_exchange_wait_merge, or vs_exchange_wait + vs_merge
 * beyond G * k = 256); finally `cur_stream` is made to wait for the merge.  `local_flags` / `local_blocks`
 * are this step's flag words / block area in THIS rank's exchange buffer.  Frees the ticket. */
VS_API int vs_exchange_result(vs_store* s, vs_ticket* ticket, const void* src_block, int64_t block_bytes,
                              void* const* peer_dst, void* const* peer_flag, int G, uint32_t step,
                              void* counter, const void* local_flags, const void* local_blocks,
                              int B, int k, float* out_scores, int32_t* out_ids, void* xs_stream,
                              void* cur_stream);

/* K5 rescore_fp32 -- exact fp32 scores (same arithmetic as the fp32 scan) for `kc`
 * candidate ids per query, sorted, best `k` written out.  cand_ids: (B, kc) device. */
VS_API int vs_rescore(vs_store* s, const float* q, int B, const int32_t* cand_ids, int kc,
                      int k, float* out_scores, int32_t* out_ids, void* stream);

/* Stand-alone helpers mirroring performance/mlx_optimized.py on device arrays ------- */

/* normalize_vectors (:110-125): out[i,:] = x[i,:] / max(||x[i,:]||, 1e-8) */
VS_API int vs_normalize_rows(int device, const float* x, int64_t n, int dim, float* out,
                             void* stream);
/* full (B, N) score matrix -- compute_cosine_similarity_{single,batch} (:26-88),
 * compute_euclidean_distance (:139-148), compute_dot_product (:150-156).
 * For API parity only; the search path never materialises this matrix. */
VS_API int vs_score_matrix(int device, int metric, const float* q, int B, const float* db,
                           int64_t n, int dim, float* out, void* stream);

/* Test hook for K3: the full (B, count) matrix of tensor-core scores
 * half(prep(q)) . half(shadow row) (fp16 for cosine, bf16 otherwise) with fp32 accumulation, exactly what the GEMM path's
 * epilogue filters.  `out` is a (B, count) fp32 DEVICE buffer.  Not used by the search path. */
VS_API int vs_debug_gemm_scores(vs_store* s, const float* q, int B, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200VS_H */
