// K2 scan_topk: HBM-bound streaming scan of the database with a fused metric epilogue and
// an on-chip top-k; the (B, N) score matrix of the reference
// (service/optimized_vector_store.py:41, performance/mlx_optimized.py:86) is never written.
#pragma once
#include "common.cuh"

namespace vs {

constexpr int kMaxScanWarps = 12;   // upper bound on consumer warps per CTA (register budget)
constexpr int kDefScanWarps = 12;   // default consumer warps (8 when the k-lists are large)
constexpr int kMaxTmaWarps = 12;    // TMA variant: register budget of a 1-CTA/SM kernel
constexpr int kMaxQB = 8;           // queries scored per pass over the database

// consumer warps for a query block of `qb` with k-entry lists: 12 while the per-warp lists
// (warps x qb x k x 8 B of shared memory) stay within 32 KB, else 8
inline int scan_warps_for(int qb, int k) {
  return (size_t)kDefScanWarps * qb * k * 8 <= 32 * 1024 ? kDefScanWarps : 8;
}

struct ScanParams {
  const void* db;          // fp32 rows (vec = 4 floats) or 16-bit shadow rows (vec = 8), 16 B vectors
  int64_t n;               // rows visible to this search
  int vec_per_row;         // 16-byte vectors per row
  const float* norms;      // max(||x||,1e-8) per row (fp32 cosine only)
  const float* q;          // prepared queries of this pass: (QB, ldq) floats, see prep_queries
  int ldq;                 // floats per prepared query
  int nb;                  // live queries in this pass (<= QB)
  int k;
  int epilogue;            // VS_METRIC_* of the epilogue (cosine divides by the row norm)
  const uint32_t* row_mask;
  float* part_key;         // (QB, nlists, k) per-CTA partial lists, sorted best first
  int32_t* part_id;
  uint32_t* tau;           // (QB,) shared pruning threshold, enc_key()-encoded, init enc(-inf)
  int nlists;              // gridDim.x
  int warps;               // consumer warps per CTA (set by launch_scan)
  // TMA variant only
  int tile_rows;           // rows per staged tile = warps * rows_per_warp
  int stages;
};

// host launcher: scans `db` for queries [0, nb) of the prepared block and leaves
// per-warp lists + tau behind; returns the number of lists per query via *nlists.
int launch_scan(const ScanParams& base, int qb, bool l2, int fmt, bool use_tma, int num_sms,
                int* nlists_out, size_t* part_elems_out, bool dry_run, cudaStream_t stream);

// prepare queries: cosine -> q / max(||q||, 1e-8); pad to ldq; bf16 database -> split into
// two 4-float planes per 16-byte vector so that shared-memory reads are conflict-free.
// Also resets the shared thresholds tau[0..B) to enc(-inf) when `tau` is given.
int launch_prep_queries(const float* q, int B, int dim, int metric, int ldq, bool bf16_planes,
                        float scale, float* out, float* qnorm_out, uint32_t* tau,
                        cudaStream_t stream);

// K4: per query, select the best `k` of `per_query` candidates (id < 0 = empty slot) into
// out[b*out_stride ..]; slots [k, out_stride) are filled with id -1 / score 0.
struct MergeParams {
  const float* ck;         // candidate scores / keys
  const int32_t* ci;       // candidate ids (local row numbers), < 0 = empty
  int64_t per_query;       // candidates per query
  int64_t chunk;           // candidates per contiguous chunk
  int64_t chunk_stride;    // elements between chunks of the same query
  int64_t query_stride;    // elements between queries inside a chunk
  int list_len;            // > 0: candidates are consecutive lists of list_len entries sorted
                           // best first (enables the head-sample pre-filter); 0: no structure
  int k;
  const uint32_t* tau;     // nullable: encoded lower bound on the k-th best key (pre-filter)
  int negate_in;           // candidates are distances: key = -score
  int negate_out;          // write -key (euclidean distance)
  const int32_t* id_map;   // nullable: local row -> global id (row-sharded stores)
  float* out_s;
  int32_t* out_i;
  int64_t out_stride;
  float* kth_out;          // nullable: (B,) the k-th best key (-inf when fewer than k candidates)
  const int32_t* counts;   // nullable: chunk c of query b holds counts[c * count_stride + b] valid
  int64_t count_stride;    //           candidates (its remaining slots are not read)
};
int launch_merge(const MergeParams& p, int B, cudaStream_t stream);
// scan-list layout: candidates of query b contiguous at b * per_query
int launch_merge(const float* cand_key, const int32_t* cand_id, int64_t per_query, int B, int k,
                 const uint32_t* tau, int negate_scores, float* out_scores, int32_t* out_ids,
                 int64_t out_stride, cudaStream_t stream, const int32_t* id_map = nullptr,
                 int list_len = 0);

// K5
int launch_rescore(const float* rows, int ld, int dim, const float* norms, int metric,
                   const float* qprep, int ldq, int B, const int32_t* cand_ids, int kc,
                   float* cand_keys_out, cudaStream_t stream);

}  // namespace vs
