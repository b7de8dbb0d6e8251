// K3 gemm_topk: tcgen05/TMEM GEMM over the bf16 shadow with a fused threshold top-k epilogue,
// exact fp32 rescoring (K5) and per-query certification.  See gemm_topk.cu.
#pragma once
#include "common.cuh"

#include <vector>

struct vs_store;

namespace vs {
// One enqueued block of queries of a certified GEMM search whose certification count has not
// been read yet (gemm_topk.cu).
struct PendingBlock {
  int slot = -1;
  const float* q = nullptr;
  int B = 0, kk = 0, kc_want = 0;
  int64_t n = 0;
  bool scan_tma = false;
  const uint32_t* row_mask = nullptr;
  float* out_s = nullptr;
  int32_t* out_i = nullptr;
  int64_t out_stride = 0;
};
}  // namespace vs

// vs_search_submit -> vs_search_complete (include/b200vs.h)
struct vs_ticket {
  std::vector<vs::PendingBlock> blocks;
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;     // recorded behind the search on `stream` (vs_search_submit_on; owned by the store's pool)
};

namespace vs {

// true when the GEMM path can serve this search (shadow copy present, sizes in range)
bool gemm_supported(const vs_store* s, int64_t n, int B, int kk);

// B queries against n rows; results (kk live entries per query) at out_*[b * out_stride].
// certify: prove per query that the 16-bit candidate set contains the exact fp32 top-k, and
// re-run the queries that cannot be proven (wider GEMM retry, then the exact fp32 scan).
// ticket != NULL: only enqueue; the certification checks stay pending in the ticket until
// gemm_complete.  ticket == NULL: checks (one host wait per block) happen before returning.
// row_mask / n_live: optional device bitmap of the rows that take part and its popcount.
// fp8: run over the e4m3 shadow (kind::f8f6f4), never certified (recall-reported)
int gemm_path(vs_store* s, int64_t n, const float* q, int B, int kk, bool certify, bool scan_tma, bool fp8,
              const uint32_t* row_mask, int64_t n_live, float* out_scores, int32_t* out_ids, int64_t out_stride,
              cudaStream_t stream, vs_ticket* ticket);
int gemm_complete(vs_store* s, vs_ticket* ticket);
void free_cert_slots(vs_store* s);
void free_ws_blocks(vs_store* s);

}  // namespace vs
