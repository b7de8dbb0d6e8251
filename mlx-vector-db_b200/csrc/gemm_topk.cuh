// K3 gemm_topk: tcgen05/TMEM GEMM over the bf16 shadow with a fused threshold top-k epilogue,
// exact fp32 rescoring (K5) and per-query certification.  See gemm_topk.cu.
#pragma once
#include "common.cuh"

struct vs_store;

namespace vs {

// true when the GEMM path can serve this search (shadow copy present, sizes in range)
bool gemm_supported(const vs_store* s, int64_t n, int B, int kk);

// B queries against n rows; results (kk live entries per query) at out_*[b * out_stride].
// certify: prove per query that the bf16 candidate set contains the exact fp32 top-k, and
// re-run the queries that cannot be proven through the exact fp32 scan.
// fp8: run over the e4m3 shadow (kind::f8f6f4), never certified (recall-reported)
int gemm_path(vs_store* s, int64_t n, const float* q, int B, int kk, bool certify, bool scan_tma, bool fp8,
              float* out_scores, int32_t* out_ids, int64_t out_stride, cudaStream_t stream);

}  // namespace vs
