// K4 merge_topk, K5 rescore_fp32 and the small API-parity helpers (normalize_rows,
// score_matrix).
#include "scan_topk.cuh"

namespace vs {

constexpr int kMergeCap = 4096;     // candidates sorted in shared memory
constexpr int kMergeThreads = 512;

// One CTA per query.  Candidate i of query b lives at
//   (i / chunk) * chunk_stride + b * query_stride + (i % chunk)
// (scan lists: one chunk of nlists*k entries; gathered shard results: G chunks of k).
// Candidates at or above the pruning threshold are compacted into shared memory and
// bitonic-sorted by (key desc, id asc); if more than kMergeCap survive (mass ties), fall back
// to k selection passes over global memory (always correct).  Ids are mapped local -> global
// (`id_map`) before they are compared, so the tie order is the global one.
__device__ __forceinline__ int64_t cand_addr(const MergeParams& p, int b, int64_t i) {
  const int64_t c = i / p.chunk;
  return c * p.chunk_stride + (int64_t)b * p.query_stride + (i - c * p.chunk);
}
__device__ __forceinline__ bool cand_load(const MergeParams& p, int b, int64_t i, float thr,
                                          float& key, int& id) {
  if (p.counts) {
    const int64_t c = i / p.chunk;
    if (i - c * p.chunk >= p.counts[c * p.count_stride + b]) return false;
  }
  const int64_t a = cand_addr(p, b, i);
  id = p.ci ? p.ci[a] : (int)i;     // no id array: the id is the candidate's position
  if (id < 0) return false;
  key = p.ck[a];
  if (p.negate_in) key = -key;
  if (!(key >= thr)) return false;
  if (p.id_map) id = p.id_map[id];
  return true;
}

// bitonic sort of sk[0..n2) (and si when WITH_ID) in shared memory, best first
template <bool WITH_ID>
__device__ __forceinline__ void smem_bitonic(float* sk, int* si, int n2, int tid) {
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < n2; i += kMergeThreads) {
        const int j = i ^ stride;
        if (j > i) {
          const bool fwd = (i & size) == 0;
          const float ka = sk[i], kb = sk[j];
          if (WITH_ID) {
            const int ia = si[i], ib = si[j];
            const bool swap = fwd ? better(kb, ib, ka, ia) : better(ka, ia, kb, ib);
            if (swap) { sk[i] = kb; si[i] = ib; sk[j] = ka; si[j] = ia; }
          } else {
            if (fwd ? (kb > ka) : (ka > kb)) { sk[i] = kb; sk[j] = ka; }
          }
        }
      }
      __syncthreads();
    }
  }
}

constexpr int kRankCap = 1024;   // up to this many items are ordered by counting ranks
constexpr int kHistBins = 1024;  // linear bins of the histogram select

__global__ void __launch_bounds__(kMergeThreads)
merge_topk_kernel(const MergeParams p) {
  __shared__ float sk[kMergeCap];
  __shared__ int si[kMergeCap];
  __shared__ int cnt;
  __shared__ int cut_bin;
  __shared__ int hist[kHistBins];
  __shared__ float thr_sh;
  __shared__ float red_k[kMergeThreads / 32];
  __shared__ int red_i[kMergeThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int k = p.k;
  float* os = p.out_s + (int64_t)b * p.out_stride;
  int32_t* oi = p.out_i + (int64_t)b * p.out_stride;
  float thr = p.tau ? dec_key(p.tau[b]) : VS_NEG_INF;
  if (tid == 0) { cnt = 0; thr_sh = VS_NEG_INF; }
  // Head-sample pre-filter for sorted lists: with L lists and pos = ceil(k / L), every list
  // whose pos-th entry is >= t holds at least pos candidates >= t, so t = the
  // ceil(k / pos)-th largest pos-th entry is a lower bound on the k-th best key.
  if (p.list_len > 0 && p.per_query > 2 * (int64_t)k) {
    const int64_t L = p.per_query / p.list_len;
    int pos = (int)((k + L - 1) / L);
    if (pos < 1) pos = 1;
    const int need = (k + pos - 1) / pos;
    if (L >= 2 && L <= kMergeCap && pos <= p.list_len && need <= L) {
      int n2 = 1;
      while (n2 < L) n2 <<= 1;
      for (int i = tid; i < n2; i += kMergeThreads) {
        float key = VS_NEG_INF;
        if (i < L) {
          const int64_t a = cand_addr(p, b, (int64_t)i * p.list_len + pos - 1);
          if (!p.ci || p.ci[a] >= 0) key = p.negate_in ? -p.ck[a] : p.ck[a];
        }
        sk[i] = key;
      }
      __syncthreads();
      if (L <= kRankCap) {
        for (int i = tid; i < (int)L; i += kMergeThreads) {
          const float mine = sk[i];
          int r = 0;
          for (int j = 0; j < (int)L; ++j) {
            const float o = sk[j];
            r += (o > mine) || (o == mine && j < i);
          }
          if (r == need - 1) thr_sh = mine;
        }
      } else {
        smem_bitonic<false>(sk, si, n2, tid);
        if (tid == 0) thr_sh = sk[need - 1];
      }
      __syncthreads();
      thr = fmaxf(thr, thr_sh);
    }
  }
  __syncthreads();
  // compaction with one shared-memory atomic per warp (ballot + prefix popcount)
  int64_t limit = p.per_query;
  if (p.counts && p.chunk >= p.per_query) {          // one dense list: read the valid prefix only
    const int64_t c = p.counts[b];
    if (c < limit) limit = c;
  }
  for (int64_t i0 = 0; i0 < limit; i0 += kMergeThreads) {
    const int64_t i = i0 + tid;
    float key = 0.f; int id = 0;
    const bool ok = i < limit && cand_load(p, b, i, thr, key, id);
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (bal) {
      int base = 0;
      if ((tid & 31) == 0) base = atomicAdd(&cnt, __popc(bal));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (ok) {
        const int pos = base + __popc(bal & ((1u << (tid & 31)) - 1u));
        if (pos < kMergeCap) { sk[pos] = key; si[pos] = id; }
      }
    }
  }
  __syncthreads();
  int n = cnt;
  // Histogram select: many survivors, few wanted.  Bin the keys linearly between their min
  // and max (a monotone map), find the highest bin edge with at least k keys at or above it
  // and keep only those -- typically a little more than k -- for the exact ordering below.
  if (n > 256 && n <= kMergeCap && n > 2 * k) {
    float lo = __int_as_float(0x7f800000), hi = VS_NEG_INF;
    for (int i = tid; i < n; i += kMergeThreads) { lo = fminf(lo, sk[i]); hi = fmaxf(hi, sk[i]); }
    for (int off = 16; off; off >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if ((tid & 31) == 0) { red_k[tid >> 5] = lo; red_i[tid >> 5] = __float_as_int(hi); }
    for (int i = tid; i < kHistBins; i += kMergeThreads) hist[i] = 0;
    __syncthreads();
    lo = red_k[0]; hi = __int_as_float(red_i[0]);
    for (int w = 1; w < kMergeThreads / 32; ++w) {
      lo = fminf(lo, red_k[w]);
      hi = fmaxf(hi, __int_as_float(red_i[w]));
    }
    if (hi > lo && hi - lo < 3.0e38f) {
      const float scale = (float)(kHistBins - 1) / (hi - lo);
      constexpr int kPer = kMergeCap / kMergeThreads;
      float rk[kPer];
      int ri[kPer], rb[kPer];
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int i = tid + j * kMergeThreads;
        rb[j] = -1;
        if (i < n) {
          rk[j] = sk[i]; ri[j] = si[i];
          int bin = (int)((rk[j] - lo) * scale);
          bin = bin < 0 ? 0 : (bin > kHistBins - 1 ? kHistBins - 1 : bin);
          rb[j] = bin;
          atomicAdd(&hist[bin], 1);
        }
      }
      __syncthreads();
      if (tid < 32) {   // suffix sums over 32 bins per lane, then locate the crossing bin
        constexpr int kSpan = kHistBins / 32;
        int mine = 0;
        for (int j = 0; j < kSpan; ++j) mine += hist[tid * kSpan + j];
        int above = 0;   // keys in the bins of higher lanes
        for (int l = 31; l > 0; --l) {
          const int v = __shfl_sync(0xffffffffu, mine, l);
          if (tid < l) above += v;
        }
        const bool cross = above < k && above + mine >= k;
        const unsigned who = __ballot_sync(0xffffffffu, cross);
        if (who == 0) { if (tid == 0) cut_bin = 0; }
        else if (cross) {
          int acc = above, bstar = tid * kSpan;
          for (int j = kSpan - 1; j >= 0; --j) {
            acc += hist[tid * kSpan + j];
            if (acc >= k) { bstar = tid * kSpan + j; break; }
          }
          cut_bin = bstar;
        }
        if (tid == 0) cnt = 0;
      }
      __syncthreads();
      const int cb = cut_bin;
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        if (rb[j] >= cb) {
          const int pos = atomicAdd(&cnt, 1);
          sk[pos] = rk[j]; si[pos] = ri[j];
        }
      }
      __syncthreads();
      n = cnt;
    }
  }
  int written = 0;
  if (n <= kRankCap) {
    // order by counting: rank = number of better survivors; the first k ranks are written
    for (int i = tid; i < n; i += kMergeThreads) {
      const float mk = sk[i];
      const int mi = si[i];
      int r = 0;
      for (int j = 0; j < n; ++j) r += better(sk[j], si[j], mk, mi);
      if (r < k) { os[r] = p.negate_out ? -mk : mk; oi[r] = mi; }
      if (r == k - 1 && p.kth_out) p.kth_out[b] = mk;
    }
    written = n < k ? n : k;
  } else if (n <= kMergeCap) {
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int i = n + tid; i < n2; i += kMergeThreads) { sk[i] = VS_NEG_INF; si[i] = VS_ID_SENTINEL; }
    __syncthreads();
    smem_bitonic<true>(sk, si, n2, tid);
    written = n < k ? n : k;
    for (int i = tid; i < written; i += kMergeThreads) {
      os[i] = p.negate_out ? -sk[i] : sk[i];
      oi[i] = si[i];
      if (i == k - 1 && p.kth_out) p.kth_out[b] = sk[i];
    }
  } else {
    // selection passes: next = best candidate strictly worse than the previous pick
    float last_k = __int_as_float(0x7f800000);  // +inf
    int last_i = -1;
    for (int j = 0; j < k; ++j) {
      float bk = VS_NEG_INF;
      int bi = VS_ID_SENTINEL;
      for (int64_t i = tid; i < p.per_query; i += kMergeThreads) {
        float key; int id;
        if (cand_load(p, b, i, thr, key, id) && better(last_k, last_i, key, id) &&
            better(key, id, bk, bi)) {
          bk = key; bi = id;
        }
      }
      for (int off = 16; off; off >>= 1) {
        const float ok = __shfl_xor_sync(0xffffffffu, bk, off);
        const int oi2 = __shfl_xor_sync(0xffffffffu, bi, off);
        if (better(ok, oi2, bk, bi)) { bk = ok; bi = oi2; }
      }
      if ((tid & 31) == 0) { red_k[tid >> 5] = bk; red_i[tid >> 5] = bi; }
      __syncthreads();
      bk = red_k[0]; bi = red_i[0];
      for (int w = 1; w < kMergeThreads / 32; ++w)
        if (better(red_k[w], red_i[w], bk, bi)) { bk = red_k[w]; bi = red_i[w]; }
      __syncthreads();
      if (bi == VS_ID_SENTINEL) break;
      if (tid == 0) {
        os[j] = p.negate_out ? -bk : bk; oi[j] = bi;
        if (j == k - 1 && p.kth_out) p.kth_out[b] = bk;
      }
      last_k = bk; last_i = bi;
      written = j + 1;
    }
  }
  for (int64_t i = written + tid; i < p.out_stride; i += kMergeThreads) { os[i] = 0.f; oi[i] = -1; }
  if (written < k && tid == 0 && p.kth_out) p.kth_out[b] = VS_NEG_INF;   // fewer than k candidates
}

int launch_merge(const MergeParams& p, int B, cudaStream_t stream) {
  if (B <= 0) return VS_OK;
  merge_topk_kernel<<<B, kMergeThreads, 0, stream>>>(p);
  count_launch();
  VS_CHECK_LAUNCH();
  return VS_OK;
}

int launch_merge(const float* cand_key, const int32_t* cand_id, int64_t per_query, int B, int k,
                 const uint32_t* tau, int negate_scores, float* out_scores, int32_t* out_ids,
                 int64_t out_stride, cudaStream_t stream, const int32_t* id_map, int list_len) {
  MergeParams p = {};
  p.list_len = list_len;
  p.ck = cand_key; p.ci = cand_id;
  p.per_query = per_query; p.chunk = per_query > 0 ? per_query : 1; p.chunk_stride = 0;
  p.query_stride = per_query;
  p.k = k; p.tau = tau; p.negate_in = 0; p.negate_out = negate_scores; p.id_map = id_map;
  p.out_s = out_scores; p.out_i = out_ids; p.out_stride = out_stride;
  return launch_merge(p, B, stream);
}

// K5: one warp per (query, candidate); same accumulation order as the fp32 scan, so the key
// is bit-identical to what K2 computes for that row.
__global__ void __launch_bounds__(256)
rescore_kernel(const float4* __restrict__ rows, int nvec, const float* __restrict__ norms,
               int metric, const float4* __restrict__ q, int qstride, int B,
               const int32_t* __restrict__ cand, int kc, float* __restrict__ keys) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (int64_t)B * kc) return;
  const int b = (int)(w / kc);
  const int id = cand[w];
  if (id < 0) { if (lane == 0) keys[w] = VS_NEG_INF; return; }
  const float4* x = rows + (int64_t)id * nvec;
  const float4* qq = q + (int64_t)b * qstride;
  float acc = 0.f;
  if (metric == VS_METRIC_EUCLIDEAN) {
    for (int c = lane; c < nvec; c += 32) acc = sqdiff4_acc(acc, ldg_stream(x + c), qq[c]);
  } else {
    for (int c = lane; c < nvec; c += 32) acc = dot4_acc(acc, ldg_stream(x + c), qq[c]);
  }
  const float tot = warp_sum(acc);
  float key;
  if (metric == VS_METRIC_COSINE) key = tot / __ldg(norms + id);
  else if (metric == VS_METRIC_EUCLIDEAN) key = -sqrtf(tot);
  else key = tot;
  if (lane == 0) keys[w] = key;
}

int launch_rescore(const float* rows, int ld, int dim, const float* norms, int metric,
                   const float* qprep, int ldq, int B, const int32_t* cand_ids, int kc,
                   float* cand_keys_out, cudaStream_t stream) {
  (void)dim;
  const int64_t warps = (int64_t)B * kc;
  if (warps <= 0) return VS_OK;
  const int64_t blocks = (warps + 7) / 8;
  rescore_kernel<<<(unsigned)blocks, 256, 0, stream>>>(
      reinterpret_cast<const float4*>(rows), ld / 4, norms, metric,
      reinterpret_cast<const float4*>(qprep), ldq / 4, B, cand_ids, kc, cand_keys_out);
  count_launch();
  VS_CHECK_LAUNCH();
  return VS_OK;
}

// ------------------------------------------------ API-parity helpers (not the hot path)
__global__ void normalize_rows_kernel(const float* __restrict__ x, int64_t n, int dim,
                                      float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const float* s = x + r * dim;
  float acc = 0.f;
  for (int c = lane; c < dim; c += 32) acc = fmaf(s[c], s[c], acc);
  const float nrm = fmaxf(sqrtf(warp_sum(acc)), 1e-8f);
  for (int c = lane; c < dim; c += 32) out[r * dim + c] = s[c] / nrm;
}

// out[b, r]: one warp per (row r, all B queries in turn); queries are prepared (normalised
// for cosine) by prep_queries.
__global__ void score_matrix_kernel(const float* __restrict__ q, int ldq, int B,
                                    const float* __restrict__ db, int64_t n, int dim, int metric,
                                    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const float* x = db + r * dim;
  float nrm = 1.f;
  if (metric == VS_METRIC_COSINE) {
    float a = 0.f;
    for (int c = lane; c < dim; c += 32) a = fmaf(x[c], x[c], a);
    nrm = fmaxf(sqrtf(warp_sum(a)), 1e-8f);
  }
  for (int b = 0; b < B; ++b) {
    const float* qq = q + (int64_t)b * ldq;
    float acc = 0.f;
    if (metric == VS_METRIC_EUCLIDEAN) {
      for (int c = lane; c < dim; c += 32) { const float d = x[c] - qq[c]; acc = fmaf(d, d, acc); }
    } else {
      for (int c = lane; c < dim; c += 32) acc = fmaf(x[c], qq[c], acc);
    }
    const float tot = warp_sum(acc);
    float v = tot;
    if (metric == VS_METRIC_COSINE) v = tot / nrm;
    else if (metric == VS_METRIC_EUCLIDEAN) v = sqrtf(tot);
    if (lane == 0) out[(int64_t)b * n + r] = v;
  }
}

}  // namespace vs

using namespace vs;

extern "C" {

int vs_merge(int device, int metric, const float* cand_scores, const int32_t* cand_ids, int G,
             int B, int k, int64_t group_stride, float* out_scores, int32_t* out_ids,
             void* stream_) {
  VS_REQUIRE(G >= 1 && B >= 0 && k >= 0, "G >= 1, B >= 0, k >= 0 required");
  VS_REQUIRE(metric >= 0 && metric <= 2, "bad metric");
  if (B == 0 || k == 0) return VS_OK;
  VS_REQUIRE(cand_scores && cand_ids && out_scores && out_ids, "NULL pointer");
  if (group_stride == 0) group_stride = (int64_t)B * k;
  VS_REQUIRE(group_stride >= (int64_t)B * k, "group_stride must be >= B*k");
  cudaStream_t stream = (cudaStream_t)stream_;
  VS_CUDA(cudaSetDevice(device));
  // candidates arrive as G groups of (B, k) scores in reference convention (euclidean:
  // distances); the kernel reads them in place.
  MergeParams p = {};
  p.ck = cand_scores; p.ci = cand_ids;
  p.per_query = (int64_t)G * k; p.chunk = k; p.chunk_stride = group_stride; p.query_stride = k;
  p.list_len = k;
  p.k = k; p.tau = nullptr;
  p.negate_in = p.negate_out = metric == VS_METRIC_EUCLIDEAN ? 1 : 0;
  p.id_map = nullptr;
  p.out_s = out_scores; p.out_i = out_ids; p.out_stride = k;
  return launch_merge(p, B, stream);
}

int vs_normalize_rows(int device, const float* x, int64_t n, int dim, float* out, void* stream_) {
  VS_REQUIRE(n >= 0 && dim > 0, "n >= 0 and dim > 0 required");
  if (n == 0) return VS_OK;
  VS_REQUIRE(x && out, "NULL pointer");
  VS_CUDA(cudaSetDevice(device));
  normalize_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream_>>>(x, n, dim, out);
  count_launch();
  VS_CHECK_LAUNCH();
  return VS_OK;
}

int vs_score_matrix(int device, int metric, const float* q, int B, const float* db, int64_t n,
                    int dim, float* out, void* stream_) {
  VS_REQUIRE(B >= 0 && n >= 0 && dim > 0, "bad sizes");
  VS_REQUIRE(metric >= 0 && metric <= 2, "bad metric");
  if (B == 0 || n == 0) return VS_OK;
  VS_REQUIRE(q && db && out, "NULL pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  VS_CUDA(cudaSetDevice(device));
  float* qp = nullptr;
  VS_CUDA(cudaMallocAsync((void**)&qp, (size_t)B * dim * 4, stream));
  int rc = launch_prep_queries(q, B, dim, metric, dim, false, 1.f, qp, nullptr, nullptr, stream);
  if (!rc) {
    score_matrix_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(qp, dim, B, db, n, dim, metric, out);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = cuda_fail(e, "score_matrix_kernel", __FILE__, __LINE__);
  }
  cudaFreeAsync(qp, stream);
  return rc;
}

}  // extern "C"

