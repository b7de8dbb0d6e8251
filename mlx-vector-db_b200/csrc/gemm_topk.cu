// K3 gemm_topk -- large-batch search as a dense contraction on the 5th-gen tensor cores (host side:
// plan, parameters, the small helper kernels; the tcgen05 kernel itself is in gemm_kernel.cuh).
//
//   S'[q, r] = sum_k h(Q[q, k]) * h(X[r, k])        tcgen05.mma kind::f16 (fp16 for cosine, bf16
//                                                   otherwise), fp32 accumulators in TMEM
//
// replaces `matmul(Qn, DBn.T)` + `argsort(axis=1)[:, :k]` of the reference
// (performance/mlx_optimized.py:86,235-236) without ever writing the (B, N) matrix:
//
//   pass 1  GEMM over a small strided sample of the tiles; epilogue keeps, per query, the maximum of
//           every tile.  The ks-th largest tile maximum is a lower bound tau[q] on the ks-th best
//           score of the whole database (ks distinct rows reach it); higher order statistics of the
//           sample become a ladder of trial thresholds (tau_select_kernel).
//   pass 2  GEMM over all rows; epilogue compares TMEM columns with tau[q] (one thread owns one
//           query: lane = query) and appends the rare survivors to that thread's private candidate
//           buffer.  Survivors are also counted per ladder level (global atomics); once ks rows have
//           reached a level it becomes the threshold of every CTA, so tau tightens as the pass goes.
//   then    one launch per query block: the kc best survivors by 16-bit key, K5 rescoring in exact
//           fp32 with the scan's arithmetic, final ordering, certification: the candidate set
//           provably contains the exact top-k when
//               exact_k-th  >  beta + E,   E >= |exact - 16-bit| for every row,
//           beta = the kc-th candidate's 16-bit key (or the final tau with fewer than kc survivors);
//           otherwise (or if a candidate buffer overflowed) the query is re-run -- once through K3
//           with 4x the candidates, then through the exact fp32 scan (K2).  Results are therefore
//           always the exact fp32 ones.
//
// Kernel anatomy (one CTA per SM, 384 threads): warp 0 = TMA producer, warp 1 = MMA issuer
// (one elected lane; RESIDENT: warps 1 and 3 issue alternate accumulators), warp 2 = TMEM
// allocator, warps 4-11 = epilogue (two groups of four, TMEM lane quadrant = warp % 4).
// Operands are 16-bit, K-major, 128-byte swizzled: one "chunk" is 128 rows x 64
// elements = 16 KB, loaded by one cp.async.bulk.tensor.2d.
//   RESIDENT (K <= 256): the CTA's query tiles (up to 4 x 128 queries) are loaded once and stay
//     in shared memory; database tiles of 128 rows stream through a ring; 4 TMEM accumulators
//     of 128 columns, one per query tile, let the epilogue of tile m overlap the MMAs of m+1.
//   STREAMING (any K): CTA pairs (cta_group::2); every pair serves ONE query tile pair, the first
//     K chunks of which stay resident, the rest and the database chunks stream through a ring per
//     64-wide K step; database tiles of 256 rows; 2 TMEM accumulators of 256 columns.
#include <cstdio>
#include <mutex>
#include <vector>
#include "gemm_kernel.cuh"
#include "gemm_topk.cuh"
#include "scan_topk.cuh"
#include "store.cuh"

namespace vs {

// ------------------------------------------------------------ small helper kernels
// queries -> bf16 (normalised for cosine), padded to (m_tiles*128, ld16); also the per-query
// rounding-error norm ||u - bf16(u)|| and ||u|| for the certification bound
__global__ void prep_queries_bf16_kernel(const float* __restrict__ q, int B, int dim, int metric, int ld16,
                                         int rows_padded, __nv_bfloat16* __restrict__ out,
                                         float* __restrict__ qerr, float* __restrict__ qlen,
                                         int32_t* __restrict__ overflow, int32_t* __restrict__ gcount) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= rows_padded) return;
  if (lane == 0 && overflow != nullptr) { overflow[b] = 0; gcount[b] = 0; }   // this search's per-query state
  __nv_bfloat16* dst = out + (size_t)b * ld16;
  if (b >= B) {
    for (int c = lane; c < ld16; c += 32) dst[c] = __float2bfloat16_rn(0.f);
    return;
  }
  const float* src = q + (size_t)b * dim;
  float acc = 0.f;
  for (int c = lane; c < dim; c += 32) { const float v = src[c]; acc = fmaf(v, v, acc); }
  const float nrm = fmaxf(sqrtf(warp_sum(acc)), 1e-8f);
  float e2 = 0.f, u2 = 0.f;
  for (int c = lane; c < ld16; c += 32) {
    float v = c < dim ? src[c] : 0.f;
    if (metric == VS_METRIC_COSINE) v = v / nrm;
    float vb;
    if (metric == VS_METRIC_COSINE) {     // fp16 operands for unit-norm data, bf16 otherwise
      const __half h = __float2half_rn(v);
      reinterpret_cast<__half*>(dst)[c] = h;
      vb = __half2float(h);
    } else {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      dst[c] = h;
      vb = __bfloat162float(h);
    }
    const float d = v - vb;
    e2 = fmaf(d, d, e2);
    u2 = fmaf(v, v, u2);
  }
  e2 = warp_sum(e2);
  u2 = warp_sum(u2);
  if (lane == 0) { qerr[b] = sqrtf(e2) * 1.00001f; qlen[b] = sqrtf(u2) * 1.00001f; }
}

// queries -> e4m3 of 16 * q / max(||q||, 1e-8), padded to (rows_padded, ld8)
__global__ void prep_queries_fp8_kernel(const float* __restrict__ q, int B, int dim, int ld8, int rows_padded,
                                        unsigned char* __restrict__ out, int32_t* __restrict__ overflow,
                                        int32_t* __restrict__ gcount) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= rows_padded) return;
  if (lane == 0 && overflow != nullptr) { overflow[b] = 0; gcount[b] = 0; }
  unsigned char* dst = out + (size_t)b * ld8;
  if (b >= B) {
    for (int c = lane; c < ld8; c += 32) dst[c] = 0;
    return;
  }
  const float* src = q + (size_t)b * dim;
  float acc = 0.f;
  for (int c = lane; c < dim; c += 32) { const float v = src[c]; acc = fmaf(v, v, acc); }
  const float sc = 16.f / fmaxf(sqrtf(warp_sum(acc)), 1e-8f);
  for (int c = lane; c < ld8; c += 32)
    dst[c] = (unsigned char)__nv_cvt_float_to_fp8(c < dim ? src[c] * sc : 0.f, __NV_SATFINITE, __NV_E4M3);
}

// r-th largest (1-based) of n order-encoded values in shared memory, by a bitwise search over the
// encoding: 32 counting rounds, so the cost does not depend on the data.  Whole warp.
__device__ __forceinline__ uint32_t warp_rth_largest(const uint32_t* v, int n, int r, int lane) {
  uint32_t T = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = T | (1u << bit);
    int c = 0;
    for (int i = lane; i < n; i += 32) c += v[i] >= cand;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= r) T = cand;
  }
  return T;
}

// Pass 1 -> the starting threshold and the threshold ladder of every query.  One warp per query.
//   lvl[q][0] = the ks-th largest of the query's sample maxima: ks distinct rows reach it, so it is a
//               lower bound of the ks-th best 16-bit key of the whole database (-inf with fewer than ks
//               maxima).  tau_cur[q] starts there.
//   lvl[q][1..]: the ceil(ks/2)-th, ceil(ks/4)-th, ... largest sample maximum, up to the largest one
//               (where pass 2 will have seen ks rows after about 2, 4, ... times the sample), then
//               equally spaced extrapolated levels (spacing = the mean spacing per halving of the rank
//               between lvl 0 and lvl 2, shrinking 5 % per level as a Gaussian tail's does).  ANY
//               ascending values are valid here: a level only ever becomes the threshold after pass 2
//               has counted ks rows at or above it (gemm_kernel.cuh).
// Rows past B (padding of the last query tile) get +inf everywhere.  Also clears this search's
// uncertified-query counter.
constexpr int kTauWarps = 4;
constexpr int kTauList = 320;            // sample maxima at or above lvl 0 kept for the ladder (>= kMaxCand)
__global__ void __launch_bounds__(kTauWarps * 32)
tau_select_kernel(const float* __restrict__ gmax, int n_groups, int ks, int B, int rows_padded,
                  uint32_t* __restrict__ tau_cur, float* __restrict__ lvl, int32_t* __restrict__ lvl_cnt,
                  int32_t* __restrict__ bad) {
  extern __shared__ uint32_t tsm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (blockIdx.x == 0 && threadIdx.x == 0 && bad != nullptr) bad[0] = 0;
  const int b = blockIdx.x * kTauWarps + warp;
  if (b >= rows_padded) return;
  float* lv = lvl + (int64_t)b * kLevels;
  if (lane < kLevels) lvl_cnt[(int64_t)b * kLevels + lane] = 0;
  if (b >= B) {
    if (lane < kLevels) lv[lane] = __int_as_float(0x7f800000);
    if (lane == 0) tau_cur[b] = 0xff800000u;             // enc_key(+inf)
    return;
  }
  uint32_t* v = tsm + (size_t)warp * (n_groups + kTauList + kLevels);
  uint32_t* top = v + n_groups;
  const float* src = gmax + (int64_t)b * n_groups;
  for (int i = lane; i < n_groups; i += 32) v[i] = enc_key(src[i]);
  __syncwarp();
  float L[kLevels];
#pragma unroll
  for (int j = 0; j < kLevels; ++j) L[j] = __int_as_float(0x7f800000);
  if (n_groups < ks) {
    L[0] = VS_NEG_INF;
  } else {
    // the ks-th largest maximum: up to 1024 maxima are searched from registers (32 per lane, no memory
    // traffic in the 32 counting rounds), more from shared memory
    uint32_t T0 = 0;
    if (n_groups <= 1024) {
      uint32_t r[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = lane + 32 * i < n_groups ? v[lane + 32 * i] : 0u;   // 0 sorts below every key
      for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = T0 | (1u << bit);
        int c = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) c += r[i] >= cand;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= ks) T0 = cand;
      }
    } else {
      T0 = warp_rth_largest(v, n_groups, ks, lane);
    }
    L[0] = dec_key(T0);
    // the maxima at or above lvl 0 (ks of them, more with ties), compacted
    int m = 0;
    for (int i0 = 0; i0 < n_groups; i0 += 32) {
      const int i = i0 + lane;
      const bool keep = i < n_groups && v[i] >= T0;
      const uint32_t bal = __ballot_sync(0xffffffffu, keep);
      const int pos = m + __popc(bal & ((1u << lane) - 1u));
      if (keep && pos < kTauList) top[pos] = v[i];
      m += __popc(bal);
    }
    __syncwarp();
    if (m <= kTauList) {
      // rank the kept maxima by counting (there are about ks of them): the value of rank r - 1 is the
      // r-th largest; level j takes rank ceil(ks / 2^j)
      float* lsh = reinterpret_cast<float*>(top + kTauList);          // kLevels floats per warp
      if (lane < kLevels) lsh[lane] = __int_as_float(0x7f800000);
      __syncwarp();
      for (int i = lane; i < m; i += 32) {
        const uint32_t mine = top[i];
        int r = 0;
        for (int j = 0; j < m; ++j) { const uint32_t o = top[j]; r += (o > mine) || (o == mine && j < i); }
        int rk = ks;
        for (int j = 1; j < kLevels && rk > 1; ++j) {
          rk = (rk + 1) >> 1;                                          // ceil(ks / 2^j)
          if (r == rk - 1) lsh[j] = dec_key(mine);
        }
      }
      __syncwarp();
      int nl = 1;
      {
        int rk = ks;
        for (int j = 1; j < kLevels && rk > 1; ++j) { rk = (rk + 1) >> 1; nl = j + 1; }
      }
#pragma unroll
      for (int j = 1; j < kLevels; ++j) if (j < nl) L[j] = lsh[j];
      // extrapolated levels
      float d = nl >= 3 ? 0.5f * (L[2] - L[0]) : (nl == 2 ? L[1] - L[0] : 0.f);
#pragma unroll
      for (int j = 1; j < kLevels; ++j) {
        if (j >= nl) {
          d *= 0.95f;
          L[j] = d > 0.f ? L[j - 1] + d : __int_as_float(0x7f800000);
        }
      }
#pragma unroll
      for (int j = 1; j < kLevels; ++j)                   // ascending whatever the data did
        if (!(L[j] > L[j - 1])) L[j] = __int_as_float(0x7f800000);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < kLevels; ++j) lv[j] = L[j];
    tau_cur[b] = enc_key(L[0]);
  }
}

// Candidate selection + K5 + final ordering + certification in ONE launch, one CTA per query.
//   1. the query's filter survivors (dense list written by pass 2) are staged in shared memory
//      and the kc best by 16-bit key are selected (bitwise search for the kc-th largest key
//      beta, then compaction);
//   2. warp 0 prepares the query exactly like prep_queries_kernel, the 8 warps rescore the
//      candidates in exact fp32 with K2's accumulation order (bit-identical scores);
//   3. the candidates are ranked by (key desc, global id asc) by counting; the top k are written;
//   4. certification: with u the prepared query, v a prepared row and u^, v^ their 16-bit
//      roundings,  |u.v - u^.v^| <= ||u|| max||v - v^|| + ||u - u^|| max||v^|| + slack.
//        cosine / dot : key16 = u^.v^            E = that bound
//                       certified  <=>  exact k-th  >  beta + E
//        euclidean    : key16 = 2 u^.v^ - ||v||^2 = ||u||^2 - d^2 up to E = 2 * that bound (+ the
//                       fp32 rounding of ||v||^2, ||u||^2 and d^2, folded into slack)
//                       certified  <=>  d_k^2  <  ||u||^2 - beta - E
//      beta = the kc-th candidate's 16-bit key, or the final filter threshold (tau_cur) when fewer than
//      kc rows passed the filter (no row outside the candidate set has a 16-bit key above beta).  Queries that cannot be
//      certified (or whose candidate buffers overflowed) are appended to bad[1..], bad[0] counts them.
constexpr int kFinishThreads = 256;
constexpr int kMaxCand = 256;
constexpr int kFinishStage = 1024;       // survivors staged in shared memory (the rest is read from global memory)

struct FinishParams {
  const float* q;            // (B, dim) raw queries
  int dim, ld, metric, B, k, kc;
  int64_t n_rows;
  const float* rows;         // (n, ld) fp32 master
  const float* norms;
  const int32_t* id_map;     // nullable
  const float* glist_s;      // (B, kGlobalCap) survivors of the filter: 16-bit keys
  const int32_t* glist_i;    //                 local row ids
  const int32_t* gcount;     // (B,) survivors per query (may exceed kGlobalCap: overflow is set then)
  const uint32_t* tau_cur;   // (B,) final filter thresholds, order-encoded
  const int32_t* overflow;
  const float* qerr;
  const float* qlen;
  const uint32_t* bounds;
  float slack;
  int certify;
  float* out_s;
  int32_t* out_i;
  int64_t out_stride;
  int32_t* bad;              // [0] = number of uncertified queries, [1 + i] = their indices
};

__global__ void __launch_bounds__(kFinishThreads)
select_finish_kernel(const FinishParams p) {
  extern __shared__ __align__(16) unsigned char fsm[];
  uint32_t* sk = reinterpret_cast<uint32_t*>(fsm);                         // encoded 16-bit keys (first kFinishStage)
  float4* qs = reinterpret_cast<float4*>(sk + kFinishStage);               // prepared query, ld floats
  __shared__ float keys[kMaxCand];
  __shared__ int ids[kMaxCand];
  __shared__ int sel[kMaxCand];
  __shared__ int wcount[2][kFinishThreads / 32];
  __shared__ int cnt_hi, cnt_eq, have_sh;
  __shared__ uint32_t T_sh;
  __shared__ float kth_sh, qsq_sh;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool l2 = p.metric == VS_METRIC_EUCLIDEAN;
  const int n = min(p.gcount[b], kGlobalCap);
  const float* gs = p.glist_s + (int64_t)b * kGlobalCap;
  const int32_t* gi = p.glist_i + (int64_t)b * kGlobalCap;
  for (int i = tid; i < min(n, kFinishStage); i += kFinishThreads) sk[i] = enc_key(gs[i]);
  auto key_at = [&](int i) -> uint32_t { return i < kFinishStage ? sk[i] : enc_key(gs[i]); };
  if (tid == 0) { cnt_hi = 0; cnt_eq = 0; have_sh = 0; kth_sh = VS_NEG_INF; T_sh = 0; }
  const float* src = p.q + (size_t)b * p.dim;
  if (warp == 0) {
    float acc = 0.f;
    for (int c = lane; c < p.dim; c += 32) { const float v = src[c]; acc = fmaf(v, v, acc); }
    const float tot = warp_sum(acc);
    const float nrm = fmaxf(sqrtf(tot), 1e-8f);
    float* dst = reinterpret_cast<float*>(qs);
    for (int c = lane; c < p.ld; c += 32) {
      float v = c < p.dim ? src[c] : 0.f;
      dst[c] = p.metric == VS_METRIC_COSINE ? v / nrm : v * 1.f;
    }
    if (lane == 0) qsq_sh = tot;
  }
  __syncthreads();
  // ---- 1. the kc best survivors by 16-bit key: sel[0..nsel) = their positions in the survivor list
  const bool full = n >= p.kc;
  uint32_t T = 0;                                  // full: the kc-th largest key
  if (!full) {
    for (int i = tid; i < n; i += kFinishThreads) sel[i] = i;
  } else if (n <= kFinishThreads) {
    // the usual case (the adaptive filter passes 100-200 rows): one survivor per thread, ranked by counting
    if (tid < n) {
      const uint32_t mk = sk[tid];
      int r = 0;
      for (int j = 0; j < n; ++j) { const uint32_t o = sk[j]; r += (o > mk) || (o == mk && j < tid); }
      if (r < p.kc) sel[r] = tid;
      if (r == p.kc - 1) T_sh = mk;
    }
    __syncthreads();
    T = T_sh;
  } else {
    // bitwise search for the kc-th largest key, then compaction
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = T | (1u << bit);
      int c = 0;
      for (int i = tid; i < n; i += kFinishThreads) c += key_at(i) >= cand;
      c = __reduce_add_sync(0xffffffffu, c);
      if (lane == 0) wcount[bit & 1][warp] = c;
      __syncthreads();
      int tot = 0;
#pragma unroll
      for (int w = 0; w < kFinishThreads / 32; ++w) tot += wcount[bit & 1][w];
      if (tot >= p.kc) T = cand;
    }
    for (int i = tid; i < n; i += kFinishThreads)
      if (key_at(i) > T) sel[atomicAdd(&cnt_hi, 1)] = i;
    __syncthreads();
    const int hi = cnt_hi;                         // < kc: fewer than kc keys exceed the kc-th largest
    for (int i = tid; i < n; i += kFinishThreads)
      if (key_at(i) == T) { const int pos = hi + atomicAdd(&cnt_eq, 1); if (pos < p.kc) sel[pos] = i; }
  }
  __syncthreads();
  const int nsel = full ? p.kc : n;
  // ---- 2. exact fp32 scores of the candidates (K5): four rows in flight per warp, each with K2's
  //         accumulation order (per-lane fma chain over its float4 columns, then the xor butterfly)
  const int nvec = p.ld >> 2;
  constexpr int kWarps = kFinishThreads / 32;
  for (int e0 = warp; e0 < nsel; e0 += 4 * kWarps) {
    int id[4];
    const float4* x[4];
    float acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * kWarps;
      id[u] = e < nsel ? gi[sel[e]] : gi[sel[e0]];
      x[u] = reinterpret_cast<const float4*>(p.rows) + (int64_t)id[u] * nvec;
      acc[u] = 0.f;
    }
    for (int c = lane; c < nvec; c += 32) {
      const float4 qv = qs[c];
      float4 xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) xv[u] = ldg_stream(x[u] + c);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = l2 ? sqdiff4_acc(acc[u], xv[u], qv) : dot4_acc(acc[u], xv[u], qv);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * kWarps;
      const float tot = warp_sum(acc[u]);
      float key;
      if (p.metric == VS_METRIC_COSINE) key = tot / __ldg(p.norms + id[u]);
      else if (l2) key = -sqrtf(tot);
      else key = tot;
      if (lane == 0 && e < nsel) { keys[e] = key; ids[e] = p.id_map ? p.id_map[id[u]] : id[u]; }
    }
  }
  __syncthreads();
  // ---- 3. rank by counting
  float* os = p.out_s + (int64_t)b * p.out_stride;
  int32_t* oi = p.out_i + (int64_t)b * p.out_stride;
  const int kk = (int)(p.n_rows < p.k ? p.n_rows : p.k);
  if (tid < nsel) {
    const float mk = keys[tid];
    const int mi = ids[tid];
    int r = 0;
    for (int j = 0; j < nsel; ++j) r += better(keys[j], ids[j], mk, mi);
    if (r < p.k) { os[r] = l2 ? -mk : mk; oi[r] = mi; }
    if (r == kk - 1) { kth_sh = mk; have_sh = 1; }
  }
  for (int e = (nsel < p.k ? nsel : p.k) + tid; e < p.out_stride; e += kFinishThreads) { os[e] = 0.f; oi[e] = -1; }
  __syncthreads();
  if (tid != 0 || !p.certify) return;
  // ---- 4. certification
  const float max_err = __uint_as_float(p.bounds[0]);
  const float max_len = __uint_as_float(p.bounds[1]);
  const float beta = full ? dec_key(T) : dec_key(p.tau_cur[b]);
  bool ok = p.overflow[b] == 0;
  if (ok && !(!full && p.n_rows <= p.kc)) {
    const float ql = p.qlen[b];
    if (l2) {
      const float E = 2.f * (ql * max_err + p.qerr[b] * max_len) + p.slack * (ql + max_len) * (ql + max_len);
      const float dk = kth_sh;                       // = -distance of the k-th result
      ok = have_sh && dk * dk * 1.000001f < qsq_sh - beta - E;
    } else {
      const float E = ql * max_err + p.qerr[b] * max_len + p.slack * (1.f + ql * max_len);
      ok = have_sh && kth_sh > beta + E;
    }
  }
  if (!ok) p.bad[1 + atomicAdd(p.bad, 1)] = b;
}

__global__ void fill_u32_kernel(uint32_t* p, uint32_t v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

__global__ void gather_queries_kernel(const float* __restrict__ q, int dim, const int32_t* __restrict__ idx,
                                      int n, float* __restrict__ out) {
  const int64_t total = (int64_t)n * dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = q[(int64_t)idx[i / dim] * dim + i % dim];
}
__global__ void scatter_results_kernel(const float* __restrict__ ts, const int32_t* __restrict__ ti, int k,
                                       const int32_t* __restrict__ idx, int n, float* __restrict__ out_s,
                                       int32_t* __restrict__ out_i, int64_t out_stride) {
  const int64_t total = (int64_t)n * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = (int64_t)idx[i / k] * out_stride + i % k;
    out_s[o] = ts[i];
    out_i[o] = ti[i];
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    cudaGetLastError();
  });
  return fn;
}

// (rows, K) bf16 row-major matrix, boxes of 128 rows x 64 elements, 128-byte swizzle
// fmt: 0 = bf16, 1 = fp16, 2 = e4m3 (one byte per element, 128 elements per chunk row)
static int make_map(CUtensorMap* map, const void* base, int64_t rows, int K, int fmt, int box_rows = 128) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return VS_ERR_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const int esz = fmt == 2 ? 1 : 2;
  cuuint64_t strides[1] = {(cuuint64_t)K * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = fmt == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                 : (fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")"); return VS_ERR_CUDA; }
  return VS_OK;
}

// CTAs per MMA.  Measured on B200: CTA pairs win where the kernel streams both
// operands (K > 256: +9 % at D = 768, +67 % at D = 1536).  With resident queries (K <= 256) one
// CTA per MMA is faster (profiles/r01_k3_probe_experiments.txt: 2.24 vs 3.29 ms at 10 M x 128):
// a pair's accumulator is released by eight epilogue warps in two CTAs instead of four in one.
// B200VS_GEMM_CG=1|2 forces one.
static int gemm_cta_group(int kchunks) {
  const char* e = getenv("B200VS_GEMM_CG");       // read per call: tests switch it inside one process
  const int forced = e && *e ? atoi(e) : 0;
  if (forced == 1 || forced == 2) return forced;
  return kchunks <= 4 ? 1 : 2;
}

// m_tiles: 128-row query tiles, already a multiple of cg
// l2: reserve the epilogue groups' ||x||^2 buffers (2 per group, TN floats each)
static void plan_gemm(int kchunks, int m_tiles, int cg, bool l2, GemmPlan* plan) {
  const size_t sq_res = l2 ? (size_t)kEpiGroupsRes * 2 * kResTN * 4 : 0;
  const size_t sq_str = l2 ? (size_t)kEpiGroupsStream * 2 * 256 * 4 : 0;
  // tail of the carve-up: barriers (512 B), the hit queues, euclidean's ||x||^2 buffers
  const size_t tail_res = 512 + hq_bytes(true) + sq_res, tail_str = 512 + hq_bytes(false) + sq_str;
  const size_t limit = 227 * 1024 - 1024 /*alignment*/ - std::max(tail_res, tail_str);
  const int um_tiles = m_tiles / cg;
  if (kchunks <= 4) {   // K <= 256: resident queries
    int mt = kchunks <= 2 ? 4 : 1;
    while (mt > um_tiles) mt >>= 1;
    if (mt < 1) mt = 1;
    const size_t a = (size_t)mt * kchunks * kChunkBytes;
    const size_t stage = (size_t)kchunks * (kResTN / cg) * 128;  // kResTN / cg database rows per CTA
    int stages = (int)((limit - a) / stage);
    if (stages > 4) stages = 4;
    if (stages >= 2) {
      GemmPlan pl;
      pl.mt = mt; pl.cg = cg; pl.stages = stages; pl.smem = a + stages * stage + 1024 + tail_res; pl.tn = kResTN;
      *plan = pl;
      return;
    }
  }
  // STREAMING: query and database chunks stream through a ring of one-chunk slots (16 KB per CTA
  // of a pair, 32 KB for a single CTA's 256 database rows): a K step takes a query slot and a
  // database slot, 12 slots = 6 K steps in flight.  Optional query-group layout (like RESIDENT's):
  // every unit serves ONE query tile whose first K chunks stay resident next to a 6-slot ring
  // (D = 768: 8 of 12 chunks, a third less L2 -> SM operand traffic).
  const size_t slot = (size_t)(cg == 2 ? 1 : 2) * kChunkBytes;
  const size_t lim = limit;
  int a_res = 0, m_per_unit = um_tiles;
  // Measured on the B200 (profiles/r02_k3_probe.txt): with query groups the ring is left with three
  // K steps in flight and the kernel is SLOWER (1 M x 768: 1.49 vs 1.30 ms, 1 M x 1536: 2.63 vs 2.29 ms),
  // so one group is the default; B200VS_GEMM_QGROUPS=1 selects the query-group layout.
  // A unit with a single query tile has no grouping to lose: when ALL of the tile's K chunks fit next to
  // the ring (K <= 448) they stay resident and only database chunks stream (config E, 5 M x 384 batch
  // 256: 0.9 vs 1.7 ms per append + query cycle).
  const char* e = getenv("B200VS_GEMM_QGROUPS");
  const bool single_resident = um_tiles == 1 && (int64_t)kchunks * kChunkBytes + 6 * (int64_t)slot <= (int64_t)lim;
  if ((e && *e == '1') || (single_resident && !(e && *e == '0'))) {
    m_per_unit = 1;
    a_res = (int)std::min<int64_t>(kchunks, ((int64_t)lim - 6 * (int64_t)slot) / kChunkBytes);
    if (a_res < 0) a_res = 0;
  }
  int stages = (int)((lim - (size_t)a_res * kChunkBytes) / slot);
  if (stages > 12) stages = 12;
  GemmPlan pl;
  pl.a_res_k = a_res; pl.m_per_unit = m_per_unit; pl.mt = 0; pl.cg = cg; pl.stages = stages;
  pl.smem = (size_t)a_res * kChunkBytes + stages * slot + 1024 + tail_str; pl.tn = 256;
  *plan = pl;
}

// rows per TMA box of the database operand: RESIDENT CTAs load 128 / cg rows per tile in one
// box, STREAMING CTAs 256 / cg rows as one or two boxes of 128
static int x_box_rows(const GemmPlan& plan) { return plan.mt > 0 ? std::min(128, kResTN / plan.cg) : 128; }

// units (CTAs or CTA pairs) a launch uses, its query groups and candidate lists per query
static void gemm_units(const GemmPlan& plan, int m_tiles, int n_tiles, int num_sms, int* units_out, int* ngroups_out,
                       int* lists_out) {
  int units = num_sms / plan.cg;
  const int um_tiles = m_tiles / plan.cg;
  const int span = plan.mt > 0 ? plan.mt : plan.m_per_unit;      // query tiles per group
  const int ngroups = (um_tiles + span - 1) / span;
  const int64_t want = (int64_t)n_tiles * ngroups;
  if (want < units) units = (int)want;
  if (units < ngroups) units = ngroups;
  const int lists = (units + ngroups - 1) / ngroups;
  *units_out = units; *ngroups_out = ngroups; *lists_out = lists;
}

static int launch_gemm(const GemmPlan& plan, const CUtensorMap& mq, const CUtensorMap& mx, GemmParams p,
                       int num_sms, int* lists_out, cudaStream_t stream) {
  p.stages = plan.stages;
  p.a_res_k = plan.a_res_k;
  p.m_per_unit = plan.m_per_unit > 0 ? plan.m_per_unit : 1;
  if (p.tile_stride < 1) p.tile_stride = 1;
  p.idesc = instr_desc(kTileM * plan.cg, plan.tn, p.fp16 ? 0 : 1);
  int units, lists;
  gemm_units(plan, p.m_tiles, p.n_tiles, num_sms, &units, &p.ngroups, &lists);
  if (lists_out) *lists_out = lists;
  if (p.mode == kModeMax) {
    const int per_unit = (p.n_tiles + lists - 1) / lists;            // most tiles any unit gets
    p.groups_per_unit = (per_unit + p.group_tiles - 1) / p.group_tiles;
    p.n_groups = lists * p.groups_per_unit;
  }
  const int grid = units * plan.cg;
  if (g_trace == 1)
    fprintf(stderr, "[b200vs trace] gemm mode %d nq %d m_tiles %d n_tiles %d first %d stride %d kch %d mt %d cg %d stages %d "
            "a_res %d m_per_unit %d ngroups %d lists %d grid %d smem %zu n_groups %d gpu %d gt %d l2 %d mask %d\n",
            p.mode, p.nq, p.m_tiles, p.n_tiles, p.tile_first, p.tile_stride, p.kchunks, plan.mt, plan.cg, p.stages,
            p.a_res_k, p.m_per_unit, p.ngroups, lists, grid, plan.smem, p.n_groups, p.groups_per_unit, p.group_tiles,
            p.sqnorms != nullptr, p.row_mask != nullptr);
  // the general-key kernels only where they are needed: the plain ones carry no trace of them
  if (p.sqnorms != nullptr || p.row_mask != nullptr) return launch_gemm_general(plan, mq, mx, p, grid, stream);
  return launch_gemm_plain(plan, mq, mx, p, grid, stream);
}

static int gemm_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200VS_GEMM"); v = (e && strcmp(e, "0") == 0) ? 0 : 1; }
  return v;
}
// Smallest batch AUTO sends to K3.  Measured on B200 (10M x 128): K3 takes 0.50-0.58 ms for any
// batch of 1..64 queries (it is HBM-bound on the 16-bit shadow there, 6.5 TB/s), K2 needs 0.88 ms
// for one query over the fp32 rows and ~4 ms per pass of 8 -- so every cosine / dot_product search
// goes to K3 (exact by certification); euclidean single queries keep the fp32 scan, whose direct-
// difference form needs no certification margin.  B200VS_GEMM_MIN_BATCH overrides.
static int gemm_min_batch(const vs_store* s) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B200VS_GEMM_MIN_BATCH"); v = e && *e ? atoi(e) : 0; }
  if (v > 0) return v;
  return s->metric == VS_METRIC_EUCLIDEAN ? 2 : 1;
}

// candidates kept for rescoring
// Candidates kept for the exact rescoring: the certification margin is the gap between the k-th and
// the kc-th best row.  Cosine keys are fp16 of unit vectors (error ~4e-4 against gaps of 1e-2..1e-3):
// 2k is enough.  Euclidean / dot_product keys carry the bf16 rounding of un-normalised rows (config C,
// 1 M x 1536 top-100: E ~ 17 in squared distance against a rank-100 to rank-200 gap of ~20, which
// failed for 60 % of the queries and cost a second GEMM pass): 2.5k there, capped by kMaxCand.
static int cand_count(const vs_store* s, int kk) {
  const int base = std::max(2 * kk, kk + 22);
  if (s->metric == VS_METRIC_COSINE || base > kMaxCand) return base;
  return std::min(kMaxCand, std::max(5 * kk / 2, kk + 22));
}
// rows the pass-1 threshold is guaranteed to admit (see gemm_block)
static int sample_rank(int kk) { return kk + std::max(6, kk / 2); }

bool gemm_supported(const vs_store* s, int64_t n, int B, int kk) {
  if (!gemm_enabled() || !s->shadow) return false;
  if (B < gemm_min_batch(s) || s->dim > 8192) return false;
  if (cand_count(s, kk) > kMaxCand) return false;            // k <= 128
  if (n < 65536) return false;                              // small stores: the scan is enough
  return true;
}

// The workspace of one enqueue: one recycled device block carved into 1 KB-aligned pieces.
// Blocks live in the store (vs_store::ws_blocks), not in the stream-ordered allocator: with two
// searches in flight on two streams the allocator either serialises the streams (a block freed on
// one stream is reused on the other behind an internal dependency) or allocates afresh per search.
struct Ws {
  std::vector<std::pair<void**, size_t>> items;
  vs_store* store = nullptr;
  int block = -1;
  cudaStream_t stream = nullptr;
  template <typename T> void want(T** slot, size_t count) { items.push_back({(void**)slot, count * sizeof(T)}); }
  int alloc(vs_store* s, cudaStream_t st) {
    store = s;
    stream = st;
    size_t total = 0;
    for (auto& it : items) total += (size_t)round_up((int64_t)it.second, 1024);
    if (total == 0) total = 1024;
    {
      std::lock_guard<std::mutex> g(s->ws_mu);
      int pick = -1;
      for (int pass = 0; pass < 2 && pick < 0; ++pass)
        for (size_t i = 0; i < s->ws_blocks.size() && pick < 0; ++i) {
          auto& b = s->ws_blocks[i];
          if (b.busy || b.bytes < total) continue;
          // pass 0: last used on this stream; pass 1: its last use on another stream has completed
          if (pass == 0 ? b.stream == st : cudaEventQuery(b.ev) == cudaSuccess) pick = (int)i;
        }
      cudaGetLastError();                       // cudaErrorNotReady from the queries above
      if (pick < 0) {
        vs_store::WsBlock b;
        // (a block too small for this search stays in the list for smaller ones; the list is short:
        //  one or two blocks per stream in flight and per size class reached)
        const size_t bytes = (size_t)round_up((int64_t)total + total / 8, 1 << 20);
        VS_CUDA(cudaMalloc(&b.ptr, bytes));
        b.bytes = bytes;
        VS_CUDA(cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming));
        s->ws_blocks.push_back(b);
        pick = (int)s->ws_blocks.size() - 1;
      }
      s->ws_blocks[pick].busy = true;
      block = pick;
    }
    unsigned char* base = (unsigned char*)s->ws_blocks[block].ptr;
    size_t off = 0;
    for (auto& it : items) { *it.first = base + off; off += (size_t)round_up((int64_t)it.second, 1024); }
    return VS_OK;
  }
  // everything that uses the block has been enqueued on `stream`
  ~Ws() {
    if (block < 0) return;
    std::lock_guard<std::mutex> g(store->ws_mu);
    auto& b = store->ws_blocks[block];
    cudaEventRecord(b.ev, stream);
    b.stream = stream;
    b.busy = false;
  }
};
void free_ws_blocks(vs_store* s) {
  for (auto& b : s->ws_blocks) {
    if (b.ev) cudaEventDestroy(b.ev);
    if (b.ptr) cudaFree(b.ptr);
  }
  s->ws_blocks.clear();
}

// exact fp32 scan of selected queries (defined in search.cu)
int scan_queries_exact(vs_store* s, int64_t n, const float* q, int B, int kk, bool use_tma, const uint32_t* row_mask,
                       float* out_scores, int32_t* out_ids, int64_t out_stride, cudaStream_t stream);

// ---- certification slots: what outlives the enqueue of a certified GEMM block until its check.
// A slot owns a pinned int the uncertified-query count is copied to, the event recorded behind
// that copy and the device list of uncertified queries; slots are recycled, never freed before
// vs_destroy.
static int acquire_slot(vs_store* s, int* out) {
  std::lock_guard<std::mutex> g(s->slot_mu);
  for (size_t i = 0; i < s->slots.size(); ++i)
    if (!s->slots[i].busy) { s->slots[i].busy = true; *out = (int)i; return VS_OK; }
  vs_store::CertSlot c;
  VS_CUDA(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
  VS_CUDA(cudaMallocHost((void**)&c.h_bad, sizeof(int)));
  VS_CUDA(cudaMalloc((void**)&c.d_bad, (size_t)(1 + kMaxQueriesPerLaunch) * 4));
  c.busy = true;
  s->slots.push_back(c);
  *out = (int)s->slots.size() - 1;
  return VS_OK;
}
static void release_slot(vs_store* s, int slot) {
  std::lock_guard<std::mutex> g(s->slot_mu);
  s->slots[slot].busy = false;
}
void free_cert_slots(vs_store* s) {
  for (auto& c : s->slots) {
    if (c.done) cudaEventDestroy(c.done);
    if (c.h_bad) cudaFreeHost(c.h_bad);
    if (c.d_bad) cudaFree(c.d_bad);
  }
  s->slots.clear();
}

static int gemm_block_complete(vs_store* s, const PendingBlock& pb, cudaStream_t stream);

// Enqueue one block: prep -> pass 1 -> tau -> pass 2 -> select/rescore/certify.  Never blocks.
// certify: the uncertified-query count travels to the slot's pinned int behind the last kernel;
// *pending describes what gemm_block_complete needs to finish the block.
// kc_want: candidates kept for rescoring (0 = default for kk).
static int gemm_block_enqueue(vs_store* s, int64_t n, const float* q, int B, int kk, bool certify, bool scan_tma,
                              const uint32_t* row_mask, int64_t n_live, float* out_scores, int32_t* out_ids,
                              int64_t out_stride, int kc_want, bool fp8, cudaStream_t stream, PendingBlock* pending) {
  pending->slot = -1;
  const bool l2 = s->metric == VS_METRIC_EUCLIDEAN;
  // fp8: e4m3 shadow, 128 elements per 128-byte chunk row; otherwise the 16-bit shadow, 64 per row
  const int K = fp8 ? s->ld8 : s->ld16;
  const int kch = fp8 ? K / 128 : K / kChunkK;
  const int cg = gemm_cta_group(kch);
  const int m_tiles = ((B + kTileM * cg - 1) / (kTileM * cg)) * cg;     // a multiple of cg
  const int rows_padded = m_tiles * kTileM;
  const int kc = (int)std::min<int64_t>(kc_want > 0 ? kc_want : cand_count(s, kk), n);
  GemmPlan plan;
  plan_gemm(kch, m_tiles, cg, l2, &plan);
  const int tn = plan.tn;
  const int n_tiles = (int)((n + tn - 1) / tn);
  // Pass 1 (the kernel in MAX mode over a sample of the tiles, spread over the whole row range)
  // only SEEDS the filter: tau_select turns the per-tile maxima of the sample into a starting
  // threshold (the ks-th largest maximum: ks distinct rows reach it, so it is a lower bound of the
  // ks-th best 16-bit key of the database) and a ladder of higher trial thresholds that pass 2
  // climbs as soon as it has itself counted ks rows at a level.  The filter therefore passes about
  // ks (log2(1 / f) + 1) rows per query however small the sample fraction f is, and the sample can
  // be small: eight tiles per unit (measured on the B200, profiles/r02_k3_probe.txt: 10 M x 128 is
  // fastest with ~800 sample tiles = 1 %, a 1.25 M-row shard with ~600 = 6 %; fewer tiles leave the
  // ladder's first levels too low, more cost pass-1 time), at least 4 ks tiles.  ks only has to exceed k by a margin (the filter may pass FEWER than kc rows: then
  // no row outside the candidate set exceeds the final threshold and the certification uses that);
  // a wide retry (kc_want > 0) asks for all of its kc candidates.
  // With a row mask only n_live rows take part: the sample must hold enough of THOSE.
  int ks = kc_want > 0 ? kc : std::min(kc, sample_rank(kk));
  if (const char* e = getenv("B200VS_GEMM_KS")) { if (*e) ks = std::max(1, std::min(kc, atoi(e))); }   // diagnostic
  const double live_frac = std::max(1e-9, std::min(1.0, (double)n_live / (double)n));
  const int full_tiles = (int)(n / tn);
  int f_units, f_ngroups, f_lists;
  gemm_units(plan, m_tiles, n_tiles, s->num_sms, &f_units, &f_ngroups, &f_lists);
  int64_t want_tiles = std::max<int64_t>((int64_t)(4 * ks / live_frac), 8 * (int64_t)f_lists);
  if (const char* e = getenv("B200VS_GEMM_SAMPLE")) { if (*e) want_tiles = (int64_t)(atof(e) * (double)n / tn); }   // diagnostic
  // at most 4096 maxima per query (tau_select's shared memory), counted in whole rounds of the units
  int s_tiles = (int)std::min<int64_t>(want_tiles, full_tiles);
  if (s_tiles > 4096 / f_lists * f_lists) s_tiles = 4096 / f_lists * f_lists;
  // every `s_stride`-th tile, not a prefix: on a time-ordered or clustered ingest a prefix can be
  // unlike the rest and give a useless starting threshold
  const int s_stride = s_tiles > 0 ? std::max(1, full_tiles / s_tiles) : 1;
  int s_units, s_ngroups, s_lists;
  gemm_units(plan, m_tiles, s_tiles, s->num_sms, &s_units, &s_ngroups, &s_lists);
  const int gt = 1;                                     // one maximum per sample tile
  const int s_groups = s_lists * ((s_tiles + s_lists - 1) / s_lists);
  const bool sampled = s_tiles >= 2 * ks && s_groups <= 4096 && live_frac >= 0.2;
  // Too few rows for a useful threshold (fewer than 2 ks sample tiles, or a filter that leaves
  // less than a fifth of the rows): every row would be a candidate and the per-thread buffers
  // would overflow.  The exact scan serves such a search directly (for the uncertified modes
  // too: its recall is 1).
  if (!sampled) {
    s->fallbacks.fetch_add(B);
    return scan_queries_exact(s, n, q, B, kk, scan_tma, row_mask, out_scores, out_ids, out_stride, stream);
  }

  int slot = -1;
  if (certify) { if (int rc = acquire_slot(s, &slot)) return rc; }
  int32_t* bad = certify ? s->slots[slot].d_bad : nullptr;
  struct SlotGuard {   // an error return before the hand-over gives the slot back
    vs_store* s; int slot; bool keep = false;
    ~SlotGuard() { if (slot >= 0 && !keep) release_slot(s, slot); }
  } guard{s, slot};

  const int max_lists = s->num_sms;
  Ws ws;
  __nv_bfloat16* qb; float *qerr, *qlen, *lvl, *gmax, *cs, *gls; int32_t *ci, *ccnt, *ovf, *gli, *gcnt, *lcnt;
  uint32_t* tau;
  ws.want(&qb, (size_t)rows_padded * K);
  ws.want(&qerr, (size_t)rows_padded);
  ws.want(&qlen, (size_t)rows_padded);
  ws.want(&tau, (size_t)rows_padded);
  ws.want(&lvl, (size_t)rows_padded * kLevels);
  ws.want(&lcnt, (size_t)rows_padded * kLevels);
  ws.want(&gmax, (size_t)rows_padded * s_groups);
  ws.want(&cs, (size_t)max_lists * rows_padded * kCandCap);
  ws.want(&ci, (size_t)max_lists * rows_padded * kCandCap);
  ws.want(&ccnt, (size_t)max_lists * rows_padded);
  ws.want(&gls, (size_t)rows_padded * kGlobalCap);
  ws.want(&gli, (size_t)rows_padded * kGlobalCap);
  ws.want(&ovf, (size_t)rows_padded);      // cleared by the prep kernel, like gcnt
  ws.want(&gcnt, (size_t)rows_padded);
  if (int rc = ws.alloc(s, stream)) return rc;

  CUtensorMap mq, mx;
  const bool fp16 = s->metric == VS_METRIC_COSINE;
  const int fmt = fp8 ? 2 : (fp16 ? 1 : 0);
  if (int rc = make_map(&mq, qb, rows_padded, K, fmt)) return rc;
  if (int rc = make_map(&mx, fp8 ? s->shadow8_rows.ptr() : s->shadow_rows.ptr(), n, K, fmt, x_box_rows(plan))) return rc;

  if (fp8)   // (the buffer is sized for 2 bytes per element; e4m3 uses half of it)
    prep_queries_fp8_kernel<<<(rows_padded + 7) / 8, 256, 0, stream>>>(q, B, s->dim, K, rows_padded,
                                                                     reinterpret_cast<unsigned char*>(qb), ovf, gcnt);
  else
    prep_queries_bf16_kernel<<<(rows_padded + 7) / 8, 256, 0, stream>>>(q, B, s->dim, s->metric, K, rows_padded, qb,
                                                                      qerr, qlen, ovf, gcnt);
  count_launch();
  VS_CHECK_LAUNCH();

  GemmParams p = {};
  p.kchunks = kch; p.n_rows = n; p.m_tiles = m_tiles; p.nq = B; p.fp16 = (fp16 || fp8) ? 1 : 0; p.fp8 = fp8 ? 1 : 0;
  p.cand_score = cs; p.cand_id = ci; p.cand_cnt = ccnt; p.overflow = ovf;
  // Rows that must reach a ladder level before it becomes the threshold.  The final threshold then
  // sits near the adapt_rank-th best key, and the certification margin is the gap between the k-th
  // and about the adapt_rank-th best row: ks is enough where the 16-bit error E is small against the
  // spacing of the top scores (cosine: fp16 of unit vectors); euclidean / dot_product keys carry
  // bf16 rounding of un-normalised rows (config C: E ~ 17 against a rank-100 to rank-150 gap of ~11),
  // so there the ladder stops at the kc-th best and beta is the kc-th candidate as before.
  p.tau_cur = tau; p.lvl = lvl; p.lvl_cnt = lcnt;
  p.adapt_rank = s->metric == VS_METRIC_COSINE ? ks : kc;
  if (const char* e = getenv("B200VS_GEMM_RANK")) {      // diagnostic: ks | kc | a number
    if (!strcmp(e, "ks")) p.adapt_rank = ks; else if (!strcmp(e, "kc")) p.adapt_rank = kc; else if (*e) p.adapt_rank = std::max(1, atoi(e));
  }
  if (const char* e = getenv("B200VS_GEMM_ADAPT")) { if (*e == '0') p.adapt_rank = 1 << 30; }   // diagnostic: fixed threshold
  p.glist_s = gls; p.glist_i = gli; p.gcount = gcnt;
  p.sqnorms = l2 ? (const float*)s->sqnorms.ptr() : nullptr;
  p.row_mask = row_mask;

  if (plan.mt == 0)   // STREAMING keeps its running candidate counts in global memory
    VS_CUDA(cudaMemsetAsync(ccnt, 0, (size_t)max_lists * rows_padded * 4, stream));
  // pass 1: per-query maxima of the sample tiles -> starting threshold + ladder
  p.mode = kModeMax; p.n_tiles = s_tiles; p.tile_stride = s_stride; p.gmax = gmax; p.group_tiles = gt;
  if (int rc = launch_gemm(plan, mq, mx, p, s->num_sms, nullptr, stream)) return rc;
  {
    static std::once_flag once;
    std::call_once(once, [] {
      cudaFuncSetAttribute(tau_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    });
    tau_select_kernel<<<(rows_padded + kTauWarps - 1) / kTauWarps, kTauWarps * 32,
                        (size_t)kTauWarps * (s_groups + kTauList + kLevels) * 4, stream>>>(gmax, s_groups, ks, B, rows_padded, tau,
                                                                                 lvl, lcnt, bad);
  }
  count_launch();
  VS_CHECK_LAUNCH();
  // diagnostic (timing only, results are wrong): no row passes the filter, so pass 2 never
  // takes its rare path
  if (const char* e = getenv("B200VS_GEMM_TAU_INF")) {
    if (*e == '1') fill_u32_kernel<<<(rows_padded + 255) / 256, 256, 0, stream>>>(tau, 0xff800000u, rows_padded);
  }
  // pass 2: threshold filter over all rows
  p.mode = kModeFilter; p.n_tiles = n_tiles; p.tile_stride = 1; p.gmax = nullptr;
#ifdef VS_GEMM_DEBUG_MODES
  if (const char* e = getenv("B200VS_GEMM_DBGMODE")) { if (*e == '3' || *e == '4') p.mode = atoi(e); }
#endif
  if (int rc = launch_gemm(plan, mq, mx, p, s->num_sms, nullptr, stream)) return rc;
  // candidate selection + K5 + final ordering + certification, one CTA per query
  {
    FinishParams f = {};
    f.q = q; f.dim = s->dim; f.ld = s->ld; f.metric = s->metric; f.B = B; f.k = kk; f.kc = kc; f.n_rows = n;
    f.rows = (const float*)s->rows.ptr(); f.norms = (const float*)s->norms.ptr(); f.id_map = s->id_map();
    f.glist_s = gls; f.glist_i = gli; f.gcount = gcnt;
    f.tau_cur = tau; f.overflow = ovf; f.qerr = qerr; f.qlen = qlen; f.bounds = s->bounds;
    // fp32 accumulation error of the tensor core and of the exact kernels, relative to ||u|| ||v||
    // (tests/test_gemm_gpu.py measures the tensor core's share against float64)
    f.slack = (l2 ? 8.f : 4.f) * (float)s->dim * 5.9604645e-8f + 1e-6f;
    f.certify = certify ? 1 : 0;
    f.out_s = out_scores; f.out_i = out_ids; f.out_stride = out_stride; f.bad = bad;
    const size_t smem = (size_t)kFinishStage * 4 + (size_t)s->ld * 4;
    static std::once_flag once;
    std::call_once(once, [] {
      cudaFuncSetAttribute(select_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    });
    select_finish_kernel<<<B, kFinishThreads, smem, stream>>>(f);
    count_launch();
    VS_CHECK_LAUNCH();
  }
  if (!certify) return VS_OK;
  VS_CUDA(cudaMemcpyAsync(s->slots[slot].h_bad, bad, 4, cudaMemcpyDeviceToHost, stream));
  VS_CUDA(cudaEventRecord(s->slots[slot].done, stream));
  guard.keep = true;
  pending->slot = slot; pending->q = q; pending->B = B; pending->kk = kk; pending->kc_want = kc_want;
  pending->n = n; pending->scan_tma = scan_tma; pending->row_mask = row_mask;
  pending->out_s = out_scores; pending->out_i = out_ids; pending->out_stride = out_stride;
  return VS_OK;
}

// Wait for the block's certification count; re-run what could not be certified -- once through
// K3 with 4x the candidates (a wider 16-bit margin), then through the exact fp32 scan -- and
// scatter those results over the block's output rows.  The common case (count 0) only waits.
static int gemm_block_complete(vs_store* s, const PendingBlock& pb, cudaStream_t stream) {
  if (pb.slot < 0) return VS_OK;
  vs_store::CertSlot& c = s->slots[pb.slot];
  cudaError_t e = cudaEventSynchronize(c.done);
  const int h_bad = *c.h_bad;
  if (e != cudaSuccess) { release_slot(s, pb.slot); return cuda_fail(e, "cudaEventSynchronize", __FILE__, __LINE__); }
  if (h_bad == 0) { release_slot(s, pb.slot); return VS_OK; }
  struct Bufs {
    float* gq = nullptr; float* ts = nullptr; int32_t* ti = nullptr; cudaStream_t st;
    ~Bufs() { if (gq) cudaFreeAsync(gq, st); if (ts) cudaFreeAsync(ts, st); if (ti) cudaFreeAsync(ti, st); }
  } bufs;
  bufs.st = stream;
  struct Rel { vs_store* s; int slot; ~Rel() { release_slot(s, slot); } } rel{s, pb.slot};
  const int kk = pb.kk;
  const int32_t* bad = c.d_bad + 1;
  VS_CUDA(cudaMallocAsync((void**)&bufs.gq, (size_t)h_bad * s->dim * 4, stream));
  VS_CUDA(cudaMallocAsync((void**)&bufs.ts, (size_t)h_bad * kk * 4, stream));
  VS_CUDA(cudaMallocAsync((void**)&bufs.ti, (size_t)h_bad * kk * 4, stream));
  gather_queries_kernel<<<std::min(1184, (h_bad * s->dim + 255) / 256), 256, 0, stream>>>(pb.q, s->dim, bad, h_bad, bufs.gq);
  count_launch();
  VS_CHECK_LAUNCH();
  int rc;
  const int kc = (int)std::min<int64_t>(cand_count(s, kk), pb.n);
  const int kc_retry = std::min(4 * kc, kMaxCand);
  if (pb.kc_want == 0 && kc_retry > kc && (int64_t)kc_retry * 8 <= pb.n) {
    s->retries.fetch_add(h_bad);
    PendingBlock again;
    rc = gemm_block_enqueue(s, pb.n, bufs.gq, h_bad, kk, true, pb.scan_tma, pb.row_mask, pb.n, bufs.ts, bufs.ti, kk,
                            kc_retry, false, stream, &again);
    if (!rc) rc = gemm_block_complete(s, again, stream);
  } else {
    s->fallbacks.fetch_add(h_bad);
    rc = scan_queries_exact(s, pb.n, bufs.gq, h_bad, kk, pb.scan_tma, pb.row_mask, bufs.ts, bufs.ti, kk, stream);
  }
  if (rc) return rc;
  scatter_results_kernel<<<std::min(1184, (h_bad * kk + 255) / 256), 256, 0, stream>>>(bufs.ts, bufs.ti, kk, bad, h_bad,
                                                                                     pb.out_s, pb.out_i, pb.out_stride);
  count_launch();
  VS_CHECK_LAUNCH();
  return VS_OK;
}

// Enqueue the whole search (blocks of kMaxQueriesPerLaunch queries).  With `ticket` the
// certification checks are left pending in it (gemm_complete finishes them); without, every
// block is completed before the next one is enqueued.
int gemm_path(vs_store* s, int64_t n, const float* q, int B, int kk, bool certify, bool scan_tma, bool fp8,
              const uint32_t* row_mask, int64_t n_live, float* out_scores, int32_t* out_ids, int64_t out_stride,
              cudaStream_t stream, vs_ticket* ticket) {
  if (fp8) {
    if (!s->shadow8) { set_error("store was created without an fp8 shadow copy (VS_SHADOW_FP8)"); return VS_ERR_STATE; }
    certify = false;            // e4m3 rounding is far above any top-k margin: recall-reported variant
  } else if (!s->shadow) { set_error("store was created without a 16-bit shadow copy"); return VS_ERR_STATE; }
  if (cand_count(s, kk) > kMaxCand) { set_error("invalid argument: k too large for the GEMM path (k <= 128)"); return VS_ERR_INVALID; }
  // the fp8 variant keeps 4x the candidates for the exact rescoring
  const int kc_want = fp8 ? std::min(4 * cand_count(s, kk), kMaxCand) : 0;
  for (int b0 = 0; b0 < B; b0 += kMaxQueriesPerLaunch) {
    const int nb = std::min(kMaxQueriesPerLaunch, B - b0);
    PendingBlock pb;
    if (int rc = gemm_block_enqueue(s, n, q + (size_t)b0 * s->dim, nb, kk, certify, scan_tma, row_mask, n_live,
                                    out_scores + (int64_t)b0 * out_stride, out_ids + (int64_t)b0 * out_stride,
                                    out_stride, kc_want, fp8, stream, &pb))
      return rc;
    if (pb.slot < 0) continue;
    if (ticket) ticket->blocks.push_back(pb);
    else if (int rc = gemm_block_complete(s, pb, stream)) return rc;
  }
  return VS_OK;
}

int gemm_complete(vs_store* s, vs_ticket* ticket) {
  int rc = VS_OK;
  for (auto& pb : ticket->blocks) {
    const int r = gemm_block_complete(s, pb, ticket->stream);
    if (r && !rc) rc = r;
  }
  ticket->blocks.clear();
  return rc;
}

// test / API helper: the full (B, n) bf16 tensor-core score matrix (kModeDump)
int gemm_dump_scores(vs_store* s, int64_t n, const float* q, int B, float* out, int64_t ld, cudaStream_t stream) {
  const int K = s->ld16;
  const int kch = K / kChunkK;
  const int cg = gemm_cta_group(kch);
  const int m_tiles = ((B + kTileM * cg - 1) / (kTileM * cg)) * cg;
  const int rows_padded = m_tiles * kTileM;
  GemmPlan plan;
  plan_gemm(kch, m_tiles, cg, false, &plan);
  Ws ws;
  __nv_bfloat16* qb; float *qerr, *qlen;
  ws.want(&qb, (size_t)rows_padded * K);
  ws.want(&qerr, (size_t)rows_padded);
  ws.want(&qlen, (size_t)rows_padded);
  if (int rc = ws.alloc(s, stream)) return rc;
  CUtensorMap mq, mx;
  const bool fp16 = s->metric == VS_METRIC_COSINE;
  if (int rc = make_map(&mq, qb, rows_padded, K, fp16 ? 1 : 0)) return rc;
  if (int rc = make_map(&mx, s->shadow_rows.ptr(), n, K, fp16 ? 1 : 0, x_box_rows(plan))) return rc;
  prep_queries_bf16_kernel<<<(rows_padded + 7) / 8, 256, 0, stream>>>(q, B, s->dim, s->metric, K, rows_padded, qb,
                                                                    qerr, qlen, nullptr, nullptr);
  count_launch();
  VS_CHECK_LAUNCH();
  GemmParams p = {};
  p.kchunks = kch; p.n_rows = n; p.m_tiles = m_tiles; p.nq = B; p.fp16 = fp16 ? 1 : 0;
  p.mode = kModeDump; p.n_tiles = (int)((n + plan.tn - 1) / plan.tn); p.dump = out; p.dump_ld = ld;
  return launch_gemm(plan, mq, mx, p, s->num_sms, nullptr, stream);
}

}  // namespace vs

using namespace vs;

extern "C" int vs_debug_gemm_scores(vs_store* s, const float* q, int B, float* out, void* stream_) {
  VS_REQUIRE(s != nullptr && q != nullptr && out != nullptr, "NULL pointer");
  VS_REQUIRE(B > 0, "B must be > 0");
  if (!s->shadow) { set_error("store was created without a bf16 shadow copy"); return VS_ERR_STATE; }
  VS_CUDA(cudaSetDevice(s->device));
  const int64_t n = s->count.load(std::memory_order_acquire);
  VS_REQUIRE(n > 0, "store is empty");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (s->append_done && s->append_stream.load() != stream) VS_CUDA(cudaStreamWaitEvent(stream, s->append_done, 0));
  return gemm_dump_scores(s, n, q, B, out, n, stream);
}
