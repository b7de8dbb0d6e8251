// vs_search / vs_search_host / vs_rescore: kernel selection and workspace plumbing.
#include <cstdlib>
#include <cstring>
#include "scan_topk.cuh"
#include "gemm_topk.cuh"
#include "store.cuh"
#include "exchange.cuh"

namespace vs {

// one stream-ordered allocation carved into 256 B-aligned pieces
struct Workspace {
  struct Item { void** slot; size_t bytes; };
  std::vector<Item> items;
  unsigned char* base = nullptr;
  cudaStream_t stream = nullptr;
  template <typename T> void want(T** slot, size_t count) {
    items.push_back({reinterpret_cast<void**>(slot), count * sizeof(T)});
  }
  int alloc(cudaStream_t st) {
    stream = st;
    size_t total = 0;
    for (auto& it : items) total += (size_t)round_up((int64_t)it.bytes, 256);
    if (total == 0) return VS_OK;
    VS_CUDA(cudaMallocAsync((void**)&base, total, st));
    size_t off = 0;
    for (auto& it : items) { *it.slot = base + off; off += (size_t)round_up((int64_t)it.bytes, 256); }
    return VS_OK;
  }
  ~Workspace() { if (base) cudaFreeAsync(base, stream); }
};

__global__ void fill_empty_kernel(float* s, int32_t* ids, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) { s[i] = 0.f; ids[i] = -1; }
}

static bool default_tma() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200VS_SCAN");
    // measured on B200 (profiles/r01_scan_sweep.txt): direct 128-bit loads beat the
    // TMA-staged ring at every shape, so LDG is the default; B200VS_SCAN=tma selects the ring
    v = (e && strcmp(e, "tma") == 0) ? 1 : 0;
  }
  return v == 1;
}

// K2 path: scan (fp32 master or 16-bit shadow) -> merge.  Results for query b land in
// out_*[b * out_stride ..], kk live entries each.
static int scan_path(vs_store* s, int64_t n, const float* q, int B, int kk, bool bf16,
                     bool use_tma, const uint32_t* row_mask, float* out_scores, int32_t* out_ids,
                     int64_t out_stride, bool map_ids, cudaStream_t stream) {
  const int ldq = bf16 ? s->ld16 : s->ld;
  int qb_max = 8;
  while (qb_max > 1 && (size_t)8 * qb_max * kk * 8 > 64 * 1024) qb_max >>= 1;
  if ((size_t)8 * kk * 8 > 64 * 1024) {
    set_error("invalid argument: k must be <= 1024");
    return VS_ERR_INVALID;
  }
  while (qb_max > 1 && qb_max / 2 >= B) qb_max >>= 1;

  ScanParams p = {};
  p.db = bf16 ? s->shadow_rows.ptr() : s->rows.ptr();
  p.n = n;
  p.vec_per_row = bf16 ? s->ld16 / 8 : s->ld / 4;
  const bool l2 = s->metric == VS_METRIC_EUCLIDEAN;
  // 16-bit shadow: fp16 for cosine (unit-norm rows), bf16 otherwise (store.cu, K1)
  const int fmt = !bf16 ? 0 : (s->metric == VS_METRIC_COSINE ? 2 : 1);
  // the cosine bf16 shadow is stored normalised: its epilogue is a plain dot product
  p.epilogue = (bf16 && s->metric == VS_METRIC_COSINE) ? VS_METRIC_DOT : s->metric;
  p.norms = (!bf16 && s->metric == VS_METRIC_COSINE) ? (const float*)s->norms.ptr() : nullptr;
  p.ldq = ldq;
  p.k = kk;
  p.row_mask = row_mask;

  // sizes of the partial lists for every query-block width this call will use
  size_t part_elems = 0;
  for (int qb = 1; qb <= qb_max; qb <<= 1) {
    int nl = 0; size_t pe = 0;
    p.nb = qb;
    if (int rc = launch_scan(p, qb, l2, fmt, use_tma, s->num_sms, &nl, &pe, true, stream)) return rc;
    if (pe > part_elems) part_elems = pe;
  }
  Workspace ws;
  float* qprep; uint32_t* tau; float* part_key; int32_t* part_id;
  ws.want(&qprep, (size_t)B * ldq);
  ws.want(&tau, (size_t)B + kMaxQB);
  ws.want(&part_key, part_elems);
  ws.want(&part_id, part_elems);
  if (int rc = ws.alloc(stream)) return rc;
  if (int rc = launch_prep_queries(q, B, s->dim, s->metric, ldq, bf16, 1.f, qprep, nullptr, tau, stream))
    return rc;
  p.part_key = part_key;
  p.part_id = part_id;
  for (int b0 = 0; b0 < B; b0 += qb_max) {
    const int nb = B - b0 < qb_max ? B - b0 : qb_max;
    int qb = 1;
    while (qb < nb) qb <<= 1;
    p.q = qprep + (size_t)b0 * ldq;
    p.nb = nb;
    p.tau = tau + b0;
    int nl = 0;
    if (int rc = launch_scan(p, qb, l2, fmt, use_tma, s->num_sms, &nl, nullptr, false, stream)) return rc;
    if (int rc = launch_merge(part_key, part_id, (int64_t)nl * kk, nb, kk, tau + b0, l2 ? 1 : 0,
                              out_scores + (int64_t)b0 * out_stride, out_ids + (int64_t)b0 * out_stride,
                              out_stride, stream, map_ids ? s->id_map() : nullptr, kk))
      return rc;
  }
  return VS_OK;
}

// exact fp32 scan of a query set (the GEMM path's fallback for uncertified queries)
int scan_queries_exact(vs_store* s, int64_t n, const float* q, int B, int kk, bool use_tma, const uint32_t* row_mask,
                       float* out_scores, int32_t* out_ids, int64_t out_stride, cudaStream_t stream) {
  return scan_path(s, n, q, B, kk, false, use_tma, row_mask, out_scores, out_ids, out_stride, true, stream);
}

// exact fp32 rescoring of (B, kc) candidate ids + final ordering into (B, out_stride)
static int rescore_path(vs_store* s, const float* q, int B, const int32_t* cand, int kc, int kk,
                        float* out_scores, int32_t* out_ids, int64_t out_stride,
                        cudaStream_t stream) {
  Workspace ws;
  float* qprep; float* keys;
  ws.want(&qprep, (size_t)B * s->ld);
  ws.want(&keys, (size_t)B * kc);
  if (int rc = ws.alloc(stream)) return rc;
  if (int rc = launch_prep_queries(q, B, s->dim, s->metric, s->ld, false, 1.f, qprep, nullptr, nullptr, stream))
    return rc;
  if (int rc = launch_rescore((const float*)s->rows.ptr(), s->ld, s->dim, (const float*)s->norms.ptr(),
                              s->metric, qprep, s->ld, B, cand, kc, keys, stream))
    return rc;
  return launch_merge(keys, cand, kc, B, kk, nullptr, s->metric == VS_METRIC_EUCLIDEAN ? 1 : 0,
                      out_scores, out_ids, out_stride, stream, s->id_map());
}

}  // namespace vs

using namespace vs;

extern "C" {

// Common body of vs_search / vs_search_submit.  ticket == NULL: certification checks are done
// before returning (one host wait per GEMM block); otherwise they stay pending in the ticket.
static int search_impl(vs_store* s, const float* q, int B, int k, int flags, const uint32_t* row_mask,
                       int64_t mask_live, float* out_scores, int32_t* out_ids, cudaStream_t stream,
                       vs_ticket* ticket) {
  VS_REQUIRE(s != nullptr, "store is NULL");
  VS_REQUIRE(B >= 0, "B must be >= 0");
  if (B == 0 || k <= 0) return VS_OK;
  VS_REQUIRE(q && out_scores && out_ids, "NULL pointer");
  VS_CUDA(cudaSetDevice(s->device));
  const int64_t n = s->count.load(std::memory_order_acquire);
  if (s->append_done && s->append_stream.load(std::memory_order_acquire) != stream)
    VS_CUDA(cudaStreamWaitEvent(stream, s->append_done, 0));
  if (n == 0) {
    const int64_t total = (int64_t)B * k;
    fill_empty_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 1024), 256, 0, stream>>>(
        out_scores, out_ids, total);
    count_launch();
    VS_CHECK_LAUNCH();
    return VS_OK;
  }
  const int kk = (int)std::min<int64_t>(k, n);
  int mode = flags & VS_SEARCH_MODE_MASK;
  bool use_tma = default_tma();
  if (flags & VS_SEARCH_TMA) use_tma = true;
  if (flags & VS_SEARCH_LDG) use_tma = false;
  // rows taking part: the mask's popcount when the caller knows it (the GEMM path sizes its
  // sample from it; unknown -> the masked scan)
  const int64_t n_live = row_mask == nullptr ? n : mask_live;
  if (mode == VS_SEARCH_AUTO) {
    mode = VS_SEARCH_SCAN_FP32;
    if (gemm_supported(s, n, B, kk) && (row_mask == nullptr || (n_live >= 0 && n_live * 5 >= n))) mode = VS_SEARCH_GEMM;
  }
  switch (mode) {
    case VS_SEARCH_SCAN_FP32:
      return scan_path(s, n, q, B, kk, false, use_tma, row_mask, out_scores, out_ids, k, true, stream);
    case VS_SEARCH_SCAN_BF16: {
      if (!s->shadow) { set_error("store was created without a 16-bit shadow copy"); return VS_ERR_STATE; }
      // over-fetch, then exact fp32 rescoring of the candidates (K5)
      int kc = (int)std::min<int64_t>(n, std::max(2 * kk, kk + 32));
      if (kc > 1024) kc = (int)std::min<int64_t>(n, 1024);
      if (kc < kk) { set_error("invalid argument: k too large for the bf16 candidate scan"); return VS_ERR_INVALID; }
      Workspace ws;
      float* cs; int32_t* ci;
      ws.want(&cs, (size_t)B * kc);
      ws.want(&ci, (size_t)B * kc);
      if (int rc = ws.alloc(stream)) return rc;
      if (int rc = scan_path(s, n, q, B, kc, true, use_tma, row_mask, cs, ci, kc, false, stream)) return rc;
      return rescore_path(s, q, B, ci, kc, kk, out_scores, out_ids, k, stream);
    }
    case VS_SEARCH_GEMM:
    case VS_SEARCH_GEMM_NOCERT:
    case VS_SEARCH_GEMM_FP8:
      if (row_mask != nullptr && n_live < 0) {
        set_error("invalid argument: the GEMM path needs mask_live (the number of set bits) with a row_mask");
        return VS_ERR_INVALID;
      }
      return gemm_path(s, n, q, B, kk, mode == VS_SEARCH_GEMM, use_tma, mode == VS_SEARCH_GEMM_FP8, row_mask, n_live,
                       out_scores, out_ids, k, stream, ticket);
    default:
      set_error("invalid argument: unknown search mode");
      return VS_ERR_INVALID;
  }
}

int vs_search(vs_store* s, const float* q, int B, int k, int flags, const uint32_t* row_mask, int64_t mask_live,
              float* out_scores, int32_t* out_ids, void* stream_) {
  return search_impl(s, q, B, k, flags, row_mask, mask_live, out_scores, out_ids, (cudaStream_t)stream_, nullptr);
}

int vs_search_submit(vs_store* s, const float* q, int B, int k, int flags, const uint32_t* row_mask,
                     int64_t mask_live, float* out_scores, int32_t* out_ids, void* stream_, vs_ticket** ticket_out) {
  VS_REQUIRE(ticket_out != nullptr, "ticket_out is NULL");
  *ticket_out = nullptr;
  vs_ticket* t = new vs_ticket();
  t->stream = (cudaStream_t)stream_;
  const int rc = search_impl(s, q, B, k, flags, row_mask, mask_live, out_scores, out_ids, t->stream, t);
  if (rc) {                      // blocks already enqueued still hold slots: finish them, keep the first error
    gemm_complete(s, t);
    delete t;
    return rc;
  }
  *ticket_out = t;
  return VS_OK;
}

// an event of the store's pool (32 events, round-robin: a wait enqueued on an event keeps the state
// the event had at that moment, so re-recording it later is safe)
static cudaEvent_t pool_event(vs_store* s) {
  std::lock_guard<std::mutex> g(s->ev_mu);
  if (s->ev_pool.size() < 32) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    s->ev_pool.push_back(e);
    return e;
  }
  return s->ev_pool[s->ev_next++ % s->ev_pool.size()];
}

int vs_search_submit_on(vs_store* s, const float* q, int B, int k, int flags, const uint32_t* row_mask,
                        int64_t mask_live, float* out_scores, int32_t* out_ids, void* cur_stream_, void* search_stream_,
                        vs_ticket** ticket_out) {
  VS_REQUIRE(s != nullptr && ticket_out != nullptr, "NULL pointer");
  *ticket_out = nullptr;
  VS_CUDA(cudaSetDevice(s->device));
  cudaStream_t cur = (cudaStream_t)cur_stream_, ss = (cudaStream_t)search_stream_;
  if (ss != cur) {                       // the queries (and an earlier reader of the outputs) are on `cur`
    cudaEvent_t e = pool_event(s);
    VS_REQUIRE(e != nullptr, "out of CUDA events");
    VS_CUDA(cudaEventRecord(e, cur));
    VS_CUDA(cudaStreamWaitEvent(ss, e, 0));
  }
  vs_ticket* t = new vs_ticket();
  t->stream = ss;
  const int rc = search_impl(s, q, B, k, flags, row_mask, mask_live, out_scores, out_ids, ss, t);
  if (rc) { gemm_complete(s, t); delete t; return rc; }
  t->done = pool_event(s);
  if (t->done == nullptr) { gemm_complete(s, t); delete t; set_error("out of CUDA events"); return VS_ERR_CUDA; }
  VS_CUDA(cudaEventRecord(t->done, ss));
  *ticket_out = t;
  return VS_OK;
}

int vs_exchange_result(vs_store* s, vs_ticket* ticket, const void* src_block, int64_t block_bytes, void* const* peer_dst,
                       void* const* peer_flag, int G, uint32_t step, void* counter, const void* local_flags,
                       const void* local_blocks, int B, int k, float* out_scores, int32_t* out_ids, void* xs_stream_,
                       void* cur_stream_) {
  VS_REQUIRE(s != nullptr && ticket != nullptr, "NULL pointer");
  VS_CUDA(cudaSetDevice(s->device));
  cudaStream_t xs = (cudaStream_t)xs_stream_, cur = (cudaStream_t)cur_stream_;
  const int64_t redo0 = s->retries.load() + s->fallbacks.load();
  int rc = gemm_complete(s, ticket);                 // waits for THIS search's certification counts
  const bool redone = s->retries.load() + s->fallbacks.load() != redo0;
  cudaEvent_t after = ticket->done;
  cudaStream_t ss = ticket->stream;
  delete ticket;
  if (rc) return rc;
  if (redone || after == nullptr) {                  // re-run queries wrote their rows behind `done`
    after = pool_event(s);
    VS_REQUIRE(after != nullptr, "out of CUDA events");
    VS_CUDA(cudaEventRecord(after, ss));
  }
  if (xs != ss) VS_CUDA(cudaStreamWaitEvent(xs, after, 0));
  if ((rc = vs_exchange_push(s->device, src_block, block_bytes, peer_dst, peer_flag, G, step, counter, xs))) return rc;
  if ((int64_t)G * k <= 256) {
    rc = vs_exchange_wait_merge(s->device, s->metric, local_flags, G, step, local_blocks, block_bytes / 4, B, k,
                                out_scores, out_ids, xs);
  } else {
    if ((rc = vs_exchange_wait(s->device, local_flags, G, step, xs))) return rc;
    rc = vs_merge(s->device, s->metric, (const float*)local_blocks, (const int32_t*)local_blocks + (int64_t)B * k, G, B, k,
                  block_bytes / 4, out_scores, out_ids, xs);
  }
  if (rc) return rc;
  if (cur != xs) {                                   // the caller consumes the results on its own stream
    cudaEvent_t e = pool_event(s);
    VS_REQUIRE(e != nullptr, "out of CUDA events");
    VS_CUDA(cudaEventRecord(e, xs));
    VS_CUDA(cudaStreamWaitEvent(cur, e, 0));
  }
  return VS_OK;
}

int vs_search_complete(vs_store* s, vs_ticket* ticket) {
  VS_REQUIRE(s != nullptr, "store is NULL");
  if (ticket == nullptr) return VS_OK;
  VS_CUDA(cudaSetDevice(s->device));
  const int rc = gemm_complete(s, ticket);
  delete ticket;
  return rc;
}

int vs_rescore(vs_store* s, const float* q, int B, const int32_t* cand_ids, int kc, int k,
               float* out_scores, int32_t* out_ids, void* stream_) {
  VS_REQUIRE(s != nullptr, "store is NULL");
  VS_REQUIRE(B >= 0 && kc >= 0 && k >= 0, "negative size");
  if (B == 0 || k == 0) return VS_OK;
  VS_REQUIRE(q && cand_ids && out_scores && out_ids, "NULL pointer");
  VS_REQUIRE(kc >= 1, "kc must be >= 1");
  cudaStream_t stream = (cudaStream_t)stream_;
  VS_CUDA(cudaSetDevice(s->device));
  return rescore_path(s, q, B, cand_ids, kc, std::min(k, kc), out_scores, out_ids, k, stream);
}

static bool host_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

int vs_search_host(vs_store* s, const float* q_host, int B, int k, int flags,
                   const uint32_t* row_mask_dev, int64_t mask_live, float* out_scores_host, int32_t* out_ids_host) {
  VS_REQUIRE(s != nullptr, "store is NULL");
  VS_REQUIRE(B >= 0, "B must be >= 0");
  if (B == 0 || k <= 0) return VS_OK;
  VS_REQUIRE(q_host && out_scores_host && out_ids_host, "NULL pointer");
  VS_CUDA(cudaSetDevice(s->device));
  std::lock_guard<std::mutex> g(s->host_mu);
  cudaStream_t stream = s->host_stream;
  const size_t qbytes = (size_t)B * s->dim * 4, obytes = (size_t)B * k * 4;
  Workspace ws;
  float* dq; float* ds; int32_t* di;
  ws.want(&dq, (size_t)B * s->dim);
  ws.want(&ds, (size_t)B * k);
  ws.want(&di, (size_t)B * k);
  if (int rc = ws.alloc(stream)) return rc;
  // Page-locked caller buffers are copied from / to directly; pageable ones go through the store's
  // pinned staging buffers when they fit (true async copies), else through the driver's staging.
  const bool q_direct = host_pinned(q_host) || qbytes > s->pinned_bytes;
  const bool o_direct = (host_pinned(out_scores_host) && host_pinned(out_ids_host)) || 2 * obytes > s->pinned_bytes;
  if (q_direct) {
    VS_CUDA(cudaMemcpyAsync(dq, q_host, qbytes, cudaMemcpyHostToDevice, stream));
  } else {
    memcpy(s->pinned_in, q_host, qbytes);
    VS_CUDA(cudaMemcpyAsync(dq, s->pinned_in, qbytes, cudaMemcpyHostToDevice, stream));
  }
  // The whole search and the copies back are enqueued before the host waits for anything: the
  // certification count is checked AFTER the copies are in flight, and only a search that had to
  // re-run queries (rare) copies its results a second time.
  vs_ticket ticket;
  ticket.stream = stream;
  unsigned char* po = (unsigned char*)s->pinned_out;
  auto copy_back = [&]() -> int {
    if (o_direct) {
      VS_CUDA(cudaMemcpyAsync(out_scores_host, ds, obytes, cudaMemcpyDeviceToHost, stream));
      VS_CUDA(cudaMemcpyAsync(out_ids_host, di, obytes, cudaMemcpyDeviceToHost, stream));
    } else {
      VS_CUDA(cudaMemcpyAsync(po, ds, obytes, cudaMemcpyDeviceToHost, stream));
      VS_CUDA(cudaMemcpyAsync(po + obytes, di, obytes, cudaMemcpyDeviceToHost, stream));
    }
    return VS_OK;
  };
  int rc = search_impl(s, dq, B, k, flags, row_mask_dev, mask_live, ds, di, stream, &ticket);
  if (!rc) rc = copy_back();
  const int64_t redo0 = s->retries.load() + s->fallbacks.load();
  const int rc2 = gemm_complete(s, &ticket);          // waits for the certification counts (if any)
  if (!rc) rc = rc2;
  if (!rc && s->retries.load() + s->fallbacks.load() != redo0) rc = copy_back();
  if (rc) { cudaStreamSynchronize(stream); return rc; }
  VS_CUDA(cudaStreamSynchronize(stream));
  if (!o_direct) {
    memcpy(out_scores_host, po, obytes);
    memcpy(out_ids_host, po + obytes, obytes);
  }
  return VS_OK;
}

}  // extern "C"
