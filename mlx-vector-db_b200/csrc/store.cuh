// Host-side state of one store: growable HBM arenas + row count.
#pragma once
#include <cuda.h>
#include <deque>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace vs {

// A device buffer that grows in place.  Default implementation: CUDA virtual memory
// management -- one address range reserved up front, physical 2 MB-granular chunks mapped
// behind it as rows arrive, so the base pointer never changes (kernels in flight and TMA
// descriptors stay valid; appends never copy the old rows, unlike the reference's
// mx.concatenate, service/optimized_vector_store.py:102).
// B200VS_ARENA=malloc selects a cudaMalloc + copy-on-grow arena instead (diagnostic).
class Arena {
 public:
  int init(int device, size_t max_bytes);
  int ensure(size_t bytes, cudaStream_t stream);   // make [0, bytes) usable
  void destroy();
  void* ptr() const { return reinterpret_cast<void*>(base_); }
  size_t mapped() const { return mapped_; }

 private:
  int device_ = 0;
  bool vmm_ = true;
  CUdeviceptr base_ = 0;
  size_t reserved_ = 0, mapped_ = 0, gran_ = 0;
  std::vector<std::pair<CUmemGenericAllocationHandle, size_t>> chunks_;
};

}  // namespace vs

struct vs_store {
  int device = 0;
  int dim = 0;        // D
  int ld = 0;         // fp32 row stride in floats (D rounded up to 4 -> 16 B aligned rows)
  int ld16 = 0;       // 16-bit shadow row stride in elements (D rounded up to 64 -> 128 B)
  int ld8 = 0;        // fp8 shadow row stride in elements = bytes (D rounded up to 128)
  int shadow8 = 0;    // keep an e4m3 shadow copy (cosine only): 16 * x/||x||, recall-reported searches
  int metric = 0;
  int shadow = 0;
  int num_sms = 148;
  int64_t max_rows = 0;
  std::atomic<int64_t> count{0};
  std::atomic<int64_t> fallbacks{0};   // queries re-run through the exact scan by the GEMM path
  std::atomic<int64_t> retries{0};     // queries retried with 4x the candidates first
  std::mutex mu;                 // one writer at a time (append / reset)
  cudaEvent_t append_done = nullptr;   // recorded after the last append's kernels
  std::atomic<cudaStream_t> append_stream{nullptr};   // read by concurrent searches
  // certification slots of the GEMM path (gemm_topk.cu): pinned count + event + device list
  struct CertSlot { cudaEvent_t done = nullptr; int* h_bad = nullptr; int32_t* d_bad = nullptr; bool busy = false; };
  std::mutex slot_mu;
  std::deque<CertSlot> slots;    // deque: growing never moves a slot another thread is using
  // K3 workspaces (gemm_topk.cu): device blocks that are recycled instead of going through the
  // stream-ordered allocator per search.  A block is handed to an enqueue on the stream it was last
  // used on (stream order makes that safe) or once its event has completed (another stream).
  struct WsBlock { void* ptr = nullptr; size_t bytes = 0; cudaStream_t stream = nullptr; cudaEvent_t ev = nullptr; bool busy = false; };
  std::mutex ws_mu;
  std::deque<WsBlock> ws_blocks;
  // cross-stream ordering events of vs_search_submit_on / vs_exchange_result, recycled round-robin
  std::mutex ev_mu;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_next = 0;
  // vs_search_host: private stream + pinned staging, serialised by host_mu
  std::mutex host_mu;
  cudaStream_t host_stream = nullptr;
  void* pinned_in = nullptr;
  void* pinned_out = nullptr;
  size_t pinned_bytes = 0;
  vs::Arena rows;                // fp32 master, (N, ld)
  vs::Arena norms;               // max(||x||, 1e-8), (N,)
  vs::Arena sqnorms;             // ||x||^2, (N,)
  vs::Arena shadow_rows;         // 16-bit, (N, ld16): fp16 of x/max(||x||,1e-8) for cosine, bf16 of x otherwise
  vs::Arena shadow8_rows;        // e4m3, (N, ld8): 16 * x/max(||x||,1e-8)
  uint32_t* bounds = nullptr;    // device: float bits of max ||v - bf16(v)||, max ||bf16(v)||
  vs::Arena gids;                // int32 global id per local row (row-sharded stores only)
  bool mapped = false;           // true once an append supplied global ids
  const int32_t* id_map() const { return mapped ? (const int32_t*)gids.ptr() : nullptr; }
};
