// Candidate exchange between the shards of one box over NVLink peer memory (new: the reference
// is single-device).
//
// Every rank owns an exchange buffer that all ranks of the box have mapped (symmetric memory; the
// host side, b200vs/sharded.py, obtains the peer pointers from torch.distributed's symmetric-memory
// rendezvous).  After its local search a rank PUSHES its packed (2, B, k) candidate block straight
// into slot [step % depth][rank] of every rank's buffer with plain 128-bit stores over NVLink
// and then publishes the step number in that rank's flag word for it.  The receiving side waits
// (a one-warp kernel polling ITS OWN memory) until all ranks' flags show the step and runs the K4
// merge over its local copy.  Compared with the NCCL all-gather this replaces, there is no
// rendezvous between the ranks' kernels, no staging copy and one host call per step; the 80 KB a
// rank sends per batch of 1024 queries cost under a microsecond of link time.
#include <algorithm>
#include <cstdint>
#include "common.cuh"
#include "exchange.cuh"

namespace vs {

constexpr int kMaxPeers = 16;
struct PeerPtrs {
  void* dst[kMaxPeers];        // where this rank's block goes in every rank's buffer
  uint32_t* flag[kMaxPeers];   // this rank's flag word in every rank's buffer
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// src: `vec` 16-byte words.  The last block to finish publishes the step.
__global__ void __launch_bounds__(256)
exchange_push_kernel(const uint4* __restrict__ src, int64_t vec, PeerPtrs peers, int G, uint32_t step,
                     uint32_t* __restrict__ counter) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < vec; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = src[i];
    for (int g = 0; g < G; ++g) reinterpret_cast<uint4*>(peers.dst[g])[i] = v;
  }
  __threadfence_system();                  // this thread's peer stores before the flag
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(counter, 1u) == gridDim.x - 1) {
      *counter = 0;                        // ready for the next push (pushes are stream-ordered)
      __threadfence_system();
      for (int g = 0; g < G; ++g) st_release_sys(peers.flag[g], step);
    }
  }
}

// flags: the G flag words of this step's slot in THIS rank's buffer.  Steps only grow, so a
// flag at or above `step` means that rank's block of this step has landed.
__global__ void exchange_wait_kernel(const uint32_t* __restrict__ flags, int G, uint32_t step) {
  const int g = threadIdx.x;
  if (g >= G) return;
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(flags + g) - step) < 0) {
    __nanosleep(200);
    // a rank that never arrives (crashed process) must not hang the box: ~10 s, then fail loudly
    if (clock64() - t0 > 20000000000ll) __trap();
  }
}

// Wait for every rank's block of this step, then merge the G x k candidates of each query into its
// global top-k: one warp per query, candidates in registers, ranks by counting over warp shuffles
// (key desc, id asc -- K4's order).  No shared memory and few registers, so these CTAs co-reside
// with the next search's GEMM CTAs (which own nearly all of an SM's shared memory) instead of
// waiting for a gap: the exchange is off the critical path.  n = G * k <= 32 * kMergeSlots.
constexpr int kMergeSlots = 8;
constexpr int kMergeWarps = 4;
__global__ void __launch_bounds__(kMergeWarps * 32)
exchange_wait_merge_kernel(const uint32_t* __restrict__ flags, int G, uint32_t step, const int32_t* __restrict__ blocks,
                           int64_t block_words, int B, int k, int l2, float* __restrict__ out_s,
                           int32_t* __restrict__ out_i) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * kMergeWarps + (threadIdx.x >> 5);
  if (lane < G) {
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(flags + lane) - step) < 0) {
      __nanosleep(200);
      if (clock64() - t0 > 20000000000ll) __trap();      // a rank that never arrives: fail loudly (~10 s)
    }
  }
  __syncwarp();
  if (b >= B) return;
  const int n = G * k;
  float key[kMergeSlots];
  int id[kMergeSlots];
#pragma unroll
  for (int r = 0; r < kMergeSlots; ++r) {
    const int c = r * 32 + lane;                          // candidate c = (rank g, position j)
    key[r] = VS_NEG_INF;
    id[r] = VS_ID_SENTINEL;
    if (c < n) {
      const int g = c / k, j = c - g * k;
      const int32_t* blk = blocks + (int64_t)g * block_words;     // (2, B, k): scores then ids
      const int ci = __ldcg(blk + (int64_t)B * k + (int64_t)b * k + j);
      if (ci >= 0) {
        const float sc = __int_as_float(__ldcg(blk + (int64_t)b * k + j));
        key[r] = l2 ? -sc : sc;
        id[r] = ci;
      }
    }
  }
  int rank[kMergeSlots];
#pragma unroll
  for (int r = 0; r < kMergeSlots; ++r) rank[r] = 0;
  const int slots = (n + 31) >> 5;
  for (int r2 = 0; r2 < slots; ++r2) {
    float ok = VS_NEG_INF; int oi = VS_ID_SENTINEL;
#pragma unroll
    for (int r = 0; r < kMergeSlots; ++r) if (r == r2) { ok = key[r]; oi = id[r]; }
    for (int l = 0; l < 32; ++l) {
      const float k2 = __shfl_sync(0xffffffffu, ok, l);
      const int i2 = __shfl_sync(0xffffffffu, oi, l);
#pragma unroll
      for (int r = 0; r < kMergeSlots; ++r) rank[r] += better(k2, i2, key[r], id[r]);
    }
  }
  float* os = out_s + (int64_t)b * k;
  int32_t* oi_ = out_i + (int64_t)b * k;
  int valid = 0;
#pragma unroll
  for (int r = 0; r < kMergeSlots; ++r) {
    const bool live = id[r] != VS_ID_SENTINEL;
    valid += __popc(__ballot_sync(0xffffffffu, live));
    if (live && rank[r] < k) { os[rank[r]] = l2 ? -key[r] : key[r]; oi_[rank[r]] = id[r]; }
  }
  for (int j = valid + lane; j < k; j += 32) { os[j] = 0.f; oi_[j] = -1; }
}

}  // namespace vs

using namespace vs;

extern "C" {

int vs_exchange_push(int device, const void* src, int64_t bytes, void* const* peer_dst, void* const* peer_flag,
                     int G, uint32_t step, void* counter, void* stream_) {
  VS_REQUIRE(src && peer_dst && peer_flag && counter, "NULL pointer");
  VS_REQUIRE(G >= 1 && G <= kMaxPeers, "1 <= G <= 16");
  VS_REQUIRE(bytes > 0 && bytes % 16 == 0, "bytes must be a positive multiple of 16");
  VS_CUDA(cudaSetDevice(device));
  PeerPtrs pp = {};
  for (int g = 0; g < G; ++g) { pp.dst[g] = peer_dst[g]; pp.flag[g] = (uint32_t*)peer_flag[g]; }
  const int64_t vec = bytes / 16;
  const int blocks = (int)std::min<int64_t>((vec + 255) / 256, 32);
  exchange_push_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>((const uint4*)src, vec, pp, G, step, (uint32_t*)counter);
  count_launch();
  VS_CHECK_LAUNCH();
  return VS_OK;
}

int vs_exchange_wait(int device, const void* flags, int G, uint32_t step, void* stream_) {
  VS_REQUIRE(flags != nullptr, "NULL pointer");
  VS_REQUIRE(G >= 1 && G <= kMaxPeers, "1 <= G <= 16");
  VS_CUDA(cudaSetDevice(device));
  exchange_wait_kernel<<<1, 32, 0, (cudaStream_t)stream_>>>((const uint32_t*)flags, G, step);
  count_launch();
  VS_CHECK_LAUNCH();
  return VS_OK;
}

int vs_exchange_wait_merge(int device, int metric, const void* flags, int G, uint32_t step, const void* blocks,
                           int64_t block_words, int B, int k, float* out_scores, int32_t* out_ids, void* stream_) {
  VS_REQUIRE(flags && blocks && out_scores && out_ids, "NULL pointer");
  VS_REQUIRE(G >= 1 && G <= kMaxPeers, "1 <= G <= 16");
  VS_REQUIRE(B >= 1 && k >= 1, "B, k must be >= 1");
  VS_REQUIRE((int64_t)G * k <= 32 * kMergeSlots, "G * k must be <= 256 (use vs_exchange_wait + vs_merge beyond)");
  VS_CUDA(cudaSetDevice(device));
  exchange_wait_merge_kernel<<<(B + kMergeWarps - 1) / kMergeWarps, kMergeWarps * 32, 0, (cudaStream_t)stream_>>>(
      (const uint32_t*)flags, G, step, (const int32_t*)blocks, block_words, B, k, metric == VS_METRIC_EUCLIDEAN ? 1 : 0,
      out_scores, out_ids);
  count_launch();
  VS_CHECK_LAUNCH();
  return VS_OK;
}

}  // extern "C"
