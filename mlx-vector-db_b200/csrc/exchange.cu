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

}  // extern "C"
