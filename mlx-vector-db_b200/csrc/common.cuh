// Shared device/host helpers for libb200vs (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <atomic>
#include <string>

#include "../../include/b200vs.h"

namespace vs {

// ---------------------------------------------------------------- host side
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
extern std::atomic<int64_t> g_launches;
// B200VS_TRACE=1 (diagnostic): every kernel launch is followed by a device synchronisation and a
// line on stderr naming the call site, so a kernel that never returns can be identified.
void trace_launch(const char* file, int line);
extern int g_trace;                       // -1 = not read yet
inline void count_launch_at(const char* file, int line) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (g_trace != 0) trace_launch(file, line);
}
#define count_launch() ::vs::count_launch_at(__FILE__, __LINE__)

// Optional CUDA-event bracket around the dominant kernels (vs_profile / vs_profile_read):
// bench.py uses it to time the scan / GEMM kernel alone on the stream it is launched on.
enum { kProfScan = 0, kProfGemm = 1, kProfKinds = 2 };
struct ProfScope {
  ProfScope(int kind, cudaStream_t stream);
  ~ProfScope();
  cudaStream_t stream_;
  cudaEvent_t stop_ = nullptr;
};

#define VS_CUDA(expr)                                                        \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) return ::vs::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define VS_CHECK_LAUNCH()                                                    \
  do {                                                                       \
    cudaError_t _e = cudaGetLastError();                                     \
    if (_e != cudaSuccess) return ::vs::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

#define VS_REQUIRE(cond, msg)                                                \
  do {                                                                       \
    if (!(cond)) { ::vs::set_error(std::string("invalid argument: ") + (msg)); return VS_ERR_INVALID; } \
  } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// --------------------------------------------------------------- device side
#define VS_NEG_INF (__int_as_float(0xff800000))
#define VS_ID_SENTINEL 0x7fffffff

// Monotone float <-> uint encoding so atomicMax on uint orders like the floats.
__device__ __forceinline__ uint32_t enc_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_key(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
#define VS_ENC_NEG_INF 0x007fffffu   // enc_key(-inf)

// Total order of results: larger key first, equal keys -> lower id first
// (stable argsort of the negated scores, service/optimized_vector_store.py:176-181).
__device__ __forceinline__ bool better(float ka, int ia, float kb, int ib) {
  return (ka > kb) || (ka == kb && ia < ib);
}

// The one fp32 accumulation order used by every exact kernel (scan and rescoring), so
// that both produce bit-identical scores: per lane an fma chain over its float4 columns in
// ascending order, then an xor-butterfly 16,8,4,2,1 across lanes.
__device__ __forceinline__ float dot4_acc(float acc, const float4& x, const float4& q) {
  acc = fmaf(x.x, q.x, acc);
  acc = fmaf(x.y, q.y, acc);
  acc = fmaf(x.z, q.z, acc);
  acc = fmaf(x.w, q.w, acc);
  return acc;
}
__device__ __forceinline__ float sqdiff4_acc(float acc, const float4& x, const float4& q) {
  float d;
  d = x.x - q.x; acc = fmaf(d, d, acc);
  d = x.y - q.y; acc = fmaf(d, d, acc);
  d = x.z - q.z; acc = fmaf(d, d, acc);
  d = x.w - q.w; acc = fmaf(d, d, acc);
  return acc;
}
__device__ __forceinline__ float warp_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v;
}

// streaming 128-bit load: read-only path, do not allocate in L1
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// final score from the reduced accumulator (metric-specific epilogue); returns the KEY
// (larger = better) -- euclidean keys are negated distances.
template <int METRIC>
__device__ __forceinline__ float key_from_acc(float acc, float clamped_norm) {
  if (METRIC == VS_METRIC_COSINE) return acc / clamped_norm;     // query pre-normalised
  if (METRIC == VS_METRIC_EUCLIDEAN) return -sqrtf(acc);
  return acc;
}

}  // namespace vs
