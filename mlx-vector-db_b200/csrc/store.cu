// Store lifetime, HBM arenas and K1 append_norm.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include "store.cuh"
#include "gemm_topk.cuh"

namespace vs {

// ------------------------------------------------------------------ errors
static thread_local std::string t_error;
std::atomic<int64_t> g_launches{0};
int g_trace = -1;
void trace_launch(const char* file, int line) {
  if (g_trace < 0) { const char* e = getenv("B200VS_TRACE"); g_trace = (e && *e == '1') ? 1 : 0; }
  if (g_trace == 0) return;
  const char* base = strrchr(file, '/');
  fprintf(stderr, "[b200vs trace] launch #%lld at %s:%d ...", (long long)g_launches.load(), base ? base + 1 : file, line);
  fflush(stderr);
  const cudaError_t e = cudaDeviceSynchronize();
  fprintf(stderr, " %s\n", e == cudaSuccess ? "done" : cudaGetErrorString(e));
  fflush(stderr);
}

void set_error(const std::string& msg) { t_error = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e),
           file, line, what);
  t_error = buf;
  cudaGetLastError();  // clear the sticky-less error state
  return e == cudaErrorMemoryAllocation ? VS_ERR_OOM : VS_ERR_CUDA;
}

// --------------------------------------------------------------- profiling
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
struct ProfPair { cudaEvent_t a, b; };
static std::vector<ProfPair> g_prof[kProfKinds];

static std::vector<ProfPair> g_prof_pool;   // recycled event pairs (creating events costs host time)

ProfScope::ProfScope(int kind, cudaStream_t stream) : stream_(stream) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfPair p;
  std::lock_guard<std::mutex> g(g_prof_mu);
  if (!g_prof_pool.empty()) {
    p = g_prof_pool.back();
    g_prof_pool.pop_back();
  } else {
    if (cudaEventCreate(&p.a) != cudaSuccess) return;
    if (cudaEventCreate(&p.b) != cudaSuccess) { cudaEventDestroy(p.a); return; }
  }
  cudaEventRecord(p.a, stream);
  stop_ = p.b;
  g_prof[kind].push_back(p);
}
ProfScope::~ProfScope() {
  if (stop_) cudaEventRecord(stop_, stream_);
}

// --------------------------------------------------- driver API via cudart
// libcuda is not linked: the library must load (and export its symbols) on a box without a
// driver; the entry points are resolved at first use through the runtime.
struct DriverApi {
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*MemAddressFree)(CUdeviceptr, size_t);
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*MemRelease)(CUmemGenericAllocationHandle);
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*MemUnmap)(CUdeviceptr, size_t);
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*GetErrorString)(CUresult, const char**);
  bool ok = false;
};

static int load_driver(DriverApi** out) {
  static DriverApi api;
  static std::once_flag once;
  static std::string why;
  std::call_once(once, [] {
    struct { const char* name; void** slot; } syms[] = {
        {"cuMemGetAllocationGranularity", (void**)&api.MemGetAllocationGranularity},
        {"cuMemAddressReserve", (void**)&api.MemAddressReserve},
        {"cuMemAddressFree", (void**)&api.MemAddressFree},
        {"cuMemCreate", (void**)&api.MemCreate},
        {"cuMemRelease", (void**)&api.MemRelease},
        {"cuMemMap", (void**)&api.MemMap},
        {"cuMemUnmap", (void**)&api.MemUnmap},
        {"cuMemSetAccess", (void**)&api.MemSetAccess},
        {"cuGetErrorString", (void**)&api.GetErrorString},
    };
    for (auto& s : syms) {
      cudaDriverEntryPointQueryResult q;
      cudaError_t e = cudaGetDriverEntryPoint(s.name, s.slot, cudaEnableDefault, &q);
      if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || *s.slot == nullptr) {
        why = std::string("driver entry point not found: ") + s.name;
        cudaGetLastError();
        return;
      }
    }
    api.ok = true;
  });
  if (!api.ok) { set_error(why); return VS_ERR_CUDA; }
  *out = &api;
  return VS_OK;
}

static int cu_fail(DriverApi* d, CUresult r, const char* what) {
  const char* s = nullptr;
  d->GetErrorString(r, &s);
  set_error(std::string("driver error in ") + what + ": " + (s ? s : "?"));
  return r == CUDA_ERROR_OUT_OF_MEMORY ? VS_ERR_OOM : VS_ERR_CUDA;
}

// ------------------------------------------------------------------- Arena
int Arena::init(int device, size_t max_bytes) {
  device_ = device;
  const char* mode = getenv("B200VS_ARENA");
  vmm_ = !(mode && strcmp(mode, "malloc") == 0);
  if (!vmm_) { reserved_ = max_bytes; return VS_OK; }
  DriverApi* d;
  if (int rc = load_driver(&d)) return rc;
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  CUresult r = d->MemGetAllocationGranularity(&gran_, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
  if (r != CUDA_SUCCESS) return cu_fail(d, r, "cuMemGetAllocationGranularity");
  if (gran_ == 0) gran_ = 2u << 20;
  reserved_ = (size_t)round_up((int64_t)max_bytes, (int64_t)gran_);
  r = d->MemAddressReserve(&base_, reserved_, 0, 0, 0);
  if (r != CUDA_SUCCESS) return cu_fail(d, r, "cuMemAddressReserve");
  return VS_OK;
}

int Arena::ensure(size_t bytes, cudaStream_t stream) {
  if (bytes <= mapped_) return VS_OK;
  if (bytes > reserved_) { set_error("store is full (max_rows reached)"); return VS_ERR_OOM; }
  if (!vmm_) {
    size_t want = bytes > mapped_ * 2 ? bytes : mapped_ * 2;
    if (want > reserved_) want = reserved_;
    void* p = nullptr;
    VS_CUDA(cudaMalloc(&p, want));
    if (base_) {
      VS_CUDA(cudaMemcpyAsync(p, (void*)base_, mapped_, cudaMemcpyDeviceToDevice, stream));
      VS_CUDA(cudaStreamSynchronize(stream));
      VS_CUDA(cudaFree((void*)base_));
    }
    base_ = (CUdeviceptr)p;
    mapped_ = want;
    return VS_OK;
  }
  DriverApi* d;
  if (int rc = load_driver(&d)) return rc;
  // grow by at least 25 % (and 64 MB) so a stream of small appends maps few chunks
  size_t want = bytes;
  size_t geo = mapped_ + mapped_ / 4;
  if (geo > want) want = geo;
  if (want < mapped_ + (64u << 20)) want = mapped_ + (64u << 20);
  want = (size_t)round_up((int64_t)want, (int64_t)gran_);
  if (want > reserved_) want = reserved_;
  size_t chunk = want - mapped_;
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device_;
  CUmemGenericAllocationHandle h;
  CUresult r = d->MemCreate(&h, chunk, &prop, 0);
  if (r != CUDA_SUCCESS && want > (size_t)round_up((int64_t)bytes, (int64_t)gran_)) {
    // geometric head-room did not fit: retry with exactly what is needed
    want = (size_t)round_up((int64_t)bytes, (int64_t)gran_);
    chunk = want - mapped_;
    r = d->MemCreate(&h, chunk, &prop, 0);
  }
  if (r != CUDA_SUCCESS) return cu_fail(d, r, "cuMemCreate");
  r = d->MemMap(base_ + mapped_, chunk, 0, h, 0);
  if (r != CUDA_SUCCESS) { d->MemRelease(h); return cu_fail(d, r, "cuMemMap"); }
  CUmemAccessDesc acc = {};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device_;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = d->MemSetAccess(base_ + mapped_, chunk, &acc, 1);
  if (r != CUDA_SUCCESS) {
    d->MemUnmap(base_ + mapped_, chunk);
    d->MemRelease(h);
    return cu_fail(d, r, "cuMemSetAccess");
  }
  chunks_.push_back({h, chunk});
  mapped_ = want;
  return VS_OK;
}

void Arena::destroy() {
  if (!vmm_) {
    if (base_) cudaFree((void*)base_);
  } else if (base_) {
    DriverApi* d;
    if (load_driver(&d) == VS_OK) {
      size_t off = 0;
      for (auto& c : chunks_) {
        d->MemUnmap(base_ + off, c.second);
        d->MemRelease(c.first);
        off += c.second;
      }
      d->MemAddressFree(base_, reserved_);
    }
  }
  chunks_.clear();
  base_ = 0;
  mapped_ = reserved_ = 0;
}

// --------------------------------------------------------- K1 append_norm
// One warp per appended row: copy into the master arena (padded stride), row norm with the
// 1e-8 clamp of service/optimized_vector_store.py:36-38, ||x||^2, optional 16-bit shadow row
// (fp16 of x/||x|| for cosine, bf16 of x otherwise) and the two maxima the GEMM path's
// certification needs: max ||v - v^|| and max ||v^|| over all shadow rows (kept per warp, one
// atomicMax per warp at the end).  VEC: 128-bit accesses when dim % 4 == 0.
__device__ __forceinline__ uint32_t pack16(float a, float b, bool fp16, float& ra, float& rb) {
  if (fp16) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    ra = f.x; rb = f.y;
    return *reinterpret_cast<const uint32_t*>(&h);
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  ra = __bfloat162float(h.x); rb = __bfloat162float(h.y);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// 16-bit shadow of 8 consecutive prepared components (fp16 for cosine, bf16 otherwise) + the
// squared rounding error and squared length of the rounded values
__device__ __forceinline__ uint4 shadow8(const float4& a, const float4& b, bool fp16, float& e2, float& s2) {
  float q0, q1, q2, q3, q4, q5, q6, q7, t;
  uint4 o;
  o.x = pack16(a.x, a.y, fp16, q0, q1);
  o.y = pack16(a.z, a.w, fp16, q2, q3);
  o.z = pack16(b.x, b.y, fp16, q4, q5);
  o.w = pack16(b.z, b.w, fp16, q6, q7);
  t = a.x - q0; e2 = fmaf(t, t, e2); s2 = fmaf(q0, q0, s2);
  t = a.y - q1; e2 = fmaf(t, t, e2); s2 = fmaf(q1, q1, s2);
  t = a.z - q2; e2 = fmaf(t, t, e2); s2 = fmaf(q2, q2, s2);
  t = a.w - q3; e2 = fmaf(t, t, e2); s2 = fmaf(q3, q3, s2);
  t = b.x - q4; e2 = fmaf(t, t, e2); s2 = fmaf(q4, q4, s2);
  t = b.y - q5; e2 = fmaf(t, t, e2); s2 = fmaf(q5, q5, s2);
  t = b.z - q6; e2 = fmaf(t, t, e2); s2 = fmaf(q6, q6, s2);
  t = b.w - q7; e2 = fmaf(t, t, e2); s2 = fmaf(q7, q7, s2);
  return o;
}
__device__ __forceinline__ float4 scale4(const float4& v, float r) { return make_float4(v.x * r, v.y * r, v.z * r, v.w * r); }

// Rows are read ONCE into registers (no second pass for the shadow) and a row is spread over a
// lane GROUP of G lanes, each owning chunks of 8 consecutive floats (two 128-bit loads, one
// 128-bit shadow store), so that 32 / G rows per warp reduce together in log2(G) shuffle steps:
//   G = 8,  C = 2: rows of up to 128 floats, 4 rows per warp step (x kAppendSteps steps in flight)
//   G = 32, C = 2 / 4 / 8: rows of up to 512 / 1024 / 2048 floats, one row per warp step
// The shadow of a cosine store holds x * (1 / ||x||) (one reciprocal per row); the certification
// bound is measured on exactly that value (+2e-7 for its distance to x / ||x||).
constexpr int kAppendSteps = 2;
template <int G, int C>
__global__ void __launch_bounds__(256)
append_norm_reg_kernel(const float* __restrict__ src, int64_t src_ld, float* __restrict__ rows,
                       int ld, int dim, int64_t n0, int64_t m, float* __restrict__ norms,
                       float* __restrict__ sqnorms, __nv_bfloat16* __restrict__ shadow, int ld16,
                       int normalize_shadow, int32_t* __restrict__ gids, int64_t gid0,
                       uint32_t* __restrict__ bounds) {
  constexpr int RPW = 32 / G;                       // rows per warp step
  const int lane = threadIdx.x & 31;
  const int sub = lane % G, grp = lane / G;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const bool fp16 = normalize_shadow != 0;
  float e_max = 0.f, s_max = 0.f;
  for (int64_t r0 = warp * (RPW * kAppendSteps); r0 < m; r0 += nwarps * (RPW * kAppendSteps)) {
    float4 v[kAppendSteps][C][2];
#pragma unroll
    for (int st = 0; st < kAppendSteps; ++st) {
      const int64_t r = r0 + st * RPW + grp;
      const float4* s4 = reinterpret_cast<const float4*>(src + (r < m ? r : r0) * src_ld);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int e0 = 8 * (sub + G * c);           // first float of this lane's chunk
        v[st][c][0] = e0 < dim ? ldg_stream(s4 + (e0 >> 2)) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[st][c][1] = e0 + 4 < dim ? ldg_stream(s4 + (e0 >> 2) + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int st = 0; st < kAppendSteps; ++st) {
      const int64_t r = r0 + st * RPW + grp;
      const bool live = r < m;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4& x = v[st][c][h];
          acc = fmaf(x.x, x.x, acc); acc = fmaf(x.y, x.y, acc); acc = fmaf(x.z, x.z, acc); acc = fmaf(x.w, x.w, acc);
        }
#pragma unroll
      for (int off = G / 2; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      const float nrm = fmaxf(sqrtf(acc), 1e-8f);
      const float rcp = fp16 ? 1.f / nrm : 1.f;
      float* d = rows + (n0 + r) * (int64_t)ld;
      const bool copy = live && d != src + r * src_ld;
      float e2 = 0.f, s2 = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int e0 = 8 * (sub + G * c);
        if (copy && e0 < dim) reinterpret_cast<float4*>(d)[e0 >> 2] = v[st][c][0];
        if (copy && e0 + 4 < dim) reinterpret_cast<float4*>(d)[(e0 >> 2) + 1] = v[st][c][1];
        if (shadow != nullptr && e0 < ld16) {       // zero padded up to ld16 (loads beyond dim gave 0)
          const uint4 o = shadow8(scale4(v[st][c][0], rcp), scale4(v[st][c][1], rcp), fp16, e2, s2);
          if (live) *reinterpret_cast<uint4*>(shadow + (n0 + r) * (int64_t)ld16 + e0) = o;
        }
      }
      if (live && sub == 0) {
        norms[n0 + r] = nrm;
        sqnorms[n0 + r] = acc;
        if (gids != nullptr) gids[n0 + r] = (int32_t)(gid0 + r);
      }
      if (shadow != nullptr) {
#pragma unroll
        for (int off = G / 2; off; off >>= 1) {
          e2 += __shfl_xor_sync(0xffffffffu, e2, off);
          s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        }
        if (live) { e_max = fmaxf(e_max, sqrtf(e2)); s_max = fmaxf(s_max, sqrtf(s2)); }
      }
    }
  }
  if (shadow != nullptr && sub == 0) {   // non-negative floats order like their bit patterns
    atomicMax(bounds + 0, __float_as_uint(e_max * 1.00001f + 2e-7f));
    atomicMax(bounds + 1, __float_as_uint(s_max * 1.00001f));
  }
}

template <bool VEC>
__global__ void __launch_bounds__(256)
append_norm_kernel(const float* __restrict__ src, int64_t src_ld, float* __restrict__ rows,
                   int ld, int dim, int64_t n0, int64_t m, float* __restrict__ norms,
                   float* __restrict__ sqnorms, __nv_bfloat16* __restrict__ shadow, int ld16,
                   int normalize_shadow, int32_t* __restrict__ gids, int64_t gid0,
                   uint32_t* __restrict__ bounds) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float e_max = 0.f, s_max = 0.f;
  for (int64_t r = warp; r < m; r += nwarps) {
    const float* s = src + r * src_ld;
    float* d = rows + (n0 + r) * (int64_t)ld;
    const bool in_place = (s == d);
    float acc = 0.f;
    if (VEC) {
      const float4* s4 = reinterpret_cast<const float4*>(s);
      float4* d4 = reinterpret_cast<float4*>(d);
      const int nvec = dim >> 2;
      // same per-lane order as the scalar path is not required: norms are defined by this kernel
      for (int c = lane; c < nvec; c += 32) {
        const float4 v = s4[c];
        acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc);
        acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
        if (!in_place) d4[c] = v;
      }
    } else {
      for (int c = lane; c < ld; c += 32) {
        float v = c < dim ? s[c] : 0.f;
        acc = fmaf(v, v, acc);
        if (!in_place) d[c] = v;
      }
    }
    const float tot = warp_sum(acc);
    const float nrm = fmaxf(sqrtf(tot), 1e-8f);
    if (lane == 0) {
      norms[n0 + r] = nrm;
      sqnorms[n0 + r] = tot;
      if (gids != nullptr) gids[n0 + r] = (int32_t)(gid0 + r);
    }
    if (shadow != nullptr) {
      // cosine: unit-norm rows fit fp16's range and keep 11 significant bits (8x smaller
      // rounding error than bf16 -> far tighter certification); raw rows keep bf16's range
      const bool fp16 = normalize_shadow != 0;
      float e2 = 0.f, s2 = 0.f;
      if (VEC) {
        const float4* s4 = reinterpret_cast<const float4*>(s);
        uint2* sh2 = reinterpret_cast<uint2*>(shadow + (n0 + r) * (int64_t)ld16);
        const int nvec = dim >> 2;
        for (int c = lane; c < (ld16 >> 2); c += 32) {
          float4 v = c < nvec ? s4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
          if (fp16) { v.x = v.x / nrm; v.y = v.y / nrm; v.z = v.z / nrm; v.w = v.w / nrm; }
          float r0, r1, r2, r3;
          uint2 o;
          o.x = pack16(v.x, v.y, fp16, r0, r1);
          o.y = pack16(v.z, v.w, fp16, r2, r3);
          sh2[c] = o;
          float t;
          t = v.x - r0; e2 = fmaf(t, t, e2); s2 = fmaf(r0, r0, s2);
          t = v.y - r1; e2 = fmaf(t, t, e2); s2 = fmaf(r1, r1, s2);
          t = v.z - r2; e2 = fmaf(t, t, e2); s2 = fmaf(r2, r2, s2);
          t = v.w - r3; e2 = fmaf(t, t, e2); s2 = fmaf(r3, r3, s2);
        }
      } else {
        __nv_bfloat16* sh = shadow + (n0 + r) * (int64_t)ld16;
        for (int c = lane; c < ld16; c += 32) {
          float v = c < dim ? s[c] : 0.f;
          if (fp16) v = v / nrm;
          float vb;
          if (fp16) {
            const __half h = __float2half_rn(v);
            reinterpret_cast<__half*>(sh)[c] = h;
            vb = __half2float(h);
          } else {
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            sh[c] = h;
            vb = __bfloat162float(h);
          }
          const float dlt = v - vb;
          e2 = fmaf(dlt, dlt, e2);
          s2 = fmaf(vb, vb, s2);
        }
      }
      e_max = fmaxf(e_max, sqrtf(warp_sum(e2)));
      s_max = fmaxf(s_max, sqrtf(warp_sum(s2)));
    }
  }
  if (shadow != nullptr && lane == 0) {   // non-negative floats order like their bit patterns
    atomicMax(bounds + 0, __float_as_uint(e_max * 1.00001f));
    atomicMax(bounds + 1, __float_as_uint(s_max * 1.00001f));
  }
}

// e4m3 shadow rows for the fp8 GEMM candidates: 16 * x / max(||x||, 1e-8) (the factor 16 keeps
// unit-norm components of D up to a few thousand inside e4m3's normal range), 4 elements per lane
__global__ void __launch_bounds__(256)
shadow8_kernel(const float* __restrict__ rows, int ld, int dim, const float* __restrict__ norms, int64_t n0,
               int64_t m, unsigned char* __restrict__ out, int ld8) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp; r < m; r += nwarps) {
    const float* x = rows + (n0 + r) * (int64_t)ld;
    const float sc = 16.f / norms[n0 + r];
    uint32_t* o = reinterpret_cast<uint32_t*>(out + (n0 + r) * (int64_t)ld8);
    for (int c4 = lane; c4 < (ld8 >> 2); c4 += 32) {
      uint32_t w = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = 4 * c4 + j;
        const float v = c < dim ? x[c] * sc : 0.f;
        w |= (uint32_t)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3) << (8 * j);
      }
      o[c4] = w;
    }
  }
}

}  // namespace vs

using namespace vs;

// ------------------------------------------------------------------- C-ABI
extern "C" {

const char* vs_last_error(void) { return t_error.c_str(); }
const char* vs_version(void) { return "b200vs 0.1 (sm_100a)"; }
int64_t vs_launch_count(void) { return g_launches.load(); }

int vs_profile(int enable) {
  g_prof_on.store(enable ? 1 : 0);
  return VS_OK;
}

int vs_profile_read(int kind, double* total_ms, int64_t* launches) {
  VS_REQUIRE(kind >= 0 && kind < kProfKinds, "unknown profile kind");
  VS_REQUIRE(total_ms && launches, "NULL pointer");
  std::lock_guard<std::mutex> g(g_prof_mu);
  double sum = 0.0;
  int64_t n = 0;
  for (auto& p : g_prof[kind]) {
    float ms = 0.f;
    if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      sum += ms;
      ++n;
    }
    g_prof_pool.push_back(p);
  }
  cudaGetLastError();
  g_prof[kind].clear();
  *total_ms = sum;
  *launches = n;
  return VS_OK;
}

int vs_create(int device, int dim, int metric, int shadow, int64_t max_rows, vs_store** out) {
  VS_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  VS_REQUIRE(dim > 0 && dim <= 65536, "dim must be in [1, 65536]");
  VS_REQUIRE(metric >= 0 && metric <= 2, "metric must be 0 (cosine), 1 (euclidean) or 2 (dot)");
  VS_REQUIRE(shadow >= 0 && shadow <= (VS_SHADOW_BF16 | VS_SHADOW_FP8), "shadow must be a VS_SHADOW_* mask");
  VS_REQUIRE(!(shadow & VS_SHADOW_FP8) || metric == VS_METRIC_COSINE, "the fp8 shadow copy serves cosine only");
  VS_REQUIRE(max_rows >= 0 && max_rows < (int64_t)0x7fffffff, "max_rows must be < 2^31");
  int ndev = 0;
  VS_CUDA(cudaGetDeviceCount(&ndev));
  VS_REQUIRE(device >= 0 && device < ndev, "no such CUDA device");
  VS_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  VS_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("libb200vs is built for sm_100a (B200) only; device is sm_" +
              std::to_string(prop.major) + std::to_string(prop.minor));
    return VS_ERR_STATE;
  }
  vs_store* s = new vs_store();
  s->device = device;
  s->dim = dim;
  s->ld = (int)round_up(dim, 4);
  s->ld16 = (int)round_up(dim, 64);
  s->metric = metric;
  s->shadow = (shadow & VS_SHADOW_BF16) ? 1 : 0;
  s->shadow8 = (shadow & VS_SHADOW_FP8) ? 1 : 0;
  s->ld8 = (int)round_up(dim, 128);
  shadow = s->shadow;
  s->num_sms = prop.multiProcessorCount;
  const size_t row_bytes = (size_t)s->ld * 4 + 8 + (shadow ? (size_t)s->ld16 * 2 : 0) + (s->shadow8 ? s->ld8 : 0);
  if (max_rows == 0) {
    max_rows = (int64_t)((double)prop.totalGlobalMem * 0.92 / (double)row_bytes);
    if (max_rows > 0x7ffffff0) max_rows = 0x7ffffff0;
  }
  s->max_rows = max_rows;
  int rc = s->rows.init(device, (size_t)max_rows * s->ld * 4);
  if (!rc) rc = s->norms.init(device, (size_t)max_rows * 4 + 64);
  if (!rc) rc = s->sqnorms.init(device, (size_t)max_rows * 4);
  if (!rc && shadow) rc = s->shadow_rows.init(device, (size_t)max_rows * s->ld16 * 2);
  if (!rc && s->shadow8) rc = s->shadow8_rows.init(device, (size_t)max_rows * s->ld8);
  if (!rc) rc = s->gids.init(device, (size_t)max_rows * 4);
  if (rc) { vs_destroy(s); return rc; }
  // keep stream-ordered workspace memory cached in the pool between searches
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  cudaGetLastError();
  s->pinned_bytes = 1u << 20;
  cudaError_t e = cudaEventCreateWithFlags(&s->append_done, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->bounds, 16);
  if (e == cudaSuccess) e = cudaMemset(s->bounds, 0, 16);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->host_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMallocHost(&s->pinned_in, s->pinned_bytes);
  if (e == cudaSuccess) e = cudaMallocHost(&s->pinned_out, s->pinned_bytes);
  if (e != cudaSuccess) { vs_destroy(s); return cuda_fail(e, "store stream/pinned setup", __FILE__, __LINE__); }
  *out = s;
  return VS_OK;
}

int vs_destroy(vs_store* s) {
  if (!s) return VS_OK;
  cudaSetDevice(s->device);
  cudaDeviceSynchronize();
  s->rows.destroy();
  s->norms.destroy();
  s->sqnorms.destroy();
  s->shadow_rows.destroy();
  s->shadow8_rows.destroy();
  s->gids.destroy();
  free_cert_slots(s);
  free_ws_blocks(s);
  for (cudaEvent_t e : s->ev_pool) cudaEventDestroy(e);
  if (s->bounds) cudaFree(s->bounds);
  if (s->append_done) cudaEventDestroy(s->append_done);
  if (s->host_stream) cudaStreamDestroy(s->host_stream);
  if (s->pinned_in) cudaFreeHost(s->pinned_in);
  if (s->pinned_out) cudaFreeHost(s->pinned_out);
  delete s;
  cudaGetLastError();
  return VS_OK;
}

int64_t vs_count(const vs_store* s) { return s ? s->count.load(std::memory_order_acquire) : 0; }
int64_t vs_fallback_count(const vs_store* s) { return s ? s->fallbacks.load() : 0; }
int64_t vs_retry_count(const vs_store* s) { return s ? s->retries.load() : 0; }

int64_t vs_memory_bytes(const vs_store* s) {
  if (!s) return 0;
  return (int64_t)(s->rows.mapped() + s->norms.mapped() + s->sqnorms.mapped() +
                   s->shadow_rows.mapped() + s->shadow8_rows.mapped() + s->gids.mapped());
}

int vs_reset(vs_store* s) {
  VS_REQUIRE(s != nullptr, "store is NULL");
  std::lock_guard<std::mutex> g(s->mu);
  s->count.store(0, std::memory_order_release);
  s->mapped = false;
  VS_CUDA(cudaSetDevice(s->device));
  VS_CUDA(cudaMemset(s->bounds, 0, 16));
  return VS_OK;
}

__global__ void iota_kernel(int32_t* out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) out[i] = (int32_t)i;
}

static int append_impl(vs_store* s, const float* rows, int64_t m, int rows_on_device,
                       int64_t first_global_id, void* stream_) {
  VS_REQUIRE(s != nullptr, "store is NULL");
  VS_REQUIRE(m >= 0, "m must be >= 0");
  if (m == 0) return VS_OK;
  VS_REQUIRE(rows != nullptr, "rows is NULL");
  cudaStream_t stream = (cudaStream_t)stream_;
  std::lock_guard<std::mutex> g(s->mu);
  VS_CUDA(cudaSetDevice(s->device));
  const int64_t n0 = s->count.load(std::memory_order_acquire);
  const int64_t n1 = n0 + m;
  if (n1 > s->max_rows) { set_error("store is full (max_rows reached)"); return VS_ERR_OOM; }
  if (int rc = s->rows.ensure((size_t)n1 * s->ld * 4, stream)) return rc;
  // + slack: the TMA scan copies norms in whole 16-byte granules past row n
  if (int rc = s->norms.ensure((size_t)n1 * 4 + 64, stream)) return rc;
  if (int rc = s->sqnorms.ensure((size_t)n1 * 4, stream)) return rc;
  if (s->shadow)
    if (int rc = s->shadow_rows.ensure((size_t)n1 * s->ld16 * 2, stream)) return rc;
  if (s->shadow8)
    if (int rc = s->shadow8_rows.ensure((size_t)n1 * s->ld8, stream)) return rc;
  const bool with_ids = first_global_id >= 0;
  if (s->mapped && !with_ids) {
    set_error("store maps local rows to global ids: append with vs_append_ids");
    return VS_ERR_STATE;
  }
  if (with_ids) {
    VS_REQUIRE(first_global_id + m <= (int64_t)0x7fffffff, "global ids must be < 2^31");
    if (int rc = s->gids.ensure((size_t)n1 * 4, stream)) return rc;
    if (!s->mapped && n0 > 0) {   // rows appended so far keep their identity ids
      iota_kernel<<<(unsigned)std::min<int64_t>((n0 + 255) / 256, 1184), 256, 0, stream>>>(
          (int32_t*)s->gids.ptr(), n0);
      count_launch();
      VS_CHECK_LAUNCH();
    }
  }

  float* master = (float*)s->rows.ptr();
  const int64_t piece = std::max<int64_t>(1, (int64_t)(256u << 20) / ((int64_t)s->dim * 4));
  for (int64_t off = 0; off < m; off += piece) {
    const int64_t mm = std::min(piece, m - off);
    const float* src = rows + off * s->dim;
    float* staging = nullptr;
    const float* ksrc = src;
    int64_t ksrc_ld = s->dim;
    if (!rows_on_device) {
      if (s->ld == s->dim) {  // land directly in the arena, normalise in place
        float* dst = master + (n0 + off) * s->ld;
        VS_CUDA(cudaMemcpyAsync(dst, src, (size_t)mm * s->dim * 4, cudaMemcpyHostToDevice, stream));
        ksrc = dst;
        ksrc_ld = s->ld;
      } else {
        VS_CUDA(cudaMallocAsync((void**)&staging, (size_t)mm * s->dim * 4, stream));
        VS_CUDA(cudaMemcpyAsync(staging, src, (size_t)mm * s->dim * 4, cudaMemcpyHostToDevice, stream));
        ksrc = staging;
      }
    }
    const int warps_per_block = 8;
    int64_t blocks = (mm + warps_per_block - 1) / warps_per_block;
    const int64_t cap = (int64_t)s->num_sms * 8;      // 8 CTAs of 8 warps per SM: a grid multiple of 148
    if (blocks > cap) blocks = cap;
    const bool vec = (s->dim & 3) == 0 && (ksrc_ld & 3) == 0 && ((uintptr_t)ksrc & 15) == 0;
    auto kern = vec ? append_norm_kernel<true> : append_norm_kernel<false>;
    // register-resident fast paths: dim % 4 == 0 rows with 16-byte aligned source rows
    if (vec && s->dim <= 128) {
      kern = append_norm_reg_kernel<8, 2>;
      const int per_block = warps_per_block * 4 * kAppendSteps;
      blocks = std::min<int64_t>((mm + per_block - 1) / per_block, cap);
    } else if (vec && s->dim <= 2048) {
      kern = s->dim <= 512 ? append_norm_reg_kernel<32, 2>
                           : (s->dim <= 1024 ? append_norm_reg_kernel<32, 4> : append_norm_reg_kernel<32, 8>);
      const int per_block = warps_per_block * kAppendSteps;
      blocks = std::min<int64_t>((mm + per_block - 1) / per_block, cap);
    }
    kern<<<(unsigned)blocks, 256, 0, stream>>>(
        ksrc, ksrc_ld, master, s->ld, s->dim, n0 + off, mm, (float*)s->norms.ptr(),
        (float*)s->sqnorms.ptr(), s->shadow ? (__nv_bfloat16*)s->shadow_rows.ptr() : nullptr,
        s->ld16, s->metric == VS_METRIC_COSINE ? 1 : 0,
        with_ids ? (int32_t*)s->gids.ptr() : nullptr, with_ids ? first_global_id + off : 0,
        s->bounds);
    count_launch();
    VS_CHECK_LAUNCH();
    if (s->shadow8) {
      shadow8_kernel<<<(unsigned)std::min<int64_t>((mm + 7) / 8, cap), 256, 0, stream>>>(
          master, s->ld, s->dim, (const float*)s->norms.ptr(), n0 + off, mm,
          (unsigned char*)s->shadow8_rows.ptr(), s->ld8);
      count_launch();
      VS_CHECK_LAUNCH();
    }
    if (staging) VS_CUDA(cudaFreeAsync(staging, stream));
  }
  if (!rows_on_device) VS_CUDA(cudaStreamSynchronize(stream));  // host buffer may be reused
  // searches on other streams wait for this event before reading the new rows
  VS_CUDA(cudaEventRecord(s->append_done, stream));
  s->append_stream.store(stream, std::memory_order_release);
  if (with_ids) s->mapped = true;
  s->count.store(n1, std::memory_order_release);
  return VS_OK;
}

int vs_append(vs_store* s, const float* rows, int64_t m, int rows_on_device, void* stream) {
  return append_impl(s, rows, m, rows_on_device, -1, stream);
}

int vs_append_ids(vs_store* s, const float* rows, int64_t m, int rows_on_device,
                  int64_t first_global_id, void* stream) {
  VS_REQUIRE(first_global_id >= 0, "first_global_id must be >= 0");
  return append_impl(s, rows, m, rows_on_device, first_global_id, stream);
}

int vs_read_rows(vs_store* s, int64_t first, int64_t m, float* out, int out_on_device,
                 void* stream_) {
  VS_REQUIRE(s != nullptr, "store is NULL");
  const int64_t n = s->count.load(std::memory_order_acquire);
  VS_REQUIRE(first >= 0 && m >= 0 && first + m <= n, "row range out of bounds");
  if (m == 0) return VS_OK;
  VS_REQUIRE(out != nullptr, "out is NULL");
  cudaStream_t stream = (cudaStream_t)stream_;
  VS_CUDA(cudaSetDevice(s->device));
  // rows appended on another stream: wait for their K1 (the copy below must not overtake it)
  if (s->append_done && s->append_stream.load(std::memory_order_acquire) != stream)
    VS_CUDA(cudaStreamWaitEvent(stream, s->append_done, 0));
  const float* src = (const float*)s->rows.ptr() + first * s->ld;
  VS_CUDA(cudaMemcpy2DAsync(out, (size_t)s->dim * 4, src, (size_t)s->ld * 4, (size_t)s->dim * 4,
                            (size_t)m, out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost,
                            stream));
  if (!out_on_device) VS_CUDA(cudaStreamSynchronize(stream));
  return VS_OK;
}

}  // extern "C"
