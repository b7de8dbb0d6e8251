// K2 scan_topk -- see scan_topk.cuh.  Two staging variants of the same scan:
//   LDG: each warp streams its rows with 128-bit ld.global.nc loads (R rows in flight);
//   TMA: a producer thread stages whole row tiles into shared memory with cp.async.bulk
//        (1-D TMA, mbarrier complete_tx) through a multi-stage ring; 8 consumer warps score
//        the rows out of shared memory.
// Both keep, per warp and query, a sorted k-entry list in shared memory guarded by a
// threshold; a grid-wide threshold (atomicMax on an order-preserving encoding) prunes rows
// that can no longer enter the global top-k.  Row r is skipped iff key < tau, so ties at the
// threshold survive and the final (key desc, id asc) order is deterministic.
#include <cstdlib>
#include <cuda_fp16.h>
#include "scan_topk.cuh"

namespace vs {

// ----------------------------------------------------------------- helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D TMA: global -> shared::cta, completion signalled on an mbarrier as transaction bytes
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// Sorted insert of (key, id) into a k-entry list (best first) shared by one warp.
// Returns the list's k-th key afterwards (-inf while the list is not full).
__device__ __forceinline__ float warp_insert(float* lk, int* li, int k, float key, int id,
                                             int lane) {
  int pos = 0;
  for (int base = 0; base < k; base += 32) {
    const int e = base + lane;
    const bool b = (e < k) && better(lk[e], li[e], key, id);
    const unsigned bal = __ballot_sync(0xffffffffu, b);
    pos += __popc(bal);
    if (bal != 0xffffffffu) break;
  }
  if (pos < k) {
    for (int base = ((k - 1) >> 5) << 5; base >= 0 && base + 32 > pos; base -= 32) {
      const int e = base + lane;
      const bool mv = (e >= pos) && (e < k - 1);
      float tk = 0.f;
      int ti = 0;
      if (mv) { tk = lk[e]; ti = li[e]; }
      __syncwarp();
      if (mv) { lk[e + 1] = tk; li[e + 1] = ti; }
      __syncwarp();
    }
    if (lane == 0) { lk[pos] = key; li[pos] = id; }
    __syncwarp();
  }
  return lk[k - 1];
}

// Reduce V per-lane partial sums across the warp.  Reduce-scatter for the first log2(V)
// butterfly steps (halving the live values each time), plain butterfly afterwards:
// V/2+V/4+..+1+(5-log2 V) shuffles instead of 5*V.  On return v[0] is the total of value
// index (lane >> (5 - log2 V)); per value the additions happen in exactly the order of a
// plain xor-butterfly 16,8,4,2,1, so results are bit-identical to warp_sum().
template <int V>
__device__ __forceinline__ void reduce_scatter(float (&v)[V], int lane) {
  int n = V;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    if (n > 1) {
      n >>= 1;
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < V / 2; ++i) {
        if (i < n) {
          const float send = upper ? v[i] : v[i + n];
          const float keep = upper ? v[i + n] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
    }
  }
}
template <int V> struct Log2 { static constexpr int value = 1 + Log2<V / 2>::value; };
template <> struct Log2<1> { static constexpr int value = 0; };

// bf16 pair -> two fp32 (exact)
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// FMT: 0 = fp32 rows, 1 = bf16 shadow rows, 2 = fp16 shadow rows (cosine: unit-norm rows)
template <bool L2, int FMT>
struct Acc;  // accumulate one 16-byte database vector against the query

template <bool L2>
struct Acc<L2, 0> {
  using Vec = float4;
  static constexpr int kQPerVec = 1;  // float4 query vectors per database vector
  __device__ static __forceinline__ Vec load_global(const void* base, int64_t idx) {
    return ldg_stream(reinterpret_cast<const float4*>(base) + idx);
  }
  __device__ static __forceinline__ float run(float acc, const Vec& x, const float4* q, int c,
                                              int /*plane_stride*/) {
    const float4 qv = q[c];
    return L2 ? sqdiff4_acc(acc, x, qv) : dot4_acc(acc, x, qv);
  }
};
__device__ __forceinline__ float2 h2f(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
template <bool L2>
struct Acc<L2, 2> {
  using Vec = uint4;
  __device__ static __forceinline__ Vec load_global(const void* base, int64_t idx) {
    return ldg_stream_u4(reinterpret_cast<const uint4*>(base) + idx);
  }
  __device__ static __forceinline__ float run(float acc, const Vec& x, const float4* q, int c,
                                              int plane_stride) {
    const float4 q0 = q[c];
    const float4 q1 = q[plane_stride + c];
    const float2 a = h2f(x.x), b = h2f(x.y), cc = h2f(x.z), d = h2f(x.w);
    const float4 x0 = make_float4(a.x, a.y, b.x, b.y);
    const float4 x1 = make_float4(cc.x, cc.y, d.x, d.y);
    acc = L2 ? sqdiff4_acc(acc, x0, q0) : dot4_acc(acc, x0, q0);
    return L2 ? sqdiff4_acc(acc, x1, q1) : dot4_acc(acc, x1, q1);
  }
};
template <bool L2>
struct Acc<L2, 1> {
  using Vec = uint4;
  __device__ static __forceinline__ Vec load_global(const void* base, int64_t idx) {
    return ldg_stream_u4(reinterpret_cast<const uint4*>(base) + idx);
  }
  __device__ static __forceinline__ float run(float acc, const Vec& x, const float4* q, int c,
                                              int plane_stride) {
    const float4 q0 = q[c];
    const float4 q1 = q[plane_stride + c];
    const float4 x0 = make_float4(bf_lo(x.x), bf_hi(x.x), bf_lo(x.y), bf_hi(x.y));
    const float4 x1 = make_float4(bf_lo(x.z), bf_hi(x.z), bf_lo(x.w), bf_hi(x.w));
    acc = L2 ? sqdiff4_acc(acc, x0, q0) : dot4_acc(acc, x0, q0);
    return L2 ? sqdiff4_acc(acc, x1, q1) : dot4_acc(acc, x1, q1);
  }
};

// Shared-memory carve-up common to both variants.
struct ScanSmem {
  float4* q;     // (QB, ldq/4)
  float* lk;     // (warps, QB, k)
  int* li;
};
__device__ __forceinline__ ScanSmem carve(unsigned char* smem, int qb, int ldq, int k, int warps) {
  ScanSmem s;
  s.q = reinterpret_cast<float4*>(smem);
  s.lk = reinterpret_cast<float*>(smem + (size_t)qb * ldq * 4);
  s.li = reinterpret_cast<int*>(s.lk + (size_t)warps * qb * k);
  return s;
}
static size_t scan_fixed_smem(int qb, int ldq, int k, int warps) {
  size_t b = (size_t)qb * ldq * 4 + (size_t)warps * qb * k * 8;
  return (size_t)round_up((int64_t)b, 128);
}

template <int QB>
__device__ __forceinline__ void scan_prologue(const ScanParams& p, const ScanSmem& s, int nthreads) {
  const int qv = p.ldq >> 2;
  const float4* gq = reinterpret_cast<const float4*>(p.q);
  for (int i = threadIdx.x; i < QB * qv; i += nthreads) {
    const int b = i / qv;
    s.q[i] = b < p.nb ? gq[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int i = threadIdx.x; i < p.warps * QB * p.k; i += nthreads) {
    s.lk[i] = VS_NEG_INF;
    s.li[i] = VS_ID_SENTINEL;
  }
}

// After the reduction lane `l` owns value idx = l >> (5 - log2 V): row r = idx / QB of the
// group, query b = idx % QB.  Score it, test it against the threshold and insert survivors.
template <int QB, int R>
__device__ __forceinline__ void scan_epilogue(const ScanParams& p, const ScanSmem& s, float total,
                                              int64_t row0, float nrm, bool row_ok, int lane,
                                              int warp, float& my_tau) {
  constexpr int V = QB * R;
  constexpr int SH = 5 - Log2<V>::value;
  const int idx = lane >> SH;
  const int my_b = idx % QB;
  const bool rep = (lane & ((1 << SH) - 1)) == 0;
  float key;
  if (p.epilogue == VS_METRIC_COSINE) key = total / nrm;
  else if (p.epilogue == VS_METRIC_EUCLIDEAN) key = -sqrtf(total);
  else key = total;
  const bool pass = rep && row_ok && (my_b < p.nb) && (key >= my_tau);
  unsigned m = __ballot_sync(0xffffffffu, pass);
  while (m) {
    const int src = __ffs(m) - 1;
    m &= m - 1;
    const float ks = __shfl_sync(0xffffffffu, key, src);
    const int sidx = src >> SH;
    const int bs = sidx % QB;
    const int id = (int)(row0 + sidx / QB);
    const size_t off = ((size_t)warp * QB + bs) * p.k;
    const float kth = warp_insert(s.lk + off, s.li + off, p.k, ks, id, lane);
    if (my_b == bs) my_tau = fmaxf(my_tau, kth);
    if (lane == 0 && kth > VS_NEG_INF) atomicMax(p.tau + bs, enc_key(kth));
  }
}

// End of the scan: the consumer warps of a CTA merge their sorted k-lists into ONE sorted
// k-list per query (a `warps`-way merge: lane l walks list l, a shuffle butterfly picks the
// best head k times), so the grid leaves gridDim.x lists per query behind instead of
// gridDim.x * warps.
template <int QB>
__device__ __forceinline__ void scan_block_merge(const ScanParams& p, const ScanSmem& s, int lane,
                                                 int warp) {
  asm volatile("bar.sync 1, %0;" ::"r"(p.warps * 32) : "memory");   // consumer warps only
  for (int b = warp; b < p.nb; b += p.warps) {
    const bool has = lane < p.warps;
    const size_t src = ((size_t)(has ? lane : 0) * QB + b) * p.k;
    const size_t dst = ((size_t)b * p.nlists + blockIdx.x) * p.k;
    int head = 0;
    for (int j = 0; j < p.k; ++j) {
      float bk = VS_NEG_INF;
      int bi = VS_ID_SENTINEL, bl = lane;
      if (has && head < p.k) { bk = s.lk[src + head]; bi = s.li[src + head]; }
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        const float ok = __shfl_xor_sync(0xffffffffu, bk, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        const int ol = __shfl_xor_sync(0xffffffffu, bl, off);
        if (better(ok, oi, bk, bi)) { bk = ok; bi = oi; bl = ol; }
      }
      if (lane == bl && bi != VS_ID_SENTINEL) ++head;
      if (lane == 0) {
        p.part_key[dst + j] = bk;
        p.part_id[dst + j] = bi == VS_ID_SENTINEL ? -1 : bi;
      }
    }
  }
}

__device__ __forceinline__ bool mask_bit(const uint32_t* mask, int64_t row) {
  return mask == nullptr || ((__ldg(mask + (row >> 5)) >> (row & 31)) & 1u);
}

// ------------------------------------------------------------ LDG variant
template <int QB, int R, bool L2, int FMT>
__global__ void __launch_bounds__(kMaxScanWarps * 32)
scan_topk_ldg_kernel(const ScanParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  using A = Acc<L2, FMT>;
  constexpr int V = QB * R;
  constexpr int SH = 5 - Log2<V>::value;
  const ScanSmem s = carve(smem, QB, p.ldq, p.k, p.warps);
  scan_prologue<QB>(p, s, blockDim.x);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * p.warps + warp;
  const int64_t nw = (int64_t)gridDim.x * p.warps;
  const int nvec = p.vec_per_row;
  const int qstride = p.ldq >> 2;           // float4 per prepared query
  const int plane = qstride >> 1;           // bf16: second 4-float plane
  const int my_r = (lane >> SH) / QB;
  const int my_b = (lane >> SH) % QB;
  float my_tau = VS_NEG_INF;

  for (int64_t row0 = gw * R; row0 < p.n; row0 += nw * R) {
    // per-lane epilogue inputs fetched early so their latency hides behind the row loads
    const int64_t my_row = row0 + my_r;
    const bool row_ok = my_row < p.n && mask_bit(p.row_mask, my_row);
    float nrm = 1.f;
    if (p.epilogue == VS_METRIC_COSINE && p.norms != nullptr && my_row < p.n) nrm = __ldg(p.norms + my_row);
    const uint32_t gtau = __ldcg(p.tau + my_b);

    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    int64_t base[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = row0 + r < p.n ? row0 + r : row0;   // clamp: result discarded
      base[r] = row * nvec;
    }
#pragma unroll 2
    for (int c = lane; c < nvec; c += 32) {
      typename A::Vec x[R];
#pragma unroll
      for (int r = 0; r < R; ++r) x[r] = A::load_global(p.db, base[r] + c);
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int b = 0; b < QB; ++b)
          acc[r * QB + b] = A::run(acc[r * QB + b], x[r], s.q + b * qstride, c, plane);
    }
    reduce_scatter<V>(acc, lane);
    my_tau = fmaxf(my_tau, dec_key(gtau));
    scan_epilogue<QB, R>(p, s, acc[0], row0, nrm, row_ok, lane, warp, my_tau);
  }
  __syncwarp();
  scan_block_merge<QB>(p, s, lane, warp);
}

// ------------------------------------------------------------ TMA variant
// Warp p.warps is the producer; warps 0..p.warps-1 consume.  Ring of p.stages tiles of
// p.tile_rows rows (+ their clamped norms for fp32 cosine, a second bulk copy on the same
// barrier); full[s] (count 1 + tx bytes) / empty[s] (count p.warps) mbarriers.
template <int QB, int R, bool L2, int FMT>
__global__ void __launch_bounds__((kMaxTmaWarps + 1) * 32, 1)
scan_topk_tma_kernel(const ScanParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  using A = Acc<L2, FMT>;
  constexpr int V = QB * R;
  constexpr int SH = 5 - Log2<V>::value;
  const int nwarps = p.warps;
  const ScanSmem s = carve(smem, QB, p.ldq, p.k, nwarps);
  const size_t fixed = ((size_t)QB * p.ldq * 4 + (size_t)nwarps * QB * p.k * 8 + 127) / 128 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + fixed);       // full[stages], empty[stages]
  unsigned char* tiles = smem + fixed + 128;                        // 16 barriers max
  const int row_bytes = p.vec_per_row * 16;
  const uint32_t tile_bytes = (uint32_t)p.tile_rows * row_bytes;
  const bool stage_norms = p.epilogue == VS_METRIC_COSINE && p.norms != nullptr;
  const uint32_t norm_bytes = stage_norms ? (uint32_t)p.tile_rows * 4 : 0;   // multiple of 16
  const uint32_t stage_bytes = tile_bytes + norm_bytes;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  scan_prologue<QB>(p, s, blockDim.x);
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(smem_u32(bars + i), 1);
      mbar_init(smem_u32(bars + p.stages + i), nwarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int64_t ntiles = (p.n + p.tile_rows - 1) / p.tile_rows;

  if (warp == nwarps) {
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(smem_u32(bars + p.stages + st), ph ^ 1u);
        const int64_t first = t * p.tile_rows;
        const int64_t rows = p.n - first < p.tile_rows ? p.n - first : p.tile_rows;
        const uint32_t bytes = (uint32_t)(rows * row_bytes);
        // norms: whole 16 B granules (the arena keeps slack past row n, see store.cu)
        const uint32_t nbytes = stage_norms ? (uint32_t)((rows * 4 + 15) & ~15) : 0;
        const uint32_t full = smem_u32(bars + st);
        unsigned char* dst = tiles + (size_t)st * stage_bytes;
        mbar_expect_tx(full, bytes + nbytes);
        tma_load_1d(smem_u32(dst), reinterpret_cast<const unsigned char*>(p.db) + first * row_bytes,
                    bytes, full);
        if (stage_norms) tma_load_1d(smem_u32(dst + tile_bytes), p.norms + first, nbytes, full);
        if (++st == p.stages) { st = 0; ph ^= 1u; }
      }
    }
    return;
  }

  const int nvec = p.vec_per_row;
  const int qstride = p.ldq >> 2;
  const int plane = qstride >> 1;
  const int my_r = (lane >> SH) / QB;
  const int my_b = (lane >> SH) % QB;
  const int rows_per_warp = p.tile_rows / nwarps;   // multiple of R
  float my_tau = VS_NEG_INF;
  int st = 0;
  uint32_t ph = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    // the grid-wide threshold is refreshed once per tile; its L2 latency overlaps the wait
    const uint32_t gtau = __ldcg(p.tau + my_b);
    mbar_wait(smem_u32(bars + st), ph);
    my_tau = fmaxf(my_tau, dec_key(gtau));
    const unsigned char* stage = tiles + (size_t)st * stage_bytes;
    const typename A::Vec* tile = reinterpret_cast<const typename A::Vec*>(stage);
    const float* tnorm = reinterpret_cast<const float*>(stage + tile_bytes);
    for (int rr = 0; rr < rows_per_warp; rr += R) {
      const int lrow0 = warp * rows_per_warp + rr;            // row within the tile
      const int64_t row0 = t * p.tile_rows + lrow0;
      if (row0 >= p.n) break;                                  // warp-uniform
      const int64_t my_row = row0 + my_r;
      const bool row_ok = my_row < p.n && mask_bit(p.row_mask, my_row);
      float nrm = 1.f;
      if (stage_norms && my_row < p.n) nrm = tnorm[lrow0 + my_r];
      float acc[V];
#pragma unroll
      for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll 2
      for (int c = lane; c < nvec; c += 32) {
        typename A::Vec x[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          // rows past n inside the last tile were not copied: read row lrow0 instead
          const int lr = (row0 + r < p.n) ? lrow0 + r : lrow0;
          x[r] = tile[(size_t)lr * nvec + c];
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int b = 0; b < QB; ++b)
            acc[r * QB + b] = A::run(acc[r * QB + b], x[r], s.q + b * qstride, c, plane);
      }
      reduce_scatter<V>(acc, lane);
      scan_epilogue<QB, R>(p, s, acc[0], row0, nrm, row_ok, lane, warp, my_tau);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(bars + p.stages + st));
    if (++st == p.stages) { st = 0; ph ^= 1u; }
  }
  __syncwarp();
  scan_block_merge<QB>(p, s, lane, warp);
}

// ---------------------------------------------------------------- launcher
// tuning knobs (diagnostic): B200VS_SCAN_WARPS, B200VS_SCAN_R, B200VS_SCAN_MAXWARPS
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e && *e ? atoi(e) : dflt;
}

template <int QB, int R, bool L2, int FMT>
static int launch_one(ScanParams p, bool use_tma, int num_sms, int* nlists_out, bool dry_run,
                      cudaStream_t stream) {
  const size_t fixed = scan_fixed_smem(QB, p.ldq, p.k, p.warps);
  if (!use_tma) {
    auto kern = scan_topk_ldg_kernel<QB, R, L2, FMT>;
    const size_t smem = fixed;
    if (smem > 48 * 1024)
      VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    VS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, p.warps * 32, smem));
    if (per_sm < 1) { set_error("scan kernel does not fit on an SM (k or dim too large)"); return VS_ERR_INVALID; }
    const int max_warps = env_int("B200VS_SCAN_MAXWARPS", 32);
    if (per_sm * p.warps > max_warps) per_sm = max_warps / p.warps > 0 ? max_warps / p.warps : 1;
    int64_t groups = (p.n + R - 1) / R;
    int64_t blocks = (groups + p.warps - 1) / p.warps;
    const int64_t cap = (int64_t)num_sms * per_sm;
    if (blocks > cap) blocks = cap;
    p.nlists = (int)blocks;
    *nlists_out = p.nlists;
    if (dry_run) return VS_OK;
    {
      ProfScope prof(kProfScan, stream);
      kern<<<(unsigned)blocks, p.warps * 32, smem, stream>>>(p);
    }
    count_launch();
    VS_CHECK_LAUNCH();
    return VS_OK;
  }
  auto kern = scan_topk_tma_kernel<QB, R, L2, FMT>;
  const bool stage_norms = p.epilogue == VS_METRIC_COSINE && p.norms != nullptr;
  const size_t smem = fixed + 128 +
                      (size_t)p.stages * ((size_t)p.tile_rows * p.vec_per_row * 16 + (stage_norms ? p.tile_rows * 4 : 0));
  VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t ntiles = (p.n + p.tile_rows - 1) / p.tile_rows;
  int64_t blocks = ntiles < num_sms ? ntiles : num_sms;
  p.nlists = (int)blocks;
  *nlists_out = p.nlists;
  if (dry_run) return VS_OK;
  {
    ProfScope prof(kProfScan, stream);
    kern<<<(unsigned)blocks, (p.warps + 1) * 32, smem, stream>>>(p);
  }
  count_launch();
  VS_CHECK_LAUNCH();
  return VS_OK;
}

template <int QB, bool L2, int FMT>
static int launch_r(const ScanParams& p, int r, bool use_tma, int num_sms, int* nl, bool dry,
                    cudaStream_t st) {
  switch (r) {   // QB * R <= 32 accumulators per lane
    case 1: return launch_one<QB, 1, L2, FMT>(p, use_tma, num_sms, nl, dry, st);
    case 2: return launch_one<QB, 2, L2, FMT>(p, use_tma, num_sms, nl, dry, st);
    case 4: return launch_one<QB, 4, L2, FMT>(p, use_tma, num_sms, nl, dry, st);
    case 8: if constexpr (QB <= 4) return launch_one<QB, 8, L2, FMT>(p, use_tma, num_sms, nl, dry, st); break;
    case 16: if constexpr (QB <= 2) return launch_one<QB, 16, L2, FMT>(p, use_tma, num_sms, nl, dry, st); break;
  }
  set_error("internal: unsupported rows-per-warp");
  return VS_ERR_INVALID;
}

template <bool L2, int FMT>
static int launch_qb(const ScanParams& p, int qb, int r, bool use_tma, int num_sms, int* nl,
                     bool dry, cudaStream_t st) {
  switch (qb) {
    case 1: return launch_r<1, L2, FMT>(p, r, use_tma, num_sms, nl, dry, st);
    case 2: return launch_r<2, L2, FMT>(p, r, use_tma, num_sms, nl, dry, st);
    case 4: return launch_r<4, L2, FMT>(p, r, use_tma, num_sms, nl, dry, st);
    case 8: return launch_r<8, L2, FMT>(p, r, use_tma, num_sms, nl, dry, st);
  }
  set_error("internal: unsupported query block");
  return VS_ERR_INVALID;
}

int launch_scan(const ScanParams& base, int qb, bool l2, int fmt, bool use_tma, int num_sms,
                int* nlists_out, size_t* part_elems_out, bool dry_run, cudaStream_t stream) {
  ScanParams p = base;
  // consumer warps: 12 while the per-warp lists stay small, else 8
  p.warps = scan_warps_for(qb, p.k);
  const int ew = env_int("B200VS_SCAN_WARPS", 0);
  if (ew > 0 && ew <= kMaxScanWarps && (size_t)ew * qb * p.k * 8 <= 64 * 1024) p.warps = ew;
  // rows in flight per warp: V = QB*R accumulators per lane (16 by default, up to 32)
  int r = qb == 1 ? 16 : 32 / qb;      // measured: 10M x 128 batch 8 runs 1.5x faster at V = 32
  const int er = env_int("B200VS_SCAN_R", 0);
  if (er > 0 && er * qb <= 32 && er <= 16 && (er & (er - 1)) == 0) r = er;
  const int row_bytes = p.vec_per_row * 16;
  if (use_tma && p.warps > kMaxTmaWarps) p.warps = kMaxTmaWarps;
  if (use_tma) {
    // tile = warps x rows_per_warp rows, about 32 KB; at least 2 stages must fit
    const size_t fixed = scan_fixed_smem(qb, p.ldq, p.k, p.warps) + 128;
    const size_t budget = 200 * 1024 > fixed ? 200 * 1024 - fixed : 0;
    int rpw = 8;
    while (rpw > 1 && (size_t)p.warps * rpw * row_bytes > 32 * 1024) rpw >>= 1;
    if (r > rpw) r = rpw;
    const size_t tile = (size_t)p.warps * rpw * (row_bytes + 4);
    int stages = (int)(budget / tile);
    if (stages > 6) stages = 6;
    if (stages < 2) use_tma = false;   // rows too wide to stage: stream them directly
    p.tile_rows = p.warps * rpw;
    p.stages = stages;
  }
  int rc;
  if (l2) rc = fmt ? launch_qb<true, 1>(p, qb, r, use_tma, num_sms, nlists_out, dry_run, stream)
                   : launch_qb<true, 0>(p, qb, r, use_tma, num_sms, nlists_out, dry_run, stream);
  else if (fmt == 2) rc = launch_qb<false, 2>(p, qb, r, use_tma, num_sms, nlists_out, dry_run, stream);
  else rc = fmt ? launch_qb<false, 1>(p, qb, r, use_tma, num_sms, nlists_out, dry_run, stream)
                : launch_qb<false, 0>(p, qb, r, use_tma, num_sms, nlists_out, dry_run, stream);
  if (rc == VS_OK && part_elems_out) *part_elems_out = (size_t)qb * (*nlists_out) * p.k;
  return rc;
}

// ------------------------------------------------------------ prep_queries
// One warp per query.  cosine: q / max(||q||, 1e-8) (service/optimized_vector_store.py:34-39);
// other metrics: q * scale.  Pads to ldq; optional two-plane layout for the bf16 scan.
__global__ void prep_queries_kernel(const float* __restrict__ q, int B, int dim, int metric,
                                    int ldq, int planes, float scale, float* __restrict__ out,
                                    float* __restrict__ qnorm_out, uint32_t* __restrict__ tau) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* src = q + (size_t)b * dim;
  float acc = 0.f;
  for (int c = lane; c < dim; c += 32) { const float v = src[c]; acc = fmaf(v, v, acc); }
  const float tot = warp_sum(acc);
  const float nrm = fmaxf(sqrtf(tot), 1e-8f);
  if (lane == 0) {
    if (qnorm_out) qnorm_out[b] = nrm;
    if (tau) tau[b] = VS_ENC_NEG_INF;
  }
  float* dst = out + (size_t)b * ldq;
  const int half = ldq >> 1;
  for (int c = lane; c < ldq; c += 32) {
    float v = c < dim ? src[c] : 0.f;
    v = metric == VS_METRIC_COSINE ? v / nrm : v * scale;
    int o = c;
    if (planes) o = ((c & 7) >= 4 ? half : 0) + ((c >> 3) << 2) + (c & 3);
    dst[o] = v;
  }
}

int launch_prep_queries(const float* q, int B, int dim, int metric, int ldq, bool bf16_planes,
                        float scale, float* out, float* qnorm_out, uint32_t* tau,
                        cudaStream_t stream) {
  const int wpb = 4;
  prep_queries_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, stream>>>(
      q, B, dim, metric, ldq, bf16_planes ? 1 : 0, scale, out, qnorm_out, tau);
  count_launch();
  VS_CHECK_LAUNCH();
  return VS_OK;
}

}  // namespace vs
