// K3 gemm_topk -- large-batch search as a dense contraction on the 5th-gen tensor cores.
//
//   S'[q, r] = sum_k bf16(Q[q, k]) * bf16(X[r, k])        tcgen05.mma kind::f16, fp32 in TMEM
//
// replaces `matmul(Qn, DBn.T)` + `argsort(axis=1)[:, :k]` of the reference
// (performance/mlx_optimized.py:86,235-236) without ever writing the (B, N) matrix:
//
//   pass 1  GEMM over a sample of the rows; epilogue keeps, per query, the maximum of every
//           128-row tile.  The kc-th largest tile maximum is a lower bound tau[q] on the kc-th
//           best score of the whole database (kc distinct rows reach it).
//   pass 2  GEMM over all rows; epilogue compares TMEM columns with tau[q] (one thread owns
//           one query: lane = query) and appends the rare survivors to that thread's private
//           candidate buffer.
//   then    K4 merge -> top-kc candidates by bf16 score; K5 rescoring in exact fp32 with the
//           scan's arithmetic; final K4 merge -> top-k; certification: the candidate set
//           provably contains the exact top-k when
//               exact_k-th  >  bf16_kc-th + E,   E >= |exact - bf16| for every row,
//           otherwise (or if a candidate buffer overflowed) the query is re-run through the
//           exact fp32 scan (K2).  Results are therefore always the exact fp32 ones.
//
// Kernel anatomy (one CTA per SM, 384 threads): warp 0 = TMA producer, warp 1 = MMA issuer
// (one elected lane; RESIDENT: warps 1 and 3 issue alternate accumulators), warp 2 = TMEM
// allocator, warps 4-11 = epilogue (two groups of four, TMEM lane quadrant = warp % 4).
// Operands are 16-bit, K-major, 128-byte swizzled: one "chunk" is 128 rows x 64
// elements = 16 KB, loaded by one cp.async.bulk.tensor.2d.
//   RESIDENT (K <= 256): the CTA's query tiles (up to 4 x 128 queries) are loaded once and stay
//     in shared memory; database tiles of 128 rows stream through a ring; 4 TMEM accumulators
//     of 128 columns, one per query tile, let the epilogue of tile m overlap the MMAs of m+1.
//   STREAMING (any K): query and database chunks both stream through the ring per 64-wide K
//     step; database tiles of 256 rows; 2 TMEM accumulators of 256 columns.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>
#include "gemm_topk.cuh"
#include "scan_topk.cuh"
#include "store.cuh"

namespace vs {

#ifndef VS_EPI_GROUPS
#define VS_EPI_GROUPS 2
#endif
#ifndef VS_EPI_GROUPS_RES
#define VS_EPI_GROUPS_RES 2
#endif
// epilogue warp groups (4 warps each; query tile mt belongs to group mt % groups), separately
// for the STREAMING (K > 256) and RESIDENT (K <= 256) variants.  Measured at 10 M x 128, batch
// 1024 (profiles/r01_k3_probe_experiments.txt): four groups are 5 % SLOWER than two for RESIDENT,
// with one MMA-issuing warp (2.37 vs 2.24-2.33 ms) and with two (2.21 vs 2.10 ms).
constexpr int kEpiGroupsStream = VS_EPI_GROUPS;
constexpr int kEpiGroupsRes = VS_EPI_GROUPS_RES;
__host__ __device__ constexpr int epi_groups(bool resident) { return resident ? kEpiGroupsRes : kEpiGroupsStream; }
// 4 control warps + the epilogue warps
__host__ __device__ constexpr int gemm_threads(bool resident) { return 128 + 128 * epi_groups(resident); }
constexpr int kTileM = 128;              // queries per m-tile = TMEM lanes
constexpr int kChunkK = 64;              // bf16 elements per 128-byte swizzled row
constexpr int kChunkBytes = 128 * 128;   // 128 rows x 128 B
constexpr int kTmemCols = 512;
constexpr int kCandCap = 64;             // candidate slots per (CTA, query)
constexpr int kMaxGroupTiles = 4;        // pass 1: tiles (of one CTA) per maximum
constexpr int kGlobalCap = 4096;         // candidate slots per query over all CTAs (= K4's capacity)
constexpr int kMaxQueriesPerLaunch = 2048;
#ifndef VS_RES_ISSUERS
#define VS_RES_ISSUERS 2
#endif
constexpr int kResIssuers = VS_RES_ISSUERS;   // RESIDENT: warps that issue MMAs (1: warp 1; 2: warps 1 and 3)
#ifndef VS_RES_TN
#define VS_RES_TN 128
#endif
constexpr int kResTN = VS_RES_TN;        // RESIDENT: database rows per tile (MMA N), 128 or 256

enum { kModeFilter = 0, kModeMax = 1, kModeDump = 2,
       // diagnostic builds only (-DVS_GEMM_DEBUG_MODES, timing experiments, results are not usable):
       kModeNop = 3,    // epilogue releases every accumulator unread: the pure TMA + MMA pipeline
       kModeHalf = 4 }; // the filter epilogue over the first half of every accumulator's columns

struct GemmParams {
  int kchunks;              // K / 64
  int64_t n_rows;           // database rows visible
  int n_tiles;              // database tiles this launch covers (tile = TN rows)
  int tile_stride;          // launch tile t is database tile (tile_first + t) * tile_stride (pass 1 samples
  int tile_first;           //   the whole row range with a stride; pass 2 may run as two ranges)
  int m_tiles;              // query tiles in the batch
  int ngroups;              // RESIDENT: query groups (CTA c serves group c % ngroups)
  int nq;                   // live queries
  int mode;
  int fp16;                 // operands are fp16 (cosine) instead of bf16
  int fp8;                  // operands are e4m3: tcgen05.mma kind::f8f6f4 (K = 32 per MMA, 128 per chunk)
  int stages;               // ring depth
  uint32_t idesc;           // tcgen05 instruction descriptor (operand format, M, N)
  const float* tau;         // (nq,) filter threshold (kModeFilter)
  float* cand_score;        // (lists, m_tiles*128, kCandCap)
  int32_t* cand_id;
  int32_t* cand_cnt;        // (lists, m_tiles*128) candidates per list, zeroed by the host (STREAMING keeps
                            // its running counts here, RESIDENT writes them at the end)
  int32_t* overflow;        // (m_tiles*128,) set to 1 when a buffer overflowed
  float* glist_s;           // (m_tiles*128, kGlobalCap) dense per-query candidate lists: at the end of the
  int32_t* glist_i;         //   kernel every thread moves its private candidates here (one atomicAdd on
  int32_t* gcount;          //   gcount[q] per (CTA, query)), so K4 reads contiguous entries only
  float* gmax;              // kModeMax: (m_tiles*128, n_groups) maxima of groups of kMaxGroupTiles tiles
  int n_groups;             //   = units_in_group * groups_per_unit
  int groups_per_unit;
  int group_tiles;          //   tiles (of one unit) per maximum: kMaxGroupTiles, or 1 when tiles are scarce
  float* dump;              // kModeDump: (nq, dump_ld)
  int64_t dump_ld;
  const float* sqnorms;     // euclidean: ||x||^2 per row; the epilogue turns the accumulator s into the key
                            //   2 s - ||x||^2 (= ||q||^2 - d^2: larger = closer).  NULL: key = s
  const uint32_t* row_mask; // nullable: bit r set = row r takes part (metadata filter pushed into the GEMM)
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// One lane of a converged warp, chosen by `elect.sync`.  The single-thread roles (TMA producer,
// MMA issuer) branch on THIS rather than on `lane == 0`: with a lane-id test the compiler cannot
// tell that exactly one thread is active and wraps every warp-uniform instruction (UTCHMMA,
// UTCBAR, UTMALDG) in an ELECT / BRA.U.ANY serialisation loop -- measured ~110 clk per MMA for
// the issuing thread, more than the 64 clk an M = 128, N = 128, K = 16 MMA takes to execute.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1
// [46,48), layout_type=2 [61,64))
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, both operands K-major,
// operand format 0 = fp16, 1 = bf16
__host__ __device__ constexpr uint32_t instr_desc(int m, int n, int fmt) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// -------------------------------------------------------------------- the kernel
// MT > 0: RESIDENT with MT query tiles per CTA.  MT == 0: STREAMING.
// CG = 1: one CTA per MMA (M = 128).  CG = 2: a CTA pair (cluster of 2, `cta_group::2`) shares
// every MMA: M = 256 = 128 query rows from each CTA, N = 256 database rows of which each CTA
// loads and holds half; the leader (cluster rank 0) issues the MMAs for both, accumulator rows
// land in each CTA's own TMEM.  Halves the shared-memory operand reads per MMA and the
// database bytes each SM pulls through TMA.
//   MMA N (database rows per tile) TN: RESIDENT 128 (four accumulators of 128 columns, so the
//   epilogue of one overlaps the MMAs of the next three), STREAMING 256.  RESIDENT with CG = 2
//   therefore issues M = 256, N = 128 MMAs: each CTA loads 64 database rows per tile.
template <int CG> struct CgOps;
template <> struct CgOps<1> {
  __device__ static __forceinline__ void alloc(uint32_t dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __device__ static __forceinline__ void dealloc(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(kTmemCols) : "memory");
  }
  __device__ static __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    tc_mma(d, a, b, idesc, acc);
  }
  __device__ static __forceinline__ void mma_f8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    tc_mma_f8(d, a, b, idesc, acc);
  }
  __device__ static __forceinline__ void commit(uint32_t bar) { tc_commit(bar); }
  __device__ static __forceinline__ void load(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    tma_load_2d(dst, map, c0, c1, bar);
  }
};
template <> struct CgOps<2> {
  __device__ static __forceinline__ void alloc(uint32_t dst) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  __device__ static __forceinline__ void dealloc(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(kTmemCols) : "memory");
  }
  __device__ static __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  }
  __device__ static __forceinline__ void mma_f8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  }
  // arrive on the barrier at this shared-memory offset in BOTH CTAs of the pair
  __device__ static __forceinline__ void commit(uint32_t bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(bar), "h"((uint16_t)3) : "memory");
  }
  // the transaction bytes are credited to the LEADER's barrier (peer bit of the address cleared)
  __device__ static __forceinline__ void load(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
  }
};
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void bar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(rank) : "memory");
}

template <int MT, int MODE, int CG>
__global__ void __launch_bounds__(gemm_threads(MT > 0), 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x,
                 const GemmParams p) {
  constexpr bool RES = MT > 0;
  constexpr int kEpiGroups = epi_groups(RES);
  constexpr bool FILT = MODE == kModeFilter || MODE == kModeHalf;
  constexpr int TN = RES ? kResTN : 256;                  // MMA N = database rows per tile
  constexpr int TN_LOCAL = TN / CG;                       // rows of the tile this CTA loads
  constexpr int SLOTS = kTmemCols / TN;
  constexpr int B_CHUNK_BYTES = TN_LOCAL * 128;
  using Ops = CgOps<CG>;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment for the 128B-swizzle atoms (same offset in both CTAs of a pair)
  unsigned char* smem = smem_raw + ((1024 - (s_u32(smem_raw) & 1023)) & 1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kch = p.kchunks;
  const int kstep = p.fp8 ? 128 : kChunkK;                // elements per 128-byte chunk row
  const int crank = CG == 2 ? (int)cluster_rank() : 0;    // 0 = leader (issues the MMAs)

  // ---- carve-up
  unsigned char* a_res = smem;                                         // RES: MT*kch chunks
  const size_t a_bytes = RES ? (size_t)MT * kch * kChunkBytes : 0;
  const size_t stage_bytes = RES ? (size_t)kch * B_CHUNK_BYTES : (size_t)kChunkBytes + B_CHUNK_BYTES;
  unsigned char* ring = smem + a_bytes;
  unsigned char* tail = ring + (size_t)p.stages * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  // bars: [0]=a_full, [1..S]=full, [1+S..2S]=empty, then acc_full[SLOTS], acc_empty[SLOTS]
  const uint32_t bar_a = s_u32(bars);
  const uint32_t bar_full = s_u32(bars + 1);
  const uint32_t bar_empty = s_u32(bars + 1 + p.stages);
  const uint32_t bar_accf = s_u32(bars + 1 + 2 * p.stages);
  const uint32_t bar_acce = s_u32(bars + 1 + 2 * p.stages + SLOTS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + 2 * p.stages + 2 * SLOTS);
  // euclidean only: ||x||^2 of the current tile's rows, per epilogue group, double-buffered
  float* sq_base = reinterpret_cast<float*>(tail + 512);

  // ---- work assignment, in units of one CTA (CG=1) or one CTA pair (CG=2)
  // query tiles are counted per unit: unit tile m covers the 128-row tiles m*CG + crank
  const int unit = blockIdx.x / CG;
  const int n_units = gridDim.x / CG;
  const int um_tiles = p.m_tiles / CG;                                 // m_tiles is a multiple of CG
  const int ngroups = RES ? p.ngroups : 1;
  const int group = unit % ngroups;
  const int uig = unit / ngroups;                                      // unit index inside its group
  const int units_in_group = (n_units - group + ngroups - 1) / ngroups;
  const int m_first = RES ? group * MT : 0;
  const int m_count = RES ? min(MT, um_tiles - m_first) : um_tiles;
  int my_tiles = 0;
  if (uig < p.n_tiles) my_tiles = (p.n_tiles - uig + units_in_group - 1) / units_in_group;

  if (threadIdx.x == 0) {
    bar_init(bar_a, 1);
    // a RESIDENT stage is released by every MMA-issuing warp (one tcgen05.commit each)
    for (int i = 0; i < p.stages; ++i) { bar_init(bar_full + 8 * i, 1); bar_init(bar_empty + 8 * i, RES ? kResIssuers : 1); }
    for (int i = 0; i < SLOTS; ++i) { bar_init(bar_accf + 8 * i, 1); bar_init(bar_acce + 8 * i, 4 * CG); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) Ops::alloc(s_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();          // both CTAs' barriers exist before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================ TMA producer
    // Every CTA loads its own operands; with CG = 2 the bytes of both CTAs are credited to the
    // leader's full barriers, which the leader arms with the pair's total.
    if (my_tiles > 0 && m_count > 0 && elect_one()) {
      if (RES) {
        if (crank == 0) bar_expect_tx(bar_a, (uint32_t)(CG * m_count * kch * kChunkBytes));
        for (int mt = 0; mt < m_count; ++mt)
          for (int kc = 0; kc < kch; ++kc)
            Ops::load(s_u32(a_res + (size_t)(mt * kch + kc) * kChunkBytes), &map_q, kc * kstep,
                      ((m_first + mt) * CG + crank) * kTileM, bar_a);
      }
      int st = 0;
      uint32_t ph = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int nt = (p.tile_first + uig + i * units_in_group) * p.tile_stride;
        const int row0 = nt * TN + crank * TN_LOCAL;                  // first database row this CTA loads
        if (RES) {
          bar_wait(bar_empty + 8 * st, ph ^ 1u);
          if (crank == 0) bar_expect_tx(bar_full + 8 * st, (uint32_t)(CG * kch * B_CHUNK_BYTES));
          for (int kc = 0; kc < kch; ++kc) {
            unsigned char* dst = ring + (size_t)st * stage_bytes + (size_t)kc * B_CHUNK_BYTES;
            Ops::load(s_u32(dst), &map_x, kc * kstep, row0, bar_full + 8 * st);
            if (TN_LOCAL == 256) Ops::load(s_u32(dst + kChunkBytes), &map_x, kc * kstep, row0 + 128, bar_full + 8 * st);
          }
          if (++st == p.stages) { st = 0; ph ^= 1u; }
        } else {
          for (int mt = 0; mt < m_count; ++mt)
            for (int kc = 0; kc < kch; ++kc) {
              bar_wait(bar_empty + 8 * st, ph ^ 1u);
              if (crank == 0) bar_expect_tx(bar_full + 8 * st, (uint32_t)(CG * (kChunkBytes + B_CHUNK_BYTES)));
              unsigned char* dst = ring + (size_t)st * stage_bytes;
              Ops::load(s_u32(dst), &map_q, kc * kstep, (mt * CG + crank) * kTileM, bar_full + 8 * st);
              Ops::load(s_u32(dst + kChunkBytes), &map_x, kc * kstep, row0, bar_full + 8 * st);
              if (TN_LOCAL == 256)   // 256 rows = two boxes of 128
                Ops::load(s_u32(dst + 2 * kChunkBytes), &map_x, kc * kstep, row0 + 128, bar_full + 8 * st);
              if (++st == p.stages) { st = 0; ph ^= 1u; }
            }
        }
      }
    }
  } else if (warp == 1 || (RES && kResIssuers == 2 && warp == 3)) {
    // ================================================================ MMA issuer (leader CTA)
    // RESIDENT at K <= 256 is bound by how fast ONE thread can issue: ~117 SASS instructions per
    // accumulator (barrier wait, two descriptors per K chunk, 8 MMAs, commit) on a single warp's
    // dependent uniform-datapath chain take 700-800 clk, the 8 MMAs execute in 512 (ncu: the
    // issuing warp 85 % busy, tensor pipe 63 %).  So two warps issue, taking alternate accumulators
    // (`it` parity: with 4 query tiles each warp always owns the same two TMEM slots); MMAs of
    // different accumulators are independent, both warps read the same shared-memory operands and
    // each releases the stage with its own commit.
    const int issuer = warp == 1 ? 0 : 1;
    if (crank == 0 && my_tiles > 0 && m_count > 0 && elect_one()) {
      const uint32_t idesc = p.idesc;
      if (RES) { bar_wait(bar_a, 0); tc_fence_after(); }
      int st = 0;
      uint32_t ph = 0;
      int it = 0;                                       // (n-tile, m-tile) sequence number
      for (int i = 0; i < my_tiles; ++i) {
        if (RES) {
          bar_wait(bar_full + 8 * st, ph);
          tc_fence_after();
        }
        for (int mt = 0; mt < m_count; ++mt, ++it) {
          if (RES && kResIssuers == 2 && (it & 1) != issuer) continue;   // the other issuing warp's
          const int slot = it % SLOTS;
          const uint32_t aph = (uint32_t)(it / SLOTS) & 1u;
          bar_wait(bar_acce + 8 * slot, aph ^ 1u);      // epilogues drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(slot * TN);
          for (int kc = 0; kc < kch; ++kc) {
            uint32_t a_addr, b_addr;
            if (RES) {
              a_addr = s_u32(a_res + (size_t)(mt * kch + kc) * kChunkBytes);
              b_addr = s_u32(ring + (size_t)st * stage_bytes + (size_t)kc * B_CHUNK_BYTES);
            } else {
              bar_wait(bar_full + 8 * st, ph);
              tc_fence_after();
              a_addr = s_u32(ring + (size_t)st * stage_bytes);
              b_addr = a_addr + kChunkBytes;
            }
            const uint64_t a_desc = smem_desc(a_addr);
            const uint64_t b_desc = smem_desc(b_addr);
            if (p.fp8) {
#pragma unroll
              for (int k = 0; k < 4; ++k)               // e4m3: +32 B per 32-element K step
                Ops::mma_f8(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                            (kc | k) != 0 ? 1u : 0u);
            } else {
#pragma unroll
              for (int k = 0; k < kChunkK / 16; ++k)    // 16-bit: +32 B per 16-element K step
                Ops::mma(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                         (kc | k) != 0 ? 1u : 0u);
            }
            if (!RES) {
              Ops::commit(bar_empty + 8 * st);          // smem stage free once these MMAs retire
              if (++st == p.stages) { st = 0; ph ^= 1u; }
            }
          }
          Ops::commit(bar_accf + 8 * slot);             // accumulator ready for the epilogues
        }
        if (RES) {
          Ops::commit(bar_empty + 8 * st);
          if (++st == p.stages) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================================================ epilogue
    // One thread owns one query row of the m-tile (TMEM lane = query).  Per accumulator the
    // warp reads 32-column chunks (double-buffered tcgen05.ld) and reduces each to its
    // maximum with 3-input FMNMX; only when some lane's maximum reaches its threshold does the
    // warp take the rare path, which re-reads the offending 8-column groups from TMEM (short
    // code: the hot loop must stay resident in the instruction cache).
    const int quad = warp & 3;                          // TMEM lane quadrant of this warp
    const int grp = (warp - 4) >> 2;                    // query tile mt belongs to group mt % kEpiGroups
    const int row = quad * 32 + lane;                   // query row inside the m-tile
    const int list = RES ? uig : unit;                  // candidate list of this unit
    const int64_t q_total = (int64_t)p.m_tiles * kTileM;
    constexpr int NST = RES ? MT : 1;
    float tau_l[NST];                                   // RESIDENT: per-thread state of its
    int cnt_l[NST];                                     // query tiles (dynamically indexed)
    float rmax_l[MODE == kModeMax ? 16 : 1];            // pass 1: running maximum per query tile
    if (MODE == kModeMax)
      for (int mt = 0; mt < 16; ++mt) rmax_l[mt] = VS_NEG_INF;
    if (RES && FILT) {
      for (int mt = 0; mt < NST; ++mt) {
        cnt_l[mt] = 0;
        const int q = ((m_first + mt) * CG + crank) * kTileM + row;
        tau_l[mt] = q < p.nq ? p.tau[q] : __int_as_float(0x7f800000);
      }
    }
    int it = 0;
    const bool l2 = p.sqnorms != nullptr;
    const int gtid = threadIdx.x & 127;                 // thread inside its epilogue group
    float* sq_grp = sq_base + grp * 2 * TN;
    float sq_next[TN / 128];
    if (l2 && my_tiles > 0) {
#pragma unroll
      for (int h = 0; h < TN / 128; ++h) {
        const int64_t r = (int64_t)(p.tile_first + uig) * p.tile_stride * TN + h * 128 + gtid;
        sq_next[h] = r < p.n_rows ? __ldg(p.sqnorms + r) : __int_as_float(0x7f800000);
      }
    }
    for (int i = 0; i < my_tiles; ++i) {
      const int nt = (p.tile_first + uig + i * units_in_group) * p.tile_stride;
      const float* sqb = sq_grp + (i & 1) * TN;
      if (l2) {
        // rows past the end of the store get ||x||^2 = +inf: their key is -inf in every mode
#pragma unroll
        for (int h = 0; h < TN / 128; ++h) sq_grp[(i & 1) * TN + h * 128 + gtid] = sq_next[h];
        asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
        if (i + 1 < my_tiles) {
#pragma unroll
          for (int h = 0; h < TN / 128; ++h) {
            const int64_t r = (int64_t)(p.tile_first + uig + (i + 1) * units_in_group) * p.tile_stride * TN + h * 128 + gtid;
            sq_next[h] = r < p.n_rows ? __ldg(p.sqnorms + r) : __int_as_float(0x7f800000);
          }
        }
      }
#pragma unroll 1
      for (int mt = 0; mt < m_count; ++mt, ++it) {
        if ((mt % kEpiGroups) != grp) continue;
        const int slot = it % SLOTS;
        const uint32_t aph = (uint32_t)(it / SLOTS) & 1u;
        const int q = ((m_first + mt) * CG + crank) * kTileM + row;
        const int64_t cbase = ((int64_t)list * q_total + q) * kCandCap;
        float t = __int_as_float(0x7f800000);
        int c = 0;
        if (FILT) {
          if (RES) { t = tau_l[mt]; c = cnt_l[mt]; }
          else {
            if (q < p.nq) t = p.tau[q];
            c = p.cand_cnt[(int64_t)list * q_total + q];
          }
        }
        bar_wait(bar_accf + 8 * slot, aph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(slot * TN);
        float rmax = MODE == kModeMax ? rmax_l[mt & 15] : VS_NEG_INF;
        float va[32], vb[32];
        constexpr int TN_READ = MODE == kModeNop ? 0 : (MODE == kModeHalf ? TN / 2 : TN);
        if (TN_READ > 0) tc_ld32(taddr, va);
#pragma unroll 1
        for (int c0 = 0; c0 < TN_READ; c0 += 64) {
          tc_wait_ld();
          tc_ld32(taddr + (uint32_t)(c0 + 32), vb);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float (&v)[32] = half == 0 ? va : vb;
            const int cc = c0 + 32 * half;
            if (half == 1) {
              tc_wait_ld();
              if (c0 + 64 < TN_READ) tc_ld32(taddr + (uint32_t)(c0 + 64), va);
            }
            if (MODE == kModeDump) {
              if (q < p.nq) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int64_t r = (int64_t)nt * TN + cc + j;
                  if (r < p.n_rows) p.dump[(int64_t)q * p.dump_ld + r] = v[j];
                }
              }
            } else {
              if (l2) {                                  // warp-uniform: key = 2 s - ||x||^2
                const float4* sq4 = reinterpret_cast<const float4*>(sqb + cc);
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 w = sq4[j4];
                  v[4 * j4 + 0] = fmaf(2.f, v[4 * j4 + 0], -w.x);
                  v[4 * j4 + 1] = fmaf(2.f, v[4 * j4 + 1], -w.y);
                  v[4 * j4 + 2] = fmaf(2.f, v[4 * j4 + 2], -w.z);
                  v[4 * j4 + 3] = fmaf(2.f, v[4 * j4 + 3], -w.w);
                }
              }
              uint32_t mword = 0xffffffffu;              // rows of this chunk that take part
              if (p.row_mask != nullptr) {
                const int64_t r0 = (int64_t)nt * TN + cc;
                mword = r0 < p.n_rows ? __ldg(p.row_mask + (r0 >> 5)) : 0u;
                if (MODE == kModeMax) {                  // pass 1 bounds the ks-th best TAKING-PART row
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = (mword >> j) & 1u ? v[j] : VS_NEG_INF;
                }
              }
              // maxima of the four 8-column groups (independent 3-input FMNMX trees)
              float g[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float a = max3(v[8 * u], v[8 * u + 1], v[8 * u + 2]);
                const float b = max3(v[8 * u + 3], v[8 * u + 4], v[8 * u + 5]);
                g[u] = max3(a, b, fmaxf(v[8 * u + 6], v[8 * u + 7]));
              }
              const float m = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
              if (MODE == kModeMax) rmax = fmaxf(rmax, m);
              if (FILT && __any_sync(0xffffffffu, m >= t)) {
                // rare path: take the hits straight from the registers of this chunk (no second
                // tcgen05.ld).  Per 8-column group one warp-uniform vote; inside, a branch-free hit
                // mask per lane and a (divergent, almost always single-trip) loop over its set bits
                // that picks the value with a select chain -- no per-value branches.  Measured at
                // 10 M x 128, batch 1024 (profiles/r02_k3_probe.txt): 1.96 ms vs 2.14 ms per search
                // with the first version, which re-read the 8-column groups from TMEM.
                const int lim = (int)min((int64_t)TN, p.n_rows - (int64_t)nt * TN) - cc;   // live columns
                const int32_t id0 = (int32_t)((int64_t)nt * TN + cc);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  if (!__any_sync(0xffffffffu, g[u] >= t)) continue;
                  uint32_t hits = 0;
#pragma unroll
                  for (int j = 0; j < 8; ++j) hits |= (v[8 * u + j] >= t ? 1u : 0u) << j;
                  const int live = lim - 8 * u;                       // columns of this group inside the store
                  hits &= live >= 8 ? 0xffu : (live <= 0 ? 0u : (1u << live) - 1u);
                  hits &= mword >> (8 * u);
                  while (hits) {
                    const int j = __ffs((int)hits) - 1;
                    hits &= hits - 1u;
                    float w = v[8 * u];
#pragma unroll
                    for (int jj = 1; jj < 8; ++jj) w = j == jj ? v[8 * u + jj] : w;
                    if (c < kCandCap) {
                      p.cand_score[cbase + c] = w;
                      p.cand_id[cbase + c] = id0 + 8 * u + j;
                    }
                    ++c;
                  }
                }
              }
            }
          }
        }
        // accumulator fully consumed: hand it back to the (leader's) MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2 && crank != 0) bar_arrive_remote(bar_acce + 8 * slot, 0);
          else bar_arrive(bar_acce + 8 * slot);
        }
        if (MODE == kModeMax) {
          // one maximum per kMaxGroupTiles consecutive tiles of this unit
          if ((i % p.group_tiles) == p.group_tiles - 1 || i == my_tiles - 1) {
            p.gmax[(int64_t)q * p.n_groups + uig * p.groups_per_unit + i / p.group_tiles] = rmax;
            rmax = VS_NEG_INF;
          }
          rmax_l[mt & 15] = rmax;
        }
        if (FILT) {
          if (RES) cnt_l[mt] = c;
          else p.cand_cnt[(int64_t)list * q_total + q] = c;
        }
      }
    }
    if (MODE == kModeMax) {   // groups this unit has no tiles for
      for (int g = (my_tiles + p.group_tiles - 1) / p.group_tiles; g < p.groups_per_unit; ++g)
        for (int mt = grp; mt < m_count; mt += kEpiGroups)
          p.gmax[(int64_t)(((m_first + mt) * CG + crank) * kTileM + row) * p.n_groups + uig * p.groups_per_unit + g] =
              VS_NEG_INF;
    }
    // move this thread's private candidates into the dense per-query lists
    if (FILT) {
      for (int mt = grp; mt < m_count; mt += kEpiGroups) {   // this warp group's query tiles
        const int q = ((m_first + mt) * CG + crank) * kTileM + row;
        if (q >= p.nq) continue;
        const int64_t cbase = ((int64_t)list * q_total + q) * kCandCap;
        int c = RES ? cnt_l[mt] : p.cand_cnt[(int64_t)list * q_total + q];
        if (c > kCandCap) { p.overflow[q] = 1; c = kCandCap; }
        if (c == 0) continue;
        const int base = atomicAdd(p.gcount + q, c);
        if (base + c > kGlobalCap) { p.overflow[q] = 1; continue; }
        for (int e = 0; e < c; ++e) {
          p.glist_s[(int64_t)q * kGlobalCap + base + e] = p.cand_score[cbase + e];
          p.glist_i[(int64_t)q * kGlobalCap + base + e] = p.cand_id[cbase + e];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();          // the peer's smem / TMEM stay alive until the leader is done
  if (warp == 2) {
    tc_fence_after();
    Ops::dealloc(tmem_base);
  }
}

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

// tau[q] = the ks-th largest of query q's pass-1 group maxima (-inf with fewer than ks groups).
// One warp per query: the values are staged in shared memory and the answer is built bit by bit
// on the order-preserving uint32 encoding (32 counting rounds), so the cost does not depend on
// the data.  Also clears this search's uncertified-query counter.
constexpr int kTauWarps = 4;
__global__ void __launch_bounds__(kTauWarps * 32)
tau_select_kernel(const float* __restrict__ gmax, int n_groups, int ks, int B, float* __restrict__ tau,
                  int32_t* __restrict__ bad) {
  extern __shared__ uint32_t tsm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (blockIdx.x == 0 && threadIdx.x == 0 && bad != nullptr) bad[0] = 0;
  const int b = blockIdx.x * kTauWarps + warp;
  if (b >= B) return;
  uint32_t* v = tsm + (size_t)warp * n_groups;
  const float* src = gmax + (int64_t)b * n_groups;
  for (int i = lane; i < n_groups; i += 32) v[i] = enc_key(src[i]);
  __syncwarp();
  uint32_t T = 0;
  if (n_groups >= ks) {
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = T | (1u << bit);
      int c = 0;
      for (int i = lane; i < n_groups; i += 32) c += v[i] >= cand;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c >= ks) T = cand;
    }
  }
  if (lane == 0) tau[b] = n_groups >= ks ? dec_key(T) : VS_NEG_INF;
}

// Between the two ranges of pass 2: the survivors of the first range (every row of it at or above
// tau) are a far larger sample than pass 1's, so their ks-th largest key is a tighter -- and still
// valid -- lower bound of the ks-th best key overall; the second range filters against it.
// One warp per query over its dense survivor list.
__global__ void __launch_bounds__(kTauWarps * 32)
tau_refine_kernel(const float* __restrict__ glist_s, const int32_t* __restrict__ gcount, int ks, int B,
                  float* __restrict__ tau) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * kTauWarps + warp;
  if (b >= B) return;
  const int n = min(gcount[b], kGlobalCap);
  if (n < ks) return;
  const float* v = glist_s + (int64_t)b * kGlobalCap;
  uint32_t T = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = T | (1u << bit);
    int c = 0;
    for (int i = lane; i < n; i += 32) c += enc_key(v[i]) >= cand;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= ks) T = cand;
  }
  if (lane == 0) tau[b] = fmaxf(tau[b], dec_key(T));
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
//      beta = the kc-th candidate's 16-bit key, or tau when fewer than kc rows passed the filter
//      (no row outside the candidate set has a 16-bit key above beta).  Queries that cannot be
//      certified (or whose candidate buffers overflowed) are appended to bad[1..], bad[0] counts them.
constexpr int kFinishThreads = 256;
constexpr int kMaxCand = 256;

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
  const float* tau;
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
  uint32_t* sk = reinterpret_cast<uint32_t*>(fsm);                         // encoded 16-bit keys
  int* si = reinterpret_cast<int*>(sk + kGlobalCap);
  float4* qs = reinterpret_cast<float4*>(si + kGlobalCap);                 // prepared query, ld floats
  __shared__ float keys[kMaxCand];
  __shared__ int ids[kMaxCand];
  __shared__ int sel[kMaxCand];
  __shared__ int wcount[2][kFinishThreads / 32];
  __shared__ int cnt_hi, cnt_eq, have_sh;
  __shared__ float kth_sh, qsq_sh;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool l2 = p.metric == VS_METRIC_EUCLIDEAN;
  const int n = min(p.gcount[b], kGlobalCap);
  for (int i = tid; i < n; i += kFinishThreads) {
    sk[i] = enc_key(p.glist_s[(int64_t)b * kGlobalCap + i]);
    si[i] = p.glist_i[(int64_t)b * kGlobalCap + i];
  }
  if (tid == 0) { cnt_hi = 0; cnt_eq = 0; have_sh = 0; kth_sh = VS_NEG_INF; }
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
  // ---- 1. the kc best survivors by 16-bit key
  const bool full = n >= p.kc;
  uint32_t T = 0;
  if (full) {
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = T | (1u << bit);
      int c = 0;
      for (int i = tid; i < n; i += kFinishThreads) c += sk[i] >= cand;
      c = __reduce_add_sync(0xffffffffu, c);
      if (lane == 0) wcount[bit & 1][warp] = c;
      __syncthreads();
      int tot = 0;
#pragma unroll
      for (int w = 0; w < kFinishThreads / 32; ++w) tot += wcount[bit & 1][w];
      if (tot >= p.kc) T = cand;
    }
  }
  for (int i = tid; i < n; i += kFinishThreads)
    if (!full || sk[i] > T) sel[atomicAdd(&cnt_hi, 1)] = i;
  __syncthreads();
  if (full) {
    const int hi = cnt_hi;                       // < kc: fewer than kc keys exceed the kc-th largest
    for (int i = tid; i < n; i += kFinishThreads)
      if (sk[i] == T) { const int pos = hi + atomicAdd(&cnt_eq, 1); if (pos < p.kc) sel[pos] = i; }
  }
  __syncthreads();
  const int nsel = full ? p.kc : n;
  // ---- 2. exact fp32 scores of the candidates
  const int nvec = p.ld >> 2;
  for (int e = warp; e < nsel; e += kFinishThreads / 32) {
    const int id = si[sel[e]];
    const float4* x = reinterpret_cast<const float4*>(p.rows) + (int64_t)id * nvec;
    float acc = 0.f;
    if (l2) { for (int c = lane; c < nvec; c += 32) acc = sqdiff4_acc(acc, ldg_stream(x + c), qs[c]); }
    else { for (int c = lane; c < nvec; c += 32) acc = dot4_acc(acc, ldg_stream(x + c), qs[c]); }
    const float tot = warp_sum(acc);
    float key;
    if (p.metric == VS_METRIC_COSINE) key = tot / __ldg(p.norms + id);
    else if (l2) key = -sqrtf(tot);
    else key = tot;
    if (lane == 0) { keys[e] = key; ids[e] = p.id_map ? p.id_map[id] : id; }
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
  const float beta = full ? dec_key(T) : p.tau[b];
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

__global__ void fill_f32_kernel(float* p, float v, int64_t n) {
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

struct GemmPlan {
  int mt;            // resident query tiles per unit (0 = streaming)
  int cg;            // CTAs per MMA: 1, or 2 (CTA pairs, cta_group::2)
  int stages;
  size_t smem;
  int tn;            // database rows per tile (MMA N)
};

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
  const size_t limit = 227 * 1024 - 1024 /*alignment*/ - 512 /*barriers*/ - (l2 ? std::max(sq_res, sq_str) : 0);
  const int um_tiles = m_tiles / cg;
  if (kchunks <= 4) {   // K <= 256: resident queries
    int mt = kchunks <= 2 ? 4 : 1;
    while (mt > um_tiles) mt >>= 1;
    if (mt < 1) mt = 1;
    const size_t a = (size_t)mt * kchunks * kChunkBytes;
    const size_t stage = (size_t)kchunks * (kResTN / cg) * 128;  // kResTN / cg database rows per CTA
    int stages = (int)((limit - a) / stage);
    if (stages > 4) stages = 4;
    if (stages >= 2) { *plan = {mt, cg, stages, a + stages * stage + 1024 + 512 + sq_res, kResTN}; return; }
  }
  // streaming: query chunk + this CTA's share of the 256-row database chunk
  const size_t stage = (size_t)(cg == 2 ? 2 : 3) * kChunkBytes;
  int stages = (int)(limit / stage);
  if (stages > 6) stages = 6;
  *plan = {0, cg, stages, stages * stage + 1024 + 512 + sq_str, 256};
}

// rows per TMA box of the database operand: RESIDENT CTAs load 128 / cg rows per tile in one
// box, STREAMING CTAs 256 / cg rows as one or two boxes of 128
static int x_box_rows(const GemmPlan& plan) { return plan.mt > 0 ? std::min(128, kResTN / plan.cg) : 128; }

template <int MT, int MODE, int CG>
static int launch_gemm_tmc(const CUtensorMap& mq, const CUtensorMap& mx, const GemmParams& p, int grid, size_t smem,
                           cudaStream_t stream) {
  auto kern = gemm_topk_kernel<MT, MODE, CG>;
  VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(gemm_threads(MT > 0));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  {
    ProfScope prof(kProfGemm, stream);
    VS_CUDA(cudaLaunchKernelEx(&cfg, kern, mq, mx, p));
  }
  count_launch();
  VS_CHECK_LAUNCH();
  return VS_OK;
}

template <int MT, int MODE>
static int launch_gemm_tm(const CUtensorMap& mq, const CUtensorMap& mx, const GemmParams& p, int cg, int grid,
                          size_t smem, cudaStream_t stream) {
  return cg == 2 ? launch_gemm_tmc<MT, MODE, 2>(mq, mx, p, grid, smem, stream)
                 : launch_gemm_tmc<MT, MODE, 1>(mq, mx, p, grid, smem, stream);
}

template <int MT>
static int launch_gemm_t(const CUtensorMap& mq, const CUtensorMap& mx, const GemmParams& p, int cg, int grid,
                         size_t smem, cudaStream_t stream) {
  switch (p.mode) {
    case kModeFilter: return launch_gemm_tm<MT, kModeFilter>(mq, mx, p, cg, grid, smem, stream);
    case kModeMax: return launch_gemm_tm<MT, kModeMax>(mq, mx, p, cg, grid, smem, stream);
#ifdef VS_GEMM_DEBUG_MODES
    case kModeNop: return launch_gemm_tm<MT, kModeNop>(mq, mx, p, cg, grid, smem, stream);
    case kModeHalf: return launch_gemm_tm<MT, kModeHalf>(mq, mx, p, cg, grid, smem, stream);
#endif
    default: return launch_gemm_tm<MT, kModeDump>(mq, mx, p, cg, grid, smem, stream);
  }
}

// units (CTAs or CTA pairs) a launch uses, its query groups and candidate lists per query
static void gemm_units(const GemmPlan& plan, int m_tiles, int n_tiles, int num_sms, int* units_out, int* ngroups_out,
                       int* lists_out) {
  int units = num_sms / plan.cg;
  int ngroups = 1, lists;
  if (plan.mt > 0) {
    const int um_tiles = m_tiles / plan.cg;
    ngroups = (um_tiles + plan.mt - 1) / plan.mt;
    const int64_t want = (int64_t)n_tiles * ngroups;
    if (want < units) units = (int)want;
    if (units < ngroups) units = ngroups;
    lists = (units + ngroups - 1) / ngroups;
  } else {
    if (n_tiles < units) units = n_tiles;
    lists = units;
  }
  *units_out = units; *ngroups_out = ngroups; *lists_out = lists;
}

static int launch_gemm(const GemmPlan& plan, const CUtensorMap& mq, const CUtensorMap& mx, GemmParams p,
                       int num_sms, int* lists_out, cudaStream_t stream) {
  p.stages = plan.stages;
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
  switch (plan.mt) {
    case 0: return launch_gemm_t<0>(mq, mx, p, plan.cg, grid, plan.smem, stream);
    case 1: return launch_gemm_t<1>(mq, mx, p, plan.cg, grid, plan.smem, stream);
    case 2: return launch_gemm_t<2>(mq, mx, p, plan.cg, grid, plan.smem, stream);
    case 4: return launch_gemm_t<4>(mq, mx, p, plan.cg, grid, plan.smem, stream);
  }
  set_error("internal: bad GEMM plan");
  return VS_ERR_INVALID;
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
static int cand_count(int kk) { return std::max(2 * kk, kk + 22); }
// rows the pass-1 threshold is guaranteed to admit (see gemm_block)
static int sample_rank(int kk) { return kk + std::max(6, kk / 2); }

bool gemm_supported(const vs_store* s, int64_t n, int B, int kk) {
  if (!gemm_enabled() || !s->shadow) return false;
  if (B < gemm_min_batch(s) || s->dim > 8192) return false;
  if (cand_count(kk) > kMaxCand) return false;            // k <= 128
  if (n < 65536) return false;                              // small stores: the scan is enough
  return true;
}

struct Ws {
  std::vector<std::pair<void**, size_t>> items;
  unsigned char* base = nullptr;
  cudaStream_t stream = nullptr;
  template <typename T> void want(T** slot, size_t count) { items.push_back({(void**)slot, count * sizeof(T)}); }
  int alloc(cudaStream_t st) {
    stream = st;
    size_t total = 0;
    for (auto& it : items) total += (size_t)round_up((int64_t)it.second, 1024);
    VS_CUDA(cudaMallocAsync((void**)&base, total ? total : 1024, st));
    size_t off = 0;
    for (auto& it : items) { *it.first = base + off; off += (size_t)round_up((int64_t)it.second, 1024); }
    return VS_OK;
  }
  ~Ws() { if (base) cudaFreeAsync(base, stream); }
};

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
  const int kc = (int)std::min<int64_t>(kc_want > 0 ? kc_want : cand_count(kk), n);
  GemmPlan plan;
  plan_gemm(kch, m_tiles, cg, l2, &plan);
  const int tn = plan.tn;
  const int n_tiles = (int)((n + tn - 1) / tn);
  // pass-1 sample.  tau[q] = the ks-th largest per-group maximum of the sample: ks distinct rows
  // reach it, so it is a lower bound of the ks-th best 16-bit key of the whole database.  ks only
  // has to exceed k by a margin (the filter may pass FEWER than kc rows: then no row outside the
  // candidate set exceeds tau and the certification uses beta = tau); a wide retry (kc_want > 0)
  // asks for all of its kc candidates.  The filter passes about ks / f rows per query, i.e. a
  // fraction 1024 ks / (f N) of the 32 x 32 epilogue chunks take the rare path (~r chunk times
  // each), while pass 1 costs ~0.8 f of a full pass: minimising 0.8 f + r 1024 ks / (f N) gives
  // f = sqrt(1280 r ks / N); r ~ 0.6 for the register-based rare path (tuned on the B200,
  // profiles/r02_k3_probe.txt).  At least 4 ks groups, at most 4096 (K4's capacity), whole tiles.
  // With a row mask only n_live rows take part: the sample must hold enough of THOSE.
  int ks = kc_want > 0 ? kc : std::min(kc, sample_rank(kk));
  if (const char* e = getenv("B200VS_GEMM_KS")) { if (*e) ks = std::max(1, std::min(kc, atoi(e))); }   // diagnostic
  const double live_frac = std::max(1e-9, std::min(1.0, (double)n_live / (double)n));
  double f = std::sqrt(768.0 * ks / ((double)n * live_frac)) ;
  // K > 256 (STREAMING): the MMAs of a tile take several times longer than its epilogue, so the
  // rare path is hidden and the sample only has to keep the survivors (about 1.2 ks / f per
  // query) within the candidate buffers (64 per CTA pair and query) and K4's capacity.
  if (kch > 4) f = std::min(f, ks / (750.0 * live_frac));
  if (const char* e = getenv("B200VS_GEMM_SAMPLE")) { if (*e) f = atof(e); }   // diagnostic override
  if (f > 0.5) f = 0.5;
  if (f < 1.0 / 128) f = 1.0 / 128;
  const int full_tiles = (int)(n / tn);
  int s_tiles = (int)std::min<int64_t>(std::max<int64_t>((int64_t)(f * (double)n) / tn, (int64_t)(4 * ks / live_frac)),
                                       4096 * kMaxGroupTiles);
  if (s_tiles > full_tiles) s_tiles = full_tiles;
  // the sample is spread over the whole row range (every `s_stride`-th tile), not a prefix: on a
  // time-ordered or clustered ingest a prefix can be unlike the rest and give a useless threshold
  const int s_stride = s_tiles > 0 ? std::max(1, full_tiles / s_tiles) : 1;
  int s_units, s_ngroups, s_lists;
  gemm_units(plan, m_tiles, s_tiles, s->num_sms, &s_units, &s_ngroups, &s_lists);
  const int gt = s_tiles / kMaxGroupTiles >= 4 * ks ? kMaxGroupTiles : 1;
  const int s_groups = s_lists * (((s_tiles + s_lists - 1) / s_lists + gt - 1) / gt);
  const bool sampled = s_tiles / gt >= 2 * ks && s_groups <= 4096 && live_frac >= 0.2;
  // Too few rows for a useful threshold (fewer than 2 ks sample groups, or a filter that leaves
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
  __nv_bfloat16* qb; float *qerr, *qlen, *tau, *gmax, *cs, *gls; int32_t *ci, *ccnt, *ovf, *gli, *gcnt;
  ws.want(&qb, (size_t)rows_padded * K);
  ws.want(&qerr, (size_t)rows_padded);
  ws.want(&qlen, (size_t)rows_padded);
  ws.want(&tau, (size_t)rows_padded);
  ws.want(&gmax, (size_t)rows_padded * s_groups);
  ws.want(&cs, (size_t)max_lists * rows_padded * kCandCap);
  ws.want(&ci, (size_t)max_lists * rows_padded * kCandCap);
  ws.want(&ccnt, (size_t)max_lists * rows_padded);
  ws.want(&gls, (size_t)rows_padded * kGlobalCap);
  ws.want(&gli, (size_t)rows_padded * kGlobalCap);
  ws.want(&ovf, (size_t)rows_padded);      // cleared by the prep kernel, like gcnt
  ws.want(&gcnt, (size_t)rows_padded);
  if (int rc = ws.alloc(stream)) return rc;

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
  p.cand_score = cs; p.cand_id = ci; p.cand_cnt = ccnt; p.overflow = ovf; p.tau = tau;
  p.glist_s = gls; p.glist_i = gli; p.gcount = gcnt;
  p.sqnorms = l2 ? (const float*)s->sqnorms.ptr() : nullptr;
  p.row_mask = row_mask;

  if (plan.mt == 0)   // STREAMING keeps its running candidate counts in global memory
    VS_CUDA(cudaMemsetAsync(ccnt, 0, (size_t)max_lists * rows_padded * 4, stream));
  // pass 1: per-query maxima of the sample tiles -> tau = ks-th largest
  p.mode = kModeMax; p.n_tiles = s_tiles; p.tile_stride = s_stride; p.gmax = gmax; p.group_tiles = gt;
  if (int rc = launch_gemm(plan, mq, mx, p, s->num_sms, nullptr, stream)) return rc;
  tau_select_kernel<<<(B + kTauWarps - 1) / kTauWarps, kTauWarps * 32, (size_t)kTauWarps * s_groups * 4, stream>>>(
      gmax, s_groups, ks, B, tau, bad);
  count_launch();
  VS_CHECK_LAUNCH();
  // diagnostic (timing only, results are wrong): no row passes the filter, so pass 2 never
  // takes its rare path
  if (const char* e = getenv("B200VS_GEMM_TAU_INF")) {
    if (*e == '1') fill_f32_kernel<<<(rows_padded + 255) / 256, 256, 0, stream>>>(tau, __builtin_inff(), rows_padded);
  }
  // pass 2: threshold filter over all rows
  p.mode = kModeFilter; p.n_tiles = n_tiles; p.tile_stride = 1; p.gmax = nullptr;
#ifdef VS_GEMM_DEBUG_MODES
  if (const char* e = getenv("B200VS_GEMM_DBGMODE")) { if (*e == '3' || *e == '4') p.mode = atoi(e); }
#endif
  // RESIDENT kernels on a large store run pass 2 as two ranges: a head of ~15 % of the tiles
  // filtered against pass 1's threshold, tau_refine (the head's survivors are a 4x larger sample
  // than pass 1's), then the rest against the tighter threshold -- the rare path of the epilogue
  // fires ~3x less often overall (profiles/r02_k3_probe.txt).  Survivors of both ranges land in the
  // same dense per-query lists; every row at or above the FINAL tau is among them.
  int head_tiles = 0;
  if (plan.mt > 0 && n_tiles >= 16384) head_tiles = (int)(0.15 * n_tiles);
  if (const char* e = getenv("B200VS_GEMM_HEAD")) { if (*e) head_tiles = (int)(atof(e) * n_tiles); }   // diagnostic
  if (head_tiles > 0 && head_tiles < n_tiles) {
    p.n_tiles = head_tiles;
    if (int rc = launch_gemm(plan, mq, mx, p, s->num_sms, nullptr, stream)) return rc;
    tau_refine_kernel<<<(B + kTauWarps - 1) / kTauWarps, kTauWarps * 32, 0, stream>>>(gls, gcnt, ks, B, tau);
    count_launch();
    VS_CHECK_LAUNCH();
    p.tile_first = head_tiles; p.n_tiles = n_tiles - head_tiles;
  }
  if (int rc = launch_gemm(plan, mq, mx, p, s->num_sms, nullptr, stream)) return rc;
  // candidate selection + K5 + final ordering + certification, one CTA per query
  {
    FinishParams f = {};
    f.q = q; f.dim = s->dim; f.ld = s->ld; f.metric = s->metric; f.B = B; f.k = kk; f.kc = kc; f.n_rows = n;
    f.rows = (const float*)s->rows.ptr(); f.norms = (const float*)s->norms.ptr(); f.id_map = s->id_map();
    f.glist_s = gls; f.glist_i = gli; f.gcount = gcnt;
    f.tau = tau; f.overflow = ovf; f.qerr = qerr; f.qlen = qlen; f.bounds = s->bounds;
    // fp32 accumulation error of the tensor core and of the exact kernels, relative to ||u|| ||v||
    // (tests/test_gemm_gpu.py measures the tensor core's share against float64)
    f.slack = (l2 ? 8.f : 4.f) * (float)s->dim * 5.9604645e-8f + 1e-6f;
    f.certify = certify ? 1 : 0;
    f.out_s = out_scores; f.out_i = out_ids; f.out_stride = out_stride; f.bad = bad;
    const size_t smem = (size_t)kGlobalCap * 8 + (size_t)s->ld * 4;
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
  const int kc = (int)std::min<int64_t>(cand_count(kk), pb.n);
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
  if (cand_count(kk) > kMaxCand) { set_error("invalid argument: k too large for the GEMM path (k <= 128)"); return VS_ERR_INVALID; }
  // the fp8 variant keeps 4x the candidates for the exact rescoring
  const int kc_want = fp8 ? std::min(4 * cand_count(kk), kMaxCand) : 0;
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
  if (int rc = ws.alloc(stream)) return rc;
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
