// K3 gemm_topk -- the tcgen05 / TMEM kernel and its launch templates (see gemm_topk.cu for the
// algorithm).  Included by the two instantiation units gemm_inst_plain.cu (GEN = 0) and
// gemm_inst_general.cu (GEN = 1), which the build compiles in parallel, and by gemm_topk.cu (host
// side: plan, parameters).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include "common.cuh"

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
constexpr int kLevels = 12;              // threshold ladder per query (adaptive filter, see tau_select_kernel)
// Hit queues (shared memory): one single-producer ring per epilogue warp, drained by warp 2, which
// climbs the threshold ladder (global atomics, ~2 L2 round trips per hit) OFF the epilogue's
// critical path.  An entry is (16-bit key, query); a full ring drops the entry -- the ladder only
// ever needs a LOWER bound of the rows at a level.
constexpr int kHqEntries = 16;
__host__ __device__ constexpr int hq_bytes(bool resident) {
  return ((4 * epi_groups(resident)) * (kHqEntries * 8 + 8) + 16 + 127) / 128 * 128;
}
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
  int stages;               // ring depth (RESIDENT: whole database tiles; STREAMING: one-chunk slots)
  int a_res_k;              // STREAMING: K chunks of the unit's query tile kept resident in shared memory
  int m_per_unit;           // STREAMING: query tiles (per unit) each unit serves; units split into
                            //   ceil(um_tiles / m_per_unit) query groups like RESIDENT's
  uint32_t idesc;           // tcgen05 instruction descriptor (operand format, M, N)
  // kModeFilter: the threshold of query q is the (order-preserving uint32 encoding of the) value in
  // tau_cur[q].  It starts at pass 1's bound and RISES while pass 2 runs: lvl[q][1..] is an ascending
  // ladder of trial thresholds, lvl_cnt[q][j] counts the rows found so far whose key reaches lvl[q][j];
  // once `adapt_rank` rows reach a level, that level is a valid lower bound of the adapt_rank-th best
  // key and becomes the threshold of every CTA (atomicMax).  Correctness never depends on the ladder:
  // the certification only needs "no row outside the candidate set has a key above the FINAL tau_cur".
  uint32_t* tau_cur;        // (m_tiles*128,)
  const float* lvl;         // (m_tiles*128, kLevels)
  int32_t* lvl_cnt;         // (m_tiles*128, kLevels), zeroed by tau_select_kernel
  int adapt_rank;
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
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
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

// GEN = 0: plain keys (cosine / dot_product, every row takes part) -- the hot loop carries no
// trace of the other cases.  GEN = 1: euclidean keys (p.sqnorms) and / or a row mask (p.row_mask).
template <int MT, int MODE, int CG, int GEN>
__global__ void __launch_bounds__(gemm_threads(MT > 0), 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x,
                 const GemmParams p) {
  constexpr bool RES = MT > 0;
  constexpr int kEpiGroups = epi_groups(RES);
  constexpr bool FILT = MODE == kModeFilter || MODE == kModeHalf;
  constexpr int TN = RES ? kResTN : 256;                  // MMA N = database rows per tile
  constexpr int TN_LOCAL = TN / CG;                       // rows of the tile this CTA loads
  constexpr int SLOTS = kTmemCols / TN;
  static_assert(SLOTS % kEpiGroups == 0, "every epilogue group owns the same number of TMEM accumulator slots");
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
  // STREAMING: the first a_res_k K chunks of the unit's query tile stay resident (only when the unit
  // serves ONE query tile); everything else streams through a ring of one-chunk slots
  constexpr size_t SLOT_BYTES = B_CHUNK_BYTES > kChunkBytes ? B_CHUNK_BYTES : kChunkBytes;
  const int a_res_k = RES ? 0 : p.a_res_k;
  const size_t a_bytes = RES ? (size_t)MT * kch * kChunkBytes : (size_t)a_res_k * kChunkBytes;
  const size_t stage_bytes = RES ? (size_t)kch * B_CHUNK_BYTES : SLOT_BYTES;
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
  // hit queues: rings, then heads, tails, the count of finished epilogue warps
  constexpr int NEW = 4 * kEpiGroups;                                  // epilogue warps
  float2* hq_ring = reinterpret_cast<float2*>(tail + 512);
  uint32_t* hq_head = reinterpret_cast<uint32_t*>(hq_ring + NEW * kHqEntries);
  uint32_t* hq_tail = hq_head + NEW;
  uint32_t* hq_done = hq_tail + NEW;
  float* sq_base = reinterpret_cast<float*>(tail + 512 + hq_bytes(RES));

  // ---- work assignment, in units of one CTA (CG=1) or one CTA pair (CG=2)
  // query tiles are counted per unit: unit tile m covers the 128-row tiles m*CG + crank
  const int unit = blockIdx.x / CG;
  const int n_units = gridDim.x / CG;
  const int um_tiles = p.m_tiles / CG;                                 // m_tiles is a multiple of CG
  const int ngroups = p.ngroups;
  const int group = unit % ngroups;
  const int uig = unit / ngroups;                                      // unit index inside its group
  const int units_in_group = (n_units - group + ngroups - 1) / ngroups;
  const int m_span = RES ? MT : p.m_per_unit;                          // query tiles per group
  const int m_first = group * m_span;
  const int m_count = min(m_span, um_tiles - m_first);
  int my_tiles = 0;
  if (uig < p.n_tiles) my_tiles = (p.n_tiles - uig + units_in_group - 1) / units_in_group;
  // accumulator classes in use (see the MMA issuer): a unit with ONE query tile has one class, which
  // then owns every TMEM slot (otherwise half of them -- and the MMA / epilogue overlap -- would idle)
  const int p_act = m_count == 1 ? 1 : kEpiGroups;
  const int spc = SLOTS / p_act;                          // slots per class

  if (threadIdx.x == 0) {
    bar_init(bar_a, 1);
    // a RESIDENT stage is released by every MMA-issuing warp (one tcgen05.commit each)
    for (int i = 0; i < p.stages; ++i) { bar_init(bar_full + 8 * i, 1); bar_init(bar_empty + 8 * i, RES ? kResIssuers : 1); }
    for (int i = 0; i < SLOTS; ++i) { bar_init(bar_accf + 8 * i, 1); bar_init(bar_acce + 8 * i, 4 * CG); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < NEW; ++i) { hq_head[i] = 0; hq_tail[i] = 0; }
    *hq_done = 0;
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
      } else if (a_res_k > 0) {
        if (crank == 0) bar_expect_tx(bar_a, (uint32_t)(CG * a_res_k * kChunkBytes));
        for (int kc = 0; kc < a_res_k; ++kc)
          Ops::load(s_u32(a_res + (size_t)kc * kChunkBytes), &map_q, kc * kstep, (m_first * CG + crank) * kTileM, bar_a);
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
              if (kc >= a_res_k) {                      // this K chunk of the queries is streamed
                bar_wait(bar_empty + 8 * st, ph ^ 1u);
                if (crank == 0) bar_expect_tx(bar_full + 8 * st, (uint32_t)(CG * kChunkBytes));
                Ops::load(s_u32(ring + (size_t)st * stage_bytes), &map_q, kc * kstep,
                          ((m_first + mt) * CG + crank) * kTileM, bar_full + 8 * st);
                if (++st == p.stages) { st = 0; ph ^= 1u; }
              }
              bar_wait(bar_empty + 8 * st, ph ^ 1u);
              if (crank == 0) bar_expect_tx(bar_full + 8 * st, (uint32_t)(CG * B_CHUNK_BYTES));
              unsigned char* dst = ring + (size_t)st * stage_bytes;
              Ops::load(s_u32(dst), &map_x, kc * kstep, row0, bar_full + 8 * st);
              if (TN_LOCAL == 256)   // 256 rows = two boxes of 128
                Ops::load(s_u32(dst + kChunkBytes), &map_x, kc * kstep, row0 + 128, bar_full + 8 * st);
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
    // (query-tile parity: each warp always fills the same TMEM slots, see `uses` below); MMAs of
    // different accumulators are independent, both warps read the same shared-memory operands and
    // each releases the stage with its own commit.
    const int issuer = warp == 1 ? 0 : 1;
    if (crank == 0 && my_tiles > 0 && m_count > 0 && elect_one()) {
      const uint32_t idesc = p.idesc;
      if (RES || a_res_k > 0) { bar_wait(bar_a, 0); tc_fence_after(); }
      int st = 0;
      uint32_t ph = 0;
      // Accumulator slots.  Query tile mt belongs to class mt % kEpiGroups: ONE epilogue group drains
      // it and (RESIDENT, two issuers) one issuing warp fills it, and the class owns the TMEM slots
      // cls, cls + p_act, ... which it uses round-robin (p_act = classes in use: 1 for a unit with a
      // single query tile, which then cycles through all slots).  Every phase of a slot's full / empty
      // barriers is therefore observed by the same waiter -- an mbarrier parity wait must not skip a
      // phase (with slots handed out by a global sequence number and an odd number of query tiles
      // per unit the groups alternated slots, a wait could match a phase two completions old, and
      // the kernel read accumulators early and finally deadlocked).
      int uses[kEpiGroups];                             // accumulators issued so far, per class
#pragma unroll
      for (int c = 0; c < kEpiGroups; ++c) uses[c] = 0;
      for (int i = 0; i < my_tiles; ++i) {
        if (RES) {
          bar_wait(bar_full + 8 * st, ph);
          tc_fence_after();
        }
        for (int mt = 0; mt < m_count; ++mt) {
          const int cls = mt % kEpiGroups;
          if (RES && kResIssuers == 2 && (cls & 1) != issuer) continue;   // the other issuing warp's
          int u = 0;
#pragma unroll
          for (int c = 0; c < kEpiGroups; ++c) if (c == cls) u = uses[c]++;
          const int slot = cls + p_act * (u % spc);
          const uint32_t aph = (uint32_t)(u / spc) & 1u;
          bar_wait(bar_acce + 8 * slot, aph ^ 1u);      // epilogues drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(slot * TN);
          for (int kc = 0; kc < kch; ++kc) {
            uint32_t a_addr, b_addr;
            int st_a_pending = -1;
            if (RES) {
              a_addr = s_u32(a_res + (size_t)(mt * kch + kc) * kChunkBytes);
              b_addr = s_u32(ring + (size_t)st * stage_bytes + (size_t)kc * B_CHUNK_BYTES);
            } else {
              int st_a = -1;                            // ring slot of a streamed query chunk
              if (kc >= a_res_k) {
                bar_wait(bar_full + 8 * st, ph);
                a_addr = s_u32(ring + (size_t)st * stage_bytes);
                st_a = st;
                if (++st == p.stages) { st = 0; ph ^= 1u; }
              } else {
                a_addr = s_u32(a_res + (size_t)kc * kChunkBytes);
              }
              bar_wait(bar_full + 8 * st, ph);
              tc_fence_after();
              b_addr = s_u32(ring + (size_t)st * stage_bytes);
              st_a_pending = st_a;
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
              if (st_a_pending >= 0) Ops::commit(bar_empty + 8 * st_a_pending);   // slots free once these MMAs retire
              Ops::commit(bar_empty + 8 * st);
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
  } else if (warp == 2 && FILT) {
    // ================================================================ threshold ladder (hit-queue consumer)
    // Lane l handles entry l of whatever a ring holds: it counts the row at every ladder level its
    // key reaches; the thread whose count completes `adapt_rank` rows at a level raises the query's
    // threshold to it for every CTA.
    uint32_t my_tail = 0;                               // lane w: how far producer warp w's ring is drained
    for (;;) {
      const uint32_t done = *reinterpret_cast<volatile uint32_t*>(hq_done);
      bool any = false;
      for (int w = 0; w < NEW; ++w) {
        uint32_t head = *reinterpret_cast<volatile uint32_t*>(hq_head + w);
        head = __shfl_sync(0xffffffffu, head, 0);
        const uint32_t tl = __shfl_sync(0xffffffffu, my_tail, w);
        const uint32_t navail = head - tl;
        if (navail == 0) continue;
        any = true;
        __threadfence_block();
        float2 e = make_float2(0.f, 0.f);
        if ((uint32_t)lane < navail) {
          const volatile float* src = reinterpret_cast<const volatile float*>(hq_ring + w * kHqEntries + ((tl + lane) % kHqEntries));
          e.x = src[0];
          e.y = src[1];
        }
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile uint32_t*>(hq_tail + w) = head;   // slots free again
        if (lane == w) my_tail = head;
        if ((uint32_t)lane < navail) {
          const float key = e.x;
          const int q = __float_as_int(e.y);
          const float4* l4 = reinterpret_cast<const float4*>(p.lvl + (int64_t)q * kLevels);
          float L[kLevels];
#pragma unroll
          for (int j4 = 0; j4 < kLevels / 4; ++j4) {
            const float4 x = __ldg(l4 + j4);
            L[4 * j4] = x.x; L[4 * j4 + 1] = x.y; L[4 * j4 + 2] = x.z; L[4 * j4 + 3] = x.w;
          }
#pragma unroll
          for (int j = 1; j < kLevels; ++j) {
            if (key >= L[j]) {
              if (atomicAdd(p.lvl_cnt + (int64_t)q * kLevels + j, 1) + 1 == p.adapt_rank)
                atomicMax(p.tau_cur + q, enc_key(L[j]));
            }
          }
        }
      }
      if (!any) {
        if (done == (uint32_t)NEW) break;
        __nanosleep(200);
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
    const int list = uig;                               // candidate list of this unit (inside its query group)
    const int64_t q_total = (int64_t)p.m_tiles * kTileM;
    constexpr int NST = RES ? MT : 1;
    int cnt_l[NST];                                     // RESIDENT: private candidates per query tile
    float rmax_l[MODE == kModeMax ? 16 : 1];            // pass 1: running maximum per query tile
    if (MODE == kModeMax)
      for (int mt = 0; mt < 16; ++mt) rmax_l[mt] = VS_NEG_INF;
    if (RES && FILT) {
      for (int mt = 0; mt < NST; ++mt) cnt_l[mt] = 0;
    }
    int my_uses = 0;                                    // accumulators of this group's class so far
    uint32_t hq_h = 0;                                  // this warp's ring: entries published so far
    float2* my_ring = hq_ring + (warp - 4) * kHqEntries;
    const bool l2 = GEN != 0 && p.sqnorms != nullptr;
    const bool masked = GEN != 0 && p.row_mask != nullptr;
    // threshold of the NEXT accumulator this thread handles, loaded one accumulator ahead (an L2
    // round trip that must not sit between "accumulator ready" and its first tcgen05.ld)
    uint32_t tau_next = 0xff800000u;                    // enc_key(+inf)
    if (FILT && m_count > grp)
      tau_next = ld_relaxed_u32(p.tau_cur + ((m_first + grp) * CG + crank) * kTileM + row);
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
      for (int mt = 0; mt < m_count; ++mt) {
        if ((mt % kEpiGroups) != grp) continue;
        const int u = my_uses++;
        const int slot = grp + p_act * (u % spc);        // the class's slots, round-robin (see the MMA issuer)
        const uint32_t aph = (uint32_t)(u / spc) & 1u;
        const int q = ((m_first + mt) * CG + crank) * kTileM + row;
        const int64_t cbase = ((int64_t)list * q_total + q) * kCandCap;
        float t = __int_as_float(0x7f800000);
        int c = 0;
        if (FILT) {
          t = dec_key(tau_next);
          c = RES ? cnt_l[mt] : p.cand_cnt[(int64_t)list * q_total + q];
        }
        bar_wait(bar_accf + 8 * slot, aph);
        tc_fence_after();
        if (FILT) {                                      // prefetch the next accumulator's threshold
          int mt_n = mt + kEpiGroups;
          if (mt_n >= m_count) mt_n = grp;
          tau_next = ld_relaxed_u32(p.tau_cur + ((m_first + mt_n) * CG + crank) * kTileM + row);
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(slot * TN);
        float rmax = MODE == kModeMax ? rmax_l[mt & 15] : VS_NEG_INF;
        float va[32], vb[32];
        // Measured on the B200 and NOT kept (profiles/r02_k3_probe.txt, block c15): with the deferral the plain
        // resident kernel needs 167 registers and ran SLOWER (10 M x 128: 2.22 vs 1.92 ms, a 1.25 M-row shard:
        // 0.347 vs 0.295 ms).  -DVS_DEFER_HITS=1 builds it.
#ifndef VS_DEFER_HITS
#define VS_DEFER_HITS 0
#endif
        constexpr bool DEFER = VS_DEFER_HITS != 0 && FILT && GEN == 0 && RES;
        float dv[DEFER ? 32 : 1];                          // the deferred chunk (see below)
        int d_cc = -1;                                     // its first column, -1 = none (warp-uniform)
        float d_m = 0.f;
        // Rare path: the hits of one 32-column chunk `vv` (columns cc_ .. cc_+31 of tile nt, chunk maximum
        // m_ per lane, taking-part mask mword_), taken straight from registers.  Per 8-column group a
        // branch-free hit mask per lane, one warp-uniform vote, and a (divergent, almost always single-trip)
        // loop over the set bits that picks the value with a select chain -- no per-value branches.
        // Measured at 10 M x 128, batch 1024 (profiles/r02_k3_probe.txt): 1.96 ms vs 2.14 ms per search with
        // the first version, which re-read the 8-column groups from TMEM.
        auto process_hits = [&](float (&vv)[32], const int cc_, const float m_, const uint32_t mword_) {
          const int lim = (int)min((int64_t)TN, p.n_rows - (int64_t)nt * TN) - cc_;   // live columns
          const int32_t id0 = (int32_t)((int64_t)nt * TN + cc_);
          // hand (best key of this chunk, query) of every lane with a hit to the ladder warp
          // (chunks that reach past the end of the store are skipped: their maximum may belong
          // to a zero-filled column)
          if (lim >= 32) {
            const uint32_t hm = __ballot_sync(0xffffffffu, m_ >= t);
            const int nh = __popc(hm);
            uint32_t tl = *reinterpret_cast<volatile uint32_t*>(hq_tail + (warp - 4));
            tl = __shfl_sync(0xffffffffu, tl, 0);
            if (hq_h + nh - tl <= (uint32_t)kHqEntries) {
              if (m_ >= t)
                my_ring[(hq_h + __popc(hm & ((1u << lane) - 1u))) % kHqEntries] = make_float2(m_, __int_as_float(q));
              hq_h += nh;
              __syncwarp();
              if (lane == 0) {
                __threadfence_block();
                *reinterpret_cast<volatile uint32_t*>(hq_head + (warp - 4)) = hq_h;
              }
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint32_t hits = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) hits |= (vv[8 * u + j] >= t ? 1u : 0u) << j;
            const int live = lim - 8 * u;                       // columns of this group inside the store
            hits &= live >= 8 ? 0xffu : (live <= 0 ? 0u : (1u << live) - 1u);
            hits &= mword_ >> (8 * u);
            if (!__any_sync(0xffffffffu, hits != 0u)) continue;
            while (hits) {
              const int j = __ffs((int)hits) - 1;
              hits &= hits - 1u;
              float w = vv[8 * u];
#pragma unroll
              for (int jj = 1; jj < 8; ++jj) w = j == jj ? vv[8 * u + jj] : w;
              if (c < kCandCap) {
                p.cand_score[cbase + c] = w;
                p.cand_id[cbase + c] = id0 + 8 * u + j;
              }
              ++c;
            }
          }
        };
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
              if (masked) {
                const int64_t r0 = (int64_t)nt * TN + cc;
                mword = r0 < p.n_rows ? __ldg(p.row_mask + (r0 >> 5)) : 0u;
                // rows that do not take part never reach a maximum (pass 1's bound and the ladder's
                // counts are about TAKING-PART rows)
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = (mword >> j) & 1u ? v[j] : VS_NEG_INF;
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
                // rare path.  (Experiment, off: the first chunk of an accumulator with a hit is only COPIED
                // and handled after the accumulator has been handed back -- see DEFER above.)
                bool handled = false;
                if constexpr (DEFER) {
                  if (d_cc < 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) dv[j] = v[j];
                    d_cc = cc;
                    d_m = m;
                    handled = true;
                  }
                }
                if (!handled) process_hits(v, cc, m, mword);
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
        if constexpr (DEFER) {
          if (d_cc >= 0) process_hits(dv, d_cc, d_m, 0xffffffffu);
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
    if (FILT) {                                           // the ladder warp may stop once all of us are here
      __syncwarp();
      if (lane == 0) atomicAdd(hq_done, 1u);
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

struct GemmPlan {
  int a_res_k = 0;   // STREAMING: resident K chunks of the unit's (single) query tile
  int m_per_unit = 0;// STREAMING: query tiles per unit (1 with query groups, else all of them)
  int mt = 0;        // resident query tiles per unit (0 = streaming)
  int cg = 1;        // CTAs per MMA: 1, or 2 (CTA pairs, cta_group::2)
  int stages = 0;
  size_t smem = 0;
  int tn = 128;      // database rows per tile (MMA N)
};

// One launch of the kernel variant (plan.mt, p.mode, plan.cg) with `grid` CTAs; defined once per
// GEN in the instantiation units.
int launch_gemm_plain(const GemmPlan& plan, const CUtensorMap& mq, const CUtensorMap& mx, const GemmParams& p, int grid,
                      cudaStream_t stream);
int launch_gemm_general(const GemmPlan& plan, const CUtensorMap& mq, const CUtensorMap& mx, const GemmParams& p, int grid,
                        cudaStream_t stream);

#ifdef VS_GEMM_INSTANTIATE
template <int MT, int MODE, int CG, int GEN>
static int launch_gemm_tmc(const CUtensorMap& mq, const CUtensorMap& mx, const GemmParams& p, int grid, size_t smem,
                           cudaStream_t stream) {
  auto kern = gemm_topk_kernel<MT, MODE, CG, GEN>;
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

template <int MT, int MODE, int GEN>
static int launch_gemm_tm(const CUtensorMap& mq, const CUtensorMap& mx, const GemmParams& p, int cg, int grid,
                          size_t smem, cudaStream_t stream) {
  return cg == 2 ? launch_gemm_tmc<MT, MODE, 2, GEN>(mq, mx, p, grid, smem, stream)
                 : launch_gemm_tmc<MT, MODE, 1, GEN>(mq, mx, p, grid, smem, stream);
}

template <int MT, int GEN>
static int launch_gemm_t(const CUtensorMap& mq, const CUtensorMap& mx, const GemmParams& p, int cg, int grid,
                         size_t smem, cudaStream_t stream) {
  switch (p.mode) {
    case kModeFilter: return launch_gemm_tm<MT, kModeFilter, GEN>(mq, mx, p, cg, grid, smem, stream);
    case kModeMax: return launch_gemm_tm<MT, kModeMax, GEN>(mq, mx, p, cg, grid, smem, stream);
#ifdef VS_GEMM_DEBUG_MODES
    case kModeNop: return launch_gemm_tm<MT, kModeNop, GEN>(mq, mx, p, cg, grid, smem, stream);
    case kModeHalf: return launch_gemm_tm<MT, kModeHalf, GEN>(mq, mx, p, cg, grid, smem, stream);
#endif
    default:
      if constexpr (GEN == 0) {
        return launch_gemm_tm<MT, kModeDump, 0>(mq, mx, p, cg, grid, smem, stream);
      } else {
        set_error("internal: the score dump has no general-key variant");
        return VS_ERR_INVALID;
      }
  }
}

template <int GEN>
static int launch_gemm_variant(const GemmPlan& plan, const CUtensorMap& mq, const CUtensorMap& mx, const GemmParams& p,
                               int grid, cudaStream_t stream) {
  switch (plan.mt) {
    case 0: return launch_gemm_t<0, GEN>(mq, mx, p, plan.cg, grid, plan.smem, stream);
    case 1: return launch_gemm_t<1, GEN>(mq, mx, p, plan.cg, grid, plan.smem, stream);
    case 2: return launch_gemm_t<2, GEN>(mq, mx, p, plan.cg, grid, plan.smem, stream);
    case 4: return launch_gemm_t<4, GEN>(mq, mx, p, plan.cg, grid, plan.smem, stream);
  }
  set_error("internal: bad GEMM plan");
  return VS_ERR_INVALID;
}
#endif  // VS_GEMM_INSTANTIATE

}  // namespace vs
