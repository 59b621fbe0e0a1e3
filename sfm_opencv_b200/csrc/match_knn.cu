// match_knn.cu -- exact k=2 nearest-neighbour search over u8 SIFT descriptors on sm_100a.
//
// Replaces cv::BFMatcher(NORM_L2)::knnMatch(query, train, knn, 2) as called by the reference
// at OpenCV_SFM/NViewReconstuct.cpp:876-877 (SIFT/L2 form: TwoViewReconstruct.cpp:159-160).
//
// d2(i,j) = |q_i|^2 + |t_j|^2 - 2 q_i.t_j with q.t from tcgen05.mma.kind::i8 (u8 x u8 -> s32,
// exact).  One persistent CTA per SM; a work item is a 256-row query block of one image pair,
// swept over all 128-row train tiles of the pair:
//   warp 0       TMA producer : the 256-row A block once per item (2 x 16 KB) and the train
//                               image as 128-row B tiles (16 KB) through an mbarrier ring
//   warps 4..7   MMA issuers  : warp (h, p) issues the 4 MMAs (128x128x32) of query half h for
//                               the tiles of parity p into TMEM accumulator [p][h]; every B
//                               byte feeds 256 query rows.  Four issuers because a
//                               tcgen05.mma / commit / mbarrier wait each stall the issuing
//                               thread for 50-100 cycles (measured, tools/exp_probe.py): one
//                               thread cannot keep the tensor pipe busy at 64 cycles per MMA.
//                               One warp per accumulator, in order: parity waits on its
//                               full/empty mbarriers must never be two phases early
//                               (512 TMEM columns = 2 buffers x 2 halves x 128)
//   warps 8..23  epilogue     : 16 warps = 2 query halves x 2 column halves x 4 TMEM lane
//                               quarters; a thread owns one query row and 64 of the 128
//                               columns of every tile.
//
// Epilogue = exact running top-2 per row, organised around the measured budget of ~1 ALU
// op per accumulator (tools/ubench.cu): the fast path looks only at RAW dot products.
// A column j can enter a row's top-2 only if |t_j|^2 - 2 q.t_j < v2 (v2 = second-best value so
// far), hence only if 2 q.t_j - n8 > -v2, where n8 = min |t|^2 over the 8-column group of j
// (precomputed at upload).  Per group of 8 columns: a 3-input-max tree (0.5 op/element),
// one IMAD and one warp vote against the rows' bound.  Only groups in
// which some row of the warp passes get their exact packed keys built (8 IMADs with the
// column keys staged in shared memory) and inserted (knock-out min trees, 13 ALU ops); that is
// harmless for the rows that did not pass.  Skipped columns provably have two predecessors
// that beat them, so results are identical to the unfiltered epilogue (SFM_KNN_MODE=0); a GPU
// test compares the two bit for bit.  The two threads that share a row (column halves)
// exchange their top-2 through shared memory once per 1024-column window and bound with the
// row's JOINT second best.  The epilogue is ALU-throughput bound: per thread
// and 64-accumulator tile it issues 32 max + 8 compare + 8 vote instructions and 15 more per
// group that hits; everything warp-uniform (addresses, barriers, loop control) runs on the
// uniform datapath.  The distance matrix never leaves the SM.
//
// Three sweep flavours, picked by the host (launch_knn2):
//   exact        (the caller asks for the raw kNN rows): the search above, first window unfiltered.
//   match-only   (match lists only, small calls): rows that fail the ratio test with what they know
//                bound with their best instead of their second best (kMatchOnly).
//   ratio-driven (match lists only, large calls; kPrune): the bound follows the ratio test itself --
//                ratio^2 * best for rows that fail with what they know, best / ratio^2 for rows that pass;
//                8 seed columns instead of unfiltered tiles, one vote per 64-column share; the few rows
//                it cannot decide are listed and recomputed exactly (match_finalize.cu).  178 k pairs/s,
//                69 % of the int8 peak on the 19,900-pair benchmark.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <type_traits>

#include "knn_epilogue.cuh"
#include "match_types.h"
#include "ptx.cuh"

namespace sfm {

// Timing experiments (anatomy modes 2..4 and the dbg bits; their results are garbage by design)
// exist only in -DSFM_EXPERIMENTS builds (tools/variants.py); the shipped library has the exact
// modes 0 / 1 and nothing else.
#ifdef SFM_EXPERIMENTS
#define SFM_DBG(bit) ((dbg & (bit)) != 0)
#else
#define SFM_DBG(bit) false
#endif

// -DSFM_CHECKS: protocol / bounds assertions in the kernel (compute-sanitizer is closed on the GPU
// pool this was developed on, profiles/rnd2_sanitizer_unavailable.txt): work-item tables, TMEM and
// shared-memory ring addresses, result rows and the tag discipline of the bound exchange are checked
// on the device and trap with a message.  The GPU tests are run once against this build
// (tools/variants.py build chk:-DSFM_CHECKS; profiles/rnd2_checks_build_tests.txt).
#ifdef SFM_CHECKS
#define SFM_ASSERT(cond, what)                                                                  \
  do {                                                                                          \
    if (!(cond)) {                                                                              \
      printf("sfm_b200 check failed: %s (block %d thread %d, %s:%d)\n", what, blockIdx.x,      \
             threadIdx.x, __FILE__, __LINE__);                                                  \
      __trap();                                                                                 \
    }                                                                                           \
  } while (0)
#else
#define SFM_ASSERT(cond, what) do { } while (0)
#endif

#ifndef SFM_STAGES
#define SFM_STAGES 6                            // tools/variants.py builds other depths with -D
#endif
constexpr int kStages = SFM_STAGES;             // B-tile ring depth (16 KB each)
constexpr int kAccBufs = 2;                     // TMEM accumulator buffers (2 x 128 columns each)
constexpr int kFirstMmaWarp = 4;                // warps 1..3 idle (warpgroup granularity)
constexpr int kMmaWarps = 4;                    // (query half, tile parity)
constexpr int kFirstEpiWarp = kFirstMmaWarp + kMmaWarps;
// The 128 columns of a tile are split over kParts epilogue threads per query row:
//   2 parts: 64 + 64 columns, 16 epilogue warps (4 per SM sub-partition), 104 registers each
//   3 parts: 48 + 48 + 32 columns, 24 epilogue warps (6 per sub-partition, the CTA is 1024 threads),
//            72 registers each -- more runnable warps to overlap the latency-bound hit path of
//            one warp with the ALU-bound max trees of the others
#ifndef SFM_EPI_PARTS
#define SFM_EPI_PARTS 2
#endif
constexpr int kParts = SFM_EPI_PARTS;
static_assert(kParts == 2 || kParts == 3, "column parts per query row");
constexpr int kEpiWarps = 8 * kParts;           // 2 halves x kParts column parts x 4 lane quarters
constexpr int kKnnThreads = (kFirstEpiWarp + kEpiWarps) * 32;       // 768 / 1024
constexpr int kRegsLaunch = kParts == 2 ? 80 : 64;                  // 65536 / threads, multiple of 8
constexpr int kRegsProd = 24;                   // setmaxnreg: producer warpgroup (warps 0..3)
constexpr int kRegsMma = 40;                    // setmaxnreg: MMA warpgroup (warps 4..7)
constexpr int kRegsEpi = kParts == 2 ? 104 : 72;                    // setmaxnreg: epilogue warpgroups
static_assert(4 * kRegsProd + 4 * kRegsMma + kEpiWarps * kRegsEpi <= (kFirstEpiWarp + kEpiWarps) * kRegsLaunch,
              "register pool of the CTA");
// part p owns columns [part_col0(p), part_col0(p) + 32 + 8 * part_bgroups(p)): a 32-column piece A
// (four 8-column groups) and a piece B of part_bgroups(p) groups
__host__ __device__ constexpr int part_col0(int p) { return kParts == 2 ? 64 * p : 48 * p; }
__host__ __device__ constexpr int part_bgroups(int p) { return kParts == 2 ? 4 : (p < 2 ? 2 : 0); }
constexpr int kWinTiles = (1 << kColBits) / kTileN;   // train tiles per packed-key window (8)
#ifndef SFM_PRUNE_COLD_TILES
#define SFM_PRUNE_COLD_TILES 2                  // unfiltered tiles at the start of a ratio-driven (kPrune) sweep
#endif
#ifndef SFM_SHARE_VOTE
#define SFM_SHARE_VOTE 1                        // ratio-driven sweep: one vote per 64-column share of a tile
#endif
#ifndef SFM_RELEASE_FIRST
#define SFM_RELEASE_FIRST 1                     // ratio-driven sweep: TMEM buffer handed back in front of piece A
#endif
#ifndef SFM_PRUNE_SEED
#define SFM_PRUNE_SEED 1                        // ratio-driven sweep: 8 seed columns per thread instead of unfiltered tiles
#endif
#ifndef SFM_COLD_WINDOWS
#define SFM_COLD_WINDOWS 1                      // windows at the start of a sweep that skip the filter
#endif
#ifndef SFM_PRE_RELEASE_MID
#define SFM_PRE_RELEASE_MID 0                   // pre-vote flavour: hand the TMEM buffer back inside piece A
#endif
constexpr int kPreVoteTiles = 128;              // exact search: sweeps of >= 16384 train rows use the chunk pre-vote
constexpr int kHalfM = kTileM / 2;              // 128 rows per MMA
constexpr int kCkSlots = 16;                    // ring of per-tile column keys (512 B each)

constexpr uint32_t kABytes = kTileM * kDim;     // 32 KB
constexpr uint32_t kAHalfBytes = kHalfM * kDim; // 16 KB
constexpr uint32_t kBBytes = kTileN * kDim;     // 16 KB
constexpr uint32_t kCkBytes = kTileN * 4;       // 512 B
constexpr uint32_t kGmBytes = kTileN / 8 * 4;   // 64 B: min |t|^2 of each 8-column group

// What the producer tells the MMA and epilogue warps about an item.
struct ItemInfo {
  int32_t ntiles;       // train tiles of the pair
  int32_t rows_valid;   // query rows of this block that exist (<= 256)
  int32_t norm_row;     // bank row of the block's first query row
  int32_t t_row0;       // bank row of the train image's first row
  int32_t pair;         // pair index and first row of the block within the pair (flagged rows, kPrune)
  int32_t row0;
  int64_t knn_row;      // first output row of the block
};

// dynamic shared memory map (offsets from a 1024-byte aligned base)
constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffB = kOffA + 2 * kABytes;
constexpr uint32_t kOffCk = kOffB + kStages * kBBytes;
constexpr uint32_t kOffGm = kOffCk + kCkSlots * kCkBytes;
constexpr uint32_t kOffInfo = kOffGm + kCkSlots * kGmBytes;
constexpr uint32_t kOffMerge = kOffInfo + 2 * sizeof(ItemInfo);        // 2 x (kParts - 1) x 256 rows x int4
constexpr uint32_t kOffShare = kOffMerge + 2 * (kParts - 1) * kTileM * 16;   // 256 rows x 32 B (8 per part)
constexpr uint32_t kOffBar = kOffShare + kTileM * 32;
constexpr uint32_t kNumBars = 2 * kStages + 4 + 4 * kAccBufs;
constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr uint32_t kKnnSmemBytes = kOffTmemPtr + 16 + 1024;   // + alignment slack
static_assert(kKnnSmemBytes <= 227 * 1024, "shared memory budget");
static_assert(sizeof(ItemInfo) == 32, "smem layout");

// Top-2 update with one piece of a tile: kG groups of 8 columns (4 for a 32-column piece);
// ck_addr / gm_addr = shared addresses of the piece's column keys and group minima.
//   kMode 0: every group is inserted (2.5 min/max + 1 IMAD per element).
//   kMode 1: only groups whose smallest possible value beats the bound of some row of the
//            warp: per group one IMAD, one compare and one vote on top of the max tree.
//            neg2 = -2 in a register ptxas cannot see through, so that the multiply-add stays
//            an IMAD on the FMA pipe instead of an IADD3 on the (limiting) ALU pipe.
//   kPreVote : one vote for the whole piece in front of the per-group votes.  A vote costs
//            about two ALU instructions; when few pieces contain a hit most of them need only
//            that one.  Exact search: 54 % of the 32-column pieces of an 8192-column sweep
//            contain a hit and it does not pay (measured: -2.6 % there, +7 % on a 65536-column
//            train image), so the kernel picks per work item.  Match-only search (35 % of the
//            pieces hit): always (+2.5 % on the 8192-column workload).  A variant that tests
//            the piece's maximum against a precomputed 32-column minimum norm first and loads /
//            tests the groups only on a hit has 6 instructions fewer per quiet piece but is
//            slower (2240 vs 2292 TOP/s): the group tests then sit serially behind the vote
//            instead of running alongside the max trees.
// `mid` runs once inside the update, after the filter votes and before the inserts (the sweep
// hands the TMEM buffer back there).
struct NoHook { __device__ __forceinline__ void operator()() const {} };
template <int kMode, bool kPreVote, int kG, class Mid = NoHook>
__device__ __forceinline__ void chunk_update(const uint32_t (&r)[8 * kG], uint32_t ck_addr,
                                             uint32_t gm_addr, int neg2, RowTop2& s, Mid mid = Mid()) {
  static_assert(kG == 2 || kG == 4, "piece of 16 or 32 columns");
  if constexpr (kMode == 0) {
    group_insert(&r[0], ck_addr, s);
    mid();
#pragma unroll
    for (int j = 1; j < kG; ++j) group_insert(&r[8 * j], ck_addr + 32 * j, s);
  } else {
    int n8[kG];
    if constexpr (kG == 4 && kParts == 2) {
      const int4 nn = lds_v4(gm_addr);
      n8[0] = nn.x; n8[1] = nn.y; n8[2] = nn.z; n8[3] = nn.w;
    } else {                              // 3 parts: a piece starts at a multiple of 8 bytes only
#pragma unroll
      for (int j = 0; j < kG; j += 2) {
        const int2 nn = lds_v2(gm_addr + 4 * j);
        n8[j] = nn.x; n8[j + 1] = nn.y;
      }
    }
    bool h[kG];
    if constexpr (kPreVote) {
      bool p[kG];
      bool any = false;
#pragma unroll
      for (int j = 0; j < kG; ++j) {
        p[j] = group_max(&r[8 * j]) * neg2 + n8[j] < s.bv;
        any |= p[j];
      }
      if (!__any_sync(0xffffffffu, any)) { mid(); return; }
#pragma unroll
      for (int j = 0; j < kG; ++j) h[j] = __any_sync(0xffffffffu, p[j]);
    } else {
#pragma unroll
      for (int j = 0; j < kG; ++j)
        h[j] = __any_sync(0xffffffffu, group_max(&r[8 * j]) * neg2 + n8[j] < s.bv);
    }
    mid();
#pragma unroll
    for (int j = 0; j < kG; ++j) {
      if (h[j]) {
        group_insert(&r[8 * j], ck_addr + 32 * j, s);
        s.bv = min(s.bv, s.m2 >> kColBits);
      }
    }
  }
}

// smallest possible value |t|^2 - 2 q.t over a piece of kG groups (group maxima against the groups'
// minimum norms): the piece holds a candidate for a row iff this is below the row's bound
template <int kG>
__device__ __forceinline__ int share_min(const uint32_t (&r)[8 * kG], uint32_t gm_addr, int neg2) {
  static_assert(kG == 4, "a 32-column piece");
  const int4 nn = lds_v4(gm_addr);
  return min(__vimin3_s32(group_max(&r[0]) * neg2 + nn.x, group_max(&r[8]) * neg2 + nn.y,
                          group_max(&r[16]) * neg2 + nn.z),
             group_max(&r[24]) * neg2 + nn.w);
}

// kMatchOnly (mode 1 only): the caller wants match lists, not the raw kNN rows.  A row whose
// current top-2 fails Lowe's ratio test for sure (d0^2 > ratio2 * d1^2, ratio2 = ratio^2 with a
// margin that covers the float sqrt and the double compare of the real test) can only become a
// match through a NEW BEST neighbour; a column between its best and second best can only make the
// test fail harder.  Such rows bound with their best value instead of their second best (decided
// once per packed-key window, no cost per hit).  If a new best does arrive the pair (new best, old
// best) is the exact top-2 again -- every skipped column was no better than the old best -- so
// rows that pass, their distances and min_dist are exactly those of the full search; for rows
// that fail, Knn2::j1 / d1 may name a column that is not the true second neighbour (it is never
// better than the true one, so the row fails the real test too).
//
// kPrune (match-only): the bound follows the RATIO TEST itself.  With (D0, D1) the squared distances of
// the two nearest columns known for the row (its own and its row partner's, exact columns):
//   F, D0 > ratio^2 D1 (fails with what is known): a column can only matter as a new nearest neighbour
//      that PASSES, i.e. below ratio^2 * D0 -- whatever lies in [ratio^2 D0, D0) would become the
//      nearest neighbour of a row that still fails (its second is at most D0).  Bound: ratio^2 D0.
//   P, D0 <= ratio^2 D1 (passes with what is known): what matters is a better nearest neighbour or a
//      column that makes the test fail, both below D0 / ratio^2.  Bound: D0 / ratio^2.
// (always together with the exact rule "nothing at or above the known second".)  For random
// descriptors the F bound is far below every distance, so a row hits only on a real candidate.
// One case is undecidable inside the sweep: a row that ends in state P with a nearest neighbour D0
// whose "fail range" [D0, D0 / ratio^2) reaches into what was skipped earlier under an F bound
// (D0 / ratio^2 > the smallest F bound the row ever used).  Such rows are appended to a list and
// recomputed exactly by recheck_rows_kernel before the filter passes read them; everything else is
// exact: a row whose true nearest neighbour was skipped fails the real test (it is >= ratio^2 times a
// known column), and a row stored as failing fails (its stored columns are real).
struct PruneParams {
  int32_t* count;       // number of flagged rows (may exceed cap: the host then repeats the call unpruned)
  int2* rows;           // (pair, row within the pair)
  int32_t cap;
  float inv_ratio2;     // (1 + margin) / ratio^2
};
template <int kMode, bool kMatchOnly = false, bool kPrune = false>
__global__ void __launch_bounds__(kKnnThreads, 1)
knn2_kernel(const __grid_constant__ CUtensorMap tmap, const int32_t* __restrict__ ckey,
            const int32_t* __restrict__ gmin8, const int32_t* __restrict__ norm, const PairDesc* __restrict__ pairs,
            const int2* __restrict__ items, int n_items, Knn2* __restrict__ knn_out, int dbg, float ratio2,
            const PruneParams prune) {
  static_assert(!kPrune || (kMatchOnly && kMode == 1), "the ratio-driven bound belongs to the match-only sweep");
  extern __shared__ uint8_t smem_raw[];
  uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // opaque to ptxas from here on: keep the base in a register instead of re-deriving it from
  // the CTA's shared window (4 ALU instructions) in front of every shared-memory access
  asm volatile("" : "+r"(smem_base));

  const uint32_t sA = smem_base + kOffA;
  const uint32_t sB = smem_base + kOffB;
  const uint32_t sCk = smem_base + kOffCk;
  const uint32_t sGm = smem_base + kOffGm;
  const uint32_t bar0 = smem_base + kOffBar;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kStages + s); };
  auto bar_a_full = [&](int b) { return bar0 + 8u * (2 * kStages + b); };
  auto bar_a_empty = [&](int b) { return bar0 + 8u * (2 * kStages + 2 + b); };
  auto bar_t_full = [&](int b, int h) { return bar0 + 8u * (2 * kStages + 4 + 2 * b + h); };
  auto bar_t_empty = [&](int b, int h) {
    return bar0 + 8u * (2 * kStages + 4 + 2 * kAccBufs + 2 * b + h);
  };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + kOffTmemPtr);
  volatile ItemInfo* info = reinterpret_cast<volatile ItemInfo*>(smem_gen + kOffInfo);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 2);                 // the two query halves' MMA commits
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_a_full(b), 1);
      // every MMA warp + every epilogue warp
      mbar_init(bar_a_empty(b), SFM_DBG(2) ? kMmaWarps : kMmaWarps + kEpiWarps);
    }
    for (int b = 0; b < kAccBufs; ++b)
      for (int h = 0; h < 2; ++h) {
        mbar_init(bar_t_full(b, h), 1);
        mbar_init(bar_t_empty(b, h), kEpiWarps / 2);
      }
    fence_mbar_init();
  }
  if (threadIdx.x < kTileM)    // tagged (best, second best) slots of the row-sharing threads
    for (int h = 0; h < kParts; ++h)
      sts_v2(smem_base + kOffShare + threadIdx.x * 32 + h * 8, 0xffffffffu, 0xffffffffu);
  if (warp == kFirstMmaWarp) {
    tmem_alloc(smem_base + kOffTmemPtr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // Producer and MMA warps run their loops with all 32 lanes and pick the issuing lane with
  // elect.sync: ptxas then keeps descriptors / barrier addresses in uniform registers instead
  // of wrapping every UTCIMMA / UTMALDG in a per-thread waterfall loop (measured: 80 cycles
  // per MMA with `if (lane == 0)`).  768 threads leave 80 registers per thread; the control
  // warpgroups hand part of their share to the epilogue warpgroups (setmaxnreg at the top of
  // each role branch, so that ptxas allocates per branch).
  if (warp < kFirstMmaWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProd));
    if (warp == 0) {
      // ===================================================== TMA producer
      uint32_t stage = 0, phase = 0, abuf = 0, aphase = 0, tile_seq = 0;
      int item = blockIdx.x;
      int2 it = item < n_items ? __ldg(items + item) : make_int2(0, 0);
      PairDesc pd = pairs[it.x];
      for (; item < n_items; item += gridDim.x) {
        const int mblk = it.y;
        const int ntiles = (pd.nt + kTileN - 1) / kTileN;
        const int t_row0 = pd.t_row0;
        SFM_ASSERT(pd.nt >= 2 && pd.nq > 0 && mblk >= 0 && mblk * kTileM < pd.nq, "work item outside its pair");
        SFM_ASSERT(pd.q_row0 >= 0 && pd.t_row0 >= 0 && pd.knn_off >= 0, "negative bank / result row");
        mbar_wait(bar_a_empty(abuf), aphase ^ 1);
        if (elect_one()) {
          info[abuf].ntiles = ntiles;
          info[abuf].rows_valid = pd.nq - mblk * kTileM;
          info[abuf].norm_row = pd.q_row0 + mblk * kTileM;
          info[abuf].t_row0 = pd.t_row0;
          info[abuf].pair = it.x;
          info[abuf].row0 = mblk * kTileM;
          info[abuf].knn_row = pd.knn_off + static_cast<int64_t>(mblk) * kTileM;
          mbar_arrive_expect_tx(bar_a_full(abuf), kABytes);
          tma_load_2d(sA + abuf * kABytes, &tmap, bar_a_full(abuf), 0, pd.q_row0 + mblk * kTileM);
          tma_load_2d(sA + abuf * kABytes + kAHalfBytes, &tmap, bar_a_full(abuf), 0,
                      pd.q_row0 + mblk * kTileM + kHalfM);
        }
        __syncwarp();
        abuf ^= 1;
        if (abuf == 0) aphase ^= 1;
        // fetch the next item's tables now; the loads land while this item's tiles stream
        const int nxt = item + gridDim.x;
        if (nxt < n_items) {
          it = __ldg(items + nxt);
          pd = pairs[it.x];
        }
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(bar_empty(stage), phase ^ 1);
          const int row = t_row0 + t * kTileN;
          if (elect_one()) {
            if (SFM_DBG(1)) {                    // timing experiment: no operand traffic
              mbar_arrive(bar_full(stage));
            } else {
              mbar_arrive_expect_tx(bar_full(stage), kBBytes + kCkBytes + kGmBytes);
              tma_load_2d(sB + stage * kBBytes, &tmap, bar_full(stage), 0, row);
              bulk_load_1d(sCk + (tile_seq % kCkSlots) * kCkBytes, ckey + row, kCkBytes,
                           bar_full(stage));
              bulk_load_1d(sGm + (tile_seq % kCkSlots) * kGmBytes, gmin8 + row / 8, kGmBytes,
                           bar_full(stage));
            }
          }
          __syncwarp();
          ++tile_seq;
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp < kFirstEpiWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsMma));
    {
      // ===================================================== MMA issuers: (half mh, parity mp)
      const uint32_t mh = (warp - kFirstMmaWarp) & 1, mp = (warp - kFirstMmaWarp) >> 1;
      constexpr uint32_t idesc = make_idesc_u8(kHalfM, kTileN);
      const uint32_t d_tmem = tmem_base + mp * (2 * kTileN) + mh * kTileN;
      uint32_t seq = 0, abuf = 0, aphase = 0;     // seq = tiles of this CTA so far
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        mbar_wait(bar_a_full(abuf), aphase);
        const uint32_t ntiles = info[abuf].ntiles;
        const uint64_t a_desc = make_smem_desc_sw128(sA + abuf * kABytes + mh * kAHalfBytes);
        const uint32_t end = seq + ntiles;
        uint32_t ts = seq + ((seq & 1) != mp);     // first tile of this warp's parity
        if (ts >= end) {                           // single-tile item of the other parity
          if (elect_one()) mbar_arrive(bar_a_empty(abuf));
          __syncwarp();
        }
        for (; ts < end; ts += 2) {
          const uint32_t stage = ts % kStages;
          mbar_wait(bar_full(stage), (ts / kStages) & 1);
          if (!SFM_DBG(2)) mbar_wait(bar_t_empty(mp, mh), ((ts >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint64_t b_desc = make_smem_desc_sw128(sB + stage * kBBytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kDim / 32; ++k) {
              // advance 32 bytes along K inside the 128-byte swizzle atom: +2 in (addr >> 4)
              umma_i8(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, k > 0);
            }
            umma_commit(bar_t_full(mp, mh));
            umma_commit(bar_empty(stage));
            if (ts + 2 >= end) umma_commit(bar_a_empty(abuf));   // this warp's last tile
          }
          __syncwarp();
        }
        seq = end;
        abuf ^= 1;
        if (abuf == 0) aphase ^= 1;
      }
    }
  } else {
    // ===================================================== epilogue: running top-2 per row
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpi));
    const int e = warp - kFirstEpiWarp;            // 4 consecutive warps cover the 4 quarters
    const int half = (e >> 2) / kParts;            // which 128-row half of the block
    const int part = (e >> 2) % kParts;            // which column part of every tile
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int row_in_blk = half * kHalfM + quarter * 32 + lane;
    const int col0 = kParts == 2 ? 64 * part : 48 * part;           // == part_col0(part)
    // Loop invariants that ptxas would otherwise re-derive from %tid in front of every use
    // (8 ALU instructions each time): pin them in registers.
    uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + half * kTileN + col0;
    uint32_t bar_tf = bar_t_full(0, half), bar_te = bar_t_empty(0, half);   // + 16 * buffer
    uint32_t ck_base = sCk + col0 * 4;
    uint32_t gm_base = sGm + col0 / 8 * 4;
    int neg2 = -2;
    asm volatile("" : "+r"(t_addr), "+r"(bar_tf), "+r"(bar_te), "+r"(ck_base), "+r"(gm_base),
                 "+r"(neg2));
    // warp-uniform by construction (functions of the warp index): a broadcast lets ptxas see
    // it and keep the address arithmetic of the tile loop on the uniform datapath
    t_addr = __shfl_sync(0xffffffffu, t_addr, 0);
    bar_tf = __shfl_sync(0xffffffffu, bar_tf, 0);
    bar_te = __shfl_sync(0xffffffffu, bar_te, 0);
    ck_base = __shfl_sync(0xffffffffu, ck_base, 0);
    gm_base = __shfl_sync(0xffffffffu, gm_base, 0);
    const uint32_t merge_row = smem_base + kOffMerge + row_in_blk * 16;
    const uint32_t share_row = smem_base + kOffShare + row_in_blk * 32;
    const int pair_bar = 1 + half * 4 + quarter;   // named barrier of the kParts warps of a row set
    uint32_t buf = 0, bphase = 0, abuf = 0, mslot = 0, tile_seq = 0, item_seq = 0;
    for (int item = blockIdx.x; item < n_items && !SFM_DBG(2); item += gridDim.x) {
      RowTop2 st = {INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX};
      int ntiles = 1, rows_valid = 0, norm_row = 0;
      [[maybe_unused]] int pair_id = 0, row0 = 0;
      [[maybe_unused]] int smin = INT32_MAX;       // kPrune: smallest F bound this thread skipped under
      int64_t knn_row = 0;
      if constexpr (kMode <= 1) {
        // ---- sweep: a thread's share of a tile is a 32-column piece A and (except for the last
        // of three parts) a piece B; the tcgen05.ld of piece B is in flight while piece A is
        // processed, and the buffer is released between the two
        mbar_wait(bar_tf + 16 * buf, bphase);
        tc_fence_after();
        ntiles = info[abuf].ntiles;
        ntiles = __shfl_sync(0xffffffffu, ntiles, 0);
        rows_valid = info[abuf].rows_valid;
        norm_row = info[abuf].norm_row;
        knn_row = info[abuf].knn_row;
        if constexpr (kPrune) {
          pair_id = info[abuf].pair;
          row0 = info[abuf].row0;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a_empty(abuf));
        abuf ^= 1;
        SFM_ASSERT(ntiles >= 1 && rows_valid >= 1 && knn_row >= 0, "item info not published");
        SFM_ASSERT(buf < kAccBufs && (tile_seq % kCkSlots) < kCkSlots, "accumulator / key ring index");
        // |q|^2 of this thread's row (padding rows of the bank carry a sentinel: any value works)
        [[maybe_unused]] int nq2 = 0;
        if constexpr (kMatchOnly) nq2 = row_in_blk < rows_valid ? __ldg(norm + norm_row + row_in_blk) : 0;
        uint32_t ra[32];
        tmem_ld_x32(t_addr + buf * (2 * kTileN), ra);
        tmem_ld_wait();
        // the sweep in compiled flavours (insert-all / filtered, with / without the piece-level
        // pre-vote, groups of piece B); long train images take the pre-vote one
        auto sweep = [&](auto mode_tag, auto pre_tag, auto bg_tag, int w_first, int w_last) {
          constexpr int kM = decltype(mode_tag)::value;      // 0: insert every group, 1: filtered
          constexpr bool kPre = decltype(pre_tag)::value;
          constexpr int kBG = decltype(bg_tag)::value;       // groups of piece B: 4, 2 or 0
        uint32_t rb[kBG > 0 ? 8 * kBG : 8];
          // ratio-driven sweep: quiet tiles are the rule, so ONE vote covers the thread's whole
          // 64-column share; a share with a hit falls back to the piece / group tests
          constexpr bool kQuiet = kPrune && kM == 1 && kPre && kBG == 4;
          constexpr bool kShareVote = SFM_SHARE_VOTE && kQuiet;
          constexpr bool kReleaseFirst = SFM_RELEASE_FIRST && kQuiet;
        // tiles in windows of kWinTiles (one packed-key window): the window bookkeeping sits
        // behind the inner loop, not behind a per-tile test
        // a window ends at the next multiple of kWinTiles (its keys carry the column modulo 1024) or
        // where the sweep ends -- the cold sweep may stop inside a window
        for (int w0 = w_first, wend; w0 < w_last; w0 = wend) {
        wend = min((w0 / kWinTiles + 1) * kWinTiles, w_last);
        for (int t = w0; t < wend; ++t) {
          if constexpr (kBG == 4) tmem_ld_x32(t_addr + buf * (2 * kTileN) + 32, rb);   // piece B in flight
          if constexpr (kBG == 2) tmem_ld_x16(t_addr + buf * (2 * kTileN) + 32, rb);
          const uint32_t slot = tile_seq % kCkSlots;
          const uint32_t ck_addr = ck_base + slot * kCkBytes;
          const uint32_t gm_addr = gm_base + slot * kGmBytes;
          auto release = [&] {                                  // tile t is out of TMEM
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            mbar_arrive_elected(bar_te + 16 * buf);
          };
          if constexpr (kReleaseFirst) release();               // both pieces in registers: two tiles of slack
          if constexpr (kShareVote) {
            const int va = share_min<4>(ra, gm_addr, neg2);
            if constexpr (!kReleaseFirst) release();
            const int vb = share_min<4>(rb, gm_addr + 16, neg2);
            if (__any_sync(0xffffffffu, min(va, vb) < st.bv)) {
              chunk_update<kM, kPre, 4>(ra, ck_addr, gm_addr, neg2, st);
              chunk_update<kM, kPre, 4>(rb, ck_addr + 128, gm_addr + 16, neg2, st);
            }
          } else if constexpr (kBG == 0) {
            release();                                          // piece A is the whole share
            chunk_update<kM, kPre, 4>(ra, ck_addr, gm_addr, neg2, st);
          } else if constexpr (kPre && !SFM_PRE_RELEASE_MID) {
            chunk_update<kM, kPre, 4>(ra, ck_addr, gm_addr, neg2, st);
            if constexpr (!kReleaseFirst) release();
          } else {
            // released from inside piece A (after its votes, before its inserts): piece B has
            // landed by then, and the MMA of tile t+2 starts a third of a tile earlier (+0.9 %)
            chunk_update<kM, kPre, 4>(ra, ck_addr, gm_addr, neg2, st, release);
          }
          const uint32_t nbuf = buf ^ 1, nphase = bphase ^ buf; // phase flips when buf wraps to 0
          if constexpr (kBG > 0 && !kShareVote) chunk_update<kM, kPre, kBG>(rb, ck_addr + 128, gm_addr + 16, neg2, st);
          // Tile t+1 is asked for only now, not before piece B: the warps that share a TMEM
          // buffer drift apart by their hit counts, and the MMA of tile t+1 starts when the
          // slowest of them released tile t-1 -- waiting half a tile later keeps 44 % of the
          // waits from sleeping (measured: +5 %, more than the exposed tcgen05.ld costs)
          if (t + 1 < ntiles) {
            mbar_wait(bar_tf + 16 * nbuf, nphase);
            tc_fence_after();
            tmem_ld_x32(t_addr + nbuf * (2 * kTileN), ra);      // piece A of tile t+1
            tmem_ld_wait();
          }
          ++tile_seq;
          buf = nbuf;
          bphase = nphase;
        }
          {
            // close the 1024-column window, then tighten the bound, also with the row
            // partners' top-2 (ties with the partners' columns go by index: + 1)
            close_window(st, (w0 / kWinTiles * kWinTiles) * kTileN);
            int bound = st.g2v;
            if (kMode == 1) {
              // the row's second best over ALL column parts bounds what can still enter: second
              // smallest of the union of the parts' (best, second).
              // Each 32-bit word carries its own tag (this CTA's item counter, 10 bits), so the
              // exchange needs neither a barrier nor an atomic 8-byte access: a word of an older
              // item, or a pair torn between two items, fails the tag test and is ignored (the
              // partners are never more than one item away: bar.sync at every item end).  Values
              // are < 2^21 in magnitude (a missing second best is clamped to 2^21 - 1).
              constexpr int kNone = (1 << 21) - 1;
              const int tag = static_cast<int>(item_seq & 1023u);
              sts_v2(share_row + part * 8, min(st.g1v, kNone) * 1024 + tag, min(st.g2v, kNone) * 1024 + tag);
              int j1 = min(st.g1v, kNone), j2 = min(st.g2v, kNone);
#pragma unroll
              for (int o = 1; o < kParts; ++o) {
                const int op = part + o >= kParts ? part + o - kParts : part + o;
                const int2 w = lds_v2(share_row + op * 8);   // any earlier value of this item is valid
                SFM_ASSERT((w.x & 1023) == (w.y & 1023) || ((((w.x & 1023) - (w.y & 1023)) & 1023) == 1) ||
                               ((((w.y & 1023) - (w.x & 1023)) & 1023) == 1),
                           "bound exchange: words more than one item apart");
                if (((w.x & 1023) == tag) & ((w.y & 1023) == tag)) merge_top2(j1, j2, w.x >> 10, w.y >> 10);
              }
              if (j2 < kNone) {
                bound = min(bound, j2 + 1);
                if constexpr (kMatchOnly) {
                  // squared distances are exact integers < 2^22: exact in float
                  const float d0 = static_cast<float>(j1 + nq2), d1 = static_cast<float>(j2 + nq2);
                  if constexpr (kPrune) {
                    if (d0 > ratio2 * d1) {                  // F: only a passing new nearest neighbour matters
                      const int tf = static_cast<int>(ratio2 * d0) + 1 - nq2;
                      bound = min(bound, tf);
                      smin = min(smin, tf);
                    } else {                                 // P: better neighbours and whatever makes the test fail
                      bound = min(bound, static_cast<int>(d0 * prune.inv_ratio2) + 2 - nq2);
                    }
                  } else {
                    if (d0 > ratio2 * d1) bound = min(bound, j1);
                  }
                }
              }
            }
            st.bv = bound;
          }
        }
        };
        // The first windows of a sweep run unfiltered: while the top-2 of a row is still cold
        // 65-100 % of the groups hit for some row of the warp, and testing them first (max tree,
        // compare, vote, branch) costs more than it saves.
        // (the ratio-driven bound needs two known columns, not a warm top-2: a short cold sweep)
        constexpr bool kSeed = kPrune && SFM_PRUNE_SEED != 0;
        constexpr int kCold = kMode != 1 || kSeed ? 0 : (kPrune ? SFM_PRUNE_COLD_TILES : SFM_COLD_WINDOWS * kWinTiles);
        const int cold = min(kCold, ntiles);
        if constexpr (kSeed) {
          // No unfiltered tiles at all: the ratio-driven bound only needs two known columns, whichever.
          // The first group of the thread's share of tile 0 is inserted unconditionally, the bound
          // follows from those 8 columns alone, and the filtered sweep starts at tile 0 (looking at a
          // column twice is harmless: keys are unique and the knock-out insert keeps the top-2 of a SET).
          group_insert(&ra[0], ck_base + (tile_seq % kCkSlots) * kCkBytes, st);
          constexpr int kNone = (1 << 21) - 1;               // a padding row's value (kNormPad)
          const int v1 = st.m1 >> kColBits, v2 = st.m2 >> kColBits;
          int bound = v2;
          if (v2 < kNone) {
            const float d0 = static_cast<float>(v1 + nq2), d1 = static_cast<float>(v2 + nq2);
            if (d0 > ratio2 * d1) {
              const int tf = static_cast<int>(ratio2 * d0) + 1 - nq2;
              bound = min(bound, tf);
              smin = min(smin, tf);
            } else {
              bound = min(bound, static_cast<int>(d0 * prune.inv_ratio2) + 2 - nq2);
            }
          }
          st.bv = bound;
        }
        auto sweeps = [&](auto bg_tag) {
          if constexpr (kCold > 0) sweep(std::integral_constant<int, 0>{}, std::false_type{}, bg_tag, 0, cold);
          if (kMatchOnly || ntiles >= kPreVoteTiles) sweep(std::integral_constant<int, kMode>{}, std::true_type{}, bg_tag, cold, ntiles);
          else sweep(std::integral_constant<int, kMode>{}, std::false_type{}, bg_tag, cold, ntiles);
        };
        if constexpr (kParts == 2) {
          sweeps(std::integral_constant<int, 4>{});
        } else {
          if (part < 2) sweeps(std::integral_constant<int, 2>{});
          else sweeps(std::integral_constant<int, 0>{});
        }
      } else {
#ifdef SFM_EXPERIMENTS
        // ---- timing experiments only (results are garbage):
        // 2 = drain TMEM, 3 = handshake only, 4 = drain + max tree + compare
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(bar_tf + 16 * buf, bphase);
          tc_fence_after();
          if (t == 0) {
            ntiles = info[abuf].ntiles;
            rows_valid = info[abuf].rows_valid;
            norm_row = info[abuf].norm_row;
            knn_row = info[abuf].knn_row;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_a_empty(abuf));
            abuf ^= 1;
          }
          const uint32_t ta = t_addr + buf * (2 * kTileN);
          if constexpr (kMode == 2 || kMode == 4) {
            uint32_t r0[32], r1[32];
            tmem_ld_x32(ta, r0);
            tmem_ld_x32(ta + 32, r1);
            tmem_ld_wait();
            st.g1v = min(st.g1v, static_cast<int>(r0[0] ^ r1[31]));
            if constexpr (kMode == 4) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int g0 = group_max(&r0[8 * j]), g1 = group_max(&r1[8 * j]);
                if (g0 * neg2 < st.bv) { st.g1i = g0; ++st.g2v; }
                if (g1 * neg2 < st.bv) { st.g1i = g1; ++st.g2i; }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_te + 16 * buf);
          ++tile_seq;
          if (++buf == kAccBufs) {
            buf = 0;
            bphase ^= 1;
          }
        }
#endif
      }
      ++item_seq;
      // merge the column parts of the row: parts 1.. hand their top-2 over to part 0
      const uint32_t slot = merge_row + mslot * ((kParts - 1) * kTileM * 16);
      mslot ^= 1;
      if (part > 0) sts_v4(slot + (part - 1) * (kTileM * 16), st.g1v, st.g1i, st.g2v, st.g2i);
      if constexpr (kPrune) {
        if (part > 0) sts_v2(share_row + 16, static_cast<uint32_t>(smin), 0u);   // bytes 16.. of the row's slot are free
      }
      asm volatile("bar.sync %0, %1;" ::"r"(pair_bar), "n"(32 * kParts) : "memory");
      if (part == 0) {
#pragma unroll
        for (int o = 0; o < kParts - 1; ++o) {
          const int4 w = lds_v4(slot + o * (kTileM * 16));
          insert_vi(st, w.x, w.y);
          insert_vi(st, w.z, w.w);
        }
        if (row_in_blk < rows_valid) {
          const int nq2 = __ldg(norm + norm_row + row_in_blk);
          if constexpr (kPrune) {
            // the row ends in state P with a fail range that reaches into what an F bound skipped:
            // not decidable from what was kept -- recheck_rows_kernel recomputes the row
            smin = min(smin, lds_v2(share_row + 16).x);
            const float d0 = static_cast<float>(st.g1v + nq2), d1 = static_cast<float>(st.g2v + nq2);
            if (d0 <= ratio2 * d1 && static_cast<int>(d0 * prune.inv_ratio2) + 2 - nq2 > smin) {
              const int idx = atomicAdd(prune.count, 1);
              if (idx < prune.cap) prune.rows[idx] = make_int2(pair_id, row0 + row_in_blk);
            }
          }
          SFM_ASSERT(st.g1i >= 0 && st.g2i >= 0 && st.g1i != st.g2i && st.g1v <= st.g2v && st.g1v + nq2 >= 0,
                     "top-2 of a row is not two distinct, ordered train rows");
          Knn2 out;
          out.j0 = st.g1i;
          out.j1 = st.g2i;
          out.d0 = st.g1v + nq2;
          out.d1 = st.g2v + nq2;
          *reinterpret_cast<int4*>(&knn_out[knn_row + row_in_blk]) =
              *reinterpret_cast<int4*>(&out);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kFirstMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// -------------------------------------------------------------------------------------
// Bare tensor-pipe probe: back-to-back 128x256x32 u8 MMAs on every SM, operands resident
// in shared memory (contents irrelevant), no epilogue.  Gives the measured int8 peak.
constexpr uint32_t kProbeA = 128 * kDim, kProbeB = 256 * kDim;
// variant (timing experiments on the single issuing thread, N = 128 only):
//   0 MMAs only   1 + tcgen05.commit every 4 MMAs   2 + commit every 8 MMAs
//   3 + commit every 8 and a try_wait on a completed mbarrier every 8 MMAs
//   4 two issuing threads (warps 0 and 2), each half of the MMAs, commit every 4
template <int kN>
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters, int variant) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base, sB = smem_base + kProbeA;
  const uint32_t bar = smem_base + kProbeA + kProbeB;
  volatile uint32_t* tmem_ptr_smem =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kProbeA + kProbeB + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (kProbeA + kProbeB) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_gen)[i] = 0x01010101u * (i & 3);
  const uint32_t bar_dummy = bar + 32, bar_done = bar + 40, bar2 = bar + 48;
  if (warp == 0 && lane == 0) {
    mbar_init(bar, 1);
    mbar_init(bar_dummy, 0x7fff);     // never completes: sink for experiment commits
    mbar_init(bar_done, 1);
    mbar_init(bar2, 1);
    fence_mbar_init();
    mbar_arrive(bar_done);            // phase 0 complete: try_wait(parity 0) succeeds at once
  }
  if (warp == 1) {
    tmem_alloc(smem_base + kProbeA + kProbeB + 16, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (warp == 0) {
    // whole warp runs the loop; elect.sync picks the issuing lane, which lets ptxas keep the
    // descriptors in uniform registers without a per-instruction waterfall loop
    constexpr uint32_t idesc = make_idesc_u8(128, kN);
    const uint64_t a_desc = make_smem_desc_sw128(sA), b_desc = make_smem_desc_sw128(sB);
    const int n = variant == 4 ? iters : iters * (256 / kN);
    for (int i = 0; i < n; ++i) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_i8(tmem_base + (i & 1) * 256, a_desc + 2 * k, b_desc + 2 * k, idesc, k > 0);
        if (variant == 1 || variant == 4) umma_commit(bar_dummy);
        if ((variant == 2 || variant == 3) && (i & 1)) umma_commit(bar_dummy);
      }
      __syncwarp();
      if (variant == 3 && (i & 1)) mbar_wait(bar_done, 0);
    }
    if (elect_one()) umma_commit(bar);
    __syncwarp();
    mbar_wait(bar, 0);
  }
  if (variant == 4 && warp == 2 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_u8(128, kN);
    const uint64_t a_desc = make_smem_desc_sw128(sA), b_desc = make_smem_desc_sw128(sB);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_i8(tmem_base + 128 + (i & 1) * 256, a_desc + 2 * k, b_desc + 2 * k, idesc, k > 0);
      umma_commit(bar_dummy);
    }
    umma_commit(bar2);
    mbar_wait(bar2, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// -------------------------------------------------------------------------------------
// host-side launchers (called from capi.cu)

// mode 0: unfiltered exact top-2 epilogue; mode 1: threshold-filtered (default, same results).
// The shared-memory opt-in is a per-device attribute of the kernel: it is set on every launch
// (a host-side table lookup), so contexts on several devices -- one per host thread, as
// INTEGRATION.md recommends -- each get it.
template <int kMode, bool kMatchOnly, bool kPrune = false>
static cudaError_t launch_knn2_mode(const CUtensorMap& tmap, const int32_t* ckey, const int32_t* gmin8,
                                    const int32_t* norm, const PairDesc* pairs, const int2* items,
                                    int n_items, Knn2* knn_out, int grid, int dbg, float ratio2,
                                    cudaStream_t stream, const PruneParams& prune = PruneParams{nullptr, nullptr, 0, 0.f}) {
  cudaError_t e = cudaFuncSetAttribute(knn2_kernel<kMode, kMatchOnly, kPrune>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, kKnnSmemBytes);
  if (e != cudaSuccess) return e;
  knn2_kernel<kMode, kMatchOnly, kPrune><<<grid, kKnnThreads, kKnnSmemBytes, stream>>>(
      tmap, ckey, gmin8, norm, pairs, items, n_items, knn_out, dbg, ratio2, prune);
  return cudaGetLastError();
}

bool knn2_mode_valid(int mode) {
#ifdef SFM_EXPERIMENTS
  return (mode & 15) >= 0 && (mode & 15) <= 4 && (mode >> 4) >= 0 && (mode >> 4) <= 3;
#else
  return mode == 0 || mode == 1;
#endif
}

// match_ratio > 0: the caller needs match lists only (see kMatchOnly); 0: exact kNN rows.
// flag_count / flag_rows / flag_cap (nullable): list of the rows the ratio-driven sweep (kPrune) could not
// decide; with a list the match-only call prunes, without one it runs the plain match-only sweep.
cudaError_t launch_knn2(int mode, const CUtensorMap& tmap, const int32_t* ckey,
                        const int32_t* gmin8, const int32_t* norm, const PairDesc* pairs, const int2* items,
                        int n_items, Knn2* knn_out, int n_sms, double match_ratio, int32_t* flag_count,
                        int2* flag_rows, int flag_cap, cudaStream_t stream) {
  if (!knn2_mode_valid(mode)) return cudaErrorInvalidValue;
  const int grid = n_items < n_sms ? n_items : n_sms;
  if (grid <= 0) return cudaSuccess;
#ifdef SFM_EXPERIMENTS
  const int dbg = mode >> 4;   // timing experiments (results invalid), see tools/exp_modes.py
  mode &= 15;
  if (mode == 4) return launch_knn2_mode<4, false>(tmap, ckey, gmin8, norm, pairs, items, n_items, knn_out, grid, dbg, 0.f, stream);
  if (mode == 3) return launch_knn2_mode<3, false>(tmap, ckey, gmin8, norm, pairs, items, n_items, knn_out, grid, dbg, 0.f, stream);
  if (mode == 2) return launch_knn2_mode<2, false>(tmap, ckey, gmin8, norm, pairs, items, n_items, knn_out, grid, dbg, 0.f, stream);
#else
  const int dbg = 0;
#endif
  if (mode == 0) return launch_knn2_mode<0, false>(tmap, ckey, gmin8, norm, pairs, items, n_items, knn_out, grid, dbg, 0.f, stream);
  if (match_ratio > 0.0 && match_ratio <= 1.0) {
    // the real test is (double)sqrtf(d0) > ratio * (double)sqrtf(d1): two float roundings of
    // 6e-8 each, squared; a relative margin of 1e-5 on ratio^2 keeps "fails for sure" sure
    const float ratio2 = static_cast<float>(match_ratio * match_ratio * (1.0 + 1e-5));
    if (flag_count != nullptr && flag_rows != nullptr && flag_cap > 0) {
      const PruneParams prune = {flag_count, flag_rows, flag_cap,
                                 static_cast<float>((1.0 + 1e-5) / (match_ratio * match_ratio))};
      return launch_knn2_mode<1, true, true>(tmap, ckey, gmin8, norm, pairs, items, n_items, knn_out, grid, dbg, ratio2,
                                             stream, prune);
    }
    return launch_knn2_mode<1, true>(tmap, ckey, gmin8, norm, pairs, items, n_items, knn_out, grid, dbg, ratio2, stream);
  }
  return launch_knn2_mode<1, false>(tmap, ckey, gmin8, norm, pairs, items, n_items, knn_out, grid, dbg, 0.f, stream);
}

// iters > 0: 128x256x32 MMAs; iters < 0: -(n << 4 | variant): the same work as 128x128x32
// MMAs with the issue-thread experiment `variant` (see i8_peak_kernel)
cudaError_t launch_i8_peak(int iters, int n_sms, cudaStream_t stream) {
  const uint32_t smem = kProbeA + kProbeB + 64 + 1024;
  cudaError_t e = cudaFuncSetAttribute(i8_peak_kernel<256>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(i8_peak_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  if (iters > 0) i8_peak_kernel<256><<<n_sms, 128, smem, stream>>>(iters, 0);
  else i8_peak_kernel<128><<<n_sms, 128, smem, stream>>>((-iters) >> 4, (-iters) & 15);
  return cudaGetLastError();
}

}  // namespace sfm
