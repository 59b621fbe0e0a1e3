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
// column keys staged in shared memory) and inserted (20 min/max); that is harmless for the
// rows that did not pass.  Skipped columns provably have two predecessors that beat them,
// so results are identical to the unfiltered epilogue (SFM_KNN_MODE=0); a GPU test compares
// the two bit for bit.  The two threads that share a row (column halves) exchange their
// second-best value through shared memory once per 512-column window to tighten thr.
// The distance matrix never leaves the SM.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "match_types.h"
#include "ptx.cuh"

namespace sfm {

constexpr int kStages = 6;                      // B-tile ring depth (16 KB each)
constexpr int kAccBufs = 2;                     // TMEM accumulator buffers (2 x 128 columns each)
constexpr int kFirstMmaWarp = 4;                // warps 1..3 idle (warpgroup granularity)
constexpr int kMmaWarps = 4;                    // (query half, tile parity)
constexpr int kFirstEpiWarp = kFirstMmaWarp + kMmaWarps;
constexpr int kEpiWarps = 16;                   // 2 halves x 2 column halves x 4 lane quarters
constexpr int kKnnThreads = (kFirstEpiWarp + kEpiWarps) * 32;       // 768
constexpr int kRegsCtl = 48;                    // setmaxnreg: producer / MMA warpgroups
constexpr int kRegsEpi = 96;                    // setmaxnreg: epilogue warpgroups
static_assert(8 * kRegsCtl + 16 * kRegsEpi <= 24 * 80, "register pool of the CTA (768 x 80)");
constexpr int kHalfM = kTileM / 2;              // 128 rows per MMA
constexpr int kColsPerThread = kTileN / 2;      // 64 columns of each tile per epilogue thread
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
  int32_t nt_min;       // min |t|^2 over the train image
  int32_t pad;
  int64_t knn_row;      // first output row of the block
};

// dynamic shared memory map (offsets from a 1024-byte aligned base)
constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffB = kOffA + 2 * kABytes;
constexpr uint32_t kOffCk = kOffB + kStages * kBBytes;
constexpr uint32_t kOffGm = kOffCk + kCkSlots * kCkBytes;
constexpr uint32_t kOffInfo = kOffGm + kCkSlots * kGmBytes;
constexpr uint32_t kOffMerge = kOffInfo + 2 * sizeof(ItemInfo);        // 2 x 256 rows x int4
constexpr uint32_t kOffShare = kOffMerge + 2 * kTileM * 16;            // 256 rows x 2 x int2
// mode 5 (drain warps): per epilogue warp an event ring + control words, per epilogue
// thread the drain-owned running top-2
constexpr int kExactTiles = 2;                             // tiles of a sweep done unfiltered
constexpr int kQueueSlots = 32;                            // events per ring (power of two)
constexpr uint32_t kQAcc = 0;                              // [32 slots][8] raw accumulators
constexpr uint32_t kQMeta = kQAcc + kQueueSlots * 32;      // [32] bank row | owner lane << 26
constexpr uint32_t kQCommit = kQMeta + kQueueSlots * 4;    // events published by the epilogue warp
constexpr uint32_t kQRead = kQCommit + 4;                  // events consumed by the drain warp
constexpr uint32_t kQRow0 = kQRead + 4;                    // bank row of the train image
constexpr uint32_t kQReserve = kQRow0 + 4;                 // ring positions handed out so far
constexpr uint32_t kQBytes = kQReserve + 4;
constexpr int kRowBits = 26;                               // bank rows < 2^26
constexpr uint32_t kOffQueue = kOffShare + kTileM * 16;
constexpr uint32_t kOffState = kOffQueue + kEpiWarps * kQBytes;        // 512 x int4
constexpr uint32_t kOffDone = kOffState + kEpiWarps * 32 * 16;
constexpr uint32_t kOffBar = kOffDone + 16;
constexpr uint32_t kNumBars = 2 * kStages + 4 + 4 * kAccBufs;
constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr uint32_t kKnnSmemBytes = kOffTmemPtr + 16 + 1024;   // + alignment slack
static_assert(kKnnSmemBytes <= 227 * 1024, "shared memory budget");
static_assert(sizeof(ItemInfo) == 32 && kQBytes % 16 == 0, "smem layout");

// (a1 <= a2), (b1 <= b2) -> the two smallest of the four, sorted.
__device__ __forceinline__ void merge_top2(int& a1, int& a2, int b1, int b2) {
  const int t = max(a1, b1);
  a1 = min(a1, b1);
  a2 = __vimin3_s32(t, a2, b2);
}

// Packed key of accumulator r (= q.t) and column key ck = (|t|^2 << 9) | (train row & 511):
// ((|t|^2 - 2 q.t) << 9) | column, one IMAD; orders like (distance, lower column first).
__device__ __forceinline__ int make_key(uint32_t r, int ck) {
  return static_cast<int>(r) * -(2 << kColBits) + ck;   // wraps, true value fits
}

// Exact top-2 of 8 keys merged into (m1, m2): pair-sort + merge tree, 20 min/max.
__device__ __forceinline__ void insert8(const int* k, int& m1, int& m2) {
  int lo[4], hi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    lo[j] = min(k[2 * j], k[2 * j + 1]);
    hi[j] = max(k[2 * j], k[2 * j + 1]);
  }
  merge_top2(lo[0], hi[0], lo[2], hi[2]);
  merge_top2(lo[1], hi[1], lo[3], hi[3]);
  merge_top2(lo[0], hi[0], lo[1], hi[1]);
  merge_top2(m1, m2, lo[0], hi[0]);
}

// Running state of one epilogue thread (one query row, half of the columns).
struct RowTop2 {
  int g1v, g1i, g2v, g2i;   // best / second best of the finished windows: value = |t|^2 - 2 q.t
  int m1, m2;               // top-2 of the current 512-column window as packed keys
  int thr;                  // a group matters iff 2 max(q.t) - min|t|^2 > thr; thr = -second best
};

__device__ __forceinline__ bool lex_lt(int v, int i, int gv, int gi) {
  return (v < gv) | ((v == gv) & (i < gi));
}

// lexicographic (value, index) insertion into the running top-2: order independent,
// branch free; (INT32_MAX, INT32_MAX) is a no-op
__device__ __forceinline__ void insert_vi(RowTop2& s, int v, int i) {
  const bool b1 = lex_lt(v, i, s.g1v, s.g1i);
  const bool b2 = lex_lt(v, i, s.g2v, s.g2i);
  s.g2v = b1 ? s.g1v : (b2 ? v : s.g2v);
  s.g2i = b1 ? s.g1i : (b2 ? i : s.g2i);
  s.g1v = b1 ? v : s.g1v;
  s.g1i = b1 ? i : s.g1i;
}

// exact keys of the 8 columns of group j (column keys from the shared-memory ring) -> (m1, m2)
__device__ __forceinline__ void group_insert(const uint32_t* a, uint32_t ck_addr, RowTop2& s) {
  const int4 c0 = lds_v4(ck_addr), c1 = lds_v4(ck_addr + 16);
  int k[8];
  k[0] = make_key(a[0], c0.x); k[1] = make_key(a[1], c0.y);
  k[2] = make_key(a[2], c0.z); k[3] = make_key(a[3], c0.w);
  k[4] = make_key(a[4], c1.x); k[5] = make_key(a[5], c1.y);
  k[6] = make_key(a[6], c1.z); k[7] = make_key(a[7], c1.w);
  insert8(k, s.m1, s.m2);
}

// ---- mode 5: the exact work is done by drain warps, the epilogue warps only filter --------
// Epilogue side: groups that pass the bound are appended (raw accumulators + bank row +
// owner lane) to the warp's ring; slots are assigned with a ballot, `wq` counts the events
// of this warp (warp-uniform register), the drain's read counter gives back-pressure.
__device__ __forceinline__ void queue_publish(uint32_t qa, uint32_t wq, int lane) {
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    sts_32_volatile(qa + kQCommit, wq);
  }
}

// Straight-line, branch-free append of one 8-column group: when the group's score beats the
// bound, the lane reserves a ring position with a shared-memory atomic and stores its event;
// every instruction is predicated, nothing diverges.  A position that would overrun the
// drain's (cached) read position is not written -- the caller detects that from the counter
// afterwards and replays the tile through the waiting slow path.
__device__ __forceinline__ void group_append_pred(int score, int thr, uint32_t qa, uint32_t rd,
                                                  const uint32_t* r, uint32_t meta) {
  static_assert(kQAcc == 0, "ring layout");
  // every temporary is defined on both paths (pos = rd when the group does not pass), so
  // ptxas has nothing to preserve across the predicated instructions
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .u32 pos, d, sl, a, m;\n\t"
      "setp.gt.s32 p, %0, %1;\n\t"
      "mov.u32 pos, %3;\n\t"
      "@p atom.shared.add.u32 pos, [%2+%13], %15;\n\t"
      "sub.u32 d, pos, %3;\n\t"
      "setp.lt.and.u32 q, d, 32, p;\n\t"
      "and.b32 sl, pos, 31;\n\t"
      "mad.lo.u32 a, sl, 32, %2;\n\t"
      "mad.lo.u32 m, sl, 4, %2;\n\t"
      "@q st.shared.v4.b32 [a], {%4, %5, %6, %7};\n\t"
      "@q st.shared.v4.b32 [a+16], {%8, %9, %10, %11};\n\t"
      "@q st.shared.b32 [m+%14], %12;\n\t"
      "}"
      ::"r"(score), "r"(thr), "r"(qa), "r"(rd), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
        "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(meta), "n"(kQReserve), "n"(kQMeta),
        "r"(1)
      : "memory");
}

__device__ __forceinline__ void filter_chunk_pred(const uint32_t (&r)[32], uint32_t gm_addr,
                                                  int thr, uint32_t qa, uint32_t rd,
                                                  uint32_t meta0) {
  const int4 nn = lds_v4(gm_addr);
  const int n8[4] = {nn.x, nn.y, nn.z, nn.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int a = __vimax3_s32(r[8 * j + 0], r[8 * j + 1], r[8 * j + 2]);
    const int b = __vimax3_s32(r[8 * j + 3], r[8 * j + 4], r[8 * j + 5]);
    const int gm = max(__vimax3_s32(a, b, r[8 * j + 6]), static_cast<int>(r[8 * j + 7]));
    group_append_pred(gm * 2 - n8[j], thr, qa, rd, &r[8 * j], meta0 + 8 * j);
  }
}

// Replay of a chunk whose events did not all fit: the reservations of the failed attempt are
// rolled back by the caller; here every group waits for ring space (warp-uniform), so
// nothing can be lost.
__device__ __forceinline__ void chunk_filter_append_slow(const uint32_t (&r)[32],
                                                         uint32_t gm_addr, uint32_t qa,
                                                         uint32_t& wq, uint32_t meta0, int lane,
                                                         int thr) {
  const int4 nn = lds_v4(gm_addr);
  const int n8[4] = {nn.x, nn.y, nn.z, nn.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int a = __vimax3_s32(r[8 * j + 0], r[8 * j + 1], r[8 * j + 2]);
    const int b = __vimax3_s32(r[8 * j + 3], r[8 * j + 4], r[8 * j + 5]);
    const int gm = max(__vimax3_s32(a, b, r[8 * j + 6]), static_cast<int>(r[8 * j + 7]));
    const bool p = gm * 2 - n8[j] > thr;
    const uint32_t hit = __ballot_sync(0xffffffffu, p);
    if (hit) {
      const uint32_t n = __popc(hit);
      uint32_t spins = 0;
      while (static_cast<int>(wq + n - static_cast<uint32_t>(lds_32_volatile(qa + kQRead))) >
             kQueueSlots) {
        queue_publish(qa, wq, lane);
        spin_guard(spins, 1);
      }
      if (p) {
        const uint32_t slot = (wq + __popc(hit & ((1u << lane) - 1u))) & (kQueueSlots - 1);
        sts_v4(qa + kQAcc + slot * 32, r[8 * j + 0], r[8 * j + 1], r[8 * j + 2], r[8 * j + 3]);
        sts_v4(qa + kQAcc + slot * 32 + 16, r[8 * j + 4], r[8 * j + 5], r[8 * j + 6],
               r[8 * j + 7]);
        sts_32(qa + kQMeta + slot * 4, meta0 + 8 * j);
      }
      wq += n;
    }
  }
}

// Drain side: one batch of up to 32 committed events, one per lane, possibly from several
// rings: lane's event = ring at qa, position pos; state_base = shared address of the 32
// running top-2 entries of that ring's epilogue warp; skey = a key unique per ring.
__device__ __forceinline__ void drain_batch(bool has, uint32_t qa, uint32_t state_base,
                                            uint32_t pos, uint32_t skey, int lane,
                                            const int32_t* __restrict__ ckey) {
  int v1 = INT32_MAX, i1 = INT32_MAX, v2 = INT32_MAX, i2 = INT32_MAX;
  uint32_t owner = 0;
  if (has) {
    const uint32_t slot = pos & (kQueueSlots - 1);
    const uint32_t meta = static_cast<uint32_t>(lds_32(qa + kQMeta + slot * 4));
    const int row = static_cast<int>(meta & ((1u << kRowBits) - 1));
    owner = meta >> kRowBits;
    const int4 a0 = lds_v4(qa + kQAcc + slot * 32), a1 = lds_v4(qa + kQAcc + slot * 32 + 16);
    const int4 c0 = __ldg(reinterpret_cast<const int4*>(ckey + row));
    const int4 c1 = __ldg(reinterpret_cast<const int4*>(ckey + row) + 1);
    int k[8];
    k[0] = make_key(a0.x, c0.x); k[1] = make_key(a0.y, c0.y);
    k[2] = make_key(a0.z, c0.z); k[3] = make_key(a0.w, c0.w);
    k[4] = make_key(a1.x, c1.x); k[5] = make_key(a1.y, c1.y);
    k[6] = make_key(a1.z, c1.z); k[7] = make_key(a1.w, c1.w);
    int e1 = INT32_MAX, e2 = INT32_MAX;
    insert8(k, e1, e2);
    const int base = (row - lds_32(qa + kQRow0)) & ~((1 << kColBits) - 1);
    v1 = e1 >> kColBits; i1 = base + (e1 & ((1 << kColBits) - 1));
    v2 = e2 >> kColBits; i2 = base + (e2 & ((1 << kColBits) - 1));
  }
  // events of the same owner are applied one per round (read-modify-write of its entry)
  const uint32_t peers = __match_any_sync(0xffffffffu, has ? skey * 32u + owner : 0x10000u + lane);
  const int rank = __popc(peers & ((1u << lane) - 1u));
  const int rounds = __reduce_max_sync(0xffffffffu, has ? __popc(peers) : 0);
  for (int r = 0; r < rounds; ++r) {
    if (has && rank == r) {
      const uint32_t sa = state_base + owner * 16;
      const int4 cur = lds_v4(sa);
      RowTop2 s = {cur.x, cur.y, cur.z, cur.w, 0, 0, 0};
      insert_vi(s, v1, i1);
      insert_vi(s, v2, i2);
      sts_v4(sa, s.g1v, s.g1i, s.g2v, s.g2i);
    }
    __syncwarp();
  }
}

// Top-2 update with the thread's 64 columns of one tile (two 32-column chunks);
// ck_addr = shared address of their keys.
//   kMode 0: every group is inserted (2.5 min/max + 1 IMAD per element).
//   kMode 1: only groups whose raw maximum beats the bound of some row of the warp.  All
//            eight 3-input-max trees and votes are issued before any insert, so they overlap.
template <int kMode>
__device__ __forceinline__ void tile_update(const uint32_t (&r0)[32], const uint32_t (&r1)[32],
                                            uint32_t ck_addr, uint32_t gm_addr, RowTop2& s) {
  if constexpr (kMode == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) group_insert(&r0[8 * j], ck_addr + 32 * j, s);
#pragma unroll
    for (int j = 0; j < 4; ++j) group_insert(&r1[8 * j], ck_addr + 128 + 32 * j, s);
  } else {
    bool h[8];
    const int4 n0 = lds_v4(gm_addr), n1 = lds_v4(gm_addr + 16);
    const int n8[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t* r = j < 4 ? &r0[8 * j] : &r1[8 * (j - 4)];
      const int a = __vimax3_s32(r[0], r[1], r[2]);
      const int b = __vimax3_s32(r[3], r[4], r[5]);
      const int gm = max(__vimax3_s32(a, b, r[6]), static_cast<int>(r[7]));
      h[j] = __any_sync(0xffffffffu, gm * 2 - n8[j] > s.thr);
    }
    bool any = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (h[j]) {
        group_insert(j < 4 ? &r0[8 * j] : &r1[8 * (j - 4)], ck_addr + 32 * j, s);
        any = true;
      }
    }
    if (any) {
      // the window's second best also bounds what can still enter (values, not keys)
      const int w2 = s.m2 >> kColBits;
      if (w2 < (1 << 22)) s.thr = max(s.thr, -w2);
    }
  }
}

// The unfiltered first tiles of a sweep in mode 5, kept out of line so that the filtered
// steady-state loop stays small in the instruction cache.
__device__ __noinline__ void exact_tile_mode5(uint32_t ta, uint32_t bar_empty_addr, int lane,
                                              uint32_t ck_addr, uint32_t gm_addr, RowTop2& st) {
  uint32_t r0[32], r1[32];
  tmem_ld_x32(ta, r0);
  tmem_ld_x32(ta + 32, r1);
  tmem_ld_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar_empty_addr);
  tile_update<0>(r0, r1, ck_addr, gm_addr, st);
}

template <int kMode>
__global__ void __launch_bounds__(kKnnThreads, 1)
knn2_kernel(const __grid_constant__ CUtensorMap tmap, const int32_t* __restrict__ ckey,
            const int32_t* __restrict__ gmin8, const int32_t* __restrict__ norm, const PairDesc* __restrict__ pairs,
            const int2* __restrict__ items, int n_items, Knn2* __restrict__ knn_out, int dbg) {
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
      mbar_init(bar_a_empty(b), (dbg & 2) ? kMmaWarps : kMmaWarps + kEpiWarps);
    }
    for (int b = 0; b < kAccBufs; ++b)
      for (int h = 0; h < 2; ++h) {
        mbar_init(bar_t_full(b, h), 1);
        mbar_init(bar_t_empty(b, h), kEpiWarps / 2);
      }
    fence_mbar_init();
  }
  if (threadIdx.x < kTileM)    // (item tag, second-best value) slots of the row-sharing threads
    sts_v4(smem_base + kOffShare + threadIdx.x * 16, 0xffffffffu, 0, 0xffffffffu, 0);
  if (threadIdx.x < kEpiWarps * 32)
    sts_v4(smem_base + kOffState + threadIdx.x * 16, INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX);
  if (threadIdx.x < kEpiWarps) {
    sts_v4(smem_base + kOffQueue + threadIdx.x * kQBytes + kQCommit, 0, 0, 0, 0);
    if (threadIdx.x == 0) sts_32(smem_base + kOffDone, 0);
  }
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
  if (warp < kFirstEpiWarp) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsCtl));
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
        mbar_wait(bar_a_empty(abuf), aphase ^ 1);
        if (elect_one()) {
          info[abuf].ntiles = ntiles;
          info[abuf].rows_valid = pd.nq - mblk * kTileM;
          info[abuf].norm_row = pd.q_row0 + mblk * kTileM;
          info[abuf].t_row0 = pd.t_row0;
          info[abuf].nt_min = pd.nt_min;
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
            if (dbg & 1) {                       // timing experiment: no operand traffic
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
    } else if (warp < kFirstMmaWarp) {
      // ===================================================== drain warps (mode 5 only)
      if constexpr (kMode == 5) {
        const int d = warp - 1;                   // serves the epilogue warps e with e % 3 == d
        uint32_t rd[6] = {0, 0, 0, 0, 0, 0};
        uint32_t idle = 0;
        while (!(dbg & 8)) {
          // gather up to 32 committed events over this warp's rings, one per lane
          int take[6], total = 0;
#pragma unroll
          for (int qi = 0; qi < 6; ++qi) {
            const int e = d + 3 * qi;
            take[qi] = 0;
            if (e < kEpiWarps) {
              const uint32_t c = static_cast<uint32_t>(
                  lds_32_volatile(smem_base + kOffQueue + e * kQBytes + kQCommit));
              take[qi] = min(static_cast<int>(c - rd[qi]), 32 - total);
              total += take[qi];
            }
          }
          if (total == 0) {
            if (lds_32_volatile(smem_base + kOffDone) == kEpiWarps) break;
            __nanosleep(100);
            spin_guard(idle, 3);
            continue;
          }
          idle = 0;
          __threadfence_block();
          uint32_t qa = 0, sb = 0, pos = 0, skey = 0;
          int first = 0;
#pragma unroll
          for (int qi = 0; qi < 6; ++qi) {
            if (lane >= first && lane < first + take[qi]) {
              const int e = d + 3 * qi;
              qa = smem_base + kOffQueue + e * kQBytes;
              sb = smem_base + kOffState + e * 32 * 16;
              pos = rd[qi] + (lane - first);
              skey = e;
            }
            first += take[qi];
          }
          drain_batch(lane < total, qa, sb, pos, skey, lane, ckey);
          __syncwarp();
          __threadfence_block();
#pragma unroll
          for (int qi = 0; qi < 6; ++qi) {
            if (take[qi] > 0) {
              rd[qi] += take[qi];
              if (lane == 0)
                sts_32_volatile(smem_base + kOffQueue + (d + 3 * qi) * kQBytes + kQRead, rd[qi]);
            }
          }
        }
      }
    } else if (warp >= kFirstMmaWarp) {
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
          if (!(dbg & 2)) mbar_wait(bar_t_empty(mp, mh), ((ts >> 1) & 1) ^ 1);
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
    const int half = e >> 3;                       // which 128-row half of the block
    const int chalf = (e >> 2) & 1;                // which 64-column half of every tile
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int row_in_blk = half * kHalfM + quarter * 32 + lane;
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                            half * kTileN + chalf * kColsPerThread;
    const uint32_t merge_addr = smem_base + kOffMerge + row_in_blk * 16;
    const uint32_t share_own = smem_base + kOffShare + row_in_blk * 16 + chalf * 8;
    const uint32_t share_other = smem_base + kOffShare + row_in_blk * 16 + (chalf ^ 1) * 8;
    const int pair_bar = 1 + half * 4 + quarter;   // named barrier of the two column halves
    // mode 5: this warp's event ring, this thread's and its row partner's drain-owned entries
    const uint32_t qa = smem_base + kOffQueue + e * kQBytes;
    const uint32_t state_own = smem_base + kOffState + (e * 32 + lane) * 16;
    const uint32_t state_other = smem_base + kOffState + ((e ^ 4) * 32 + lane) * 16;
    uint32_t wq = 0, rdc = 0;                      // events appended / known consumed (uniform)
    uint32_t buf = 0, bphase = 0, abuf = 0, mslot = 0, tile_seq = 0;
    for (int item = blockIdx.x; item < n_items && !(dbg & 2); item += gridDim.x) {
      RowTop2 st = {INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX, INT32_MIN};
      int ntiles = 1, rows_valid = 0, norm_row = 0, nt_min = 0, t_row0 = 0;
      int64_t knn_row = 0;
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(bar_t_full(buf, half), bphase);
        tc_fence_after();
        if (t == 0) {
          ntiles = info[abuf].ntiles;
          rows_valid = info[abuf].rows_valid;
          norm_row = info[abuf].norm_row;
          nt_min = info[abuf].nt_min;
          t_row0 = info[abuf].t_row0;
          knn_row = info[abuf].knn_row;
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(bar_a_empty(abuf));
            if (kMode == 5) sts_32(qa + kQRow0, t_row0);   // published with the first events
          }
          abuf ^= 1;
        }
        const uint32_t ta = t_addr + buf * (2 * kTileN);
        if constexpr (kMode >= 2 && kMode <= 4) {
          // timing experiments only (results are garbage):
          // 2 = drain TMEM, 3 = handshake only, 4 = drain + max tree + compare
          if constexpr (kMode == 2) {
            uint32_t r0[32], r1[32];
            tmem_ld_x32(ta, r0);
            tmem_ld_x32(ta + 32, r1);
            tmem_ld_wait();
            st.g1v = min(st.g1v, static_cast<int>(r0[0] ^ r1[31]));
          }
          if constexpr (kMode == 4) {
            // fast-path anatomy: dbg bits add the ingredients of the real epilogue one by one
            //   32: load / wait per chunk instead of both loads then one wait
            //   64: bound from shared memory + per-group norms (LDS, IMAD)
            //  128: real divergent branch with stores per group (never taken)
            //  256: per-chunk __syncwarp + vote + shared-memory counter read
            int thr = st.thr;
            int n8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const uint32_t slot = tile_seq % kCkSlots;
            const uint32_t gm_addr = sGm + slot * kGmBytes + chalf * (kColsPerThread / 8 * 4);
            if (dbg & 64) {
              const int og = lds_32(state_own + 8), pg = lds_32(state_other + 8);
              thr = max(thr, -min(og, pg));
              const int4 n0 = lds_v4(gm_addr), n1 = lds_v4(gm_addr + 16);
              n8[0] = n0.x; n8[1] = n0.y; n8[2] = n0.z; n8[3] = n0.w;
              n8[4] = n1.x; n8[5] = n1.y; n8[6] = n1.z; n8[7] = n1.w;
            }
            uint32_t r0[32], r1[32];
            tmem_ld_x32(ta, r0);
            if (dbg & 32) tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              if (c == 1) {
                tmem_ld_x32(ta + 32, r1);
                tmem_ld_wait();
              }
              bool ovf = false;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t* r = c == 0 ? &r0[8 * j] : &r1[8 * j];
                const int a = __vimax3_s32(r[0], r[1], r[2]);
                const int b = __vimax3_s32(r[3], r[4], r[5]);
                const int gm = max(__vimax3_s32(a, b, r[6]), static_cast<int>(r[7]));
                if (dbg & 128) {
                  if (gm * 2 - n8[4 * c + j] > thr + 0x40000000) {     // never true
                    sts_v4(qa + kQAcc, r[0], r[1], r[2], r[3]);
                    sts_v4(qa + kQAcc + 16, r[4], r[5], r[6], r[7]);
                    ovf = true;
                  }
                } else if (gm * 2 - n8[4 * c + j] > thr) {
                  st.g1i = gm; ++st.g2v;
                }
              }
              if (dbg & 256) {
                __syncwarp();
                if (__any_sync(0xffffffffu, ovf)) st.g2i++;
                st.g1v += lds_32(qa + kQReserve);
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_t_empty(buf, half));
        } else {
          const uint32_t slot = tile_seq % kCkSlots;
          const uint32_t ck_addr = sCk + slot * kCkBytes + chalf * (kColsPerThread * 4);
          const uint32_t gm_addr = sGm + slot * kGmBytes + chalf * (kColsPerThread / 8 * 4);
          if constexpr (kMode == 5) {
            if (t >= kExactTiles || (dbg & 16)) {
              // filtered tile: one 32-register chunk at a time (the append path needs the
              // registers), bound = second best of the row so far (either column half), read
              // from the drain-owned entries; any earlier value of this item is a valid bound
              const int og = lds_32(state_own + 8), pg = lds_32(state_other + 8);
              int bound = og;
              if (pg < (1 << 22)) bound = min(bound, pg + 1);
              const int thr = (dbg & 4) ? INT32_MAX : (bound < (1 << 22) ? -bound : INT32_MIN);
              const uint32_t meta0 =
                  static_cast<uint32_t>(t_row0 + t * kTileN + chalf * kColsPerThread) |
                  (static_cast<uint32_t>(lane) << kRowBits);
              uint32_t r0[32], r1[32];
              tmem_ld_x32(ta, r0);
              tmem_ld_x32(ta + 32, r1);
              tmem_ld_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_t_empty(buf, half));
              filter_chunk_pred(r0, gm_addr, thr, qa, rdc, meta0);
              filter_chunk_pred(r1, gm_addr + 16, thr, qa, rdc, meta0 + 32);
              __syncwarp();
              uint32_t nres = static_cast<uint32_t>(lds_32_volatile(qa + kQReserve));
              if (static_cast<int>(nres - rdc) > kQueueSlots) {
                // some events did not fit behind the cached read position: drop this tile's
                // events (positions >= wq, unpublished) and replay it with the waiting path
                __syncwarp();
                if (lane == 0) sts_32(qa + kQReserve, wq);
                __syncwarp();
                chunk_filter_append_slow(r0, gm_addr, qa, wq, meta0, lane, thr);
                chunk_filter_append_slow(r1, gm_addr + 16, qa, wq, meta0 + 32, lane, thr);
                if (lane == 0) sts_32(qa + kQReserve, wq);
                __syncwarp();
                nres = wq;
              }
              if (nres != wq || static_cast<int>(nres - rdc) > kQueueSlots / 2) {
                wq = nres;
                queue_publish(qa, wq, lane);
                rdc = static_cast<uint32_t>(lds_32_volatile(qa + kQRead));
              }
            } else {
              // the first tiles of a sweep are done here, unfiltered; their result seeds
              // the entry the drain warp continues with
              exact_tile_mode5(ta, bar_t_empty(buf, half), lane, ck_addr, gm_addr, st);
              if (t == kExactTiles - 1 || t == ntiles - 1) {
                insert_vi(st, st.m1 >> kColBits, st.m1 & ((1 << kColBits) - 1));
                insert_vi(st, st.m2 >> kColBits, st.m2 & ((1 << kColBits) - 1));
                if (ntiles > kExactTiles) sts_v4(state_own, st.g1v, st.g1i, st.g2v, st.g2i);
              }
            }
          } else {
          // pull this thread's 64 accumulators out of TMEM and hand the buffer back at once
          uint32_t r0[32], r1[32];
          tmem_ld_x32(ta, r0);
          tmem_ld_x32(ta + 32, r1);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_t_empty(buf, half));
          tile_update<kMode>(r0, r1, ck_addr, gm_addr, st);
          if ((t & 3) == 3 || t == ntiles - 1) {
            // close the 512-column window: merge its packed top-2 into the (value, index)
            // pairs, then tighten the bound, also with the row partner's second best
            const int base = (t & ~3) * kTileN;
            insert_vi(st, st.m1 >> kColBits, base + (st.m1 & ((1 << kColBits) - 1)));
            insert_vi(st, st.m2 >> kColBits, base + (st.m2 & ((1 << kColBits) - 1)));
            st.m1 = INT32_MAX;
            st.m2 = INT32_MAX;
            int bound = st.g2v;
            if (kMode == 1) {
              sts_v2(share_own, item, st.g2v);
              const int2 o = lds_v2(share_other);   // any earlier value of this item is valid
              if (o.x == item && o.y < (1 << 22)) bound = min(bound, o.y + 1);
            }
            st.thr = bound < (1 << 22) ? -bound : INT32_MIN;
          }
          }
        }
        ++tile_seq;
        if (++buf == kAccBufs) {
          buf = 0;
          bphase ^= 1;
        }
      }
      if (kMode == 5 && ntiles > kExactTiles) {
        // wait until the drain warp has applied every event of this sweep, take the entry
        // over and reset it for the next item
        queue_publish(qa, wq, lane);
        uint32_t spins = 0;
        while (static_cast<uint32_t>(lds_32_volatile(qa + kQRead)) != wq) spin_guard(spins, 2);
        rdc = wq;
        __threadfence_block();
        const int4 f = lds_v4(state_own);
        st.g1v = f.x; st.g1i = f.y; st.g2v = f.z; st.g2i = f.w;
        sts_v4(state_own, INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX);
      }
      // merge the two column halves of the row: the upper half hands its top-2 over
      const uint32_t slot = merge_addr + mslot * (kTileM * 16);
      mslot ^= 1;
      if (chalf == 1) sts_v4(slot, st.g1v, st.g1i, st.g2v, st.g2i);
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      if (chalf == 0) {
        const int4 o = lds_v4(slot);
        insert_vi(st, o.x, o.y);
        insert_vi(st, o.z, o.w);
        if (row_in_blk < rows_valid) {
          const int nq2 = __ldg(norm + norm_row + row_in_blk);
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
    if (kMode == 5) {
      __syncwarp();
      if (lane == 0) atoms_add(smem_base + kOffDone, 1);   // lets the drain warps leave
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

// mode 0: unfiltered exact top-2 epilogue; mode 1: threshold-filtered (default, same results)
cudaError_t launch_knn2(int mode, const CUtensorMap& tmap, const int32_t* ckey,
                        const int32_t* gmin8, const int32_t* norm, const PairDesc* pairs, const int2* items,
                        int n_items, Knn2* knn_out, int n_sms, cudaStream_t stream) {
  const int dbg = mode >> 4;   // timing experiments (results invalid), see tools/exp_modes.py
  mode &= 15;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(knn2_kernel<0>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, kKnnSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(knn2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kKnnSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(knn2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kKnnSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(knn2_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kKnnSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(knn2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kKnnSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(knn2_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kKnnSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int grid = n_items < n_sms ? n_items : n_sms;
  if (grid <= 0) return cudaSuccess;
  if (mode == 5)
    knn2_kernel<5><<<grid, kKnnThreads, kKnnSmemBytes, stream>>>(tmap, ckey, gmin8, norm, pairs, items,
                                                                 n_items, knn_out, dbg);
  else if (mode == 4)
    knn2_kernel<4><<<grid, kKnnThreads, kKnnSmemBytes, stream>>>(tmap, ckey, gmin8, norm, pairs, items,
                                                                 n_items, knn_out, dbg);
  else if (mode == 2)
    knn2_kernel<2><<<grid, kKnnThreads, kKnnSmemBytes, stream>>>(tmap, ckey, gmin8, norm, pairs, items,
                                                                 n_items, knn_out, dbg);
  else if (mode == 3)
    knn2_kernel<3><<<grid, kKnnThreads, kKnnSmemBytes, stream>>>(tmap, ckey, gmin8, norm, pairs, items,
                                                                 n_items, knn_out, dbg);
  else if (mode == 0)
    knn2_kernel<0><<<grid, kKnnThreads, kKnnSmemBytes, stream>>>(tmap, ckey, gmin8, norm, pairs, items,
                                                                 n_items, knn_out, dbg);
  else
    knn2_kernel<1><<<grid, kKnnThreads, kKnnSmemBytes, stream>>>(tmap, ckey, gmin8, norm, pairs, items,
                                                                 n_items, knn_out, dbg);
  return cudaGetLastError();
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
