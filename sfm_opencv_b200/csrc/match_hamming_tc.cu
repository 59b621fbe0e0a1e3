// match_hamming_tc.cu -- cv::NORM_HAMMING2 k=2 nearest neighbours on the sm_100a tensor cores.
//
// The LIVE reference path (AKAZE 61-byte descriptors, `BFMatcher matcher(NORM_HAMMING2)`,
// OpenCV_SFM/NViewReconstuct.cpp:797, :875-877) as an exact integer contraction:
// NORM_HAMMING2 counts the 2-bit cells in which two descriptors differ.  Map the four values of a
// cell to the vertices of a regular tetrahedron in {+1,-1}^3,
//     0 -> (+,+,+)   1 -> (+,-,-)   2 -> (-,+,-)   3 -> (-,-,+)
// so that two cells have dot product 3 when equal and -1 when different.  Over the C = 4 * bytes
// cells of a descriptor:  dot = 3 (C - D) - D = 3 C - 4 D, i.e.  D = (3 C - dot) / 4  exactly.
// A 61-byte descriptor becomes a 732-dimensional s8 vector, stored as 768 bytes = 6 K-chunks of 128
// (zero padded: zeros add nothing), and q.t comes from tcgen05.mma.kind::i8 (s8 x s8 -> s32).
// 3 dimensions per cell is the minimum: the 4 x 4 matrix alpha [a == b] + beta has rank >= 3.
//
// One persistent CTA per SM; a work item is a 128-row query block of one image pair:
//   warp 0        TMA producer: the block's 6 A chunks once per item (96 KB, resident), then the
//                 train image as 128-row x 128-byte B chunks (16 KB) through an mbarrier ring,
//                 plus 512 B of column keys per tile
//   warps 4..6    MMA issuers, tile ts belongs to issuer ts % 3: 6 chunks x 4 MMAs (128x128x32) per
//                 tile into one of 4 TMEM accumulators (4 x 128 columns); the issuers hide each
//                 other's commit / wait bubbles (tools/exp_probe.py).  Each issuer has its own ring
//                 of 2 B stages: with one shared ring an issuer would wait for a stage's phase two
//                 rounds ahead, and a parity wait cannot tell round r from round r + 2
//   warps 8..15   epilogue: 2 column halves x 4 TMEM lane quarters; a thread owns one query row
//                 and 64 columns of every tile and keeps the exact top-2 as packed keys
//                 ((-2 dot) << 10 | column: an integer min is OpenCV's (distance, lower index)
//                 order) with the unfiltered insert of knn_epilogue.cuh -- a tile costs the
//                 tensor pipe >= 1626 cycles (24 MMAs), three times the epilogue's share, so no
//                 filter is needed: the kernel is bound by the MMAs and by the L2 -> SM traffic of
//                 the B chunks (16 KB per 4 MMAs; a 128-row block is all that fits beside them).
// Results are the same Knn2 rows as the SIFT kernel with SQUARED distances, so the filter kernels
// (sqrtf of an exact square) are shared unchanged.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "knn_epilogue.cuh"
#include "match_types.h"
#include "ptx.cuh"

namespace sfm {

constexpr int kHtChunks = 6;                     // K-chunks of 128 bytes per descriptor (768 dims)
constexpr int kHtRowBytes = kHtChunks * kDim;    // 768
constexpr int kHtM = 128;                        // query rows per work item
#ifndef SFM_HT_ISSUERS
#define SFM_HT_ISSUERS 3
#endif
constexpr int kHtIssuers = SFM_HT_ISSUERS;       // MMA issuing warps: tile ts belongs to issuer ts % kHtIssuers
constexpr int kHtStages = 6;                     // B chunk stages (16 KB each): one ring per issuer
constexpr int kHtRing = kHtStages / kHtIssuers;  // stages per ring
static_assert(kHtIssuers == 2 || kHtIssuers == 3, "issuers");
static_assert(kHtChunks % kHtRing == 0, "a tile's chunks are requested in groups of one ring");
constexpr int kHtAccBufs = 4;                    // TMEM accumulators (128 columns each)
constexpr int kHtCkSlots = 16;                   // ring of per-tile column keys (512 B each)
constexpr int kHtMmaWarp0 = 4, kHtEpiWarp0 = 8, kHtEpiWarps = 8;
constexpr int kHtThreads = (kHtEpiWarp0 + kHtEpiWarps) * 32;   // 512
constexpr int kHtWinTiles = (1 << kColBits) / kTileN;          // tiles per packed-key window (8)

constexpr uint32_t kHtChunkBytes = kTileN * kDim;              // 16 KB (A chunk == B chunk)
constexpr uint32_t kHtOffA = 0;
constexpr uint32_t kHtOffB = kHtOffA + kHtChunks * kHtChunkBytes;
constexpr uint32_t kHtOffCk = kHtOffB + kHtStages * kHtChunkBytes;
constexpr uint32_t kHtOffInfo = kHtOffCk + kHtCkSlots * kTileN * 4;
constexpr uint32_t kHtOffMerge = kHtOffInfo + 2 * 32;          // 2 slots x 128 rows x int4
constexpr uint32_t kHtOffBar = kHtOffMerge + 2 * kHtM * 16;
constexpr uint32_t kHtNumBars = 2 * kHtStages + 2 + 2 * kHtAccBufs;
constexpr uint32_t kHtOffTmemPtr = kHtOffBar + kHtNumBars * 8;
constexpr uint32_t kHtSmemBytes = kHtOffTmemPtr + 16 + 1024;   // + alignment slack
static_assert(kHtSmemBytes <= 227 * 1024, "shared memory budget");

struct HtInfo {           // what the producer tells the other warps about an item (32 bytes)
  int32_t ntiles, rows_valid, pad0, pad1;
  int64_t knn_row;
  int64_t pad2;
};
static_assert(sizeof(HtInfo) == 32, "smem layout");

// Instruction descriptor for kind::i8, s8 x s8 -> s32, K-major A and B (see make_idesc_u8).
__host__ __device__ constexpr uint32_t make_idesc_s8(uint32_t M, uint32_t N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// Expands n binary descriptors of `bytes` bytes into 768-byte tetrahedron rows (header comment):
// one thread per (descriptor, byte) writes the 12 s8 values of the byte's four cells.
__global__ void bin_expand_tc_kernel(const uint8_t* __restrict__ src, int n, int bytes, int row0,
                                     uint8_t* __restrict__ bank) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(n) * bytes) return;
  const int r = static_cast<int>(i / bytes), b = static_cast<int>(i % bytes);
  const uint32_t v = src[i];
  // cell value c -> bytes (x, y, z): 0 -> 01 01 01, 1 -> 01 ff ff, 2 -> ff 01 ff, 3 -> ff ff 01
  uint32_t w[3] = {0, 0, 0};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t cell = (v >> (2 * c)) & 3u;
    const uint32_t x = (cell & 2u) ? 0xffu : 0x01u;                 // - for 2, 3
    const uint32_t y = (cell == 1u || cell == 3u) ? 0xffu : 0x01u;  // - for 1, 3
    const uint32_t z = (cell == 1u || cell == 2u) ? 0xffu : 0x01u;  // - for 1, 2
    const uint32_t xyz[3] = {x, y, z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int byte = 3 * c + k;                                   // 0..11 within the 12-byte group
      w[byte >> 2] |= xyz[k] << (8 * (byte & 3));
    }
  }
  uint32_t* dst = reinterpret_cast<uint32_t*>(bank + static_cast<size_t>(row0 + r) * kHtRowBytes + 12 * b);
  dst[0] = w[0];
  dst[1] = w[1];
  dst[2] = w[2];
}

// Column keys of a binary image: (0 << kColBits | row & 1023) for real rows -- HAMMING2 has no
// norm term -- and the sentinel for the padding rows of the last tile, so that a padding row is
// never selected while the train image has two real rows.
__global__ void bin_ckey_kernel(int n, int n_pad, int row0, int32_t* __restrict__ ckey) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_pad) ckey[row0 + r] = ((r < n ? 0 : kNormPad) << kColBits) | (r & ((1 << kColBits) - 1));
}

__global__ void __launch_bounds__(kHtThreads, 1)
hamming2_tc_kernel(const __grid_constant__ CUtensorMap tmap, const int32_t* __restrict__ ckey,
                   const PairDesc* __restrict__ pairs, const int2* __restrict__ items, int n_items,
                   int dot_equal, Knn2* __restrict__ knn_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base + kHtOffA, sB = smem_base + kHtOffB, sCk = smem_base + kHtOffCk;
  const uint32_t bar0 = smem_base + kHtOffBar;
  auto bar_full = [&](uint32_t s) { return bar0 + 8u * s; };
  auto bar_empty = [&](uint32_t s) { return bar0 + 8u * (kHtStages + s); };
  const uint32_t bar_a_full = bar0 + 8u * (2 * kHtStages), bar_a_empty = bar0 + 8u * (2 * kHtStages + 1);
  auto bar_t_full = [&](uint32_t b) { return bar0 + 8u * (2 * kHtStages + 2 + b); };
  auto bar_t_empty = [&](uint32_t b) { return bar0 + 8u * (2 * kHtStages + 2 + kHtAccBufs + b); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + kHtOffTmemPtr);
  volatile HtInfo* info = reinterpret_cast<volatile HtInfo*>(smem_gen + kHtOffInfo);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap);
    for (uint32_t s = 0; s < kHtStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_a_full, 1);
    mbar_init(bar_a_empty, kHtIssuers + kHtEpiWarps);   // every MMA warp + every epilogue warp
    for (uint32_t b = 0; b < kHtAccBufs; ++b) {
      mbar_init(bar_t_full(b), 1);
      mbar_init(bar_t_empty(b), kHtEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == kHtMmaWarp0) {
    tmem_alloc(smem_base + kHtOffTmemPtr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================================================== TMA producer
    uint32_t tile_seq = 0, item_seq = 0;                         // tiles / items of this CTA so far
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_seq) {
      const int2 it = __ldg(items + item);
      const PairDesc pd = pairs[it.x];
      const int ntiles = (pd.nt + kTileN - 1) / kTileN;
      const int q_row = pd.q_row0 + it.y * kHtM;
      mbar_wait(bar_a_empty, (item_seq & 1) ^ 1);                // the previous item's MMAs are done
      if (elect_one()) {
        info[item_seq & 1].ntiles = ntiles;
        info[item_seq & 1].rows_valid = pd.nq - it.y * kHtM;
        info[item_seq & 1].knn_row = pd.knn_off + static_cast<int64_t>(it.y) * kHtM;
        mbar_arrive_expect_tx(bar_a_full, kHtChunks * kHtChunkBytes);
        for (int kc = 0; kc < kHtChunks; ++kc)
          tma_load_2d(sA + kc * kHtChunkBytes, &tmap, bar_a_full, kc * kDim, q_row);
      }
      __syncwarp();
      // kHtIssuers tiles in flight, one per MMA issuer: the chunks of tiles t .. t + kHtIssuers - 1 are
      // requested interleaved, kHtRing chunks of each tile in turn, so that every issuer has operands
      // at the same time (tile after tile, the next issuer would only start when the previous one is
      // done, and a single issuing thread leaves the tensor pipe idle during its commit / wait
      // bubbles).  Inside a ring the order stays (tile, chunk), which is what the issuers expect.
      for (int t = 0; t < ntiles; t += kHtIssuers) {
        const int group = ntiles - t < kHtIssuers ? ntiles - t : kHtIssuers;
        for (int h = 0; h < kHtChunks / kHtRing; ++h) {
          for (int u = 0; u < group; ++u) {
            const uint32_t ts = tile_seq + u;
            const int row = pd.t_row0 + (t + u) * kTileN;
            for (int kc = h * kHtRing; kc < (h + 1) * kHtRing; ++kc) {
              const uint32_t cr = (ts / kHtIssuers) * kHtChunks + kc;   // chunks this tile's ring has seen
              const uint32_t stage = (ts % kHtIssuers) * kHtRing + cr % kHtRing;
              mbar_wait(bar_empty(stage), ((cr / kHtRing) & 1) ^ 1);
              if (elect_one()) {
                mbar_arrive_expect_tx(bar_full(stage), kHtChunkBytes + (kc == 0 ? kTileN * 4 : 0));
                tma_load_2d(sB + stage * kHtChunkBytes, &tmap, bar_full(stage), kc * kDim, row);
                if (kc == 0)
                  bulk_load_1d(sCk + (ts % kHtCkSlots) * (kTileN * 4), ckey + row, kTileN * 4, bar_full(stage));
              }
              __syncwarp();
            }
          }
        }
        tile_seq += group;
      }
    }
  } else if (warp >= kHtMmaWarp0 && warp < kHtMmaWarp0 + kHtIssuers) {
    // ===================================================== MMA issuers: tile ts belongs to issuer ts % kHtIssuers
    const uint32_t mp = warp - kHtMmaWarp0;
    constexpr uint32_t idesc = make_idesc_s8(kHtM, kTileN);
    uint32_t seq = 0, item_seq = 0;                              // tiles / items of this CTA so far
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_seq) {
      mbar_wait(bar_a_full, item_seq & 1);
      const uint32_t ntiles = info[item_seq & 1].ntiles;
      const uint32_t end = seq + ntiles;
      uint32_t ts = seq + (mp + kHtIssuers - seq % kHtIssuers) % kHtIssuers;   // this warp's first tile
      if (ts >= end) {                                           // short item: no tile for this warp
        if (elect_one()) mbar_arrive(bar_a_empty);
        __syncwarp();
      }
      for (; ts < end; ts += kHtIssuers) {
        const uint32_t buf = ts % kHtAccBufs;
        mbar_wait(bar_t_empty(buf), ((ts / kHtAccBufs) & 1) ^ 1);
        const uint32_t d_tmem = tmem_base + buf * kTileN;
        for (uint32_t kc = 0; kc < kHtChunks; ++kc) {
          const uint32_t cr = (ts / kHtIssuers) * kHtChunks + kc;   // this warp's ring, in order
          const uint32_t stage = mp * kHtRing + cr % kHtRing;
          mbar_wait(bar_full(stage), (cr / kHtRing) & 1);
          tc_fence_after();
          const uint64_t a_desc = make_smem_desc_sw128(sA + kc * kHtChunkBytes);
          const uint64_t b_desc = make_smem_desc_sw128(sB + stage * kHtChunkBytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kDim / 32; ++k)
              umma_i8(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kc | k) > 0);
            umma_commit(bar_empty(stage));
            if (kc == kHtChunks - 1) {
              umma_commit(bar_t_full(buf));
              if (ts + kHtIssuers >= end) umma_commit(bar_a_empty);   // this warp's last tile of the item
            }
          }
          __syncwarp();
        }
      }
      seq = end;
    }
  } else if (warp >= kHtEpiWarp0) {
    // ===================================================== epilogue: running top-2 per row
    const int e = warp - kHtEpiWarp0;
    const int part = e >> 2;                       // which 64-column half of every tile
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int row_in_blk = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + part * 64;
    const uint32_t merge_row = smem_base + kHtOffMerge + row_in_blk * 16;
    uint32_t seq = 0, item_seq = 0, mslot = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_seq) {
      RowTop2 st = {INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX, INT32_MAX};
      int ntiles = 1, rows_valid = 0;
      int64_t knn_row = 0;
      for (int t = 0; t < ntiles; ++t, ++seq) {
        const uint32_t buf = seq % kHtAccBufs;
        mbar_wait(bar_t_full(buf), (seq / kHtAccBufs) & 1);
        tc_fence_after();
        if (t == 0) {                              // published before the item's first MMA
          ntiles = info[item_seq & 1].ntiles;
          rows_valid = info[item_seq & 1].rows_valid;
          knn_row = info[item_seq & 1].knn_row;
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_a_empty);
        }
        uint32_t ra[32], rb[32];
        tmem_ld_x32(t_lane + buf * kTileN, ra);
        tmem_ld_x32(t_lane + buf * kTileN + 32, rb);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_t_empty(buf));            // tile is out of TMEM
        const uint32_t ck_addr = sCk + (seq % kHtCkSlots) * (kTileN * 4) + part * 256;
#pragma unroll
        for (int j = 0; j < 4; ++j) group_insert(&ra[8 * j], ck_addr + 32 * j, st);
#pragma unroll
        for (int j = 0; j < 4; ++j) group_insert(&rb[8 * j], ck_addr + 128 + 32 * j, st);
        if ((t & (kHtWinTiles - 1)) == kHtWinTiles - 1 || t + 1 == ntiles)
          close_window(st, (t & ~(kHtWinTiles - 1)) * kTileN);
      }
      // merge the two column halves of the row: part 1 hands its top-2 to part 0
      const uint32_t slot = merge_row + mslot * (kHtM * 16);
      mslot ^= 1;
      if (part == 1) sts_v4(slot, st.g1v, st.g1i, st.g2v, st.g2i);
      asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(64) : "memory");
      if (part == 0) {
        const int4 w = lds_v4(slot);
        insert_vi(st, w.x, w.y);
        insert_vi(st, w.z, w.w);
        if (row_in_blk < rows_valid) {
          // value = -2 dot, dot = dot_equal - 4 D  ->  D = (2 dot_equal + value) / 8, exact
          const int d0 = (2 * dot_equal + st.g1v) >> 3, d1 = (2 * dot_equal + st.g2v) >> 3;
          Knn2 out;
          out.j0 = st.g1i;
          out.j1 = st.g2i;
          out.d0 = d0 * d0;
          out.d1 = d1 * d1;
          *reinterpret_cast<int4*>(&knn_out[knn_row + row_in_blk]) = *reinterpret_cast<int4*>(&out);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kHtMmaWarp0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// -------------------------------------------------------------------------------------
cudaError_t launch_bin_expand_tc(const uint8_t* src, int n, int bytes, int row0, uint8_t* bank,
                                 int32_t* ckey, cudaStream_t s) {
  constexpr int kBinRowPad = 128;
  if (n > 0) {
    const int64_t total = static_cast<int64_t>(n) * bytes;
    bin_expand_tc_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(src, n, bytes, row0, bank);
  }
  const int n_pad = (n + kBinRowPad - 1) / kBinRowPad * kBinRowPad;
  if (n_pad > 0) bin_ckey_kernel<<<(n_pad + 127) / 128, 128, 0, s>>>(n, n_pad, row0, ckey);
  return cudaGetLastError();
}

// dot_equal = 3 * cells = 12 * descriptor bytes: the dot product of a descriptor with itself.
cudaError_t launch_hamming2_tc(const CUtensorMap& tmap, const int32_t* ckey, const PairDesc* pairs,
                               const int2* items, int n_items, int dot_equal, Knn2* knn_out, int n_sms,
                               cudaStream_t stream) {
  const int grid = n_items < n_sms ? n_items : n_sms;
  if (grid <= 0) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(hamming2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kHtSmemBytes);
  if (e != cudaSuccess) return e;
  hamming2_tc_kernel<<<grid, kHtThreads, kHtSmemBytes, stream>>>(tmap, ckey, pairs, items, n_items, dot_equal,
                                                                 knn_out);
  return cudaGetLastError();
}

}  // namespace sfm
