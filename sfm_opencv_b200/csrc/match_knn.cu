// match_knn.cu -- exact k=2 nearest-neighbour search over u8 SIFT descriptors on sm_100a.
//
// Replaces cv::BFMatcher(NORM_L2)::knnMatch(query, train, knn, 2) as called by the reference
// at OpenCV_SFM/NViewReconstuct.cpp:876-877 (SIFT/L2 form: TwoViewReconstruct.cpp:159-160).
//
// d2(i,j) = |q_i|^2 + |t_j|^2 - 2 q_i.t_j with q.t from tcgen05.mma.kind::i8 (u8 x u8 -> s32,
// exact).  One persistent CTA per SM; a work item is a 256-row query block of one image pair,
// swept over all 128-row train tiles of the pair:
//   warp 0       TMA producer : the 256-row A block once per item (2 x 16 KB), B tiles
//                               (128 train rows, 16 KB) + their column keys through a ring
//   warp 1       MMA issuer   : per B tile 2 x 4 MMAs (128x128x32): query half h -> TMEM
//                               accumulator [buf][h]; every B byte feeds 256 query rows
//   warp 2       TMEM allocator (512 columns = 2 buffers x 2 halves x 128)
//   warps 4..11  epilogue     : one thread per query row (TMEM lane); tcgen05.ld, packed
//                               (distance,index) keys, exact running top-2
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
constexpr int kCkSlots = 16;                    // ring of per-tile column keys (512 B each)
constexpr int kFirstEpiWarp = 4;
constexpr int kEpiWarps = 8;                    // 2 halves x 4 TMEM lane quarters
constexpr int kKnnThreads = (kFirstEpiWarp + kEpiWarps) * 32;
constexpr int kHalfM = kTileM / 2;              // 128 rows per MMA

constexpr uint32_t kABytes = kTileM * kDim;     // 32 KB
constexpr uint32_t kAHalfBytes = kHalfM * kDim; // 16 KB
constexpr uint32_t kBBytes = kTileN * kDim;     // 16 KB
constexpr uint32_t kCkBytes = kTileN * 4;       // 512 B

// What the producer tells the MMA and epilogue warps about an item.
struct ItemInfo {
  int32_t ntiles;       // train tiles of the pair
  int32_t rows_valid;   // query rows of this block that exist (<= 256)
  int32_t norm_row;     // bank row of the block's first query row
  int32_t pad;
  int64_t knn_row;      // first output row of the block
};

// dynamic shared memory map (offsets from a 1024-byte aligned base)
constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffB = kOffA + 2 * kABytes;
constexpr uint32_t kOffCk = kOffB + kStages * kBBytes;
constexpr uint32_t kOffInfo = kOffCk + kCkSlots * kCkBytes;
constexpr uint32_t kOffBar = kOffInfo + 2 * sizeof(ItemInfo);
constexpr uint32_t kNumBars = 2 * kStages + 4 + 4 * kAccBufs;
constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr uint32_t kKnnSmemBytes = kOffTmemPtr + 16 + 1024;   // + alignment slack

// (a1 <= a2), (b1 <= b2) -> the two smallest of the four, sorted.
__device__ __forceinline__ void merge_top2(int& a1, int& a2, int b1, int b2) {
  const int t = max(a1, b1);
  a1 = min(a1, b1);
  a2 = __vimin3_s32(t, a2, b2);
}

// Exact top-2 of one 32-column chunk: keys = ((|t|^2 - 2 q.t) << 7) | column, one IMAD each;
// a key orders like (distance, lower column first).  Pair-sort + merge tree: 2.5 min/max per
// element, all independent until the last levels.
__device__ __forceinline__ void chunk_top2(const uint32_t (&r)[32], const int4* __restrict__ ck4,
                                           int& m1, int& m2) {
  int lo[16], hi[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int4 cc = ck4[k];
    const int k0 = static_cast<int>(r[4 * k + 0]) * -256 + cc.x;
    const int k1 = static_cast<int>(r[4 * k + 1]) * -256 + cc.y;
    const int k2 = static_cast<int>(r[4 * k + 2]) * -256 + cc.z;
    const int k3 = static_cast<int>(r[4 * k + 3]) * -256 + cc.w;
    lo[2 * k] = min(k0, k1);
    hi[2 * k] = max(k0, k1);
    lo[2 * k + 1] = min(k2, k3);
    hi[2 * k + 1] = max(k2, k3);
  }
#pragma unroll
  for (int n = 8; n >= 1; n >>= 1) {
#pragma unroll
    for (int j = 0; j < n; ++j) merge_top2(lo[j], hi[j], lo[j + n], hi[j + n]);
  }
  merge_top2(m1, m2, lo[0], hi[0]);
}

__global__ void __launch_bounds__(kKnnThreads, 1)
knn2_kernel(const __grid_constant__ CUtensorMap tmap, const int32_t* __restrict__ ckey,
            const int32_t* __restrict__ norm, const PairDesc* __restrict__ pairs,
            const int32_t* __restrict__ item_prefix, int n_pairs, int n_items,
            Knn2* __restrict__ knn_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t sA = smem_base + kOffA;
  const uint32_t sB = smem_base + kOffB;
  const uint32_t sCk = smem_base + kOffCk;
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

  if (warp == 0 && lane == 0) prefetch_tensormap(&tmap);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_a_full(b), 1);
      mbar_init(bar_a_empty(b), 1 + kEpiWarps);   // MMA commit + every epilogue warp
    }
    for (int b = 0; b < kAccBufs; ++b)
      for (int h = 0; h < 2; ++h) {
        mbar_init(bar_t_full(b, h), 1);
        mbar_init(bar_t_empty(b, h), kEpiWarps / 2);
      }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_base + kOffTmemPtr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, abuf = 0, aphase = 0, tile_seq = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        // pair of this item: last p with item_prefix[p] <= item
        int lo = 0, hi = n_pairs;
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (__ldg(item_prefix + mid) <= item) lo = mid; else hi = mid;
        }
        const PairDesc pd = pairs[lo];
        const int mblk = item - __ldg(item_prefix + lo);
        const int ntiles = (pd.nt + kTileN - 1) / kTileN;
        mbar_wait(bar_a_empty(abuf), aphase ^ 1);
        info[abuf].ntiles = ntiles;
        info[abuf].rows_valid = pd.nq - mblk * kTileM;
        info[abuf].norm_row = pd.q_row0 + mblk * kTileM;
        info[abuf].knn_row = pd.knn_off + static_cast<int64_t>(mblk) * kTileM;
        mbar_arrive_expect_tx(bar_a_full(abuf), kABytes);
        tma_load_2d(sA + abuf * kABytes, &tmap, bar_a_full(abuf), 0, pd.q_row0 + mblk * kTileM);
        tma_load_2d(sA + abuf * kABytes + kAHalfBytes, &tmap, bar_a_full(abuf), 0,
                    pd.q_row0 + mblk * kTileM + kHalfM);
        abuf ^= 1;
        if (abuf == 0) aphase ^= 1;
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(bar_empty(stage), phase ^ 1);
          mbar_arrive_expect_tx(bar_full(stage), kBBytes + kCkBytes);
          const int row = pd.t_row0 + t * kTileN;
          tma_load_2d(sB + stage * kBBytes, &tmap, bar_full(stage), 0, row);
          bulk_load_1d(sCk + (tile_seq % kCkSlots) * kCkBytes, ckey + row, kCkBytes,
                       bar_full(stage));
          ++tile_seq;
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_u8(kHalfM, kTileN);
      uint32_t stage = 0, phase = 0, abuf = 0, aphase = 0, buf = 0, bphase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        mbar_wait(bar_a_full(abuf), aphase);
        const int ntiles = info[abuf].ntiles;
        const uint64_t a_desc0 = make_smem_desc_sw128(sA + abuf * kABytes);
        const uint64_t a_desc1 = make_smem_desc_sw128(sA + abuf * kABytes + kAHalfBytes);
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(bar_full(stage), phase);
          const uint64_t b_desc = make_smem_desc_sw128(sB + stage * kBBytes);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(bar_t_empty(buf, h), bphase ^ 1);
            tc_fence_after();
            const uint32_t d = tmem_base + buf * (2 * kTileN) + h * kTileN;
            const uint64_t a_desc = h ? a_desc1 : a_desc0;
#pragma unroll
            for (int k = 0; k < kDim / 32; ++k) {
              // advance 32 bytes along K inside the 128-byte swizzle atom: +2 in (addr >> 4)
              umma_i8(d, a_desc + 2 * k, b_desc + 2 * k, idesc, k > 0);
            }
            umma_commit(bar_t_full(buf, h));
          }
          umma_commit(bar_empty(stage));
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
          if (++buf == kAccBufs) {
            buf = 0;
            bphase ^= 1;
          }
        }
        umma_commit(bar_a_empty(abuf));
        abuf ^= 1;
        if (abuf == 0) aphase ^= 1;
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ===================================================== epilogue: running top-2 per row
    const int half = (warp - kFirstEpiWarp) >> 2;  // which 128-row half of the block
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int row_in_blk = half * kHalfM + quarter * 32 + lane;
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + half * kTileN;
    uint32_t buf = 0, bphase = 0, tile_seq = 0, abuf = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int g1v = INT32_MAX, g2v = INT32_MAX, g1i = -1, g2i = -1;
      int ntiles = 1, rows_valid = 0, norm_row = 0;
      int64_t knn_row = 0;
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(bar_t_full(buf, half), bphase);
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
        const int4* ck4 = reinterpret_cast<const int4*>(smem_gen + kOffCk +
                                                        (tile_seq % kCkSlots) * kCkBytes);
        int m1 = INT32_MAX, m2 = INT32_MAX;
        const uint32_t ta = t_addr + buf * (2 * kTileN);
        {
          uint32_t r0[32], r1[32];
          tmem_ld_x32(ta, r0);
          tmem_ld_x32(ta + 32, r1);
          tmem_ld_wait();
          chunk_top2(r0, ck4, m1, m2);
          tmem_ld_x32(ta + 64, r0);
          chunk_top2(r1, ck4 + 8, m1, m2);
          tmem_ld_wait();
          tmem_ld_x32(ta + 96, r1);
          chunk_top2(r0, ck4 + 16, m1, m2);
          tmem_ld_wait();
          chunk_top2(r1, ck4 + 24, m1, m2);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_t_empty(buf, half));
        // merge the tile's top-2 into the running (value, index) pairs; later tiles hold
        // larger indices, so strict '<' keeps the lower index on equal distance.
        const int base = t * kTileN;
        const int v1 = m1 >> 7, i1 = base + (m1 & 127);
        const int v2 = m2 >> 7, i2 = base + (m2 & 127);
        if (v1 < g1v) {
          g2v = g1v; g2i = g1i;
          g1v = v1;  g1i = i1;
        } else if (v1 < g2v) {
          g2v = v1;  g2i = i1;
        }
        if (v2 < g2v) {
          g2v = v2;  g2i = i2;
        }
        ++tile_seq;
        if (++buf == kAccBufs) {
          buf = 0;
          bphase ^= 1;
        }
      }
      if (row_in_blk < rows_valid) {
        const int nq2 = __ldg(norm + norm_row + row_in_blk);
        Knn2 out;
        out.j0 = g1i;
        out.j1 = g2i;
        out.d0 = g1v + nq2;
        out.d1 = g2v + nq2;
        *reinterpret_cast<int4*>(&knn_out[knn_row + row_in_blk]) = *reinterpret_cast<int4*>(&out);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// -------------------------------------------------------------------------------------
// Bare tensor-pipe probe: back-to-back 128x256x32 u8 MMAs on every SM, operands resident
// in shared memory (contents irrelevant), no epilogue.  Gives the measured int8 peak.
constexpr uint32_t kProbeA = 128 * kDim, kProbeB = 256 * kDim;
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters) {
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
  if (warp == 0 && lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
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
  if (warp == 0 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_u8(128, 256);
    const uint64_t a_desc = make_smem_desc_sw128(sA), b_desc = make_smem_desc_sw128(sB);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_i8(tmem_base + (i & 1) * 256, a_desc + 2 * k, b_desc + 2 * k, idesc, k > 0);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
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

cudaError_t launch_knn2(const CUtensorMap& tmap, const int32_t* ckey, const int32_t* norm,
                        const PairDesc* pairs, const int32_t* item_prefix, int n_pairs,
                        int n_items, Knn2* knn_out, int n_sms, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(knn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kKnnSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  const int grid = n_items < n_sms ? n_items : n_sms;
  if (grid <= 0) return cudaSuccess;
  knn2_kernel<<<grid, kKnnThreads, kKnnSmemBytes, stream>>>(tmap, ckey, norm, pairs, item_prefix,
                                                            n_pairs, n_items, knn_out);
  return cudaGetLastError();
}

cudaError_t launch_i8_peak(int iters, int n_sms, cudaStream_t stream) {
  const uint32_t smem = kProbeA + kProbeB + 64 + 1024;
  cudaError_t e =
      cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  i8_peak_kernel<<<n_sms, 128, smem, stream>>>(iters);
  return cudaGetLastError();
}

}  // namespace sfm
