// match_knn.cu -- exact k=2 nearest-neighbour search over u8 SIFT descriptors on sm_100a.
//
// Replaces cv::BFMatcher(NORM_L2)::knnMatch(query, train, knn, 2) as called by the reference
// at OpenCV_SFM/NViewReconstuct.cpp:876-877 (SIFT/L2 form: TwoViewReconstruct.cpp:159-160).
//
// d2(i,j) = |q_i|^2 + |t_j|^2 - 2 q_i.t_j with q.t from tcgen05.mma.kind::i8 (u8 x u8 -> s32,
// exact).  One CTA per SM, persistent over work items (pair, 128-row query tile):
//   warp 0      TMA producer  : A tile once per item, B tiles (256 train rows) through a ring
//   warp 1      MMA issuer    : 4 x (128x256x32) MMAs per B tile into one of two TMEM buffers
//   warp 2      TMEM allocator
//   warps 4..   epilogue      : tcgen05.ld, packed (distance,index) keys, running top-2
// The distance matrix never leaves the SM.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "match_types.h"
#include "ptx.cuh"

namespace sfm {

constexpr int kStages = 5;                      // B-tile ring depth (32 KB each)
constexpr int kAccBufs = 2;                     // TMEM accumulator buffers (256 columns each)
constexpr int kCkSlots = 8;                     // ring of per-tile column keys (1 KB each)
constexpr int kFirstEpiWarp = 4;
constexpr int kEpiWarps = 4;
constexpr int kKnnThreads = (kFirstEpiWarp + kEpiWarps) * 32;

constexpr uint32_t kABytes = kTileM * kDim;     // 16 KB
constexpr uint32_t kBBytes = kTileN * kDim;     // 32 KB
constexpr uint32_t kCkBytes = kTileN * 4;       // 1 KB

// dynamic shared memory map (offsets from a 1024-byte aligned base)
constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffB = kOffA + 2 * kABytes;
constexpr uint32_t kOffCk = kOffB + kStages * kBBytes;
constexpr uint32_t kOffBar = kOffCk + kCkSlots * kCkBytes;
constexpr uint32_t kNumBars = 2 * kStages + 4 + 2 * kAccBufs;
constexpr uint32_t kOffTmemPtr = kOffBar + kNumBars * 8;
constexpr uint32_t kKnnSmemBytes = kOffTmemPtr + 16 + 1024;   // + alignment slack

__device__ __forceinline__ void top2_insert(int key, int& m1, int& m2) {
  m2 = min(m2, max(m1, key));
  m1 = min(m1, key);
}

__global__ void __launch_bounds__(kKnnThreads, 1)
knn2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const int32_t* __restrict__ ckey, const int32_t* __restrict__ norm,
            const PairDesc* __restrict__ pairs, const WorkItem* __restrict__ items, int n_items,
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
  auto bar_t_full = [&](int b) { return bar0 + 8u * (2 * kStages + 4 + b); };
  auto bar_t_empty = [&](int b) { return bar0 + 8u * (2 * kStages + 4 + kAccBufs + b); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + kOffTmemPtr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_a_full(b), 1);
      mbar_init(bar_a_empty(b), 1);
    }
    for (int b = 0; b < kAccBufs; ++b) {
      mbar_init(bar_t_full(b), 1);
      mbar_init(bar_t_empty(b), kEpiWarps);
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
        const WorkItem it = items[item];
        const PairDesc pd = pairs[it.pair];
        mbar_wait(bar_a_empty(abuf), aphase ^ 1);
        mbar_arrive_expect_tx(bar_a_full(abuf), kABytes);
        tma_load_2d(sA + abuf * kABytes, &tmap_a, bar_a_full(abuf), 0,
                    pd.q_row0 + it.mtile * kTileM);
        abuf ^= 1;
        if (abuf == 0) aphase ^= 1;
        const int ntiles = (pd.nt + kTileN - 1) / kTileN;
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(bar_empty(stage), phase ^ 1);
          mbar_arrive_expect_tx(bar_full(stage), kBBytes + kCkBytes);
          const int row = pd.t_row0 + t * kTileN;
          tma_load_2d(sB + stage * kBBytes, &tmap_b, bar_full(stage), 0, row);
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
      constexpr uint32_t idesc = make_idesc_u8(kTileM, kTileN);
      uint32_t stage = 0, phase = 0, abuf = 0, aphase = 0, buf = 0, bphase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const WorkItem it = items[item];
        const PairDesc pd = pairs[it.pair];
        const int ntiles = (pd.nt + kTileN - 1) / kTileN;
        mbar_wait(bar_a_full(abuf), aphase);
        const uint64_t a_desc = make_smem_desc_sw128(sA + abuf * kABytes);
        for (int t = 0; t < ntiles; ++t) {
          mbar_wait(bar_t_empty(buf), bphase ^ 1);
          mbar_wait(bar_full(stage), phase);
          tc_fence_after();
          const uint64_t b_desc = make_smem_desc_sw128(sB + stage * kBBytes);
#pragma unroll
          for (int k = 0; k < kDim / 32; ++k) {
            // advance 32 bytes along K inside the 128-byte swizzle atom: +2 in (addr >> 4)
            umma_i8(tmem_base + buf * kTileN, a_desc + 2 * k, b_desc + 2 * k, idesc, k > 0);
          }
          umma_commit(bar_empty(stage));
          umma_commit(bar_t_full(buf));
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
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t buf = 0, bphase = 0, tile_seq = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const WorkItem it = items[item];
      const PairDesc pd = pairs[it.pair];
      const int ntiles = (pd.nt + kTileN - 1) / kTileN;
      int g1v = INT32_MAX, g2v = INT32_MAX, g1i = -1, g2i = -1;
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(bar_t_full(buf), bphase);
        tc_fence_after();
        const int4* ck4 = reinterpret_cast<const int4*>(smem_gen + kOffCk +
                                                        (tile_seq % kCkSlots) * kCkBytes);
        int m1 = INT32_MAX, m2 = INT32_MAX;
#pragma unroll 1
        for (int c = 0; c < kTileN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_x32(tmem_base + lane_addr + buf * kTileN + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int4 cc = ck4[c * 8 + k];
            // key = ((|t|^2 - 2 q.t) << 8) | column : one IMAD, orders like (distance, index)
            top2_insert(static_cast<int>(r[4 * k + 0]) * -512 + cc.x, m1, m2);
            top2_insert(static_cast<int>(r[4 * k + 1]) * -512 + cc.y, m1, m2);
            top2_insert(static_cast<int>(r[4 * k + 2]) * -512 + cc.z, m1, m2);
            top2_insert(static_cast<int>(r[4 * k + 3]) * -512 + cc.w, m1, m2);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_t_empty(buf));
        // merge the tile's top-2 into the running (value, index) pairs; later tiles hold
        // larger indices, so strict '<' keeps the lower index on equal distance.
        const int base = t * kTileN;
        const int v1 = m1 >> 8, i1 = base + (m1 & 255);
        const int v2 = m2 >> 8, i2 = base + (m2 & 255);
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
      const int qrow = it.mtile * kTileM + row_in_tile;
      if (qrow < pd.nq) {
        const int nq2 = norm[pd.q_row0 + qrow];
        Knn2 out;
        out.j0 = g1i;
        out.j1 = g2i;
        out.d0 = g1v + nq2;
        out.d1 = g2v + nq2;
        *reinterpret_cast<int4*>(&knn_out[pd.knn_off + qrow]) = *reinterpret_cast<int4*>(&out);
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
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base, sB = smem_base + kABytes;
  const uint32_t bar = smem_base + kABytes + kBBytes;
  volatile uint32_t* tmem_ptr_smem =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kABytes + kBBytes + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < (kABytes + kBBytes) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_gen)[i] = 0x01010101u * (i & 3);
  if (warp == 0 && lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_base + kABytes + kBBytes + 16, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (warp == 0 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_u8(kTileM, kTileN);
    const uint64_t a_desc = make_smem_desc_sw128(sA), b_desc = make_smem_desc_sw128(sB);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_i8(tmem_base + (i & 1) * kTileN, a_desc + 2 * k, b_desc + 2 * k, idesc, k > 0);
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
// host-side launchers (called from context.cu)

cudaError_t launch_knn2(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const int32_t* ckey,
                        const int32_t* norm, const PairDesc* pairs, const WorkItem* items,
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
  knn2_kernel<<<grid, kKnnThreads, kKnnSmemBytes, stream>>>(tmap_a, tmap_b, ckey, norm, pairs,
                                                            items, n_items, knn_out);
  return cudaGetLastError();
}

cudaError_t launch_i8_peak(int iters, int n_sms, cudaStream_t stream) {
  const uint32_t smem = kABytes + kBBytes + 64 + 1024;
  cudaError_t e =
      cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  i8_peak_kernel<<<n_sms, 128, smem, stream>>>(iters);
  return cudaGetLastError();
}

}  // namespace sfm
