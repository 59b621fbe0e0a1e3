// match_hamming.cu -- exact k=2 nearest neighbours under cv::NORM_HAMMING2 on sm_100a.
//
// The LIVE reference path: AKAZE (MLDB, 61-byte) descriptors matched with
// `BFMatcher matcher(NORM_HAMMING2); matcher.knnMatch(query, train, knn_matches, 2)`
// (OpenCV_SFM/NViewReconstuct.cpp:797, :875-877).  cv::normHamming(a, b, n, cellSize = 2)
// counts the 2-bit cells in which a and b differ: per byte popcount((x | x >> 1) & 0x55),
// x = a ^ b.  Results are ordered by (distance, lower train index first).
//
// This is integer / bit work for the CUDA cores (XOR, shift, LOP3, POPC), not a tensor-core
// shape: 16 words per descriptor pair, ~4.3 instructions per word.
//   * bank: every descriptor is one 64-byte row (bytes beyond the descriptor length are zero
//     and add nothing to a distance), images padded to kBinRowPad rows;
//   * hamming2_knn_kernel: a block = 128 query rows (one per thread, descriptor in registers)
//     x one split of the train image; train rows stream through shared memory in 64-row tiles
//     (cp.async double buffer) and are read as warp-wide broadcasts; a thread keeps its exact
//     top-2 as packed keys (distance << 20 | train index), so an integer min IS OpenCV's order;
//   * splitting the train image over blocks keeps all SMs busy for small query sets; the
//     partial top-2s of a row are merged by hamming2_merge_kernel, which writes the same
//     Knn2 rows as the SIFT kernel with SQUARED distances, so that the filter kernels
//     (sqrtf of an exact square) are shared unchanged.
#include <cuda_runtime.h>
#include <stdint.h>

#include "match_types.h"

namespace sfm {

constexpr int kBinWords = 16;        // 64-byte rows
constexpr int kBinQRows = 128;       // query rows per block (one per thread)
constexpr int kBinTile = 64;         // train rows per shared-memory tile
constexpr int kBinIdxBits = 20;      // packed key = distance << 20 | train index

// Copies n descriptors of `bytes` bytes each into 64-byte bank rows (tail bytes stay zero:
// the bank is memset before).
__global__ void bin_pack_kernel(const uint8_t* __restrict__ src, int n, int bytes, int row0,
                                uint8_t* __restrict__ bank) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(n) * bytes) return;
  const int r = static_cast<int>(i / bytes), b = static_cast<int>(i % bytes);
  bank[static_cast<size_t>(row0 + r) * 64 + b] = src[i];
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// NORM_HAMMING2 distance of two 64-byte rows.  Two words share one POPC: the cell flags of an
// even word sit on even bit positions, those of the following odd word on odd positions.
__device__ __forceinline__ int hamming2_row(const uint32_t (&q)[kBinWords], const uint4* t4) {
  int d = 0;
#pragma unroll
  for (int v = 0; v < kBinWords / 4; ++v) {
    const uint4 t = t4[v];
    const uint32_t x0 = q[4 * v + 0] ^ t.x, x1 = q[4 * v + 1] ^ t.y;
    const uint32_t x2 = q[4 * v + 2] ^ t.z, x3 = q[4 * v + 3] ^ t.w;
    const uint32_t c01 = ((x0 | (x0 >> 1)) & 0x55555555u) | ((x1 | (x1 << 1)) & 0xAAAAAAAAu);
    const uint32_t c23 = ((x2 | (x2 >> 1)) & 0x55555555u) | ((x3 | (x3 << 1)) & 0xAAAAAAAAu);
    d += __popc(c01) + __popc(c23);
  }
  return d;
}

// partial[(knn_row) * n_splits + split] = packed top-2 of that train split
__global__ void __launch_bounds__(kBinQRows)
hamming2_knn_kernel(const uint8_t* __restrict__ bank, const PairDesc* __restrict__ pairs,
                    const int2* __restrict__ items, int n_splits, int2* __restrict__ partial) {
  __shared__ __align__(16) uint4 s_t[2][kBinTile * (kBinWords / 4)];
  const int2 it = items[blockIdx.x];
  const PairDesc pd = pairs[it.x];
  const int split = blockIdx.y;
  const int row = it.y * kBinQRows + threadIdx.x;            // query row within the image
  // train rows of this split: whole tiles, the last split takes what is left
  const int tiles = (pd.nt + kBinTile - 1) / kBinTile;
  const int per = (tiles + n_splits - 1) / n_splits;
  const int t0 = min(split * per, tiles) * kBinTile;
  const int t1 = min(min((split + 1) * per, tiles) * kBinTile, pd.nt);

  uint32_t q[kBinWords];
  {
    const int qr = min(row, pd.nq - 1);                      // clamp: padding threads still help load
    const uint4* src = reinterpret_cast<const uint4*>(bank + static_cast<size_t>(pd.q_row0 + qr) * 64);
#pragma unroll
    for (int v = 0; v < kBinWords / 4; ++v) {
      const uint4 w = __ldg(src + v);
      q[4 * v + 0] = w.x; q[4 * v + 1] = w.y; q[4 * v + 2] = w.z; q[4 * v + 3] = w.w;
    }
  }
  int m1 = INT32_MAX, m2 = INT32_MAX;
  const uint8_t* tbase = bank + static_cast<size_t>(pd.t_row0) * 64;
  auto load_tile = [&](int buf, int r0) {
    // 64 rows x 64 B = 256 x 16 B: two 16-byte pieces per thread; rows past the image are
    // padding rows of the bank (readable) and are never scored
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(&s_t[buf][0]));
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int piece = threadIdx.x + k * kBinQRows;
      cp_async16(dst + piece * 16, tbase + static_cast<size_t>(r0) * 64 + piece * 16);
    }
    cp_async_commit();
  };
  int buf = 0;
  if (t0 < t1) load_tile(0, t0);
  for (int r0 = t0; r0 < t1; r0 += kBinTile) {
    if (r0 + kBinTile < t1) {
      load_tile(buf ^ 1, r0 + kBinTile);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int n = min(kBinTile, t1 - r0);
    const uint4* tile = &s_t[buf][0];
    if (n == kBinTile) {
#pragma unroll 4
      for (int r = 0; r < kBinTile; ++r) {
        const int key = (hamming2_row(q, tile + r * (kBinWords / 4)) << kBinIdxBits) + (r0 + r);
        const int t = max(m1, key);
        m1 = min(m1, key);
        m2 = min(m2, t);
      }
    } else {
      for (int r = 0; r < n; ++r) {
        const int key = (hamming2_row(q, tile + r * (kBinWords / 4)) << kBinIdxBits) + (r0 + r);
        const int t = max(m1, key);
        m1 = min(m1, key);
        m2 = min(m2, t);
      }
    }
    __syncthreads();
    buf ^= 1;
  }
  if (row < pd.nq)
    partial[(pd.knn_off + row) * n_splits + split] = make_int2(m1, m2);
}

// Merges the per-split top-2s of each query row; distances are stored SQUARED so that the
// shared filter kernels' sqrtf((float)d) returns the integer distance exactly.
__global__ void hamming2_merge_kernel(const int2* __restrict__ partial, int64_t n_rows, int n_splits,
                                      Knn2* __restrict__ knn) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_rows) return;
  int m1 = INT32_MAX, m2 = INT32_MAX;
  for (int s = 0; s < n_splits; ++s) {
    const int2 p = partial[i * n_splits + s];
    const int t = max(m1, p.x);
    m1 = min(m1, p.x);
    m2 = min(min(m2, t), p.y);
  }
  const int d0 = m1 >> kBinIdxBits, d1 = m2 >> kBinIdxBits;
  Knn2 out;
  out.j0 = m1 & ((1 << kBinIdxBits) - 1);
  out.j1 = m2 & ((1 << kBinIdxBits) - 1);
  out.d0 = d0 * d0;
  out.d1 = d1 * d1;
  *reinterpret_cast<int4*>(&knn[i]) = *reinterpret_cast<int4*>(&out);
}

// ------------------------------------------------------------------------------- launchers
cudaError_t launch_bin_pack(const uint8_t* src, int n, int bytes, int row0, uint8_t* bank,
                            cudaStream_t s) {
  const int64_t tot = static_cast<int64_t>(n) * bytes;
  if (tot > 0)
    bin_pack_kernel<<<static_cast<unsigned>((tot + 255) / 256), 256, 0, s>>>(src, n, bytes, row0, bank);
  return cudaGetLastError();
}

cudaError_t launch_hamming2_knn(const uint8_t* bank, const PairDesc* pairs, const int2* items,
                                int n_items, int n_splits, int2* partial, int64_t n_rows,
                                Knn2* knn, cudaStream_t s) {
  if (n_items > 0) {
    const dim3 grid(static_cast<unsigned>(n_items), static_cast<unsigned>(n_splits));
    hamming2_knn_kernel<<<grid, kBinQRows, 0, s>>>(bank, pairs, items, n_splits, partial);
  }
  if (n_rows > 0)
    hamming2_merge_kernel<<<static_cast<unsigned>((n_rows + 255) / 256), 256, 0, s>>>(partial, n_rows,
                                                                                    n_splits, knn);
  return cudaGetLastError();
}

}  // namespace sfm
