// match_hamming.cu -- exact k=2 nearest neighbours under cv::NORM_HAMMING2 on sm_100a.
//
// The LIVE reference path: AKAZE (MLDB, 61-byte) descriptors matched with
// `BFMatcher matcher(NORM_HAMMING2); matcher.knnMatch(query, train, knn_matches, 2)`
// (OpenCV_SFM/NViewReconstuct.cpp:797, :875-877).  cv::normHamming(a, b, n, cellSize = 2)
// counts the 2-bit cells in which a and b differ: per byte popcount((x | x >> 1) & 0x55),
// x = a ^ b.  Results are ordered by (distance, lower train index first).
//
// This is integer / bit work for the CUDA cores (LOP3, POPC), not a tensor-core shape.
//   * bank: every descriptor (<= 64 bytes, zero padded: zero bytes add nothing to a distance)
//     is stored EXPANDED to 128 bytes so that the inner loop needs no shifts and no masks.
//     With m = 0x55555555 and x = a ^ b, the cell flags of a word are
//         (x | x >> 1) & m  =  ((a & m) ^ (b & m)) | (((a >> 1) & m) ^ ((b >> 1) & m)),
//     i.e. two 3-input LOP3 on pre-masked operands.  Odd words use the mirrored form on the
//     odd bit positions, (a & ~m, (a << 1) & ~m), so that the flags of an (even, odd) word pair
//     interleave in one register and share one POPC: 4 LOP3 + 1 POPC per two words, 2.75
//     instructions per word instead of 4.3.  Row layout: word w of the descriptor becomes the
//     pair (E[2w], E[2w+1]) = (a & m, (a >> 1) & m) for even w, (a & ~m, (a << 1) & ~m) for odd w;
//     images are padded to kBinRowPad rows;
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

constexpr int kBinWords = 32;        // 128-byte expanded rows (16 descriptor words x 2)
constexpr int kBinQRows = 128;       // query rows per block (one per thread)
constexpr int kBinTile = 64;         // train rows per shared-memory tile
constexpr int kBinIdxBits = 20;      // packed key = distance << 20 | train index

// Expands n descriptors of `bytes` bytes each into 128-byte bank rows (see the header comment);
// one thread per descriptor word, bytes beyond the descriptor read as zero.
__global__ void bin_pack_kernel(const uint8_t* __restrict__ src, int n, int bytes, int row0,
                                uint8_t* __restrict__ bank) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<int64_t>(n) * 16) return;
  const int r = static_cast<int>(i >> 4), w = static_cast<int>(i & 15);
  uint32_t a = 0;
#pragma unroll
  for (int b = 0; b < 4; ++b)
    if (4 * w + b < bytes) a |= static_cast<uint32_t>(src[static_cast<size_t>(r) * bytes + 4 * w + b]) << (8 * b);
  constexpr uint32_t m = 0x55555555u;
  uint2 e;
  if ((w & 1) == 0) e = make_uint2(a & m, (a >> 1) & m);
  else e = make_uint2(a & ~m, (a << 1) & ~m);
  reinterpret_cast<uint2*>(bank + static_cast<size_t>(row0 + r) * 128)[w] = e;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// NORM_HAMMING2 distance of two expanded rows: per descriptor word pair 4 LOP3 + 1 POPC.
__device__ __forceinline__ int hamming2_row(const uint32_t (&q)[kBinWords], const uint4* t4) {
  int d = 0;
#pragma unroll
  for (int v = 0; v < kBinWords / 4; ++v) {
    // t = (even word: b & m, (b >> 1) & m | odd word: b & ~m, (b << 1) & ~m)
    const uint4 t = t4[v];
    uint32_t c = q[4 * v + 0] ^ t.x;               // flags of the even word, even positions
    c |= q[4 * v + 1] ^ t.y;
    c |= q[4 * v + 2] ^ t.z;                       // flags of the odd word, odd positions
    c |= q[4 * v + 3] ^ t.w;
    d += __popc(c);
  }
  return d;
}

// partial[(knn_row) * n_splits + split] = packed top-2 of that train split
__global__ void __launch_bounds__(kBinQRows)
hamming2_knn_kernel(const uint8_t* __restrict__ bank, const PairDesc* __restrict__ pairs,
                    const int2* __restrict__ items, int n_splits, int2* __restrict__ partial) {
  __shared__ __align__(16) uint4 s_t[2][kBinTile * (kBinWords / 4)];   // 2 x 8 KB tiles
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
    const uint4* src = reinterpret_cast<const uint4*>(bank + static_cast<size_t>(pd.q_row0 + qr) * 128);
#pragma unroll
    for (int v = 0; v < kBinWords / 4; ++v) {
      const uint4 w = __ldg(src + v);
      q[4 * v + 0] = w.x; q[4 * v + 1] = w.y; q[4 * v + 2] = w.z; q[4 * v + 3] = w.w;
    }
  }
  int m1 = INT32_MAX, m2 = INT32_MAX;
  const uint8_t* tbase = bank + static_cast<size_t>(pd.t_row0) * 128;
  auto load_tile = [&](int buf, int r0) {
    // 64 rows x 128 B = 512 x 16 B: four 16-byte pieces per thread; rows past the image are
    // padding rows of the bank (readable) and are never scored
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(&s_t[buf][0]));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int piece = threadIdx.x + k * kBinQRows;
      cp_async16(dst + piece * 16, tbase + static_cast<size_t>(r0) * 128 + piece * 16);
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
  const int64_t tot = static_cast<int64_t>(n) * 16;
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
