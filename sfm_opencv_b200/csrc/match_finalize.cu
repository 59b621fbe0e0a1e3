// match_finalize.cu -- descriptor packing and the reference's two filter passes on the device.
//
//  pack_*            : what sfm_upload_descriptors does with descriptor_for_all[i]
//                      (CV_32F SIFT rows -> validated u8 rows + |row|^2 + packed column keys)
//  filter_count/write: match_features passes 1 and 2, OpenCV_SFM/NViewReconstuct.cpp:880-908,
//                      one block per image pair, kept matches in ascending queryIdx.
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "../../include/sfm_b200.h"
#include "match_types.h"

namespace sfm {

// flags accumulated by the pack kernels
constexpr uint32_t kFlagNotIntegral = 1u;
constexpr uint32_t kFlagRange = 2u;
constexpr uint32_t kFlagNorm = 4u;
// float sqrt is injective on integers below 2^22: require |q|^2 + |t|^2 < 2^22
constexpr int kMaxNorm = (1 << 21) - 1;

// One launch per image: a block of 4 warps owns 8 consecutive rows of the image's PADDED range
// (two rows per warp), i.e. exactly one 8-row group of the kNN filter.  src = float rows (kF32)
// or u8 rows of ONE image; dst rows are in padded bank coordinates starting at row0 (a multiple
// of 256).  Rows [n, n_pad) are padding: zero descriptor (the bank is memset), sentinel norm.
// The block also writes the group's minimum norm (gmin8), so an upload costs one kernel per image
// (it was three: pack, pad, group minima -- 600 launches in front of a 200-image step).
// kStore = false: the u8 rows are already in the bank (sfm_bank_commit's single-image form).
template <bool kF32, bool kStore>
__global__ void __launch_bounds__(128)
pack_rows_kernel(const void* __restrict__ src, int n, int n_pad, int row0,
                 uint8_t* __restrict__ desc, int32_t* __restrict__ norm,
                 int32_t* __restrict__ ckey, int32_t* __restrict__ gmin8,
                 uint32_t* __restrict__ flags) {
  __shared__ int32_t s_norm[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t bad = 0;
#pragma unroll
  for (int k2 = 0; k2 < 2; ++k2) {
    const int r = blockIdx.x * 8 + warp * 2 + k2;
    if (r >= n_pad) continue;                      // n_pad is a multiple of 8: whole blocks only
    int s = kNormPad;
    if (r < n) {
      uint32_t packed;
      if constexpr (kF32) {
        const float4 v = reinterpret_cast<const float4*>(src)[static_cast<size_t>(r) * 32 + lane];
        const float f[4] = {v.x, v.y, v.z, v.w};
        packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float x = f[k];
          if (!(x >= 0.0f && x <= 255.0f)) bad |= kFlagRange;        // also catches NaN
          else if (x != rintf(x)) bad |= kFlagNotIntegral;
          const uint32_t b = static_cast<uint32_t>(fminf(fmaxf(x, 0.0f), 255.0f));
          packed |= b << (8 * k);
        }
      } else {
        packed = reinterpret_cast<const uint32_t*>(src)[static_cast<size_t>(r) * 32 + lane];
      }
      if constexpr (kStore)
        reinterpret_cast<uint32_t*>(desc)[static_cast<size_t>(row0 + r) * 32 + lane] = packed;
      s = __dp4a(packed, packed, 0u);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (s > kMaxNorm) bad |= kFlagNorm;
    }
    if (lane == 0) {
      norm[row0 + r] = s;
      ckey[row0 + r] = (s << kColBits) | (r & ((1 << kColBits) - 1));
      s_norm[warp * 2 + k2] = s;
    }
  }
  bad = __reduce_or_sync(0xffffffffu, bad);
  if (bad && lane == 0) atomicOr(flags, bad);
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x * 8 < n_pad) {
    int m = s_norm[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) m = min(m, s_norm[k]);
    gmin8[row0 / 8 + blockIdx.x] = m;
  }
}

// Rows that arrived from another GPU (sfm_bank_commit): norms, keys and sentinels of a whole range
// of IMAGES in one launch -- one warp per bank row, the row's image found by binary search in the
// row-offset table (175 images of an 8-GPU all-gather would otherwise cost 525 tiny launches).
__global__ void commit_rows_kernel(const uint8_t* __restrict__ desc, const int32_t* __restrict__ img_row0,
                                   const int32_t* __restrict__ img_n, int first_img, int n_img,
                                   int row_begin, int row_end, int32_t* __restrict__ norm,
                                   int32_t* __restrict__ ckey, uint32_t* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int row = row_begin + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= row_end) return;
  int lo = first_img, hi = first_img + n_img - 1;          // last image whose first row is <= row
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (img_row0[mid] <= row) lo = mid; else hi = mid - 1;
  }
  const int r = row - img_row0[lo];
  int s = kNormPad;
  if (r < img_n[lo]) {
    const uint32_t packed = reinterpret_cast<const uint32_t*>(desc)[static_cast<size_t>(row) * 32 + lane];
    s = __dp4a(packed, packed, 0u);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (s > kMaxNorm && lane == 0) atomicOr(flags, kFlagNorm);
  }
  if (lane == 0) {
    norm[row] = s;
    ckey[row] = (s << kColBits) | (r & ((1 << kColBits) - 1));
  }
}

__global__ void group_min_range_kernel(const int32_t* __restrict__ norm, int row_begin, int row_end,
                                       int32_t* __restrict__ gmin8) {
  const int g = row_begin / 8 + blockIdx.x * blockDim.x + threadIdx.x;
  if (g * 8 >= row_end) return;
  const int4 a = *reinterpret_cast<const int4*>(norm + g * 8);
  const int4 b = *reinterpret_cast<const int4*>(norm + g * 8 + 4);
  gmin8[g] = min(min(min(a.x, a.y), min(a.z, a.w)), min(min(b.x, b.y), min(b.z, b.w)));
}

cudaError_t launch_commit_rows(const uint8_t* desc, const int32_t* img_row0, const int32_t* img_n,
                               int first_img, int n_img, int row_begin, int row_end, int32_t* norm,
                               int32_t* ckey, int32_t* gmin8, uint32_t* flags, cudaStream_t s) {
  const int rows = row_end - row_begin;
  if (rows <= 0) return cudaSuccess;
  commit_rows_kernel<<<(rows + 3) / 4, 128, 0, s>>>(desc, img_row0, img_n, first_img, n_img, row_begin,
                                                    row_end, norm, ckey, flags);
  group_min_range_kernel<<<(rows / 8 + 127) / 128, 128, 0, s>>>(norm, row_begin, row_end, gmin8);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------- recheck
// Rows the ratio-driven sweep of match_knn.cu (kPrune) could not decide: exact k = 2 search of the
// query row against its whole train image on the CUDA cores (a few rows in 10^4 on unrelated images,
// 0.3-3.5 % on the bundled datasets, where about half of the true matches have a fail range that
// reaches into skipped columns).
//   1. the list is counting-sorted by pair (recheck_hist / scan_counts / recheck_scatter), so that
//      rows which read the same train image sit next to each other;
//   2. recheck_rows_kernel: a block of 8 warps takes 16 consecutive rows; every run of rows of one
//      pair shares the train image through shared memory: 128-row tiles, loaded once per run with
//      coalesced 16-byte loads into rows of 144 bytes (conflict-free for the 16-byte reads below).
//      A warp owns two query rows (2 x 32 registers per lane); a lane takes every 32nd row of the tile,
//      u8 dot products by dp4a, per-lane exact top-2 in (value, index) order, shuffle merge, and the
//      row of the kNN table is overwritten.  Per (query, train) pair of rows this reads 64 bytes of
//      shared memory instead of 8 separate 128-byte lines of L1 (one warp per row on its own, as
//      first built, cost +4 ms on dataset/dog: 6 200 flagged rows x 18 000 train rows).
__device__ __forceinline__ void top2_push(int v, int i, int& v1, int& i1, int& v2, int& i2) {
  const bool b1 = (v < v1) | ((v == v1) & (i < i1));
  const bool b2 = (v < v2) | ((v == v2) & (i < i2));
  v2 = b1 ? v1 : (b2 ? v : v2);
  i2 = b1 ? i1 : (b2 ? i : i2);
  v1 = b1 ? v : v1;
  i1 = b1 ? i : i1;
}

__global__ void __launch_bounds__(256)
recheck_hist_kernel(const int2* __restrict__ rows, const int32_t* __restrict__ count, int cap,
                    int32_t* __restrict__ hist) {
  const int n = min(*count, cap);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x)
    atomicAdd(hist + rows[e].x, 1);
}

__global__ void __launch_bounds__(256)
recheck_scatter_kernel(const int2* __restrict__ rows, const int32_t* __restrict__ count, int cap,
                       const int64_t* __restrict__ offs, int32_t* __restrict__ cursor,
                       int2* __restrict__ sorted) {
  const int n = min(*count, cap);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int2 r = rows[e];
    sorted[offs[r.x] + atomicAdd(cursor + r.x, 1)] = r;
  }
}

// 16- / 4-byte asynchronous copies global -> shared; `ok` false writes zeros without reading
__device__ __forceinline__ void cp_async_16(void* dst, const void* src, bool ok) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))),
               "l"(src), "r"(ok ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_4(void* dst, const void* src, bool ok) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))),
               "l"(src), "r"(ok ? 4 : 0) : "memory");
}

constexpr int kRcWarps = 8;             // warps per block
constexpr int kRcPerWarp = 2;           // query rows per warp: every 16-byte read of a train row feeds two dot products
constexpr int kRcRows = kRcWarps * kRcPerWarp;   // rows per block
constexpr int kRcTile = 128;            // train rows per shared-memory tile
constexpr int kRcRow16 = kDim / 16 + 1; // 16-byte words per staged row: 128 bytes + 16 of padding

__global__ void __launch_bounds__(kRcWarps * 32)
recheck_rows_kernel(const uint8_t* __restrict__ desc, const int32_t* __restrict__ norm,
                    const PairDesc* __restrict__ pairs, const int2* __restrict__ rows,
                    const int32_t* __restrict__ count, int cap, Knn2* __restrict__ knn) {
  __shared__ uint4 s_tile[2][kRcTile * kRcRow16];   // double-buffered: tile t+1 lands (cp.async) under tile t
  __shared__ int32_t s_norm[2][kRcTile];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = min(*count, cap);
  // few rows: one per warp, so that more blocks (and SMs) share them; many: two per warp
  const int rpb = n <= static_cast<int>(gridDim.x) * kRcWarps ? kRcWarps : kRcRows;
  for (int chunk = blockIdx.x * rpb; chunk < n; chunk += gridDim.x * rpb) {
    const int end = min(chunk + rpb, n);
    for (int start = chunk, run; start < end; start += run) {      // block-uniform
      const int pair = rows[start].x;
      run = 1;
      while (start + run < end && rows[start + run].x == pair) ++run;
      const PairDesc pd = pairs[pair];
      // the run's rows go to the warps round-robin: row u of the run belongs to warp u % kRcWarps
      bool mine[kRcPerWarp];
      int qrow[kRcPerWarp];
      uint32_t q[kRcPerWarp][32];
      int v1[kRcPerWarp], i1[kRcPerWarp], v2[kRcPerWarp], i2[kRcPerWarp];
#pragma unroll
      for (int u = 0; u < kRcPerWarp; ++u) {
        mine[u] = warp + u * kRcWarps < run;
        qrow[u] = mine[u] ? rows[start + warp + u * kRcWarps].y : 0;
        const uint4* q4 = reinterpret_cast<const uint4*>(desc + static_cast<size_t>(pd.q_row0 + qrow[u]) * kDim);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint4 v = __ldg(q4 + k);
          q[u][4 * k] = v.x; q[u][4 * k + 1] = v.y; q[u][4 * k + 2] = v.z; q[u][4 * k + 3] = v.w;
        }
        v1[u] = i1[u] = v2[u] = i2[u] = INT32_MAX;
      }
      auto stage = [&](int t0, int b) {                              // tile [t0, t0 + 128) -> buffer b, zero-filled past nt
        for (int i = threadIdx.x; i < kRcTile * 8; i += kRcWarps * 32) {
          const int r = i >> 3, k = i & 7;
          const bool ok = t0 + r < pd.nt;
          cp_async_16(&s_tile[b][r * kRcRow16 + k],
                      desc + static_cast<size_t>(pd.t_row0 + (ok ? t0 + r : 0)) * kDim + 16 * k, ok);
        }
        if (threadIdx.x < kRcTile) {
          const bool ok = t0 + static_cast<int>(threadIdx.x) < pd.nt;
          cp_async_4(&s_norm[b][threadIdx.x], norm + pd.t_row0 + (ok ? t0 + threadIdx.x : 0), ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      __syncthreads();                                               // the previous run has left both buffers
      stage(0, 0);
      for (int t0 = 0, b = 0; t0 < pd.nt; t0 += kRcTile, b ^= 1) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                             // tile b is visible; buffer b ^ 1 has been read
        if (t0 + kRcTile < pd.nt) stage(t0 + kRcTile, b ^ 1);
        if (mine[0]) {                                               // warp-uniform; mine[1] implies mine[0]
#pragma unroll
          for (int jj = 0; jj < kRcTile / 32; ++jj) {
            const int r = lane + 32 * jj, j = t0 + r;
            uint32_t dot[kRcPerWarp] = {};
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint4 t = s_tile[b][r * kRcRow16 + k];
#pragma unroll
              for (int u = 0; u < kRcPerWarp; ++u) {
                dot[u] = __dp4a(q[u][4 * k], t.x, dot[u]);
                dot[u] = __dp4a(q[u][4 * k + 1], t.y, dot[u]);
                dot[u] = __dp4a(q[u][4 * k + 2], t.z, dot[u]);
                dot[u] = __dp4a(q[u][4 * k + 3], t.w, dot[u]);
              }
            }
            if (j < pd.nt) {
              const int tn = s_norm[b][r];
#pragma unroll
              for (int u = 0; u < kRcPerWarp; ++u)
                top2_push(tn - 2 * static_cast<int>(dot[u]), j, v1[u], i1[u], v2[u], i2[u]);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kRcPerWarp; ++u) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const int a = __shfl_xor_sync(0xffffffffu, v1[u], o), ai = __shfl_xor_sync(0xffffffffu, i1[u], o);
          const int b = __shfl_xor_sync(0xffffffffu, v2[u], o), bi = __shfl_xor_sync(0xffffffffu, i2[u], o);
          top2_push(a, ai, v1[u], i1[u], v2[u], i2[u]);
          top2_push(b, bi, v1[u], i1[u], v2[u], i2[u]);
        }
        if (mine[u] && lane == 0) {
          const int nq2 = norm[pd.q_row0 + qrow[u]];
          Knn2 out;
          out.j0 = i1[u];
          out.j1 = i2[u];
          out.d0 = v1[u] + nq2;
          out.d1 = v2[u] + nq2;
          *reinterpret_cast<int4*>(&knn[pd.knn_off + qrow[u]]) = *reinterpret_cast<int4*>(&out);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------- filter
__device__ __forceinline__ bool ratio_fails(float d0, float d1, double ratio) {
  // `knn[i][0].distance > 0.6 * knn[i][1].distance`: float promoted to double (:884, :900)
  return static_cast<double>(d0) > ratio * static_cast<double>(d1);
}

// Pass 1 + count of pass 2. One block per pair.
__global__ void __launch_bounds__(256)
filter_count_kernel(const Knn2* __restrict__ knn, const PairDesc* __restrict__ pairs, double ratio,
                    float dist_floor, float gate_mult, const float* __restrict__ min_dist_in,
                    float* __restrict__ min_dist, int32_t* __restrict__ counts) {
  __shared__ uint32_t s_min;
  __shared__ int s_cnt;
  const PairDesc pd = pairs[blockIdx.x];
  const Knn2* k = knn + pd.knn_off;
  if (threadIdx.x == 0) {
    s_min = __float_as_uint(FLT_MAX);
    s_cnt = 0;
  }
  __syncthreads();
  // min_dist_in: pass 1 was done elsewhere (a pair sharded by query rows over several GPUs:
  // min_dist couples all query rows of the pair, so it is the minimum over the shards)
  if (min_dist_in == nullptr) {
    uint32_t lmin = __float_as_uint(FLT_MAX);
    for (int i = threadIdx.x; i < pd.nq; i += blockDim.x) {
      const int4 v = *reinterpret_cast<const int4*>(&k[i]);
      const float d0 = __fsqrt_rn(static_cast<float>(v.z));
      const float d1 = __fsqrt_rn(static_cast<float>(v.w));
      if (!ratio_fails(d0, d1, ratio)) lmin = min(lmin, __float_as_uint(d0));  // d0 >= 0
    }
    lmin = __reduce_min_sync(0xffffffffu, lmin);
    if ((threadIdx.x & 31) == 0) atomicMin(&s_min, lmin);
    __syncthreads();
  }
  const float md = min_dist_in ? min_dist_in[blockIdx.x] : __uint_as_float(s_min);
  const float gate = gate_mult * fmaxf(md, dist_floor);   // 5 * max(min_dist, 10.0f), float
  int c = 0;
  for (int i = threadIdx.x; i < pd.nq; i += blockDim.x) {
    const int4 v = *reinterpret_cast<const int4*>(&k[i]);
    const float d0 = __fsqrt_rn(static_cast<float>(v.z));
    const float d1 = __fsqrt_rn(static_cast<float>(v.w));
    c += !(ratio_fails(d0, d1, ratio) || d0 > gate);
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) {
    min_dist[blockIdx.x] = md;
    counts[blockIdx.x] = s_cnt;
  }
}

// Exclusive scan of counts -> offsets[n+1] (int64). Single block; n is at most ~1e6.
__global__ void __launch_bounds__(1024)
scan_counts_kernel(const int32_t* __restrict__ counts, int n, int64_t* __restrict__ offsets) {
  __shared__ int64_t s_warp[32];
  __shared__ int64_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int64_t v = i < n ? counts[i] : 0;
    int64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int64_t w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int64_t incl = x + (warp ? s_warp[warp - 1] : 0) + s_carry;
    if (i < n) offsets[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = s_carry;
}

// Pass 2: ordered compaction of the kept matches of each pair.
__global__ void __launch_bounds__(256)
filter_write_kernel(const Knn2* __restrict__ knn, const PairDesc* __restrict__ pairs, double ratio,
                    float dist_floor, float gate_mult, const float* __restrict__ min_dist,
                    const int64_t* __restrict__ offsets, sfm_match_t* __restrict__ out,
                    int64_t out_cap) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const PairDesc pd = pairs[blockIdx.x];
  const Knn2* k = knn + pd.knn_off;
  const int64_t off = offsets[blockIdx.x];
  const float gate = gate_mult * fmaxf(min_dist[blockIdx.x], dist_floor);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int base = 0; base < pd.nq; base += blockDim.x) {
    const int i = base + threadIdx.x;
    bool keep = false;
    int4 v = make_int4(0, 0, 0, 0);
    float d0 = 0.f;
    if (i < pd.nq) {
      v = *reinterpret_cast<const int4*>(&k[i]);
      d0 = __fsqrt_rn(static_cast<float>(v.z));
      const float d1 = __fsqrt_rn(static_cast<float>(v.w));
      keep = !(ratio_fails(d0, d1, ratio) || d0 > gate);
    }
    const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(ballot);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const int rank = before + __popc(ballot & ((1u << lane) - 1u));
    if (keep && off + rank < out_cap) {
      sfm_match_t m;
      m.queryIdx = i + pd.q_first;       // index within the query IMAGE (row shards: q_first > 0)
      m.trainIdx = v.x;
      m.imgIdx = 0;
      m.distance = d0;
      *reinterpret_cast<int4*>(&out[off + rank]) = *reinterpret_cast<int4*>(&m);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += s_warp[w];
      s_base += tot;
    }
    __syncthreads();
  }
}

// Knn2 (integer squared distances) -> sfm_knn2_t (float distances, as DMatch::distance).
__global__ void knn_to_float_kernel(const Knn2* __restrict__ knn, int64_t n,
                                    sfm_knn2_t* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int4 v = *reinterpret_cast<const int4*>(&knn[i]);
  int4 o;
  o.x = v.x;
  o.y = v.y;
  o.z = __float_as_int(__fsqrt_rn(static_cast<float>(v.z)));
  o.w = __float_as_int(__fsqrt_rn(static_cast<float>(v.w)));
  *reinterpret_cast<int4*>(&out[i]) = o;
}

// Work-item table of the kNN kernels: pair p (visited in the host's processing order) owns the
// items [first, first + ceil(nq / qblock)), item k = (p, k - first).  One warp per pair.
__global__ void build_items_kernel(const int2* __restrict__ ordoff, int n_pairs,
                                   const PairDesc* __restrict__ pairs, int qblock,
                                   int2* __restrict__ items) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_pairs) return;
  const int2 po = ordoff[w];
  const int mb = (pairs[po.x].nq + qblock - 1) / qblock;
  for (int m = lane; m < mb; m += 32) items[po.y + m] = make_int2(po.x, m);
}

// ------------------------------------------------------------------------------- launchers
cudaError_t launch_build_items(const int2* ordoff, int n_pairs, const PairDesc* pairs, int qblock,
                               int2* items, cudaStream_t s) {
  if (n_pairs > 0)
    build_items_kernel<<<(n_pairs * 32 + 127) / 128, 128, 0, s>>>(ordoff, n_pairs, pairs, qblock, items);
  return cudaGetLastError();
}

// src == nullptr: derive norms / keys from the u8 rows already in the bank
cudaError_t launch_pack_rows(bool f32, const void* src, int n, int row0, uint8_t* desc,
                             int32_t* norm, int32_t* ckey, int32_t* gmin8, uint32_t* flags,
                             cudaStream_t s) {
  // 128-thread CTAs (<= 24 registers per thread): they fit next to a resident kNN CTA (768 threads,
  // 61440 of the SM's 65536 registers), so an asynchronous upload keeps packing while the
  // matching kernel owns every SM.  One launch per image: rows, padding rows and group minima.
  const int n_pad = (n + kRowPad - 1) / kRowPad * kRowPad;
  if (n_pad > 0) {
    const int grid = n_pad / 8;
    if (src == nullptr)
      pack_rows_kernel<false, false><<<grid, 128, 0, s>>>(
          desc + static_cast<size_t>(row0) * kDim, n, n_pad, row0, nullptr, norm, ckey, gmin8, flags);
    else if (f32)
      pack_rows_kernel<true, true><<<grid, 128, 0, s>>>(src, n, n_pad, row0, desc, norm, ckey, gmin8, flags);
    else
      pack_rows_kernel<false, true><<<grid, 128, 0, s>>>(src, n, n_pad, row0, desc, norm, ckey, gmin8, flags);
  }
  return cudaGetLastError();
}

// min_dist_in == nullptr: pass 1 + count of pass 2; otherwise the count under the given min_dist
cudaError_t launch_filter(const Knn2* knn, const PairDesc* pairs, int n_pairs, double ratio,
                          float dist_floor, float gate_mult, const float* min_dist_in,
                          float* min_dist, int32_t* counts, int64_t* offsets, cudaStream_t s) {
  if (n_pairs > 0)
    filter_count_kernel<<<n_pairs, 256, 0, s>>>(knn, pairs, ratio, dist_floor, gate_mult,
                                                min_dist_in, min_dist, counts);
  scan_counts_kernel<<<1, 1024, 0, s>>>(counts, n_pairs, offsets);
  return cudaGetLastError();
}

cudaError_t launch_filter_write(const Knn2* knn, const PairDesc* pairs, int n_pairs, double ratio,
                                float dist_floor, float gate_mult, const float* min_dist,
                                const int64_t* offsets, sfm_match_t* out, int64_t out_cap,
                                cudaStream_t s) {
  if (n_pairs > 0)
    filter_write_kernel<<<n_pairs, 256, 0, s>>>(knn, pairs, ratio, dist_floor, gate_mult, min_dist,
                                                offsets, out, out_cap);
  return cudaGetLastError();
}

// work: 2 * n_pairs int32 (histogram, cursors; zeroed here), offs: n_pairs + 1 int64, sorted: cap int2
cudaError_t launch_recheck_rows(const uint8_t* desc, const int32_t* norm, const PairDesc* pairs, int n_pairs,
                                const int2* rows, const int32_t* count, int cap, int32_t* work, int64_t* offs,
                                int2* sorted, Knn2* knn, int n_sms, cudaStream_t s) {
  if (n_pairs <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(work, 0, sizeof(int32_t) * 2 * static_cast<size_t>(n_pairs), s);
  if (e != cudaSuccess) return e;
  recheck_hist_kernel<<<n_sms, 256, 0, s>>>(rows, count, cap, work);
  scan_counts_kernel<<<1, 1024, 0, s>>>(work, n_pairs, offs);
  recheck_scatter_kernel<<<n_sms, 256, 0, s>>>(rows, count, cap, offs, work + n_pairs, sorted);
  recheck_rows_kernel<<<n_sms * 4, kRcWarps * 32, 0, s>>>(desc, norm, pairs, sorted, count, cap, knn);
  return cudaGetLastError();
}

cudaError_t launch_knn_to_float(const Knn2* knn, int64_t n, sfm_knn2_t* out, cudaStream_t s) {
  if (n > 0)
    knn_to_float_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(knn, n, out);
  return cudaGetLastError();
}

}  // namespace sfm
