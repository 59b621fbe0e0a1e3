// match_types.h -- device-side tables shared by the matching kernels and the host context.
#pragma once
#include <stdint.h>

namespace sfm {

constexpr int kDim = 128;          // SIFT descriptor length in bytes (u8)
constexpr int kTileM = 256;        // query rows per work item (two 128-lane TMEM halves)
constexpr int kTileN = 128;        // train rows per B tile (TMEM columns per half)
constexpr int kColBits = 10;       // packed key = (|t|^2 - 2 q.t) << 10 | (train row & 1023): a
                                   // key window is 8 train tiles; |value| < 2^21 keeps it in int32
constexpr int kRowPad = 256;       // every image is padded to a multiple of this many rows
constexpr int kNormPad = 0x1FFFFF; // norm^2 sentinel of padding rows: >= every real value |t|^2 - 2 q.t
                                   // (real norms <= 2^21 - 1) and behind them in index order, so a
                                   // padding row is never selected when nt >= 2; << kColBits fits

// One image pair of sfm_match_pairs.
struct PairDesc {
  int32_t q_row0;    // bank row of the first query row (padded bank coordinates)
  int32_t t_row0;    // first bank row of the train image
  int32_t nq;        // query descriptors
  int32_t nt;        // train descriptors
  int32_t q_first;   // index of the first query row within its image (> 0: a query-row shard)
  int32_t pad;
  int64_t knn_off;   // first row of this pair in the kNN result array
};

// Raw kNN result per query row: exact integer squared distances.
struct Knn2 {
  int32_t j0, j1;    // train indices (within the train image)
  int32_t d0, d1;    // squared L2 distances, exact
};

}  // namespace sfm
