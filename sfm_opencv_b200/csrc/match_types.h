// match_types.h -- device-side tables shared by the matching kernels and the host context.
#pragma once
#include <stdint.h>

namespace sfm {

constexpr int kDim = 128;          // SIFT descriptor length in bytes (u8)
constexpr int kTileM = 128;        // query rows per work item (TMEM lanes)
constexpr int kTileN = 256;        // train rows per MMA tile (TMEM columns)
constexpr int kRowPad = 256;       // every image is padded to a multiple of this many rows
constexpr int kNormPad = 0x7FFFFF; // norm^2 sentinel of padding rows (never selected)

// One image pair of sfm_match_pairs.
struct PairDesc {
  int32_t q_row0;    // first bank row of the query image (padded bank coordinates)
  int32_t t_row0;    // first bank row of the train image
  int32_t nq;        // query descriptors
  int32_t nt;        // train descriptors
  int64_t knn_off;   // first row of this pair in the kNN result array
};

// One work item = one 128-row query tile of one pair, swept over all train tiles.
struct WorkItem {
  int32_t pair;
  int32_t mtile;
};

// Raw kNN result per query row: exact integer squared distances.
struct Knn2 {
  int32_t j0, j1;    // train indices (within the train image)
  int32_t d0, d1;    // squared L2 distances, exact
};

}  // namespace sfm
