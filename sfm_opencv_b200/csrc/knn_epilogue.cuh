// knn_epilogue.cuh -- the exact running top-2 of one query row over packed keys, shared by the
// SIFT kernel (match_knn.cu) and the tensor-core HAMMING2 kernel (match_hamming_tc.cu).
// A packed key orders like OpenCV's (distance, lower train index first); see match_types.h.
#pragma once
#include <stdint.h>

#include "match_types.h"
#include "ptx.cuh"

namespace sfm {

// (a1 <= a2), (b1 <= b2) -> the two smallest of the four, sorted.
__device__ __forceinline__ void merge_top2(int& a1, int& a2, int b1, int b2) {
  const int t = max(a1, b1);
  a1 = min(a1, b1);
  a2 = __vimin3_s32(t, a2, b2);
}

// Packed key of accumulator r (= q.t) and column key ck = (|t|^2 << kColBits) | (train row & 1023):
// ((|t|^2 - 2 q.t) << kColBits) | column, one IMAD; orders like (distance, lower column first).
__device__ __forceinline__ int make_key(uint32_t r, int ck) {
  return static_cast<int>(r) * -(2 << kColBits) + ck;   // wraps, true value fits
}

// Exact top-2 of 8 keys merged into (m1, m2).  The new minimum is a 3-input-min tree over
// the nine candidates (4 ALU ops).  The new second is the minimum over the ten values with one
// instance of that minimum knocked out: x -> x - min - 1 as UNSIGNED sends the minimum to
// 2^32 - 1 and keeps the order of everything else (keys of a window are unique), so it is
// another 3-input unsigned-min tree (5 ops) after ten subtractions, which ptxas places on
// the otherwise idle FMA pipe (IMAD.IADD).  11 ALU ops instead of the 20 of a sorting network.
__device__ __forceinline__ void insert8(const int* k, int& m1, int& m2) {
  const int a = __vimin3_s32(k[0], k[1], k[2]);
  const int b = __vimin3_s32(k[3], k[4], k[5]);
  const int c = __vimin3_s32(k[6], k[7], m1);
  const int lo = __vimin3_s32(a, b, c);
  const uint32_t nlo = ~static_cast<uint32_t>(lo);          // x + ~lo == x - lo - 1 (mod 2^32)
  uint32_t d[10];
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] = static_cast<uint32_t>(k[i]) + nlo;
  d[8] = static_cast<uint32_t>(m1) + nlo;
  d[9] = static_cast<uint32_t>(m2) + nlo;
  const uint32_t e0 = __vimin3_u32(d[0], d[1], d[2]);
  const uint32_t e1 = __vimin3_u32(d[3], d[4], d[5]);
  const uint32_t e2 = __vimin3_u32(d[6], d[7], d[8]);
  const uint32_t e3 = __vimin3_u32(e0, e1, e2);
  const uint32_t sec = min(e3, d[9]);
  m1 = lo;
  m2 = static_cast<int>(sec - nlo);                          // sec + lo + 1
}

// Running state of one epilogue thread (one query row, half of the columns).
struct RowTop2 {
  int g1v, g1i, g2v, g2i;   // best / second best of the finished windows: value = |t|^2 - 2 q.t
  int m1, m2;               // top-2 of the current 1024-column window as packed keys
  int bv;                   // bound: a group matters iff its smallest possible value
                            // min|t|^2 - 2 max(q.t) is below bv = min(value of m2, second best
                            // of the finished windows); equal values of later columns lose on
                            // the index, so the comparison is strict
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

// exact keys of the 8 columns of a group (column keys from the shared-memory ring) -> (m1, m2)
__device__ __forceinline__ void group_insert(const uint32_t* a, uint32_t ck_addr, RowTop2& s) {
  const int4 c0 = lds_v4(ck_addr), c1 = lds_v4(ck_addr + 16);
  int k[8];
  k[0] = make_key(a[0], c0.x); k[1] = make_key(a[1], c0.y);
  k[2] = make_key(a[2], c0.z); k[3] = make_key(a[3], c0.w);
  k[4] = make_key(a[4], c1.x); k[5] = make_key(a[5], c1.y);
  k[6] = make_key(a[6], c1.z); k[7] = make_key(a[7], c1.w);
  insert8(k, s.m1, s.m2);
}

// largest raw dot product of a group of 8 accumulators: 3-input-max tree, 0.5 op per element
__device__ __forceinline__ int group_max(const uint32_t* r) {
  const int a = __vimax3_s32(r[0], r[1], r[2]);
  const int b = __vimax3_s32(r[3], r[4], r[5]);
  return max(__vimax3_s32(a, b, r[6]), static_cast<int>(r[7]));
}

// Close a packed-key window that started at train column `base`: merge its top-2 into the
// (value, index) pairs and restart the window.  Windows are closed in ascending column order, so
// every index of the window is larger than every index already in (g1, g2): on equal values the
// older entry wins and the lexicographic order reduces to a strict compare of the values.  Merge
// of two sorted pairs: 3 compares + 6 selects instead of two general insertions.
__device__ __forceinline__ void close_window(RowTop2& s, int base) {
  constexpr int kMask = (1 << kColBits) - 1;
#ifdef SFM_GENERAL_WINDOW_CLOSE
  insert_vi(s, s.m1 >> kColBits, base + (s.m1 & kMask));
  insert_vi(s, s.m2 >> kColBits, base + (s.m2 & kMask));
#else
  const int v1 = s.m1 >> kColBits, i1 = base + (s.m1 & kMask);
  const int v2 = s.m2 >> kColBits, i2 = base + (s.m2 & kMask);   // (v1, i1) <= (v2, i2)
  const bool c1 = v1 < s.g1v, c2 = v2 < s.g1v, c3 = v1 < s.g2v;
  // second of the merge: new best took the lead -> old best against the window's second;
  // otherwise the window's best against the old second
  const int sv = c1 ? (c2 ? v2 : s.g1v) : (c3 ? v1 : s.g2v);
  const int si = c1 ? (c2 ? i2 : s.g1i) : (c3 ? i1 : s.g2i);
  s.g1v = c1 ? v1 : s.g1v;
  s.g1i = c1 ? i1 : s.g1i;
  s.g2v = sv;
  s.g2i = si;
#endif
  s.m1 = INT32_MAX;
  s.m2 = INT32_MAX;
}

}  // namespace sfm
