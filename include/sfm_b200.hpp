// sfm_b200.hpp -- header-only C++11 layer over the C ABI (sfm_b200.h) with the reference's own
// function shapes, so that OpenCV_SFM/NViewReconstuct.cpp can call it with the containers it
// already holds:
//
//   match_features()            :873-913    match_features_for_all()   :850-871
//   get_matched_points()        :989-1003   reconstruct()              :1117-1159
//   ReprojectCost evaluation    :142-184    save_structure()           :186-227
//   residual-block enumeration  :1187-1211  (enumerate_observations, bundle_adjustment_residuals)
//
// No OpenCV types appear here.  The element types are template parameters that only have to be
// layout-compatible with the OpenCV value types the reference uses (checked with static_assert):
//   DMatchT  == cv::DMatch  {int queryIdx, trainIdx, imgIdx; float distance;}   16 bytes
//   Point2fT == cv::Point2f {float x, y;}                                        8 bytes
//   Point3dT == cv::Point3d {double x, y, z;}                                   24 bytes
// so std::vector<cv::DMatch>, std::vector<cv::Point2f>, std::vector<cv::Point3d> work as they are;
// the plain structs below serve callers without OpenCV (tests, examples/).
// Errors: the reference's functions return int / print; these throw sfm_b200::Error carrying the
// SFM_E_* code, except where the reference itself returns -1 (reconstruct on empty input).
#ifndef SFM_B200_HPP
#define SFM_B200_HPP

#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "sfm_b200.h"

namespace sfm_b200 {

struct DMatch { int32_t queryIdx, trainIdx, imgIdx; float distance; };
struct Point2f { float x, y; };
struct Point3d { double x, y, z; };

class Error : public std::runtime_error {
 public:
  Error(int code, const std::string& msg) : std::runtime_error(msg), code_(code) {}
  int code() const { return code_; }
 private:
  int code_;
};

// One sfm_ctx: one GPU, one host thread (the reference is single threaded).
class Context {
 public:
  explicit Context(int device = 0) {
    int err = 0;
    ctx_ = sfm_create(device, &err);
    if (!ctx_) throw Error(err, sfm_last_error(nullptr));   // no CPU fallback
  }
  ~Context() { sfm_destroy(ctx_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  sfm_ctx* get() const { return ctx_; }
  void check(int rc) const {
    if (rc != SFM_OK) throw Error(rc, sfm_last_error(ctx_));
  }
 private:
  sfm_ctx* ctx_;
};

// match_features_for_all (:850-871): descriptor_for_all[i] = rows[i] x 128 CV_32F (cv::SIFT) --
// or, with binary_bytes > 0, rows[i] x binary_bytes CV_8U matched with NORM_HAMMING2 as the live
// file does (:876).  pairs empty = the reference's consecutive schedule (i, i+1).
template <class DMatchT>
void match_features_for_all(const Context& c, const std::vector<const void*>& descriptor_for_all,
                            const std::vector<int32_t>& rows,
                            std::vector<std::vector<DMatchT>>& matches_for_all,
                            int binary_bytes = 0,
                            const std::vector<std::pair<int, int>>& pairs = {}) {
  static_assert(sizeof(DMatchT) == sizeof(sfm_match_t), "DMatchT must be layout-compatible with cv::DMatch");
  const int n = static_cast<int>(descriptor_for_all.size());
  if (binary_bytes > 0)
    c.check(sfm_upload_descriptors_bin(c.get(), n, reinterpret_cast<const uint8_t* const*>(descriptor_for_all.data()),
                                       rows.data(), binary_bytes));
  else
    c.check(sfm_upload_descriptors(c.get(), n, reinterpret_cast<const float* const*>(descriptor_for_all.data()),
                                   rows.data(), 128));
  std::vector<int32_t> pq, pt;
  if (pairs.empty())
    for (int i = 0; i + 1 < n; ++i) { pq.push_back(i); pt.push_back(i + 1); }
  else
    for (const auto& p : pairs) { pq.push_back(p.first); pt.push_back(p.second); }
  const int n_pairs = static_cast<int>(pq.size());
  std::vector<int64_t> off(n_pairs + 1, 0);
  const int rc = sfm_match_pairs(c.get(), pq.data(), pt.data(), n_pairs, 0.6 /*:884*/, 10.0f, 5.0f /*:901*/,
                                 nullptr, 0, off.data(), nullptr, nullptr);
  if (rc != SFM_OK && rc != SFM_E_CAPACITY) c.check(rc);
  std::vector<DMatchT> all(static_cast<size_t>(off[n_pairs]));
  if (!all.empty())
    c.check(sfm_fetch_matches(c.get(), reinterpret_cast<sfm_match_t*>(all.data()), static_cast<int64_t>(all.size())));
  matches_for_all.clear();
  for (int p = 0; p < n_pairs; ++p)
    matches_for_all.emplace_back(all.begin() + off[p], all.begin() + off[p + 1]);
}

// match_features(query, train, matches) (:873-913)
template <class DMatchT>
void match_features(const Context& c, const float* query, int nq, const float* train, int nt,
                    std::vector<DMatchT>& matches) {
  std::vector<std::vector<DMatchT>> all;
  match_features_for_all(c, std::vector<const void*>{query, train}, std::vector<int32_t>{nq, nt}, all);
  matches.swap(all[0]);
}

// Keypoint coordinates of every image (key_points_for_all[i][k].pt), once after feature extraction.
template <class Point2fT>
void upload_keypoints(const Context& c, const std::vector<std::vector<Point2fT>>& pts_for_all) {
  static_assert(sizeof(Point2fT) == 8, "Point2fT must be layout-compatible with cv::Point2f");
  std::vector<const float*> p;
  std::vector<int32_t> n;
  for (const auto& v : pts_for_all) {
    p.push_back(reinterpret_cast<const float*>(v.data()));
    n.push_back(static_cast<int32_t>(v.size()));
  }
  c.check(sfm_upload_keypoints(c.get(), static_cast<int>(p.size()), p.data(), n.data()));
}

// get_matched_points (:989-1003) [+ maskout_points (:943) when mask != nullptr] for pair `pair` of
// the last match_features_for_all call; n_matches = matches_for_all[pair].size().
template <class Point2fT>
void get_matched_points(const Context& c, int pair, size_t n_matches, const uint8_t* mask,
                        std::vector<Point2fT>& out_p1, std::vector<Point2fT>& out_p2) {
  static_assert(sizeof(Point2fT) == 8, "Point2fT must be layout-compatible with cv::Point2f");
  out_p1.resize(n_matches);
  out_p2.resize(n_matches);
  int64_t n = 0;
  c.check(sfm_get_matched_points(c.get(), pair, mask, reinterpret_cast<float*>(out_p1.data()),
                                 reinterpret_cast<float*>(out_p2.data()), static_cast<int64_t>(n_matches), &n));
  out_p1.resize(static_cast<size_t>(n));
  out_p2.resize(static_cast<size_t>(n));
}

// reconstruct(K, R1, T1, R2, T2, p1, p2, structure) (:1117-1159); K, R row-major 3x3, T 3-vectors
// (CV_64F in the reference).  Returns -1 on empty input like the reference, 0 otherwise.
template <class Point2fT, class Point3dT>
int reconstruct(const Context& c, const double K[9], const double R1[9], const double T1[3],
                const double R2[9], const double T2[3], const std::vector<Point2fT>& p1,
                const std::vector<Point2fT>& p2, std::vector<Point3dT>& structure) {
  static_assert(sizeof(Point2fT) == 8 && sizeof(Point3dT) == 24, "cv::Point2f / cv::Point3d layouts");
  if (p1.empty() || p2.empty() || p1.size() != p2.size()) return -1;          // :1122-1126
  float P[24];
  const double* Rs[2] = {R1, R2};
  const double* Ts[2] = {T1, T2};
  for (int v = 0; v < 2; ++v) {          // proj = fK * [R|T] as cv::gemm does on CV_32F (:1129-1143)
    float fK[9], RT[12];
    for (int i = 0; i < 9; ++i) fK[i] = static_cast<float>(K[i]);
    for (int r = 0; r < 3; ++r) {
      for (int k = 0; k < 3; ++k) RT[4 * r + k] = static_cast<float>(Rs[v][3 * r + k]);
      RT[4 * r + 3] = static_cast<float>(Ts[v][r]);
    }
    for (int r = 0; r < 3; ++r)
      for (int k = 0; k < 4; ++k) {
        volatile float a = fK[3 * r] * RT[k], b = fK[3 * r + 1] * RT[4 + k], d = fK[3 * r + 2] * RT[8 + k];
        volatile float s = a + b;
        P[12 * v + 4 * r + k] = s + d;
      }
  }
  const size_t n = p1.size();
  std::vector<float> xy(4 * n);                                               // view-major [2][n][2]
  for (size_t i = 0; i < n; ++i) {
    const float* a = reinterpret_cast<const float*>(&p1[i]);
    const float* b = reinterpret_cast<const float*>(&p2[i]);
    xy[2 * i] = a[0]; xy[2 * i + 1] = a[1];
    xy[2 * n + 2 * i] = b[0]; xy[2 * n + 2 * i + 1] = b[1];
  }
  structure.resize(n);
  c.check(sfm_triangulate_batch(c.get(), P, xy.data(), 2, static_cast<int64_t>(n), nullptr,
                                reinterpret_cast<double*>(structure.data())));
  return 0;
}

// reconstruct() on the device-resident match list of pair `pair` (mask as maskout_points).
template <class Point3dT>
int reconstruct_pair(const Context& c, int pair, size_t n_matches, const double K[9], const double R1[9],
                     const double T1[3], const double R2[9], const double T2[3], const uint8_t* mask,
                     std::vector<Point3dT>& structure) {
  static_assert(sizeof(Point3dT) == 24, "Point3dT must be layout-compatible with cv::Point3d");
  structure.resize(n_matches);
  int64_t n = 0;
  const int rc = sfm_reconstruct_pair(c.get(), pair, K, R1, T1, R2, T2, mask,
                                      reinterpret_cast<double*>(structure.data()),
                                      static_cast<int64_t>(n_matches), &n);
  if (rc == SFM_E_INVALID && n == 0) { structure.clear(); return -1; }        // "[Err]: empty 2d points."
  c.check(rc);
  structure.resize(static_cast<size_t>(n));
  return 0;
}

// All residual blocks of bundle_adjustment() (:1187-1211) in the caller's order; returns the
// HuberLoss(delta) cost 0.5 * sum rho(|r|^2) (:1184).
template <class Point3dT>
double reproject_residuals(const Context& c, const double intrinsic[4], const std::vector<double>& extrinsics6,
                           const std::vector<Point3dT>& pts3d, const std::vector<int32_t>& cam_idx,
                           const std::vector<int32_t>& pt_idx, const std::vector<float>& obs_xy,
                           std::vector<double>& residuals, double huber_delta = 4.0) {
  static_assert(sizeof(Point3dT) == 24, "Point3dT must be layout-compatible with cv::Point3d");
  residuals.resize(2 * cam_idx.size());
  double cost = 0.0;
  c.check(sfm_reproject_residuals(c.get(), intrinsic, extrinsics6.data(), static_cast<int>(extrinsics6.size() / 6),
                                  reinterpret_cast<const double*>(pts3d.data()), static_cast<int64_t>(pts3d.size()),
                                  cam_idx.data(), pt_idx.data(), obs_xy.data(), static_cast<int64_t>(cam_idx.size()),
                                  huber_delta, residuals.data(), &cost));
  return cost;
}

// The residual-block order of bundle_adjustment() (:1187-1211): for every image, for every keypoint in
// order, one observation (camera = image, point = correspond_struct_idx[img][kp], kp.pt) where the
// index is >= 0.  Host glue, no device work; fills the tables reproject_residuals() takes.
template <class Point2fT>
void enumerate_observations(const std::vector<std::vector<int>>& correspond_struct_idx,
                            const std::vector<std::vector<Point2fT>>& key_points_for_all,
                            std::vector<int32_t>& cam_idx, std::vector<int32_t>& pt_idx, std::vector<float>& obs_xy) {
  static_assert(sizeof(Point2fT) == 8, "Point2fT must be layout-compatible with cv::Point2f");
  if (correspond_struct_idx.size() != key_points_for_all.size())
    throw Error(SFM_E_INVALID, "enumerate_observations: one index vector per image");
  cam_idx.clear();
  pt_idx.clear();
  obs_xy.clear();
  for (size_t img = 0; img < correspond_struct_idx.size(); ++img) {
    const std::vector<int>& ids = correspond_struct_idx[img];
    if (ids.size() != key_points_for_all[img].size())
      throw Error(SFM_E_INVALID, "enumerate_observations: one structure index per keypoint");
    for (size_t kp = 0; kp < ids.size(); ++kp) {
      if (ids[kp] < 0) continue;
      const float* xy = reinterpret_cast<const float*>(&key_points_for_all[img][kp]);
      cam_idx.push_back(static_cast<int32_t>(img));
      pt_idx.push_back(ids[kp]);
      obs_xy.push_back(xy[0]);
      obs_xy.push_back(xy[1]);
    }
  }
}

// Every ReprojectCost block that bundle_adjustment() (:1162-1244) hands to Ceres, in its order:
// residuals [n_obs x 2], returns the HuberLoss(4) cost (:1184); *rmse = sqrt(cost / n_obs) as printed
// at :1237-1238.
template <class Point2fT, class Point3dT>
double bundle_adjustment_residuals(const Context& c, const double intrinsic[4], const std::vector<double>& extrinsics6,
                                   const std::vector<std::vector<int>>& correspond_struct_idx,
                                   const std::vector<std::vector<Point2fT>>& key_points_for_all,
                                   const std::vector<Point3dT>& structure, std::vector<double>& residuals,
                                   double* rmse = nullptr, double huber_delta = 4.0) {
  std::vector<int32_t> cam_idx, pt_idx;
  std::vector<float> obs_xy;
  enumerate_observations(correspond_struct_idx, key_points_for_all, cam_idx, pt_idx, obs_xy);
  const double cost = reproject_residuals(c, intrinsic, extrinsics6, structure, cam_idx, pt_idx, obs_xy, residuals,
                                          huber_delta);
  if (rmse) *rmse = cam_idx.empty() ? 0.0 : std::sqrt(cost / static_cast<double>(cam_idx.size()));
  return cost;
}

// save_structure(file_name, rotations, motions, structure, colors) (:186-227); rotations / motions
// flattened row-major (9 / 3 doubles per camera), colors b,g,r bytes per point.
template <class Point3dT>
void save_structure(const std::string& file_name, const std::vector<double>& rotations9,
                    const std::vector<double>& motions3, const std::vector<Point3dT>& structure,
                    const std::vector<uint8_t>& colors_bgr) {
  static_assert(sizeof(Point3dT) == 24, "Point3dT must be layout-compatible with cv::Point3d");
  const int rc = sfm_save_structure(file_name.c_str(), static_cast<int>(rotations9.size() / 9), rotations9.data(),
                                    motions3.data(), static_cast<int64_t>(structure.size()),
                                    reinterpret_cast<const double*>(structure.data()),
                                    static_cast<int64_t>(colors_bgr.size() / 3), colors_bgr.data());
  if (rc != SFM_OK) throw Error(rc, "cannot write " + file_name);
}

}  // namespace sfm_b200
#endif  // SFM_B200_HPP
