// C++ host program over include/sfm_b200.hpp: the reference's call sequence on a small bank.
//   pipeline_check <in.bin> <out.bin>
// in.bin  : int32 n_img, int32 rows[n_img], then per image rows*128 float32 descriptors and
//           rows*2 float32 keypoints, then double K[9], R1[9], T1[3], R2[9], T2[3] and the two
//           cameras as Ceres sees them: double ext[2][6] = {angle-axis, t} (:1478-1486)
// out.bin : per consecutive pair int64 n + n DMatch; int64 n_pts + n_pts Point3d (pair 0, host
//           arrays); the same from the device-resident path; int64 n_res + residual doubles + cost
// Built and run on the GPU box by tests/test_gpu_cpp_layer.py, which checks out.bin against the oracle.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sfm_b200.hpp"

namespace sb = sfm_b200;

template <class T>
static void rd(FILE* f, T* p, size_t n) {
  if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}
template <class T>
static void wr(FILE* f, const T* p, size_t n) {
  if (fwrite(p, sizeof(T), n, f) != n) { fprintf(stderr, "short write\n"); exit(2); }
}

int main(int argc, char** argv) {
  if (argc < 3) return 1;
  FILE* in = fopen(argv[1], "rb");
  if (!in) return 1;
  int32_t n_img = 0;
  rd(in, &n_img, 1);
  std::vector<int32_t> rows(n_img);
  rd(in, rows.data(), n_img);
  std::vector<std::vector<float>> desc(n_img);
  std::vector<std::vector<sb::Point2f>> kps(n_img);
  for (int i = 0; i < n_img; ++i) {
    desc[i].resize(static_cast<size_t>(rows[i]) * 128);
    rd(in, desc[i].data(), desc[i].size());
    kps[i].resize(rows[i]);
    rd(in, kps[i].data(), kps[i].size());
  }
  double K[9], R1[9], T1[3], R2[9], T2[3];
  rd(in, K, 9); rd(in, R1, 9); rd(in, T1, 3); rd(in, R2, 9); rd(in, T2, 3);
  std::vector<double> ext(12);
  rd(in, ext.data(), 12);
  fclose(in);

  try {
    sb::Context ctx(0);
    // match_features_for_all(descriptor_for_all, matches_for_all)
    std::vector<const void*> dptr;
    for (auto& d : desc) dptr.push_back(d.data());
    std::vector<std::vector<sb::DMatch>> matches_for_all;
    sb::match_features_for_all(ctx, dptr, rows, matches_for_all);
    FILE* out = fopen(argv[2], "wb");
    if (!out) return 1;
    for (auto& m : matches_for_all) {
      const int64_t n = static_cast<int64_t>(m.size());
      wr(out, &n, 1);
      wr(out, m.data(), m.size());
    }
    // get_matched_points + reconstruct on pair 0 (host arrays, as the reference passes them)
    sb::upload_keypoints(ctx, kps);
    std::vector<sb::Point2f> p1, p2;
    sb::get_matched_points(ctx, 0, matches_for_all[0].size(), nullptr, p1, p2);
    std::vector<sb::Point3d> structure, structure_dev;
    if (sb::reconstruct(ctx, K, R1, T1, R2, T2, p1, p2, structure) != 0) structure.clear();
    int64_t n = static_cast<int64_t>(structure.size());
    wr(out, &n, 1);
    wr(out, structure.data(), structure.size());
    // the same without the points leaving the device
    if (sb::reconstruct_pair(ctx, 0, matches_for_all[0].size(), K, R1, T1, R2, T2, nullptr, structure_dev) != 0)
      structure_dev.clear();
    n = static_cast<int64_t>(structure_dev.size());
    wr(out, &n, 1);
    wr(out, structure_dev.data(), structure_dev.size());
    // residual blocks of the two cameras over the new structure (camera-major order)
    const double intrinsic[4] = {K[0], K[4], K[2], K[5]};
    std::vector<int32_t> cam, pt;
    std::vector<float> obs;
    for (int c = 0; c < 2; ++c)
      for (size_t i = 0; i < structure.size(); ++i) {
        cam.push_back(c);
        pt.push_back(static_cast<int32_t>(i));
        const sb::Point2f& o = c == 0 ? p1[i] : p2[i];
        obs.push_back(o.x);
        obs.push_back(o.y);
      }
    std::vector<double> resid;
    double cost = 0.0;
    if (!structure.empty()) cost = sb::reproject_residuals(ctx, intrinsic, ext, structure, cam, pt, obs, resid);
    n = static_cast<int64_t>(resid.size());
    wr(out, &n, 1);
    wr(out, resid.data(), resid.size());
    wr(out, &cost, 1);
    fclose(out);
    // save_structure
    if (argc > 3) {
      std::vector<double> rot = {R1[0], R1[1], R1[2], R1[3], R1[4], R1[5], R1[6], R1[7], R1[8],
                                 R2[0], R2[1], R2[2], R2[3], R2[4], R2[5], R2[6], R2[7], R2[8]};
      std::vector<double> mot = {T1[0], T1[1], T1[2], T2[0], T2[1], T2[2]};
      std::vector<uint8_t> colors(structure.size() * 3, 128);
      sb::save_structure(argv[3], rot, mot, structure, colors);
    }
  } catch (const sb::Error& e) {
    fprintf(stderr, "sfm_b200 error %d: %s\n", e.code(), e.what());
    return 3;
  }
  printf("ok\n");
  return 0;
}
