// Host-only check of sfm_b200::enumerate_observations (the residual-block order of
// NViewReconstuct.cpp:1187-1211): reads "n_img, then per image n_kp and n_kp lines `idx x y`" from stdin,
// prints one line `cam pt x y` per observation.  tests/test_api_host.py compares it with the Python mirror.
#include <cstdio>
#include <vector>

#include "sfm_b200.hpp"

// compiled (not run here: it needs a GPU) so that the template stays in step with the C ABI
namespace sfm_b200 {
template double bundle_adjustment_residuals<Point2f, Point3d>(
    const Context&, const double[4], const std::vector<double>&, const std::vector<std::vector<int>>&,
    const std::vector<std::vector<Point2f>>&, const std::vector<Point3d>&, std::vector<double>&, double*, double);
}

int main() {
  int n_img = 0;
  if (std::scanf("%d", &n_img) != 1) return 2;
  std::vector<std::vector<int>> idx(n_img);
  std::vector<std::vector<sfm_b200::Point2f>> kps(n_img);
  for (int i = 0; i < n_img; ++i) {
    int n = 0;
    if (std::scanf("%d", &n) != 1) return 2;
    idx[i].resize(n);
    kps[i].resize(n);
    for (int k = 0; k < n; ++k)
      if (std::scanf("%d %f %f", &idx[i][k], &kps[i][k].x, &kps[i][k].y) != 3) return 2;
  }
  std::vector<int32_t> cam, pt;
  std::vector<float> xy;
  try {
    sfm_b200::enumerate_observations(idx, kps, cam, pt, xy);
  } catch (const sfm_b200::Error& e) {
    std::printf("error %d\n", e.code());
    return 1;
  }
  for (size_t o = 0; o < cam.size(); ++o) std::printf("%d %d %.9g %.9g\n", cam[o], pt[o], xy[2 * o], xy[2 * o + 1]);
  return 0;
}
