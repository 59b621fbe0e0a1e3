/* Plain C: include/sfm_b200.h must be a valid C header and every entry point must link.
 * Built and run by tests/test_capi_symbols.py (no GPU needed: sfm_create fails loudly, the
 * host-only writers run). */
#include <stdio.h>
#include <string.h>

#include "sfm_b200.h"

int main(int argc, char** argv) {
  int err = 0;
  /* take the address of every entry point so that the linker has to resolve it */
  const void* fns[] = {(const void*)sfm_abi_version, (const void*)sfm_create, (const void*)sfm_destroy,
                       (const void*)sfm_last_error, (const void*)sfm_strerror, (const void*)sfm_host_alloc,
                       (const void*)sfm_host_free, (const void*)sfm_upload_descriptors,
                       (const void*)sfm_upload_descriptors_u8, (const void*)sfm_upload_descriptors_async,
                       (const void*)sfm_upload_descriptors_bin, (const void*)sfm_match_pairs,
                       (const void*)sfm_fetch_matches, (const void*)sfm_match_pairs_resident,
                       (const void*)sfm_triangulate_batch, (const void*)sfm_upload_keypoints,
                       (const void*)sfm_get_matched_points, (const void*)sfm_reconstruct_pair,
                       (const void*)sfm_reproject_residuals, (const void*)sfm_reproject_jacobians,
                       (const void*)sfm_estimate_normals, (const void*)sfm_save_structure,
                       (const void*)sfm_write_ply_binary, (const void*)sfm_triangulate_batch_timed,
                       (const void*)sfm_reproject_residuals_timed, (const void*)sfm_probe_i8_peak,
                       (const void*)sfm_probe_fp64_peak, (const void*)sfm_launch_count, (const void*)sfm_last_rechecked_rows,
                       (const void*)sfm_timer_start, (const void*)sfm_timer_stop, (const void*)sfm_sync,
                       (const void*)sfm_bank_layout, (const void*)sfm_bank_upload_range,
                       (const void*)sfm_bank_commit, (const void*)sfm_bank_image_rows,
                       (const void*)sfm_bank_rows_dev, (const void*)sfm_bank_copy_peer,
                       (const void*)sfm_bank_layout_async, (const void*)sfm_bank_upload_range_async,
                       (const void*)sfm_bank_commit_async, (const void*)sfm_upload_stream,
                       (const void*)sfm_peer_export, (const void*)sfm_peer_connect,
                       (const void*)sfm_peer_disconnect, (const void*)sfm_bank_ready_async,
                       (const void*)sfm_bank_push_range_async, (const void*)sfm_bank_pull_commit_async,
                       (const void*)sfm_match_rows_begin, (const void*)sfm_match_rows_finish,
                       (const void*)sfm_ba_create, (const void*)sfm_ba_evaluate, (const void*)sfm_ba_destroy};
  size_t n = sizeof fns / sizeof fns[0], i;
  for (i = 0; i < n; ++i)
    if (!fns[i]) return 2;
  if (sfm_abi_version() != SFM_B200_ABI_VERSION) return 3;
  if (sizeof(sfm_match_t) != 16 || sizeof(sfm_knn2_t) != 16) return 4;   /* == cv::DMatch */
  {
    sfm_ctx* ctx = sfm_create(0, &err);
    if (ctx) {
      printf("device context created\n");
      sfm_destroy(ctx);
    } else {
      printf("no device: %s (%d)\n", sfm_last_error(NULL), err);
      if (err != SFM_E_NO_DEVICE && err != SFM_E_CUDA) return 5;
    }
  }
  if (argc > 1) {   /* host-only writer: one camera, two points */
    const double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, T[3] = {0, 0, 0};
    const double X[6] = {0.5, -1.25, 9.0, 1.0, 2.0, 3.0};
    const unsigned char c[6] = {1, 2, 3, 4, 5, 6};
    if (sfm_save_structure(argv[1], 1, R, T, 2, X, 2, c) != SFM_OK) return 6;
  }
  printf("ok %u entry points\n", (unsigned)n);
  return 0;
}
