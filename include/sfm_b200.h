/*
 * sfm_b200.h -- C ABI of the B200-native SfM hot path.
 *
 * Drop-in boundary behind the three call sites of the reference
 * (CaptainEven/SFM_OpenCV, OpenCV_SFM/NViewReconstuct.cpp):
 *
 *   match_features()            :873-913  (SIFT/L2 form: TwoViewReconstruct.cpp:156-194)
 *   match_features_for_all()    :850-871
 *   reconstruct()               :1117-1159 (cv::triangulatePoints + f32 de-homogenise)
 *   ReprojectCost::operator()   :142-184  (residual blocks enumerated at :1187-1211)
 *
 * The reference has no FFI of its own (one C++ translation unit), so this
 * header is what a maintainer binds instead of the OpenCV / Ceres calls; see
 * INTEGRATION.md for the replacement bodies.
 *
 * Conventions
 *   - plain C, no exceptions, no torch / OpenCV types;
 *   - every function returns 0 (SFM_OK) or a negative SFM_E_* code;
 *     sfm_last_error(ctx) gives a human-readable message;
 *   - pointers are HOST pointers unless the name ends in _dev; the library
 *     owns all device memory; calls are synchronous on return;
 *   - there is no CPU fallback: without a usable sm_100 device every compute
 *     entry point fails with SFM_E_NO_DEVICE / SFM_E_CUDA.
 *   - a context is bound to ONE device and is not thread-safe; use one
 *     context per host thread / per GPU (pairs, points and observations are
 *     independent, so callers shard them over contexts).  Several contexts on
 *     several devices may live in one process (the reference is a single
 *     process); sfm_bank_copy_peer moves packed descriptors between them.
 */
#ifndef SFM_B200_H
#define SFM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFM_B200_ABI_VERSION 2

#if defined(__GNUC__)
#define SFM_API __attribute__((visibility("default")))
#else
#define SFM_API
#endif

/* error codes */
enum {
  SFM_OK = 0,
  SFM_E_INVALID = -1,        /* null pointer, negative size, bad index ...        */
  SFM_E_NO_DEVICE = -2,      /* no CUDA device / not compute capability 10.x      */
  SFM_E_CUDA = -3,           /* a CUDA runtime / driver call failed               */
  SFM_E_DIM = -4,            /* descriptor dimension is not 128 (binary: not 1..64) */
  SFM_E_NOT_INTEGRAL = -5,   /* float descriptor holds a non-integer value        */
  SFM_E_RANGE = -6,          /* descriptor value outside 0..255, or row norm^2
                                so large that float sqrt is no longer injective   */
  SFM_E_TOO_FEW_TRAIN = -7,  /* a pair has fewer than 2 train descriptors: the
                                reference would read knn_matches[i][1] out of
                                bounds (NViewReconstuct.cpp:884)                  */
  SFM_E_CAPACITY = -8,       /* output buffer too small; offsets[] hold the need  */
  SFM_E_NOT_UPLOADED = -9,   /* sfm_match_pairs before sfm_upload_descriptors     */
  SFM_E_NOMEM = -10
};

typedef struct sfm_ctx sfm_ctx;
typedef struct sfm_ba_problem sfm_ba_problem;

/* Layout-identical to cv::DMatch {int queryIdx; int trainIdx; int imgIdx; float distance;}
 * so a result buffer can be memcpy'd into a std::vector<cv::DMatch>. */
typedef struct sfm_match_t {
  int32_t queryIdx;
  int32_t trainIdx;
  int32_t imgIdx;   /* always 0, as cv::BFMatcher::knnMatch(query, train, ...) sets it */
  float distance;   /* sqrtf((float)squared_distance), bit-exact with cv::NORM_L2       */
} sfm_match_t;

/* Raw k=2 result per query row (what knn_matches[i][0..1] hold at :877). */
typedef struct sfm_knn2_t {
  int32_t trainIdx0;
  int32_t trainIdx1;
  float distance0;
  float distance1;
} sfm_knn2_t;

/* ---- context ------------------------------------------------------------------- */

SFM_API int sfm_abi_version(void);
/* Creates a context on CUDA device `device_id`. On failure returns NULL and, if
 * err is non-null, stores the SFM_E_* code there. */
SFM_API sfm_ctx* sfm_create(int device_id, int* err);
SFM_API void sfm_destroy(sfm_ctx* ctx);
SFM_API const char* sfm_last_error(const sfm_ctx* ctx);   /* ctx may be NULL: last create error */
SFM_API const char* sfm_strerror(int code);

/* Pinned host memory helpers (optional; any host pointer is accepted everywhere,
 * pinned ones make the H2D/D2H copies asynchronous and faster). */
SFM_API void* sfm_host_alloc(size_t bytes);
SFM_API void sfm_host_free(void* p);

/* ---- matching: replaces match_features / match_features_for_all ------------------ */

/* Uploads the descriptor sets of n_img images (what extract_features() leaves in
 * descriptor_for_all, NViewReconstuct.cpp:1366).  desc_f32[i] is a row-major
 * n_desc[i] x 128 float matrix exactly as cv::SIFT produces (CV_32F, integer valued
 * 0..255).  Values are validated on the device: SFM_E_NOT_INTEGRAL / SFM_E_RANGE.
 * Replaces any previously uploaded bank. */
SFM_API int sfm_upload_descriptors(sfm_ctx* ctx, int n_img, const float* const* desc_f32,
                           const int32_t* n_desc, int dim);
/* Same, for callers that already hold uint8 descriptors (4x less PCIe traffic). */
SFM_API int sfm_upload_descriptors_u8(sfm_ctx* ctx, int n_img, const uint8_t* const* desc_u8,
                              const int32_t* n_desc, int dim);

/* Asynchronous upload: returns as soon as the copies and pack kernels are queued on the
 * context's copy stream.  desc[i] must stay valid (ideally pinned: sfm_host_alloc) until the next
 * sfm_match_pairs* call returns; that call makes each of its kernels wait only for the images
 * it reads, so matching overlaps the rest of the transfer, and it reports this upload's
 * validation result (SFM_E_NOT_INTEGRAL / SFM_E_RANGE) instead of its own results.
 * elem_bytes: 4 = CV_32F rows (what cv::SIFT produces), 1 = CV_8U rows. */
SFM_API int sfm_upload_descriptors_async(sfm_ctx* ctx, int n_img, const void* const* desc,
                                 const int32_t* n_desc, int dim, int elem_bytes);

/* Binary descriptors for the reference's LIVE configuration: AKAZE descriptors (61 bytes,
 * CV_8U) matched with BFMatcher(NORM_HAMMING2) (NViewReconstuct.cpp:797, :875-877).
 * desc_u8[i] is a row-major n_desc[i] x bytes uint8 matrix, 1 <= bytes <= 64.  After this
 * upload sfm_match_pairs / sfm_match_pairs_resident compute cv::NORM_HAMMING2 distances
 * (number of differing 2-bit cells; sfm_match_t::distance holds that integer as a float,
 * as cv::DMatch does) with the same (distance, lower train index) order and the same two
 * filter passes.  Replaces any previously uploaded bank. */
SFM_API int sfm_upload_descriptors_bin(sfm_ctx* ctx, int n_img, const uint8_t* const* desc_u8,
                               const int32_t* n_desc, int bytes);

/* ---- sharded upload: N GPUs, every image crosses PCIe once ---------------
 *
 * match_features_for_all has no cross-pair state (NViewReconstuct.cpp:857-870), so pairs are
 * sharded over GPUs; every GPU needs the whole descriptor bank, but it only has to come from
 * the host once: GPU g uploads (and validates, packs to u8) its slice of the images, and the
 * packed 128-byte rows travel GPU to GPU over NVLink -- with an all-gather on the caller's side
 * (one process per GPU: sfm_bank_rows_dev gives the buffer, e.g. ncclAllGather in place) or
 * with sfm_bank_copy_peer (one process, several contexts).
 *
 *   sfm_bank_layout(ctx, n_img, n_desc, 128)          same on every GPU: allocates, no data
 *   sfm_bank_upload_range(ctx, first, n, desc, 4)     this GPU's images, host -> bank
 *   <exchange the row ranges of the other images>     NCCL / sfm_bank_copy_peer
 *   sfm_bank_commit(ctx, first2, n2)                  images that arrived from a peer: norms,
 *                                                     keys, |row|^2 range check
 * The bank is usable (sfm_match_pairs*) once every image was uploaded or committed. */
SFM_API int sfm_bank_layout(sfm_ctx* ctx, int n_img, const int32_t* n_desc, int dim);
/* desc[k] = host rows of image first_img + k (elem_bytes 4: CV_32F, 1: CV_8U); synchronous. */
SFM_API int sfm_bank_upload_range(sfm_ctx* ctx, int first_img, int n_img, const void* const* desc,
                                  int elem_bytes);
SFM_API int sfm_bank_commit(sfm_ctx* ctx, int first_img, int n_img);
/* Rows [row0, row0 + rows) of the bank belong to image img (rows = count padded to 256; rows
 * of consecutive images are contiguous). */
SFM_API int sfm_bank_image_rows(const sfm_ctx* ctx, int img, int64_t* row0, int64_t* rows);
/* DEVICE pointer to the packed bank, uint8 [n_rows][128] (valid until the next layout /
 * upload); NULL before sfm_bank_layout. */
SFM_API void* sfm_bank_rows_dev(sfm_ctx* ctx, int64_t* n_rows);
/* Copies the packed rows of images [first_img, first_img + n_img) from src's bank (another
 * context of this process, usually on another device: cudaMemcpyPeerAsync over NVLink) into
 * dst's bank and commits them.  Both banks must have the same layout. */
SFM_API int sfm_bank_copy_peer(sfm_ctx* dst, sfm_ctx* src, int first_img, int n_img);

/* ---- staged arrival (N GPUs, upload and exchange overlapped with matching) ---------------
 *
 * The same three steps, queued on the context's upload stream without host synchronisation;
 * the caller enqueues its own peer transfers on that stream between them:
 *
 *   sfm_bank_layout_async(ctx, n_img, n_desc, 128)
 *   for every region k of the image list (2-4 regions):
 *     sfm_bank_upload_range_async(ctx, my part of region k)      host -> bank (pinned memory)
 *     <all-gather of region k on sfm_upload_stream(ctx)>          e.g. ncclAllGather in place
 *     sfm_bank_commit_async(ctx, the peers' parts of region k)
 *   sfm_match_pairs(...)                                          ONE call, as usual
 *
 * Every asynchronous range call is one arrival stage.  sfm_match_pairs visits the pairs in
 * the order their images arrive (results stay in caller order) and makes every kernel launch
 * wait only for the images it reads, so the pairs of region 0 are matched while regions 1..
 * are still crossing PCIe / NVLink; it also reports the validation result of the whole upload
 * (SFM_E_NOT_INTEGRAL / SFM_E_RANGE).  Host arrays must stay valid until it returns.
 * match_features_for_all (NViewReconstuct.cpp:857-870) has no cross-pair state, so the
 * visiting order is free. */
SFM_API int sfm_bank_layout_async(sfm_ctx* ctx, int n_img, const int32_t* n_desc, int dim);
SFM_API int sfm_bank_upload_range_async(sfm_ctx* ctx, int first_img, int n_img,
                                        const void* const* desc, int elem_bytes);
SFM_API int sfm_bank_commit_async(sfm_ctx* ctx, int first_img, int n_img);
/* The upload stream as a cudaStream_t (opaque here): work a caller enqueues on it is ordered
 * with the asynchronous uploads and commits. */
SFM_API void* sfm_upload_stream(sfm_ctx* ctx);

/* ---- peer exchange without a collective: push over NVLink, flags in peer memory ------------
 *
 * An NCCL all-gather needs SMs, and the persistent kNN kernel owns every SM for the length of a
 * launch, so a collective for region k+1 queues behind the matching of region k.  Here every
 * rank PUSHES its packed rows into the peers' banks with the copy engines (cudaMemcpyAsync to
 * peer memory mapped through CUDA IPC, or plain peer pointers inside one process) and raises a
 * flag in the peers' mailboxes; the receivers wait on their own mailbox with stream memory
 * operations (SFM_PEER_FLAGS=kernel: 1-thread kernels).  No SMs, no host synchronisation.
 *
 *   once (same layout on every rank):
 *     sfm_bank_layout(ctx, ...); sfm_peer_export(ctx, h); <exchange the handles, any transport>;
 *     sfm_peer_connect(ctx, my_rank, n_ranks, all_handles)
 *   every step, tag = 1, 2, 3, ... (the same on every rank):
 *     sfm_bank_layout_async(ctx, ...); sfm_bank_ready_async(ctx, tag)
 *     for every region k:
 *       sfm_bank_upload_range_async(ctx, my part of region k)
 *       sfm_bank_push_range_async(ctx, my part, k, tag)      waits for each peer's READY(tag)
 *       for every peer r: sfm_bank_pull_commit_async(ctx, r, r's part of region k, k, tag)
 *     sfm_match_pairs(...)
 * A peer may run ahead by less than one step: its pushes for step tag+1 wait for this rank's
 * READY(tag+1), which is raised behind this rank's own layout of step tag+1.
 * The bank must not be re-allocated between export and use (same image sizes every step).
 * Host arrays of the asynchronous calls should be pinned (sfm_host_alloc): with a pageable
 * source cudaMemcpyAsync blocks the host until the stream reaches the copy -- behind a wait for
 * a peer's flag, and forever if the same host thread is the one that has yet to raise it. */
#define SFM_PEER_HANDLE_BYTES 160
SFM_API int sfm_peer_export(sfm_ctx* ctx, void* handle /* SFM_PEER_HANDLE_BYTES */);
SFM_API int sfm_peer_connect(sfm_ctx* ctx, int my_rank, int n_ranks,
                             const void* handles /* n_ranks x SFM_PEER_HANDLE_BYTES, rank order */);
SFM_API int sfm_peer_disconnect(sfm_ctx* ctx);
SFM_API int sfm_bank_ready_async(sfm_ctx* ctx, uint32_t tag);
/* slot: 0..14, the region's index -- one flag per (slot, source rank) in every mailbox.
 * sfm_bank_pull_commit_async with n_img = 0 only waits for src_rank's flag: the parts of several
 * peers that are contiguous in the bank can then be committed by ONE sfm_bank_commit_async (a
 * region needs two: the peers in front of and behind this rank's own part). */
SFM_API int sfm_bank_push_range_async(sfm_ctx* ctx, int first_img, int n_img, int slot, uint32_t tag);
SFM_API int sfm_bank_pull_commit_async(sfm_ctx* ctx, int src_rank, int first_img, int n_img, int slot,
                                       uint32_t tag);

/* For every pair p: knnMatch(desc[pair_q[p]], desc[pair_t[p]], k=2) with NORM_L2,
 * then the reference's two filter passes (NViewReconstuct.cpp:880-908):
 *   pass 1  min_dist = min{ d0 : !(d0 > ratio*d1) }           (double compare)
 *   pass 2  keep knn[i][0] iff !(d0 > ratio*d1 || d0 > gate_mult*max(min_dist, dist_floor))
 * Reference constants: ratio 0.6, dist_floor 10.0f, gate_mult 5.
 * Kept matches of pair p are written, in ascending queryIdx, to
 * out[offsets[p] .. offsets[p+1]).  offsets has n_pairs+1 entries and is always
 * filled (also on SFM_E_CAPACITY, when out_cap < offsets[n_pairs]).
 * knn_raw (nullable) receives sum_p n_desc[pair_q[p]] rows, pair after pair.
 * min_dist (nullable) receives n_pairs floats (FLT_MAX when no row passes). */
SFM_API int sfm_match_pairs(sfm_ctx* ctx, const int32_t* pair_q, const int32_t* pair_t, int n_pairs,
                    double ratio, float dist_floor, float gate_mult,
                    sfm_match_t* out, int64_t out_cap, int64_t* offsets,
                    sfm_knn2_t* knn_raw, float* min_dist);

/* Copies the kept matches of the most recent sfm_match_pairs call (they stay resident on
 * the device) to `out`.  Lets a caller size the buffer exactly: call sfm_match_pairs with
 * out = NULL / out_cap = 0 (returns SFM_E_CAPACITY when there are matches, offsets filled),
 * allocate offsets[n_pairs] entries, then fetch -- the kNN is not recomputed. */
SFM_API int sfm_fetch_matches(sfm_ctx* ctx, sfm_match_t* out, int64_t out_cap);

/* ---- one pair sharded by QUERY ROWS over several GPUs (BASELINE config 4) --
 *
 * Query rows of a pair are independent in knnMatch, but min_dist of pass 1
 * (NViewReconstuct.cpp:880-894) couples them.  sfm_match_rows_begin matches query rows
 * [q_first[p], q_first[p] + q_count[p]) of image pair_q[p] against image pair_t[p] and returns
 * this shard's pass-1 value per pair; the caller reduces it over the shards (minimum; one float
 * per pair) and sfm_match_rows_finish runs pass 2 under the reduced value.  Fetch the kept
 * matches with sfm_fetch_matches: queryIdx counts within the query image, so the shards' lists
 * concatenated in row order ARE the list of the unsharded sfm_match_pairs call.
 * knn_raw (nullable) receives sum_p q_count[p] rows. */
SFM_API int sfm_match_rows_begin(sfm_ctx* ctx, const int32_t* pair_q, const int32_t* pair_t,
                                 const int32_t* q_first, const int32_t* q_count, int n_pairs,
                                 double ratio, float* min_dist);
SFM_API int sfm_match_rows_finish(sfm_ctx* ctx, const float* min_dist, float dist_floor,
                                  float gate_mult, int64_t* offsets, sfm_knn2_t* knn_raw);

/* Device-resident variant used to time the kernels without PCIe: runs the same
 * kernels, keeps results on the device, returns only the total number of kept
 * matches. kernel_ms (nullable) receives the CUDA-event time of the kNN kernel alone,
 * total_ms (nullable) of the whole device pipeline. */
SFM_API int sfm_match_pairs_resident(sfm_ctx* ctx, const int32_t* pair_q, const int32_t* pair_t,
                             int n_pairs, double ratio, float dist_floor, float gate_mult,
                             int64_t* total_matches, float* kernel_ms, float* total_ms);

/* ---- triangulation: replaces cv::triangulatePoints + de-homogenise in reconstruct() -- */

/* P: n_views row-major 3x4 float projection matrices, built by the caller exactly as
 *    NViewReconstuct.cpp:1129-1143 does (float32 product fK*[R|T]).
 * xy: view-major [n_views][n_pts][2] float image points (view v of point i at
 *    xy[(v*n_pts+i)*2]); for n_views==2 these are the reference's pts2d_1, pts2d_2.
 * X4 (nullable): [4][n_pts] float, the cv::triangulatePoints output layout (unit-norm
 *    homogeneous columns; sign unspecified, as in OpenCV).
 * xyz (nullable): [n_pts][3] double == std::vector<cv::Point3d>: float32 X/W widened to
 *    double (NViewReconstuct.cpp:1151-1156).
 * n_pts == 0 returns SFM_E_INVALID like the reference's -1 (:1122-1126). */
SFM_API int sfm_triangulate_batch(sfm_ctx* ctx, const float* P, const float* xy, int n_views,
                          int64_t n_pts, float* X4, double* xyz);

/* ---- match list -> points -> structure without a host round trip ----------------------
 *
 * The reference turns a pair's match list into 3-D points on the host:
 * get_matched_points (NViewReconstuct.cpp:989-1003), maskout_points (:943), the float32
 * projection build and cv::triangulatePoints + de-homogenise in reconstruct() (:1117-1159).
 * Here the match list of the last sfm_match_pairs call stays on the device and is consumed
 * there. */

/* Pixel coordinates of every image's keypoints (cv::KeyPoint::pt of key_points_for_all,
 * :1366): kp_xy[i] is n_kp[i] x 2 float; n_kp[i] must equal the descriptor count. */
SFM_API int sfm_upload_keypoints(sfm_ctx* ctx, int n_img, const float* const* kp_xy,
                         const int32_t* n_kp);

/* get_matched_points for pair `pair` (index into the last sfm_match_pairs call): out_p1[i] =
 * kp[query][match.queryIdx], out_p2[i] = kp[train][match.trainIdx], optionally compacted by
 * mask (nullable, one byte per match, kept when > 0, as maskout_points does).  n_points
 * receives the number of points; SFM_E_CAPACITY when it exceeds cap. */
SFM_API int sfm_get_matched_points(sfm_ctx* ctx, int pair, const uint8_t* mask, float* out_p1,
                           float* out_p2, int64_t cap, int64_t* n_points);

/* reconstruct(K, R1, T1, R2, T2, p1, p2, structure) on the (masked) matches of `pair`:
 * device-side gather, P = float32(K) * float32([R|T]) (:1129-1143), DLT triangulation and
 * float32 de-homogenise; structure is [n_points][3] double == std::vector<cv::Point3d>.
 * K, R1, R2 are row-major 3x3, T1, T2 3-vectors (CV_64F in the reference).  No matches left
 * is the reference's "[Err]: empty 2d points." (-1): SFM_E_INVALID. */
SFM_API int sfm_reconstruct_pair(sfm_ctx* ctx, int pair, const double K[9], const double R1[9],
                         const double T1[3], const double R2[9], const double T2[3],
                         const uint8_t* mask, double* structure, int64_t cap, int64_t* n_points);

/* ---- reprojection residuals: replaces ReprojectCost::operator() evaluation --------- */

/* intr = {fx, fy, cx, cy} (:1464-1471); ext[c] = {angle-axis(3), t(3)} (:1478-1486);
 * pts[j] = 3 doubles (&pts3d[j].x, :1209); observation k belongs to camera cam_idx[k],
 * point pt_idx[k], pixel obs_xy[2k..2k+1] (float KeyPoint::pt widened to double, :1199).
 * resid (nullable): [n_obs][2] doubles, in the caller's observation order (the reference
 * enumerates camera-major, :1187-1211).
 * huber_cost (nullable): 0.5 * sum rho(|r|^2) with ceres::HuberLoss(huber_delta) (:1184);
 * huber_delta <= 0 gives the plain 0.5*sum |r|^2. */
SFM_API int sfm_reproject_residuals(sfm_ctx* ctx, const double intr[4], const double* ext, int n_cam,
                            const double* pts, int64_t n_pts, const int32_t* cam_idx,
                            const int32_t* pt_idx, const float* obs_xy, int64_t n_obs,
                            double huber_delta, double* resid, double* huber_cost);

/* Jacobians of the same residual blocks with respect to the three parameter blocks Ceres is
 * given at NViewReconstuct.cpp:1202-1209 -- what AutoDiffCostFunction<ReprojectCost, 2, 4, 6, 3>
 * computes with Jets.  jac is [n_obs][2][13] doubles, row-major: for each of the two residual
 * rows the derivatives with respect to (fx, fy, cx, cy | angle-axis(3), t(3) | X, Y, Z).
 * resid is nullable.  iters > 0 keeps the inputs on the device, launches the kernel `iters`
 * times and reports the mean CUDA-event time per launch in ms_per_launch (nullable). */
SFM_API int sfm_reproject_jacobians(sfm_ctx* ctx, const double intr[4], const double* ext, int n_cam,
                            const double* pts, int64_t n_pts, const int32_t* cam_idx,
                            const int32_t* pt_idx, const float* obs_xy, int64_t n_obs,
                            double* resid, double* jac, int iters, float* ms_per_launch);

/* ---- the same inside a bundle-adjustment loop ------------------------------------------
 *
 * bundle_adjustment() (NViewReconstuct.cpp:1162-1244) creates its residual blocks once
 * (:1187-1211); Ceres then evaluates them per LM iteration with new extrinsics and points.
 * sfm_ba_create uploads and range-checks (on the device) the observation tables once;
 * sfm_ba_evaluate moves only ext (48 B / camera) and pts (24 B / point) to the device -- either
 * may be NULL to keep the previous values -- and returns what is asked for: resid [n_obs][2],
 * jac [n_obs][2][13] (layout of sfm_reproject_jacobians), huber_cost (see
 * sfm_reproject_residuals), kernel_ms = CUDA-event time of the kernels.  All nullable. */
SFM_API int sfm_ba_create(sfm_ctx* ctx, int n_cam, int64_t n_pts, const int32_t* cam_idx,
                          const int32_t* pt_idx, const float* obs_xy, int64_t n_obs,
                          sfm_ba_problem** out);
SFM_API int sfm_ba_evaluate(sfm_ctx* ctx, sfm_ba_problem* problem, const double intr[4],
                            const double* ext, const double* pts, double huber_delta,
                            double* resid, double* jac, double* huber_cost, float* kernel_ms);
SFM_API void sfm_ba_destroy(sfm_ctx* ctx, sfm_ba_problem* problem);

/* estimate_normals(pts3d, K, normals) (NViewReconstuct.cpp:551-599, called with K = 10 at
 * :1502) with PCAFitPlane (:601-690): per point the K nearest OTHER points (brute force), the
 * eigenvector of smallest eigenvalue of their covariance, oriented so that normal . centroid <= 0
 * and normalised.  pts and normals are [n_pts][3] doubles (std::vector<cv::Point3d>).
 * 3 <= K <= 16, n_pts > K. */
SFM_API int sfm_estimate_normals(sfm_ctx* ctx, const double* pts, int64_t n_pts, int K, double* normals);

/* ---- output files: save_structure() / write_ply_binary() ---------------------------- */

/* Writes the file save_structure() writes (NViewReconstuct.cpp:186-227) byte for byte as
 * cv::FileStorage (YAML 1.0) would: "Camera Count", "Point Count", "Rotations" (n_cam 3x3
 * CV_64F matrices, row-major), "Motions" (n_cam 3x1), "Points" (n_pts Point3d) and "Colors"
 * (n_colors Vec3b, stored b,g,r as the reference holds them).  Host-only (no device). */
SFM_API int sfm_save_structure(const char* file_name, int n_cam, const double* rotations,
                       const double* motions, int64_t n_pts, const double* structure,
                       int64_t n_colors, const uint8_t* colors);

/* write_ply_binary() (NViewReconstuct.cpp:229-294): binary little-endian PLY, 27 bytes per
 * vertex (x y z nx ny nz as float, r g b as uchar); xyz_normal is [n][6], rgb [n][3];
 * vertices holding a NaN are skipped and not counted, as in the reference.  crlf != 0 writes
 * the header lines with "\r\n", which is what the reference's text-mode stream produces on
 * its platform (and what the bundled Viewer/structure_ba.ply holds). Host-only. */
SFM_API int sfm_write_ply_binary(const char* path, int64_t n, const float* xyz_normal,
                         const uint8_t* rgb, int crlf);

/* ---- device-resident benchmarking hooks (inputs already in HBM) --------------------- */

/* Stages the inputs of sfm_triangulate_batch / sfm_reproject_residuals on the device
 * once, then runs the kernel `iters` times and reports the mean CUDA-event time per
 * launch.  Results of the last launch are copied back when the output pointers are
 * non-null.  Used by bench.py for the HBM roofline lines. */
SFM_API int sfm_triangulate_batch_timed(sfm_ctx* ctx, const float* P, const float* xy, int n_views,
                                int64_t n_pts, float* X4, double* xyz, int iters,
                                float* ms_per_launch);
SFM_API int sfm_reproject_residuals_timed(sfm_ctx* ctx, const double intr[4], const double* ext,
                                  int n_cam, const double* pts, int64_t n_pts,
                                  const int32_t* cam_idx, const int32_t* pt_idx,
                                  const float* obs_xy, int64_t n_obs, double huber_delta,
                                  double* resid, double* huber_cost, int iters,
                                  float* ms_per_launch);

/* Bare tcgen05.mma.kind::i8 issue-rate probe: `iters` back-to-back 128x256x32 u8 MMAs
 * per SM on all SMs, no epilogue.  Reports achieved tera-ops/s (2*M*N*K per MMA): the
 * measured int8 tensor peak the matching roofline is quoted against. */
SFM_API int sfm_probe_i8_peak(sfm_ctx* ctx, int iters, double* tops);

/* Bare fp64 FMA issue-rate probe (registers only): achieved TFLOP/s, FMA counted as 2.  The
 * measured ceiling of the triangulation kernel, which is fp64-pipe bound. */
SFM_API int sfm_probe_fp64_peak(sfm_ctx* ctx, int iters, double* tflops);

/* Number of kernel launches issued by this context so far (bench.py's gpu_launches). */
SFM_API int64_t sfm_launch_count(const sfm_ctx* ctx);
/* Rows of the last match-only call (knn_raw == NULL) that the ratio-driven sweep could not decide and
 * that were recomputed exactly on the CUDA cores before the filter passes (diagnostics; 0 when the
 * sweep was not used: knn_raw requested, binary descriptors, SFM_PRUNE_MODE=0, or a call below the
 * 2048 work items (256-row query blocks) from which the default SFM_PRUNE_MODE=2 uses it). */
SFM_API int64_t sfm_last_rechecked_rows(const sfm_ctx* ctx);

/* CUDA-event stopwatch on the context's own stream (torch.cuda.Event only sees torch's
 * stream): sfm_timer_start records, sfm_timer_stop records + synchronises and returns the
 * device time in ms of everything the context enqueued in between. */
SFM_API int sfm_timer_start(sfm_ctx* ctx);
SFM_API int sfm_timer_stop(sfm_ctx* ctx, float* ms);
SFM_API int sfm_sync(sfm_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SFM_B200_H */
