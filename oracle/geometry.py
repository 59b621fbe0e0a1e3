"""CPU oracle for triangulation and reprojection residuals -- TEST INFRASTRUCTURE only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.

Restates (file:line relative to the reference checkout):

* ``OpenCV_SFM/NViewReconstuct.cpp:1117-1159`` ``reconstruct``: float32 projection build
  (:1129-1143), ``cv::triangulatePoints`` (:1147), float32 de-homogenise (:1149-1156).
  2-view ancestor: ``OpenCV_SFM/TwoViewReconstruct.cpp:231-250``.
* ``OpenCV_SFM/NViewReconstuct.cpp:142-184`` ``ReprojectCost::operator()`` and the
  observation order of ``bundle_adjustment`` (:1187-1211), ``HuberLoss(4)`` (:1184),
  reported RMSE (:1237-1238).

Third-party arithmetic that is not vendored in the reference:

* ``cv::triangulatePoints`` (OpenCV calib3d, reference pins 4.4.0): per point, the 4x4 DLT
  matrix with rows ``x*P[2]-P[0]``, ``y*P[2]-P[1]`` per view is SVD'd in float64 and the
  right singular vector of the smallest singular value is returned in the points' dtype.
  :func:`triangulate_svd` restates that (and generalises to V views);
  :func:`triangulate_cv` calls the cv2 4.13 wheel of this image.  They are pinned against
  each other and against ``Viewer/structure.yml`` (first two-view block) in the tests.
* ``ceres::AngleAxisRotatePoint`` (ceres/rotation.h, version unpinned by the reference):
  restated in :func:`angle_axis_rotate`; pinned against ``cv2.Rodrigues``/``cv2.projectPoints``.
  Nothing in the reference pins residual VALUES (the console RMSE was never saved):
  residual parity is "pinned against cv2.projectPoints only".
"""
from __future__ import annotations

import numpy as np

# Hard-coded intrinsics of the reference, NViewReconstuct.cpp:1353-1356
K_REFERENCE = np.array([[2826.561, 0.0, 1835.259],
                        [0.0, 2826.519, 1370.103],
                        [0.0, 0.0, 1.0]], np.float64)


def build_projection(K, R, T) -> np.ndarray:
    """P = float32(K) * float32([R|T]), NViewReconstuct.cpp:1129-1143, with cv::gemm's
    evaluation order (pinned against cv2.gemm in tests/test_oracle_geometry.py)."""
    RT = np.empty((3, 4), np.float32)
    RT[:, :3] = np.asarray(R, np.float64).astype(np.float32)
    RT[:, 3] = np.asarray(T, np.float64).reshape(3).astype(np.float32)
    fK = np.asarray(K, np.float64).astype(np.float32)
    # cv::gemm's small-matrix path: ((a0*b0 + a1*b1) + a2*b2) in float32, no FMA
    p = fK[:, :, None] * RT[None, :, :]                  # float32 products [r, k, c]
    return ((p[:, 0, :] + p[:, 1, :]) + p[:, 2, :]).astype(np.float32)


def dlt_matrix(P: np.ndarray, xy: np.ndarray) -> np.ndarray:
    """A[n, 2v:2v+2, :] = (x*P_v[2]-P_v[0], y*P_v[2]-P_v[1]) in float64.

    P: [V,3,4]; xy: [V,N,2] -> A: [N,2V,4]
    """
    P = np.asarray(P, np.float64)
    xy = np.asarray(xy, np.float64)
    V, N = xy.shape[0], xy.shape[1]
    A = np.empty((N, 2 * V, 4), np.float64)
    for v in range(V):
        A[:, 2 * v, :] = xy[v, :, 0:1] * P[v, 2][None, :] - P[v, 0][None, :]
        A[:, 2 * v + 1, :] = xy[v, :, 1:2] * P[v, 2][None, :] - P[v, 1][None, :]
    return A


def triangulate_svd(P: np.ndarray, xy: np.ndarray, out_dtype=np.float32) -> np.ndarray:
    """Restatement of cv::triangulatePoints for V>=2 views. Returns X4 [4,N] (unit columns,
    sign arbitrary) in ``out_dtype`` (cv returns the points' dtype: float32 in the reference)."""
    A = dlt_matrix(P, xy)
    _, _, vt = np.linalg.svd(A)
    return np.ascontiguousarray(vt[:, -1, :].T).astype(out_dtype)


def triangulate_cv(P1, P2, p1, p2) -> np.ndarray:
    """cv2.triangulatePoints with the reference's argument types (float32 P, float32 2xN)."""
    import cv2
    a = np.ascontiguousarray(np.asarray(p1, np.float32).T)
    b = np.ascontiguousarray(np.asarray(p2, np.float32).T)
    return cv2.triangulatePoints(np.asarray(P1, np.float32), np.asarray(P2, np.float32), a, b)


def dehomogenize(X4: np.ndarray) -> np.ndarray:
    """``pt4d_homo /= pt4d_homo(3)`` on a ``Mat_<float>`` column, stored as Point3d
    (NViewReconstuct.cpp:1151-1156).  OpenCV implements ``Mat_<T> /= double`` as
    ``a.convertTo(a, -1, 1./b)`` (core/mat.inl.hpp): the reciprocal is taken in double, rounded to
    float, and every element is multiplied by it in float32 -- NOT divided.  Pinned by the
    reference's own bundled output: with this form 1835 of the 1847 two-view points of
    Viewer/structure.yml are reproduced bit for bit (quotient form: 758; see
    tests/test_oracle_geometry.py::test_dehomogenize_form_is_pinned_by_structure_yml)."""
    X4 = np.asarray(X4, np.float32)
    with np.errstate(divide="ignore"):
        rw = (np.float64(1.0) / X4[3].astype(np.float64)).astype(np.float32)
    xyz = (X4[:3] * rw[None, :]).astype(np.float32)
    return np.ascontiguousarray(xyz.T).astype(np.float64)


def reconstruct(K, R1, T1, R2, T2, p1, p2, use_cv: bool = True):
    """reconstruct(), NViewReconstuct.cpp:1117-1159. Returns (xyz [N,3] f64, X4 [4,N] f32)."""
    if len(p1) == 0 or len(p2) == 0:
        raise ValueError("[Err]: empty 2d points.")   # the reference returns -1 (:1122-1126)
    P1 = build_projection(K, R1, T1)
    P2 = build_projection(K, R2, T2)
    if use_cv:
        X4 = triangulate_cv(P1, P2, p1, p2)
    else:
        xy = np.stack([np.asarray(p1, np.float32), np.asarray(p2, np.float32)])
        X4 = triangulate_svd(np.stack([P1, P2]), xy)
    return dehomogenize(X4), X4


def point_rel_err(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Per-point relative error ||a-b|| / ||b|| (the parity metric for triangulated points)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.linalg.norm(a - b, axis=-1) / np.maximum(np.linalg.norm(b, axis=-1), 1e-300)


# --------------------------------------------------------------------------- residuals

def angle_axis_rotate(w: np.ndarray, X: np.ndarray) -> np.ndarray:
    """ceres::AngleAxisRotatePoint (ceres/rotation.h) in float64, broadcasting [...,3]."""
    w = np.asarray(w, np.float64)
    X = np.asarray(X, np.float64)
    theta2 = (w * w).sum(-1, keepdims=True)
    big = theta2 > np.finfo(np.float64).eps
    theta = np.sqrt(np.where(big, theta2, 1.0))
    c, s = np.cos(theta), np.sin(theta)
    wn = w * (1.0 / theta)
    wxp = np.cross(wn, X)
    tmp = (wn * X).sum(-1, keepdims=True) * (1.0 - c)
    far = X * c + wxp * s + wn * tmp
    near = X + np.cross(w, X)
    return np.where(big, far, near)


def reproject_residuals(intr, ext, pts, cam_idx, pt_idx, obs_xy) -> np.ndarray:
    """ReprojectCost::operator(), NViewReconstuct.cpp:151-183, for every observation.

    intr = (fx, fy, cx, cy); ext[c] = (angle-axis, t); returns [n_obs, 2] float64.
    """
    intr = np.asarray(intr, np.float64)
    ext = np.asarray(ext, np.float64).reshape(-1, 6)
    pts = np.asarray(pts, np.float64).reshape(-1, 3)
    e = ext[np.asarray(cam_idx)]
    X = pts[np.asarray(pt_idx)]
    p = angle_axis_rotate(e[:, :3], X) + e[:, 3:]
    x = p[:, 0] / p[:, 2]
    y = p[:, 1] / p[:, 2]
    u = intr[0] * x + intr[2]
    v = intr[1] * y + intr[3]
    obs = np.asarray(obs_xy, np.float32).astype(np.float64).reshape(-1, 2)   # Point2d(kp.pt), :1199
    return np.stack([u - obs[:, 0], v - obs[:, 1]], 1)


def reproject_jacobians(intr, ext, pts, cam_idx, pt_idx) -> np.ndarray:
    """d(residual)/d(intrinsic | extrinsic | point) of ReprojectCost (NViewReconstuct.cpp:151-183)
    in the parameter-block order of :1202-1209 -- what ceres::AutoDiffCostFunction<ReprojectCost,
    2, 4, 6, 3> evaluates.  Closed form of the same function (both AngleAxisRotatePoint
    branches); pinned in tests/test_oracle_geometry.py against cv2.projectPoints' Jacobian and
    against central differences of :func:`reproject_residuals`.  Returns [n_obs, 2, 13] float64."""
    intr = np.asarray(intr, np.float64)
    ext = np.asarray(ext, np.float64).reshape(-1, 6)
    pts = np.asarray(pts, np.float64).reshape(-1, 3)
    e = ext[np.asarray(cam_idx)]
    X = pts[np.asarray(pt_idx)]
    w, t = e[:, :3], e[:, 3:]
    n = X.shape[0]
    theta2 = (w * w).sum(-1)
    big = theta2 > np.finfo(np.float64).eps
    theta = np.sqrt(np.where(big, theta2, 1.0))
    wh = w / theta[:, None]
    s, c = np.sin(theta), np.cos(theta)
    p = angle_axis_rotate(w, X) + t
    x, y, iz = p[:, 0] / p[:, 2], p[:, 1] / p[:, 2], 1.0 / p[:, 2]
    fx, fy = intr[0], intr[1]
    A = np.zeros((n, 2, 3))
    A[:, 0, 0] = fx * iz; A[:, 0, 2] = -fx * x * iz
    A[:, 1, 1] = fy * iz; A[:, 1, 2] = -fy * y * iz
    # rotation matrix (dp/dX) and dp/dw, column by column
    I = np.eye(3)
    wxX = np.cross(wh, X)
    d = (wh * X).sum(-1)
    dpdw = np.zeros((n, 3, 3))
    R = np.zeros((n, 3, 3))
    for q in range(3):
        eq = np.broadcast_to(I[q], (n, 3))
        g = (eq - wh * wh[:, q:q + 1]) / theta[:, None]
        far = (-s * wh[:, q])[:, None] * X + s[:, None] * np.cross(g, X) + (c * wh[:, q])[:, None] * wxX \
            + (1.0 - c)[:, None] * (g * d[:, None] + wh * (g * X).sum(-1, keepdims=True)) \
            + (s * wh[:, q] * d)[:, None] * wh
        near = np.cross(eq, X)
        dpdw[:, :, q] = np.where(big[:, None], far, near)
        R[:, :, q] = angle_axis_rotate(w, eq)                 # R e_q = column q
    J = np.zeros((n, 2, 13))
    J[:, 0, 0] = x; J[:, 0, 2] = 1.0
    J[:, 1, 1] = y; J[:, 1, 3] = 1.0
    J[:, :, 4:7] = A @ dpdw
    J[:, :, 7:10] = A
    J[:, :, 10:13] = A @ R
    return J


def estimate_normals(pts3d, K: int = 10) -> np.ndarray:
    """estimate_normals (NViewReconstuct.cpp:551-599) + PCAFitPlane (:601-690): K nearest other
    points (brute force), covariance of the neighbours about their mean, eigenvector of the
    smallest eigenvalue, flipped when normal . centroid > 0 (:672-677), normalised.
    Pinned against the reference's own bundled output: the normals stored in
    Viewer/structure_ba.ply are reproduced from the points of Viewer/structure_ba.yml with zero
    float32 error (tests/test_oracle_geometry.py)."""
    X = np.asarray(pts3d, np.float64).reshape(-1, 3)
    n = X.shape[0]
    out = np.empty_like(X)
    for r0 in range(0, n, 1024):
        d = X[r0:r0 + 1024, None, :] - X[None, :, :]
        d2 = (d[:, :, 0] * d[:, :, 0] + d[:, :, 1] * d[:, :, 1]) + d[:, :, 2] * d[:, :, 2]
        d2[np.arange(d2.shape[0]), np.arange(r0, r0 + d2.shape[0])] = np.inf
        idx = np.argsort(d2, 1, kind="stable")[:, :K]
        nb = X[idx]
        mean = nb.mean(1)
        c = nb - mean[:, None, :]
        A = np.einsum("nki,nkj->nij", c, c) / K
        _, V = np.linalg.eigh(A)
        nrm = V[:, :, 0].copy()
        nrm[(nrm * mean).sum(1) > 0] *= -1
        out[r0:r0 + 1024] = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    return out


def huber_cost(resid: np.ndarray, delta: float = 4.0) -> float:
    """0.5 * sum rho(s), s = |r|^2, ceres::HuberLoss(delta): rho = s (s <= delta^2),
    2*delta*sqrt(s) - delta^2 otherwise (NViewReconstuct.cpp:1184)."""
    s = (np.asarray(resid, np.float64) ** 2).sum(1)
    if delta <= 0:
        return 0.5 * float(s.sum())
    b = delta * delta
    rho = np.where(s <= b, s, 2.0 * delta * np.sqrt(s) - b)
    return 0.5 * float(rho.sum())


def enumerate_observations(inds_2d_to_3d, keypoints_xy):
    """Residual-block order of bundle_adjustment, NViewReconstuct.cpp:1187-1211:
    for img, for kp: if idx[img][kp] >= 0 -> (img, idx, kp.pt).  Returns cam_idx, pt_idx, obs."""
    cam, pt, obs = [], [], []
    for img, (ids, kps) in enumerate(zip(inds_2d_to_3d, keypoints_xy)):
        ids = np.asarray(ids)
        sel = np.nonzero(ids >= 0)[0]
        cam.append(np.full(sel.size, img, np.int32))
        pt.append(ids[sel].astype(np.int32))
        obs.append(np.asarray(kps, np.float32).reshape(-1, 2)[sel])
    return np.concatenate(cam), np.concatenate(pt), np.concatenate(obs).astype(np.float32)
