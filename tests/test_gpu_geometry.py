"""GPU parity (through the C ABI): sfm_triangulate_batch / sfm_reproject_residuals vs the
CPU oracle.  Tolerance (BASELINE.json north_star): 1e-5 relative."""
import numpy as np
import pytest

from oracle import geometry as G
from oracle import synth

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5      # per-point ||dX||/||X||, and residuals relative to the pixel scale below


def test_two_view_vs_cv2(ctx):
    sc = synth.scene(20000, 2, seed=7)
    X4, xyz = ctx.triangulate_batch(sc["P"], sc["xy"])
    cv = G.triangulate_cv(sc["P"][0], sc["P"][1], sc["xy"][0], sc["xy"][1])
    assert X4.dtype == np.float32 and X4.shape == cv.shape
    assert np.allclose(np.linalg.norm(X4.astype(np.float64), axis=0), 1.0, atol=1e-6)
    ref = G.dehomogenize(cv)
    assert G.point_rel_err(xyz, ref).max() < REL_TOL
    # xyz is the float32 quotient widened to double, exactly as the reference stores it
    assert np.array_equal(xyz, G.dehomogenize(X4))
    # homogeneous vectors agree up to the sign OpenCV leaves unspecified
    s = np.sign((X4 * cv).sum(0))
    assert np.abs(X4 * s - cv).max() < 1e-6


@pytest.mark.parametrize("V", [2, 3, 4, 8])
def test_n_view_vs_svd(ctx, V):
    sc = synth.scene(8000, V, seed=11 + V)
    _, xyz = ctx.triangulate_batch(sc["P"], sc["xy"])
    ref = G.dehomogenize(G.triangulate_svd(sc["P"], sc["xy"]))
    assert G.point_rel_err(xyz, ref).max() < REL_TOL


@pytest.mark.parametrize("V", [2, 4])
def test_planted_mismatches(ctx, V):
    """cv::triangulatePoints is fed mismatched rays by the reference (all matches of later pairs,
    NViewReconstuct.cpp:1441): a quarter of the points get an unrelated observation in the last
    view, so that the smallest singular value of their DLT system is not small.  The kernel must
    return the SVD null vector for those too (not a fixed number of power steps)."""
    sc = synth.scene(20000, V, seed=21 + V)
    rng = np.random.default_rng(5)
    bad = rng.random(20000) < 0.25
    xy = sc["xy"].copy()
    xy[V - 1, bad] = rng.uniform([0, 0], [3648, 2736], size=(int(bad.sum()), 2)).astype(np.float32)
    X4, xyz = ctx.triangulate_batch(sc["P"], xy)
    sv = np.linalg.svd(G.dlt_matrix(sc["P"], xy), compute_uv=False)
    assert (sv[bad, 3] / sv[bad, 2]).max() > 0.3            # the case the fast path cannot do
    if V == 2:
        cv = G.triangulate_cv(sc["P"][0], sc["P"][1], xy[0], xy[1])
    else:
        cv = G.triangulate_svd(sc["P"], xy)
    ok = (sv[:, 3] < 0.999 * sv[:, 2]) & (np.abs(cv[3]) > 1e-6)
    assert ok.sum() > 19900
    assert G.point_rel_err(xyz[ok], G.dehomogenize(cv)[ok]).max() < REL_TOL
    s = np.sign((X4 * cv).sum(0))
    assert np.abs(X4 * s - cv)[:, ok].max() < 2e-6


def test_noiseless_points_are_recovered(ctx):
    sc = synth.scene(5000, 3, seed=5, noise_px=0.0)
    _, xyz = ctx.triangulate_batch(sc["P"], sc["xy"])
    assert G.point_rel_err(xyz, sc["X"]).max() < 1e-3


@pytest.mark.parametrize("n", [1, 31, 257, 1000])
def test_ragged_point_counts(ctx, n):
    sc = synth.scene(n, 2, seed=n)
    _, xyz = ctx.triangulate_batch(sc["P"], sc["xy"])
    ref, _ = G.reconstruct(G.K_REFERENCE, np.eye(3), np.zeros(3), *_rt(sc, 1),
                           sc["xy"][0], sc["xy"][1])
    assert xyz.shape == (n, 3) and G.point_rel_err(xyz, ref).max() < REL_TOL


def _rt(sc, v):
    import cv2
    R, _ = cv2.Rodrigues(sc["ext"][v, :3].reshape(3, 1))
    return R, sc["ext"][v, 3:]


def test_reconstruct_api_and_empty(ctx):
    import sfm_opencv_b200 as sfm
    sc = synth.scene(777, 2, seed=3)
    R, T = _rt(sc, 1)
    xyz = sfm.reconstruct(ctx, G.K_REFERENCE, np.eye(3), np.zeros(3), R, T, sc["xy"][0], sc["xy"][1])
    ref, _ = G.reconstruct(G.K_REFERENCE, np.eye(3), np.zeros(3), R, T, sc["xy"][0], sc["xy"][1])
    assert G.point_rel_err(xyz, ref).max() < REL_TOL
    with pytest.raises(sfm.SfmError):           # reference: "[Err]: empty 2d points." -> -1
        sfm.reconstruct(ctx, G.K_REFERENCE, np.eye(3), np.zeros(3), R, T, np.zeros((0, 2)), np.zeros((0, 2)))


def test_golden_desktop_two_view_points(ctx, golden):
    """Matches of desktop pair 0 -> E/pose by cv2 (host glue, out of scope) -> triangulate on
    the GPU vs cv2.triangulatePoints on the same inliers."""
    import cv2
    g = golden("desktop")
    m = g["match_0"]
    p1 = g["kp_0"][m[:, 0]]; p2 = g["kp_1"][m[:, 1]]
    K = G.K_REFERENCE
    f = 0.5 * (K[0, 0] + K[1, 1]); pp = (K[0, 2], K[1, 2])
    cv2.setRNGSeed(0)
    E, mask = cv2.findEssentialMat(p1, p2, f, pp, cv2.RANSAC, 0.999, 1.0)
    _, R, T, mask = cv2.recoverPose(E, p1, p2, focal=f, pp=pp, mask=mask)
    keep = mask.ravel() > 0
    assert keep.sum() > 300
    import sfm_opencv_b200 as sfm
    xyz = sfm.reconstruct(ctx, K, np.eye(3), np.zeros(3), R, T, p1[keep], p2[keep])
    ref, _ = G.reconstruct(K, np.eye(3), np.zeros(3), R, T, p1[keep], p2[keep])
    assert G.point_rel_err(xyz, ref).max() < REL_TOL


def _resid_close(r, ref):
    # 1e-5 relative; residuals are differences of ~1e3 px quantities, so floor the scale at 1e-3 px
    return np.abs(r - ref).max() <= REL_TOL * max(np.abs(ref).max(), 1e-3) and \
        (np.abs(r - ref) <= REL_TOL * np.maximum(np.abs(ref), 1e-3)).all()


@pytest.mark.parametrize("V", [2, 5])
def test_residuals_vs_oracle(ctx, V):
    n = 30000
    sc = synth.scene(n, V, seed=20 + V)
    cam, pt = synth.observations_camera_major(n, V)
    obs = sc["xy"].reshape(-1, 2)
    r, cost = ctx.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam, pt, obs, huber_delta=4.0)
    ref = G.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam, pt, obs)
    assert _resid_close(r, ref)
    assert abs(cost - G.huber_cost(ref, 4.0)) <= 1e-9 * G.huber_cost(ref, 4.0)
    _, cost0 = ctx.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam, pt, obs, huber_delta=0.0,
                                       want_resid=False)
    assert abs(cost0 - G.huber_cost(ref, 0.0)) <= 1e-9 * G.huber_cost(ref, 0.0)


def test_residuals_random_order_and_small_angle(ctx):
    rng = np.random.default_rng(4)
    n, V = 5000, 4
    sc = synth.scene(n, V, seed=31)
    sc["ext"][2, :3] = [1e-9, -2e-9, 3e-9]          # ceres small-angle branch
    k = 12345
    cam = rng.integers(0, V, k).astype(np.int32)
    pt = rng.integers(0, n, k).astype(np.int32)
    obs = rng.uniform(0, 3000, (k, 2)).astype(np.float32)   # large residuals -> Huber tail
    r, cost = ctx.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam, pt, obs)
    ref = G.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam, pt, obs)
    assert _resid_close(r, ref)
    assert abs(cost - G.huber_cost(ref, 4.0)) <= 1e-9 * G.huber_cost(ref, 4.0)


def test_residual_errors(ctx):
    import sfm_opencv_b200 as sfm
    sc = synth.scene(10, 2, seed=1)
    with pytest.raises(sfm.SfmError):
        ctx.reproject_residuals(sc["intr"], sc["ext"], sc["X"], [0, 2], [0, 1], np.zeros((2, 2)))
    with pytest.raises(sfm.SfmError):
        ctx.reproject_residuals(sc["intr"], sc["ext"], sc["X"], [0, 1], [0, 10], np.zeros((2, 2)))
    r, cost = ctx.reproject_residuals(sc["intr"], sc["ext"], sc["X"], [], [], np.zeros((0, 2)))
    assert r.shape == (0, 2) and cost == 0.0


def test_config5_size_properties(ctx):
    """BASELINE config 5 size (4M points): size-independent properties instead of the oracle:
    unit-norm homogeneous columns, noiseless round trip (triangulate the exact projections ->
    reproject -> zero residual), and linearity of the residual in the observation."""
    n = 4_000_000
    sc = synth.scene(n, 2, seed=11, noise_px=0.0)
    X4, xyz = ctx.triangulate_batch(sc["P"], sc["xy"])
    nrm = np.sqrt((X4.astype(np.float64) ** 2).sum(0))
    assert np.abs(nrm - 1.0).max() < 1e-6
    rel = G.point_rel_err(xyz, sc["X"])
    assert np.percentile(rel, 99.9) < 2e-4          # float32 projections limit the recovery
    cam, pt = synth.observations_camera_major(n, 2)
    obs = sc["xy"].reshape(-1, 2)
    r, cost = ctx.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam, pt, obs)
    assert np.abs(r).max() < 5e-3                    # observations are float32-rounded pixels
    r2, _ = ctx.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam, pt, obs + np.float32(1.0))
    shifted = (obs + np.float32(1.0)).astype(np.float64) - obs.astype(np.float64)
    assert np.abs((r - r2) - shifted).max() < 1e-9
    assert abs(cost - 0.5 * float((r ** 2).sum())) <= 1e-9 * max(1.0, cost)


def test_bundle_adjustment_residuals_api(ctx):
    """Reference-shaped call: blocks enumerated camera-major, Huber(4) cost, RMSE of :1237-1238."""
    import sfm_opencv_b200 as sfm
    sc = synth.scene(600, 3, seed=21, noise_px=3.0)
    rng = np.random.default_rng(1)
    ids = [np.where(rng.random(600) < 0.7, np.arange(600), -1) for _ in range(3)]
    kps = [sc["xy"][v] for v in range(3)]
    r, cost, rmse = sfm.bundle_adjustment_residuals(ctx, sc["intr"], sc["ext"], ids, kps, sc["X"])
    cam, pt, obs = G.enumerate_observations(ids, kps)
    rr = G.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam, pt, obs)
    assert r.shape == rr.shape and np.abs(r - rr).max() <= REL_TOL * max(1.0, np.abs(rr).max())
    c = G.huber_cost(rr, 4.0)
    assert abs(cost - c) <= 1e-9 * c and abs(rmse - np.sqrt(c / len(cam))) <= 1e-9


def test_jacobians_vs_oracle(ctx):
    """sfm_reproject_jacobians: what Ceres' autodiff derives from ReprojectCost (:1202)."""
    sc = synth.scene(3000, 4, seed=21)
    ext = sc["ext"].copy()
    ext[0, :3] = [1e-10, -2e-10, 3e-11]                       # small-angle branch camera
    cam = np.repeat(np.arange(4), 3000).astype(np.int32)
    pt = np.tile(np.arange(3000), 4).astype(np.int32)
    perm = np.random.default_rng(0).permutation(cam.size)[:11111]     # ragged count, mixed order
    cam, pt = cam[perm], pt[perm]
    obs = sc["xy"].reshape(-1, 2)[perm]
    r, J = ctx.reproject_jacobians(sc["intr"], ext, sc["X"], cam, pt, obs)
    ref_r = G.reproject_residuals(sc["intr"], ext, sc["X"], cam, pt, obs)
    ref_J = G.reproject_jacobians(sc["intr"], ext, sc["X"], cam, pt)
    assert J.shape == (11111, 2, 13)
    assert np.abs(r - ref_r).max() < REL_TOL * 1000.0
    # Jacobian entries span 1e-3 .. 1e4: relative to the row's largest entry
    scale = np.abs(ref_J).max(axis=2, keepdims=True)
    assert (np.abs(J - ref_J) / scale).max() < 1e-9
    # structural zeros and ones are exact
    assert np.array_equal(J[:, 0, [1, 3, 8]], np.zeros((11111, 3))) and np.array_equal(J[:, 0, 2], np.ones(11111))


def test_jacobians_single_observation_and_timed(ctx):
    sc = synth.scene(10, 2, seed=4)
    cam = np.array([1], np.int32); pt = np.array([7], np.int32)
    r, J = ctx.reproject_jacobians(sc["intr"], sc["ext"], sc["X"], cam, pt, sc["xy"][1, 7:8])
    ref = G.reproject_jacobians(sc["intr"], sc["ext"], sc["X"], cam, pt)
    assert np.allclose(J, ref, rtol=1e-10, atol=1e-12)
    cam = np.repeat(np.arange(2), 10).astype(np.int32); pt = np.tile(np.arange(10), 2).astype(np.int32)
    r2, J2, ms = ctx.reproject_jacobians(sc["intr"], sc["ext"], sc["X"], cam, pt, sc["xy"].reshape(-1, 2), iters=3)
    assert ms > 0 and np.allclose(J2, G.reproject_jacobians(sc["intr"], sc["ext"], sc["X"], cam, pt), rtol=1e-10, atol=1e-12)


def test_normals_reproduce_the_bundled_ply(ctx):
    """sfm_estimate_normals on the points of the reference's bundled structure_ba.yml must give
    the normals stored in its bundled structure_ba.ply (float32)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "viewer_outputs.npz"))
    X, v = g["structure_ba_yml_X"], g["structure_ba_ply_v"]
    n = ctx.estimate_normals(X, 10)
    assert n.shape == X.shape and np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-12)
    assert np.abs(n.astype(np.float32) - v[:, 3:]).max() <= 1e-6
    assert np.abs(n - G.estimate_normals(X, 10)).max() < 1e-9


@pytest.mark.parametrize("n,K", [(11, 10), (300, 3), (1000, 16), (4097, 10)])
def test_normals_vs_oracle(ctx, n, K):
    rng = np.random.default_rng(n + K)
    # points near a wavy surface in front of the camera: well-defined tangent planes
    u = rng.uniform(-3, 3, (n, 2))
    X = np.stack([u[:, 0], u[:, 1], 8 + 0.3 * np.sin(u[:, 0]) + 0.2 * np.cos(2 * u[:, 1]) + rng.normal(0, 0.01, n)], 1)
    got = ctx.estimate_normals(X, K)
    ref = G.estimate_normals(X, K)
    assert np.abs(got - ref).max() < 1e-7
    assert ((got * X).sum(1) < 0).mean() > 0.99           # oriented towards the camera


def test_normals_errors(ctx):
    import sfm_opencv_b200 as sfm
    with pytest.raises(sfm.SfmError):
        ctx.estimate_normals(np.zeros((10, 3)), 10)       # needs more than K points
    with pytest.raises(sfm.SfmError):
        ctx.estimate_normals(np.zeros((100, 3)), 40)


def test_ba_problem_handle_equals_one_shot_calls(ctx):
    """The BA-loop form (sfm_ba_create / sfm_ba_evaluate): observation tables resident, only
    cameras and points move per evaluation -- same residuals, Jacobians and Huber cost as the
    one-shot entry points, also after the parameters changed (bundle_adjustment() iterates,
    NViewReconstuct.cpp:1224)."""
    import sfm_opencv_b200 as sfm
    sc = synth.scene(6000, 3, seed=17)
    cam, pt = synth.observations_camera_major(6000, 3)
    obs = sc["xy"].reshape(-1, 2)
    pb = sfm.BAProblem(ctx, 3, 6000, cam, pt, obs)
    rng = np.random.default_rng(1)
    ext, X = sc["ext"].copy(), sc["X"].copy()
    for it in range(3):
        r, J, cost = pb.evaluate(sc["intr"], ext, X, want_jac=True)
        r1, c1 = ctx.reproject_residuals(sc["intr"], ext, X, cam, pt, obs)
        _, J1 = ctx.reproject_jacobians(sc["intr"], ext, X, cam, pt, obs, want_resid=False)
        assert np.array_equal(r, r1) and np.array_equal(J, J1) and cost == c1
        assert _resid_close(r, G.reproject_residuals(sc["intr"], ext, X, cam, pt, obs))
        if it < 2:
            ext[1:] += rng.normal(0, 1e-3, ext[1:].shape)
            X += rng.normal(0, 1e-3, X.shape)
    # only the points move: the handle keeps the cameras of the previous evaluation
    X2 = X + 0.01
    r2, _, c2 = pb.evaluate(sc["intr"], None, X2)
    r3, c3 = ctx.reproject_residuals(sc["intr"], ext, X2, cam, pt, obs)
    assert np.array_equal(r2, r3) and c2 == c3
    # cost only (what an LM step-acceptance test needs): nothing but 8 bytes comes back
    _, _, c4 = pb.evaluate(sc["intr"], None, None, want_resid=False)
    assert c4 == c3
    pb.close()


def test_observation_indices_are_checked_on_the_device(ctx):
    import sfm_opencv_b200 as sfm
    from sfm_opencv_b200 import _capi
    sc = synth.scene(100, 2, seed=2)
    cam, pt = synth.observations_camera_major(100, 2)
    obs = sc["xy"].reshape(-1, 2)
    for bad_cam, bad_pt in ((2, 0), (-1, 0), (0, 100), (0, -5)):
        c2, p2 = cam.copy(), pt.copy()
        c2[57], p2[57] = bad_cam, bad_pt
        with pytest.raises(sfm.SfmError) as e:
            ctx.reproject_residuals(sc["intr"], sc["ext"], sc["X"], c2, p2, obs)
        assert e.value.code == _capi.SFM_E_INVALID
        with pytest.raises(sfm.SfmError) as e:
            ctx.reproject_jacobians(sc["intr"], sc["ext"], sc["X"], c2, p2, obs)
        assert e.value.code == _capi.SFM_E_INVALID
        with pytest.raises(sfm.SfmError) as e:
            sfm.BAProblem(ctx, 2, 100, c2, p2, obs)
        assert e.value.code == _capi.SFM_E_INVALID
