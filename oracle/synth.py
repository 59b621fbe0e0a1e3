"""Seeded synthetic inputs of BASELINE.json configs 3-5.

Input generators only (no reference arithmetic): usable by tests and bench.py.
"""
from __future__ import annotations

import numpy as np

from .geometry import K_REFERENCE, angle_axis_rotate, build_projection


def sift_like(n: int, seed: int, dim: int = 128) -> np.ndarray:
    """SIFT-like uint8 rows: gamma(0.6) -> L2-normalise -> clip 0.2 -> renormalise ->
    rint(512*x) clipped to 0..255.  Row norm^2 ~ 262 144, max ~ 186."""
    rng = np.random.default_rng(seed)
    x = rng.gamma(0.6, 1.0, size=(n, dim)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    np.minimum(x, 0.2, out=x)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.clip(np.rint(512.0 * x), 0, 255).astype(np.uint8)


def image_bank(n_img: int, n_desc: int, shared_frac: float = 0.2, noise: int = 2,
               seed0: int = 0) -> list[np.ndarray]:
    """n_img descriptor sets; image j>0 re-uses a random ``shared_frac`` of image j-1's rows
    with small integer noise, so that the ratio test passes for some rows."""
    bank = []
    for j in range(n_img):
        d = sift_like(n_desc, seed0 + j)
        if j > 0 and shared_frac > 0:
            rng = np.random.default_rng(10_000 + seed0 + j)
            k = int(n_desc * shared_frac)
            src = rng.choice(n_desc, k, replace=False)
            dst = rng.choice(n_desc, k, replace=False)
            nz = rng.integers(-noise, noise + 1, size=(k, d.shape[1]))
            d[dst] = np.clip(bank[j - 1][src].astype(np.int32) + nz, 0, 255).astype(np.uint8)
        bank.append(d)
    return bank


def scene(n_pts: int, n_views: int, seed: int = 7, noise_px: float = 0.5):
    """Config 5: points ~ U([-4,4]x[-3,3]x[6,12]); cam 0 = identity, others angle-axis ~
    N(0,0.2^2), t ~ N(0,1); K of NViewReconstuct.cpp:1353-1356; observations = projection +
    N(0, noise_px) stored as float32; P = f32(K) f32([R|t]).

    Returns dict(P [V,3,4] f32, xy [V,N,2] f32, X [N,3] f64, ext [V,6] f64, intr [4] f64).
    """
    import cv2
    rng = np.random.default_rng(seed)
    X = np.empty((n_pts, 3), np.float64)
    X[:, 0] = rng.uniform(-4, 4, n_pts)
    X[:, 1] = rng.uniform(-3, 3, n_pts)
    X[:, 2] = rng.uniform(6, 12, n_pts)
    ext = np.zeros((n_views, 6), np.float64)
    ext[1:, :3] = rng.normal(0, 0.2, (n_views - 1, 3))
    ext[1:, 3:] = rng.normal(0, 1.0, (n_views - 1, 3))
    K = K_REFERENCE
    intr = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], np.float64)
    P = np.empty((n_views, 3, 4), np.float32)
    xy = np.empty((n_views, n_pts, 2), np.float32)
    for v in range(n_views):
        R, _ = cv2.Rodrigues(ext[v, :3].reshape(3, 1))
        P[v] = build_projection(K, R, ext[v, 3:])
        p = angle_axis_rotate(ext[v, :3][None, :], X) + ext[v, 3:][None, :]
        u = intr[0] * p[:, 0] / p[:, 2] + intr[2]
        w = intr[1] * p[:, 1] / p[:, 2] + intr[3]
        xy[v, :, 0] = (u + rng.normal(0, noise_px, n_pts)).astype(np.float32)
        xy[v, :, 1] = (w + rng.normal(0, noise_px, n_pts)).astype(np.float32)
    return dict(P=P, xy=xy, X=X, ext=ext, intr=intr)


def observations_camera_major(n_pts: int, n_views: int):
    """Camera-major observation order (NViewReconstuct.cpp:1187-1211): every camera sees
    every point."""
    cam = np.repeat(np.arange(n_views, dtype=np.int32), n_pts)
    pt = np.tile(np.arange(n_pts, dtype=np.int32), n_views)
    return cam, pt
