"""compute-sanitizer target: one small pass of every kernel family (kNN-2 in both exact modes and
the match-only mode, filter, HAMMING2, triangulation incl. the slow path, residuals, Jacobians,
normals).  Small on purpose: the sanitizer slows the warp-specialised kNN kernel by orders of magnitude."""
import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
import sfm_opencv_b200 as sfm
from oracle import synth
with sfm.Context(0) as c:
    bank = synth.image_bank(3, 600, seed0=5)
    c.upload_descriptors(bank)
    m, md, knn = c.match_pairs([(0, 1), (1, 2), (0, 2)], want_knn=True)
    m2, _, _ = c.match_pairs([(0, 1), (1, 2), (0, 2)])
    assert m.flat.tobytes() == m2.flat.tobytes()
    rng = np.random.default_rng(0)
    c.upload_descriptors([rng.integers(0, 256, (300, 61), dtype=np.uint8) for _ in range(2)], norm="hamming2")
    c.match_pairs([(0, 1)])
    sc = synth.scene(3000, 3, seed=1)
    xy = sc["xy"].copy(); xy[2, ::7] += 300.0            # mismatched rays: the slow path runs
    c.triangulate_batch(sc["P"], xy)
    cam, pt = synth.observations_camera_major(3000, 3)
    c.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam, pt, sc["xy"].reshape(-1, 2))
    c.reproject_jacobians(sc["intr"], sc["ext"], sc["X"], cam, pt, sc["xy"].reshape(-1, 2))
    c.estimate_normals(sc["X"][:500])
print("sanitize_small ok")
