"""GPU: the C++ host layer (include/sfm_b200.hpp, the reference's own function shapes) driven by a
C++ program (examples/pipeline_check.cpp) -- match_features_for_all, get_matched_points,
reconstruct (host arrays and device-resident), residual blocks, save_structure -- checked against
the CPU oracle: match lists bit-exact, geometry 1e-5."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from oracle import geometry as G
from oracle import matching as M
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MATCH = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])


def _build(tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("g++ not installed")
    exe = str(tmp_path / "pipeline_check")
    libdir = os.path.join(ROOT, "sfm_opencv_b200")
    subprocess.run(["g++", "-std=c++11", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "pipeline_check.cpp"), "-o", exe, "-L", libdir,
                    "-l:libsfm_b200.so", "-Wl,-rpath," + libdir], check=True)
    return exe


@pytest.mark.gpu
def test_cpp_layer_pipeline(tmp_path):
    exe = _build(tmp_path)
    # a small scene: 3 images whose descriptors share planted matches; keypoints = projections of
    # common 3-D points in cameras 0 and 1 (so that pair 0's matches triangulate properly)
    rng = np.random.default_rng(3)
    sizes = [900, 700, 500]
    sc = synth.scene(400, 2, seed=5)
    bank = [synth.sift_like(n, 60 + i) for i, n in enumerate(sizes)]
    bank[1][:400] = bank[0][:400]                               # 400 exact matches 0 -> 1
    bank[2][:200] = bank[1][300:500]
    kps = [rng.uniform(0, 3000, (n, 2)).astype(np.float32) for n in sizes]
    kps[0][:400] = sc["xy"][0]
    kps[1][:400] = sc["xy"][1]
    K = G.K_REFERENCE
    import cv2
    ext = sc["ext"][:2]                                         # the scene's cameras: (angle-axis, t)
    R1, _ = cv2.Rodrigues(ext[0, :3].reshape(3, 1))
    R2, _ = cv2.Rodrigues(ext[1, :3].reshape(3, 1))
    T1, T2 = ext[0, 3:], ext[1, 3:]
    blob = struct.pack("<i", 3) + np.array(sizes, "<i4").tobytes()
    for d, k in zip(bank, kps):
        blob += d.astype("<f4").tobytes() + k.astype("<f4").tobytes()
    for a in (K, R1, T1, R2, T2, ext):
        blob += np.asarray(a, "<f8").tobytes()
    (tmp_path / "in.bin").write_bytes(blob)
    out = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin"), str(tmp_path / "s.yml")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    raw = (tmp_path / "out.bin").read_bytes()
    pos = 0

    def take(dtype, count=None):
        nonlocal pos
        if count is None:
            count = struct.unpack_from("<q", raw, pos)[0]
            pos += 8
        a = np.frombuffer(raw, dtype, count, pos)
        pos += a.nbytes
        return a
    # match_features_for_all: bit-exact with the oracle
    m0 = None
    for p in range(2):
        m = take(MATCH)
        om, od, _, _, _ = M.match_features(bank[p], bank[p + 1])
        assert np.array_equal(m["queryIdx"], om[:, 0]) and np.array_equal(m["trainIdx"], om[:, 1])
        assert np.array_equal(m["distance"].view(np.uint32), od.view(np.uint32)) and (m["imgIdx"] == 0).all()
        m0 = m if p == 0 else m0
    assert len(m0) >= 400
    # reconstruct on pair 0: host-array path and device-resident path agree with the oracle
    p1, p2 = kps[0][m0["queryIdx"]], kps[1][m0["trainIdx"]]
    ref, _ = G.reconstruct(K, R1, T1, R2, T2, p1, p2)
    n_pts = struct.unpack_from("<q", raw, pos)[0]; pos += 8
    xyz_host = take("<f8", 3 * n_pts).reshape(-1, 3)
    n_dev = struct.unpack_from("<q", raw, pos)[0]; pos += 8
    xyz_dev = take("<f8", 3 * n_dev).reshape(-1, 3)
    good = np.isfinite(ref).all(1) & (np.abs(ref).max(1) < 100)   # planted matches; stray ones may be ill-posed
    assert good.sum() >= 400 and xyz_host.shape == ref.shape == xyz_dev.shape
    assert G.point_rel_err(xyz_host[good], ref[good]).max() < 1e-5
    assert np.array_equal(xyz_host, xyz_dev)
    # residual blocks + Huber cost
    n_res = struct.unpack_from("<q", raw, pos)[0]; pos += 8
    resid = take("<f8", n_res).reshape(-1, 2)
    cost = take("<f8", 1)[0]
    intr = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]])
    n = len(m0)
    cam = np.repeat(np.arange(2), n).astype(np.int32); pt = np.tile(np.arange(n), 2).astype(np.int32)
    obs = np.concatenate([p1, p2])
    rr = G.reproject_residuals(intr, ext, xyz_host, cam, pt, obs)
    ok = np.tile(good, 2)
    assert np.abs(resid[ok] - rr[ok]).max() < 1e-5 * max(1.0, np.abs(rr[ok]).max())
    assert np.isfinite(cost)
    # save_structure wrote the viewer file
    txt = (tmp_path / "s.yml").read_text()
    assert txt.startswith("%YAML:1.0\n---\nCamera Count: 2\nPoint Count: " + str(n) + "\n")
    assert txt.count("!!opencv-matrix") == 4 and txt.rstrip().endswith("- [ 128, 128, 128 ]")


def test_cpp_layer_compiles_without_a_gpu(tmp_path):
    """CPU: the header-only layer and its example build with -Wall -Werror and fail loudly
    (SFM_E_NO_DEVICE) where there is no B200 -- there is no CPU fallback behind it."""
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by test_cpp_layer_pipeline")
    blob = struct.pack("<i", 2) + np.array([3, 3], "<i4").tobytes()
    for _ in range(2):
        blob += np.zeros((3, 128), "<f4").tobytes() + np.zeros((3, 2), "<f4").tobytes()
    blob += np.zeros(9 + 9 + 3 + 9 + 3 + 12, "<f8").tobytes()
    (tmp_path / "in.bin").write_bytes(blob)
    out = subprocess.run([exe, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert out.returncode == 3 and "sfm_b200 error -2" in out.stderr
