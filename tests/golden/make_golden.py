"""Generates the committed golden fixtures from the reference's bundled datasets.

Run HERE (container with /root/reference and cv2 4.13); the fixtures travel, this script's
inputs do not.  For each dataset: SIFT_create(0,3,0.04,10) descriptors (the parameters of
OpenCV_SFM/TwoViewReconstruct.cpp:112), then for every consecutive pair
(NViewReconstuct.cpp:857-870) the reference's own library call
cv2.BFMatcher(NORM_L2).knnMatch(k=2) and the filter of NViewReconstuct.cpp:880-908.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import glob
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import matching as M  # noqa: E402

REF = "/root/reference/dataset"
OUT = os.path.dirname(os.path.abspath(__file__))


def sift_dataset(name: str):
    files = sorted(glob.glob(os.path.join(REF, name, "*.[jJ][pP][gG]")))
    sift = cv2.SIFT_create(0, 3, 0.04, 10)
    descs, kps = [], []
    for f in files:
        img = cv2.imread(f)
        kp = sift.detect(img, None)
        kp, d = sift.compute(img, kp)
        descs.append(M.as_u8(d))
        kps.append(np.array([k.pt for k in kp], np.float32))
        print(name, os.path.basename(f), d.shape, flush=True)
    return files, descs, kps


def akaze_dataset(name: str):
    """The LIVE path's features: AKAZE_create() + detect/compute (NViewReconstuct.cpp:797-814)."""
    files = sorted(glob.glob(os.path.join(REF, name, "*.[jJ][pP][gG]")))
    ak = cv2.AKAZE_create()
    descs, kps = [], []
    for f in files:
        img = cv2.imread(f)
        kp = ak.detect(img, None)
        kp, d = ak.compute(img, kp)
        descs.append(np.ascontiguousarray(d, np.uint8))
        kps.append(np.array([k.pt for k in kp], np.float32))
        print(name, "akaze", os.path.basename(f), d.shape, flush=True)
    return files, descs, kps


def main_akaze():
    """desktop / AKAZE / NORM_HAMMING2: the configuration that produced the reference's bundled
    Viewer/structure.yml (expected match counts 2186 / 1063 / 230 / 553)."""
    name = "desktop"
    files, descs, kps = akaze_dataset(name)
    out = {"n_img": np.int32(len(descs))}
    for i, (d, k) in enumerate(zip(descs, kps)):
        out[f"desc_{i}"] = d
        out[f"kp_{i}"] = k
    for i in range(len(descs) - 1):
        dist, idx = M.knn2_cv_hamming2(descs[i], descs[i + 1])   # the reference's library call
        m, d0, md = M.filter_matches(dist, idx)
        out[f"knn_dist_{i}"] = dist
        out[f"knn_idx_{i}"] = idx
        out[f"match_{i}"] = m
        out[f"match_dist_{i}"] = d0
        out[f"min_dist_{i}"] = np.float32(md)
        print(name, "akaze pair", i, "matches", len(m), "min_dist", md, "tie rows",
              int((dist[:, 0] == dist[:, 1]).sum()), flush=True)
    np.savez_compressed(os.path.join(OUT, f"{name}_akaze.npz"), **out)


def main():
    for name in ("crazyhorse", "desktop"):
        files, descs, kps = sift_dataset(name)
        out = {"n_img": np.int32(len(descs))}
        for i, (d, k) in enumerate(zip(descs, kps)):
            out[f"desc_{i}"] = d
            out[f"kp_{i}"] = k
        for i in range(len(descs) - 1):
            q = descs[i].astype(np.float32)
            t = descs[i + 1].astype(np.float32)
            dist, idx = M.knn2_cv(q, t)                      # the reference's library call
            m, d0, md = M.filter_matches(dist, idx)
            out[f"knn_dist_{i}"] = dist
            out[f"knn_idx_{i}"] = idx
            out[f"match_{i}"] = m
            out[f"match_dist_{i}"] = d0
            out[f"min_dist_{i}"] = np.float32(md)
            ties = int((dist[:, 0] == dist[:, 1]).sum())
            print(name, "pair", i, "matches", len(m), "min_dist", md, "tie rows", ties, flush=True)
        np.savez_compressed(os.path.join(OUT, f"{name}_sift.npz"), **out)


def knn_digest(dist, idx) -> np.ndarray:
    """sha256 over the raw kNN rows (idx int32 [n,2] then dist float32 [n,2]) as 32 uint8: lets a
    fixture pin EVERY kNN row of a large pair without storing 16 bytes per query."""
    import hashlib
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(idx, np.int32).tobytes())
    h.update(np.ascontiguousarray(dist, np.float32).tobytes())
    return np.frombuffer(h.digest(), np.uint8).copy()


def main_dog():
    """dataset/dog (16 images, 4.6k-21.8k SIFT descriptors each): descriptors + keypoints, and
    for the 15 consecutive pairs the filtered match list, min_dist and a digest of all kNN rows."""
    name = "dog"
    files, descs, kps = sift_dataset(name)
    out = {"n_img": np.int32(len(descs))}
    for i, d in enumerate(descs):
        out[f"desc_{i}"] = d                                 # keypoints are not needed for matching
    for i in range(len(descs) - 1):
        dist, idx = M.knn2_cv(descs[i].astype(np.float32), descs[i + 1].astype(np.float32))
        m, d0, md = M.filter_matches(dist, idx)
        out[f"knn_sha_{i}"] = knn_digest(dist, idx)
        out[f"match_{i}"] = m
        out[f"match_dist_{i}"] = d0
        out[f"min_dist_{i}"] = np.float32(md)
        print(name, "pair", i, "matches", len(m), "min_dist", md, "tie rows",
              int((dist[:, 0] == dist[:, 1]).sum()), flush=True)
    np.savez_compressed(os.path.join(OUT, f"{name}_sift.npz"), **out)


def main_all_pairs():
    """Exhaustive pair lists (i < j; north_star's schedule) on the two small bundled datasets:
    match list + kNN digest per pair from the reference's library call (descriptors are already
    in <name>_sift.npz)."""
    for name in ("crazyhorse", "desktop"):
        g = np.load(os.path.join(OUT, f"{name}_sift.npz"))
        n = int(g["n_img"])
        out = {}
        p = 0
        for a in range(n):
            for b in range(a + 1, n):
                dist, idx = M.knn2_cv(g[f"desc_{a}"].astype(np.float32), g[f"desc_{b}"].astype(np.float32))
                m, d0, md = M.filter_matches(dist, idx)
                out[f"pair_{p}"] = np.array([a, b], np.int32)
                out[f"knn_sha_{p}"] = knn_digest(dist, idx)
                out[f"match_{p}"] = m
                out[f"match_dist_{p}"] = d0
                out[f"min_dist_{p}"] = np.float32(md)
                p += 1
        out["n_pairs"] = np.int32(p)
        print(name, "all pairs", p, flush=True)
        np.savez_compressed(os.path.join(OUT, f"{name}_allpairs.npz"), **out)


if __name__ == "__main__":
    if "--dog" in sys.argv:
        main_dog()
    elif "--all-pairs" in sys.argv:
        main_all_pairs()
    else:
        if "--akaze-only" not in sys.argv:
            main()
        main_akaze()
