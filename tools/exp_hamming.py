"""GPU box: time the NORM_HAMMING2 path on the AKAZE fixture (bundled desktop dataset) and on a
synthetic 16-image all-pairs bank, next to cv2.batchDistance on the host cores."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sfm_opencv_b200 as sfm
from oracle import matching as M

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "desktop_akaze.npz"))
bank = [g[f"desc_{i}"] for i in range(int(g["n_img"]))]
out = {}
with sfm.Context(0) as c:
    c.upload_descriptors(bank, norm="hamming2")
    pairs = M.consecutive_pairs(len(bank))
    c.match_pairs_resident(pairs)
    best = min(c.match_pairs_resident(pairs)[1:] for _ in range(5))
    cells = sum(bank[a].shape[0] * bank[b].shape[0] for a, b in pairs)
    out["desktop_akaze_4pairs"] = {"knn_ms": best[0], "total_ms": best[1], "descriptor_pairs": cells,
                                   "gpairs_per_s": cells / best[0] / 1e6}
    rng = np.random.default_rng(0)
    syn = [rng.integers(0, 256, (8192, 61), dtype=np.uint8) for _ in range(16)]
    c.upload_descriptors(syn, norm="hamming2")
    ap = M.all_pairs(16)
    c.match_pairs_resident(ap)
    best = min(c.match_pairs_resident(ap)[1:] for _ in range(3))
    cells = len(ap) * 8192 * 8192
    out["synthetic_16x8192_allpairs"] = {"knn_ms": best[0], "total_ms": best[1], "image_pairs_per_s": len(ap) / best[1] * 1e3,
                                         "gpairs_per_s": cells / best[0] / 1e6}
t0 = time.time()
for a, b in M.consecutive_pairs(len(bank)):
    M.knn2_cv_hamming2(bank[a], bank[b])
out["cpu_cv2_desktop_4pairs_ms"] = (time.time() - t0) * 1e3
out["cpu_cores"] = os.cpu_count()
print(json.dumps(out))
