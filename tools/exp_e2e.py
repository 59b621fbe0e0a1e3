"""GPU box: where the end-to-end step goes (host wall clock per call)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sfm_opencv_b200 as sfm
from oracle import synth
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 200
bank = synth.image_bank(n_img, 8192)
pairs = [(i, j) for i in range(n_img) for j in range(i + 1, n_img)]
out = {}
with sfm.Context(0) as c:
    host = []
    for k, b in enumerate(bank):
        a = c.pinned_empty(b.shape, np.float32, f"d{k}"); a[...] = b; host.append(a)
    for overlap in (False, True):
        for rep in range(3):
            t0 = time.perf_counter()
            c.upload_descriptors(host, overlap=overlap)
            t1 = time.perf_counter()
            m, _, _ = c.match_pairs(pairs, copy=False)
            t2 = time.perf_counter()
        out["overlap" if overlap else "sync"] = {"upload_ms": (t1 - t0) * 1e3, "match_ms": (t2 - t1) * 1e3,
                                                   "total_ms": (t2 - t0) * 1e3}
    c.upload_descriptors(host)
    tot, kms, tms = c.match_pairs_resident(pairs)
    out["resident"] = {"knn_ms": kms, "device_total_ms": tms}
print(json.dumps(out))
