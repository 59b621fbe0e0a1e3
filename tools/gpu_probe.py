"""First-contact GPU script: tensor peak probe + a timed 8192x8192 all-pairs slice."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfm_opencv_b200 as sfm  # noqa: E402
from oracle import synth  # noqa: E402

ctx = sfm.Context(0)
out = {}
for it in (500, 4000):
    out[f"i8_peak_tops_{it}"] = ctx.probe_i8_peak(it)
n_img = int(os.environ.get("N_IMG", 12))
bank = synth.image_bank(n_img, 8192)
t = time.time(); ctx.upload_descriptors(bank); out["upload_s"] = time.time() - t
pairs = [(i, j) for i in range(n_img) for j in range(i + 1, n_img)]
for rep in range(3):
    tot, kms, tms = ctx.match_pairs_resident(pairs)
    ops = 2.0 * 8192 * 8192 * 128 * len(pairs)
    out[f"rep{rep}"] = dict(pairs=len(pairs), matches=tot, knn_ms=kms, total_ms=tms,
                            pairs_per_s=len(pairs) / (kms * 1e-3), tops=ops / (kms * 1e-3) / 1e12)
sc = synth.scene(4_000_000, 2)
_, _, ms = ctx.triangulate_batch(sc["P"], sc["xy"], iters=20)
out["tri_ms_4M_v2"] = ms
cam, pt = synth.observations_camera_major(4_000_000, 2)
_, _, ms = ctx.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam, pt, sc["xy"].reshape(-1, 2), iters=20)
out["resid_ms_8M"] = ms
print(json.dumps(out, indent=1))
