import os, sys, json
sys.path.insert(0, '/root/repo')
import numpy as np
import sfm_opencv_b200 as sfm
from oracle import synth
bank = synth.image_bank(16, 8192)
pairs = [(i, j) for i in range(16) for j in range(i + 1, 16)]
out = {}
for mode in (1, 0, 3, 4):
    os.environ["SFM_KNN_MODE"] = str(mode)
    with sfm.Context(0) as c:
        c.upload_descriptors(bank)
        c.match_pairs_resident(pairs)
        best = min(c.match_pairs_resident(pairs)[1] for _ in range(3))
        out[f"mode{mode & 15}_dbg{mode >> 4}_tops"] = 2.0 * 8192 * 8192 * 128 * len(pairs) / (best * 1e-3) / 1e12
print(json.dumps(out))
