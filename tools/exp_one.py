import os, sys, json
sys.path.insert(0, '/root/repo')
import numpy as np
import sfm_opencv_b200 as sfm
from oracle import synth
mode = int(sys.argv[1]); n_img = int(sys.argv[2]); n_desc = int(sys.argv[3])
os.environ["SFM_KNN_MODE"] = str(mode)
bank = synth.image_bank(n_img, n_desc)
pairs = [(i, j) for i in range(n_img) for j in range(i + 1, n_img)]
with sfm.Context(0) as c:
    c.upload_descriptors(bank)
    c.match_pairs_resident(pairs)
    best = min(c.match_pairs_resident(pairs)[1] for _ in range(3))
    print(json.dumps({"mode": mode, "n_img": n_img, "n_desc": n_desc,
                      "tops": 2.0 * n_desc * n_desc * 128 * len(pairs) / (best * 1e-3) / 1e12}))
