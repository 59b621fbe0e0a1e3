"""Development tool (GPU box): residual kernel, camera-major (chunk x camera order, L2 reuse of the
point table) against list order (forced by swapping two cameras' first entries), 4M points x V views."""
import json, sys
sys.path.insert(0, '/root/repo')
import numpy as np
import sfm_opencv_b200 as sfm
from oracle import synth
n = 4_000_000
with sfm.Context(0) as c:
    for V in (2, 4, 8):
        sc = synth.scene(n, V)
        cam, pt = synth.observations_camera_major(n, V)
        obs = sc["xy"].reshape(-1, 2)
        out = {"V": V}
        for name in ("camera_major", "list_order"):
            cm = cam.copy()
            if name == "list_order":
                cm[0], cm[n] = cm[n], cm[0]      # no longer sorted: the kernel walks the list as given
                ob = obs.copy(); ob[[0, n]] = ob[[n, 0]]
                p2 = pt.copy(); p2[[0, n]] = p2[[n, 0]]
            else:
                ob, p2 = obs, pt
            _, _, ms = c.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cm, p2, ob, want_cost=False, iters=20)
            b = 32 * n * V + 24 * n
            out[name] = {"ms": ms, "GBs": b / (ms * 1e-3) / 1e9}
        print(json.dumps(out), flush=True)
