"""4M-point triangulation + 8M-observation residual / Jacobian launches (BASELINE config 5), for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sfm_opencv_b200 as sfm
from oracle import synth
ctx = sfm.Context(0)
sc = synth.scene(4_000_000, 2)
cam, pt = synth.observations_camera_major(4_000_000, 2)
for _ in range(3):
    _, _, ms = ctx.triangulate_batch(sc["P"], sc["xy"], want_X4=True, want_xyz=False, iters=1)
    _, _, ms2 = ctx.reproject_residuals(sc["intr"], sc["ext"], sc["X"], cam, pt,
                                        sc["xy"].reshape(-1, 2), want_cost=False, iters=1)
    _, _, ms3 = ctx.reproject_jacobians(sc["intr"], sc["ext"], sc["X"], cam, pt,
                                        sc["xy"].reshape(-1, 2), want_resid=False, iters=1)
print("tri ms", ms, "resid ms", ms2, "jacobian ms", ms3)
