import sys, json
sys.path.insert(0, "/root/repo")
import sfm_opencv_b200 as sfm
out = {}
with sfm.Context(0) as c:
    out["n256"] = c.probe_i8_peak(4000)
    for v in range(5):
        out[f"n128_variant{v}"] = c.probe_i8_peak(-((4000 << 4) | v))
print(json.dumps(out))
