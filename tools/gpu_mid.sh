#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-mid}
timeout 600 python bench.py --workload datasets --steps 3 > gpurun_out/${TAG}_datasets.json 2> gpurun_out/${TAG}_datasets.err; echo "datasets rc=$?"
timeout 600 python bench.py --steps 3 --no-self-check --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
SMALL="python bench.py --steps 2 --warmup 3 --images 24 --no-extras --no-cpu-baseline --no-self-check"
timeout 300 $SMALL > gpurun_out/${TAG}_plain2.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn2 -s 3 -c 1 -f -o gpurun_out/${TAG}_knn2 $SMALL > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu full rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_datasets.json").read().strip().splitlines()[-1])
print("datasets", d["value"], d["cpu_baseline"]["value"], {k:(round(v["gpu_e2e_ms"],2), round(v["cpu_ms"],1), v["match_lists_equal_cv2"], v["match_lists_equal_golden"]) for k,v in d["datasets"].items()})
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
e=d["extra"]
for k in ("triangulate_4M_v2","triangulate_4M_v8"):
    print(k, e[k]["ms"], e[k].get("fp64"), e[k]["roofline"]["frac"])
print("value", d["value"], "e2e", d["e2e"]["value"], d["roofline"]["frac"])
PY
