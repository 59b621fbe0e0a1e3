#!/bin/bash
# Quick visit: all GPU tests, the datasets workload and the 24-image timing.
set -u
mkdir -p gpurun_out
TAG=${1:-q2}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --workload datasets --steps 5 > gpurun_out/${TAG}_datasets.json 2> gpurun_out/${TAG}_datasets.err; echo "datasets rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_datasets.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], {k:(v["gpu_e2e_ms"], v["match_lists_equal_cv2"]) for k,v in d["datasets"].items()})
PY
python tools/exp_one.py 1 24 8192
python tools/exp_one.py 1 2 65536
