#!/bin/bash
# All GPU tests, the datasets workload with and without the ratio-driven sweep, and its kernel launch list.
set -u
mkdir -p gpurun_out
TAG=${1:-ds}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_tests.log
for M in 0 1; do
SFM_PRUNE_MODE=$M timeout 600 python bench.py --workload datasets --steps 5 > gpurun_out/${TAG}_datasets_p$M.json 2> gpurun_out/${TAG}_datasets_p$M.err; echo "datasets prune=$M rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_datasets_p$M.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], {k:(v["gpu_e2e_ms"], v["match_lists_equal_cv2"]) for k,v in d["datasets"].items()})
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --workload datasets --steps 1 --warmup 1 > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
python tools/exp_one.py 1 24 8192
