#!/bin/bash
# Last visit of the round: all GPU tests, the datasets workload, the full bench and the reference arm.
set -u
mkdir -p gpurun_out
TAG=${1:-last}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python bench.py --workload datasets --steps 5 > gpurun_out/${TAG}_datasets.json 2> gpurun_out/${TAG}_datasets.err; echo "datasets rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_datasets.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], {k:(v["gpu_e2e_ms"], v["match_lists_equal_cv2"]) for k,v in d["datasets"].items()})
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"], d["self_check"]["all_ranks_ok"], d["extra"]["pair_65536"]["tops"])
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
