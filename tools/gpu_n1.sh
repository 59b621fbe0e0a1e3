#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-n1}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "kernel ms", d["roofline"]["kernel_ms_per_step"], d["self_check"], d["clocks"])
PY
