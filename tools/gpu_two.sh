#!/bin/bash
# Two-GPU visit: all GPU tests (none skipped: two devices in one process, two-process push exchange),
# the datasets workload, and the bench at N = 2 as the driver launches it.
set -u
mkdir -p gpurun_out
TAG=${1:-two}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --workload datasets --steps 5 > gpurun_out/${TAG}_datasets.json 2> gpurun_out/${TAG}_datasets.err; echo "datasets rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_datasets.json').read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], {k:(v["gpu_e2e_ms"], v["match_lists_equal_cv2"]) for k,v in d["datasets"].items()})
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 2500 gpurun_out/${TAG}_bench_2gpu.json | cut -c1-2500
tail -5 gpurun_out/${TAG}_bench_2gpu.err
