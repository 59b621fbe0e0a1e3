#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-prune}
timeout 900 python -m pytest tests/test_gpu_matching.py tests/test_gpu_datasets.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
for G in 0 1; do
echo "SFM_PRUNE_MODE=$G"
SFM_PRUNE_MODE=$G python tools/exp_one.py 1 24 8192
SFM_PRUNE_MODE=$G python tools/exp_one.py 1 2 65536
done
python bench.py --steps 10 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['self_check'], d['clocks']['sm_mhz'])"
