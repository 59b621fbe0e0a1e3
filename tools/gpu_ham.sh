#!/bin/bash
# tensor-core HAMMING2: parity tests on both kernels, then timings of both
set -u
mkdir -p gpurun_out
TAG=${1:-ham}
timeout 300 python -m pytest tests/test_gpu_hamming2.py -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_tests.log
for M in 0 1; do
SFM_HAMMING_MODE=$M timeout 300 python tools/exp_hamming.py > gpurun_out/${TAG}_mode$M.json 2> gpurun_out/${TAG}_mode$M.err; echo "mode $M rc=$?"; cut -c1-600 gpurun_out/${TAG}_mode$M.json; tail -3 gpurun_out/${TAG}_mode$M.err
done
