#!/bin/bash
# staged-arrival tests on one GPU + kernel variant timings
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_matching.py -m gpu -x -q -k "staged or sharded_upload or small_pair" > gpurun_out/st1_tests.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/st1_tests.log
python tools/variants.py run 24 8192 1 2>&1 | tee gpurun_out/st1_t24.log
python tools/variants.py run 2 65536 1 2>&1 | tee gpurun_out/st1_t65536.log
