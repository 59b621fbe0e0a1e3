#!/bin/bash
# Kernel-variant visit: parity tests + timings for every library under build/variants/.
set -u
mkdir -p gpurun_out
TAG=${1:-var}
for lib in build/variants/lib_*.so; do
  n=$(basename $lib .so)
  SFM_B200_LIB=$PWD/$lib timeout 900 python -m pytest tests/test_gpu_matching.py tests/test_gpu_datasets.py -m gpu -x -q > gpurun_out/${TAG}_${n}_tests.log 2>&1
  echo "$n pytest rc=$? $(tail -1 gpurun_out/${TAG}_${n}_tests.log)"
done
python tools/variants.py run 24 8192 1 2>&1 | tee gpurun_out/${TAG}_t24.log
python tools/variants.py run 2 65536 1 2>&1 | tee gpurun_out/${TAG}_t65536.log
