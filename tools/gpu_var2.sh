#!/bin/bash
# Kernel-variant visit for the ratio-driven sweep: match-only parity tests + timings for every library
# under build/variants/, then one ncu full capture of the shipped kernel on the 24-image workload.
set -u
mkdir -p gpurun_out
TAG=${1:-var2}
for lib in build/variants/lib_*.so; do
  n=$(basename $lib .so)
  SFM_B200_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_datasets.py tests/test_gpu_matching.py -m gpu -x -q \
    -k "matches_only or datasets or property or match_only or max_norm or dog or staged" > gpurun_out/${TAG}_${n}_tests.log 2>&1
  echo "$n pytest rc=$? $(tail -1 gpurun_out/${TAG}_${n}_tests.log)"
done
python tools/variants.py run 24 8192 1 2>&1 | tee gpurun_out/${TAG}_t24.log
python tools/variants.py run 2 65536 1 2>&1 | tee gpurun_out/${TAG}_t65536.log
[ -n "${NCU_LIB:-}" ] && export SFM_B200_LIB=$PWD/build/variants/$NCU_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn2 -s 1 -c 1 -f -o gpurun_out/${TAG}_knn2 \
  python tools/exp_one.py 1 24 8192 > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
