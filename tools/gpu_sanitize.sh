#!/bin/bash
# compute-sanitizer logs of tools/sanitize_small.py -> gpurun_out/<tag>_san_*.log
set -u
mkdir -p gpurun_out
TAG=${1:-r2}
python tools/sanitize_small.py > gpurun_out/${TAG}_san_plain.log 2>&1; echo "plain rc=$?"
for tool in memcheck racecheck synccheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/${TAG}_san_${tool}.log 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${TAG}_san_${tool}.log | tail -1)"
done
