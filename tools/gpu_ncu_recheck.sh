#!/bin/bash
# ncu full capture of recheck_rows_kernel on the datasets workload (sweep forced on), after a plain run.
set -u
mkdir -p gpurun_out
TAG=${1:-rc}
export SFM_PRUNE_MODE=1
CMD="python bench.py --workload datasets --steps 1 --warmup 1"
timeout 100 $CMD > gpurun_out/${TAG}_plain_rc.log 2>&1 &&
timeout 150 ncu --set full --clock-control none --import-source on -k regex:recheck_rows -s 34 -c 1 -f -o gpurun_out/${TAG}_recheck $CMD > gpurun_out/${TAG}_ncu_rc.log 2>&1
echo "ncu rc=$?"
