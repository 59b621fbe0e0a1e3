#!/bin/bash
# All GPU tests, then the ncu launch list of the 24-image bench workload (after a plain run of the same command).
set -u
mkdir -p gpurun_out
TAG=${1:-last2}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
SMALL="python bench.py --steps 2 --warmup 3 --images 24 --no-cpu-baseline --no-extras --no-self-check"
timeout 120 $SMALL > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $SMALL > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
