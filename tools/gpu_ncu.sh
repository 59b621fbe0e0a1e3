#!/bin/bash
# ncu only: launch list + full capture of the kNN kernel on the small bench config.
set -u
mkdir -p gpurun_out
TAG=${1:-n}
SMALL="python bench.py --steps 2 --warmup 3 --images 24 --no-extras --no-cpu-baseline"
$SMALL > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $SMALL > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
$SMALL > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn2 -s 3 -c 1 -f -o gpurun_out/${TAG}_knn2 $SMALL > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu full rc=$?"
tail -2 gpurun_out/${TAG}_plain.log | cut -c1-300
