#!/bin/bash
# One GPU-box visit: tests, microbenchmarks, bench, then ncu (launch list + full capture of the kNN kernel).
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
[ -x tools/ubench ] && tools/ubench > gpurun_out/${TAG}_ubench.json 2> gpurun_out/${TAG}_ubench.err
python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
SMALL="python bench.py --steps 2 --warmup 3 --images 24 --no-extras --no-cpu-baseline"
$SMALL > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $SMALL > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
$SMALL > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn2 -s 3 -c 2 -f -o gpurun_out/${TAG}_knn2 $SMALL > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
