#!/bin/bash
# Short evidence run after a kNN-kernel change: tests, bench, smoke, launch list and one full ncu
# capture of knn2_kernel (the geometry / HAMMING2 captures of tools/gpu_final.sh stay valid).
set -u
mkdir -p gpurun_out
TAG=${1:-quick}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
SMALL="python bench.py --steps 2 --warmup 3 --images 24 --no-cpu-baseline --no-extras"
$SMALL > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $SMALL > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
FULL="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline"
$FULL > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn2 -s 3 -c 1 -f -o gpurun_out/${TAG}_knn2 $FULL > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu knn2 rc=$?"
tail -2 gpurun_out/${TAG}_tests.log; cat gpurun_out/${TAG}_bench.json | cut -c1-600
