#!/bin/bash
# Round-end evidence for a change of the kNN-2 kernel only (geometry / HAMMING2 kernels and their ncu
# captures unchanged since tools/gpu_final.sh last ran): tests, full bench, reference arm, datasets
# workload, smoke, ncu launch list, ncu full capture of the kNN kernel on the full 19,900-pair launch.
# Every ncu run follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
TAG=${1:-final}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/${TAG}_bench.json | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --workload datasets --steps 5 > gpurun_out/${TAG}_datasets.json 2> gpurun_out/${TAG}_datasets.err; echo "datasets rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
SMALL="python bench.py --steps 2 --warmup 3 --images 24 --no-cpu-baseline --no-extras --no-self-check"
timeout 300 $SMALL > gpurun_out/${TAG}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $SMALL > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
FULL="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-self-check"
timeout 300 $FULL > gpurun_out/${TAG}_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn2 -s 3 -c 1 -f -o gpurun_out/${TAG}_knn2 $FULL > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu knn2 rc=$?"
ls -la gpurun_out | grep ${TAG}
