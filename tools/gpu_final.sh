#!/bin/bash
# Round-end evidence: tests, full bench, reference arm, ncu launch list, ncu full captures
# (kNN kernel on the full 19,900-pair launch; geometry kernels on the 4M-point launches).
set -u
mkdir -p gpurun_out
TAG=${1:-final}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
SMALL="python bench.py --steps 2 --warmup 3 --images 24 --no-cpu-baseline --no-extras --no-self-check"
$SMALL > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $SMALL > gpurun_out/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
FULL="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-self-check"
$FULL > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:knn2 -s 3 -c 1 -f -o gpurun_out/${TAG}_knn2 $FULL > gpurun_out/${TAG}_ncu2.log 2>&1
echo "ncu knn2 rc=$?"
GEO="python tools/geo_only.py"
$GEO > gpurun_out/${TAG}_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:triangulate_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_tri $GEO > gpurun_out/${TAG}_ncu3.log 2>&1
echo "ncu tri rc=$?"
$GEO > gpurun_out/${TAG}_plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:residual_seg_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_res $GEO > gpurun_out/${TAG}_ncu4.log 2>&1
echo "ncu res rc=$?"
ncu --set full --clock-control none --import-source on -k regex:jacobian_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_jac $GEO > gpurun_out/${TAG}_ncu5.log 2>&1
echo "ncu jac rc=$?"
HAM="python tools/exp_hamming.py"
$HAM > gpurun_out/${TAG}_hamming.json 2> gpurun_out/${TAG}_hamming.err &&
ncu --set full --clock-control none --import-source on -k regex:hamming2_knn_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_ham $HAM > gpurun_out/${TAG}_ncu6.log 2>&1
echo "ncu hamming rc=$?"
ls -la gpurun_out | grep ${TAG}
