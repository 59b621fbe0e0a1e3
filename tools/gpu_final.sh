#!/bin/bash
# Round-end evidence: tests, full bench, reference arm, datasets workload, ncu launch list, ncu full
# captures (kNN kernel on the full 19,900-pair launch; geometry kernels on the 4M-point launches;
# both HAMMING2 kernels).  Every ncu run follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
TAG=${1:-final}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
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
GEO="python tools/geo_only.py"
timeout 300 $GEO > gpurun_out/${TAG}_plain3.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:triangulate_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_tri $GEO > gpurun_out/${TAG}_ncu3.log 2>&1
echo "ncu tri rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:residual_seg_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_res $GEO > gpurun_out/${TAG}_ncu4.log 2>&1
echo "ncu res rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobian_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_jac $GEO > gpurun_out/${TAG}_ncu5.log 2>&1
echo "ncu jac rc=$?"
HAM="python tools/exp_hamming.py"
SFM_HAMMING_MODE=0 timeout 300 $HAM > gpurun_out/${TAG}_hamming_cuda.json 2> gpurun_out/${TAG}_hamming_cuda.err
timeout 300 $HAM > gpurun_out/${TAG}_hamming.json 2> gpurun_out/${TAG}_hamming.err &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hamming2_tc_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_hamtc $HAM > gpurun_out/${TAG}_ncu6.log 2>&1
echo "ncu hamming tc rc=$?"
SFM_HAMMING_MODE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:hamming2_knn_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_ham $HAM > gpurun_out/${TAG}_ncu7.log 2>&1
echo "ncu hamming rc=$?"
ls -la gpurun_out | grep ${TAG}
