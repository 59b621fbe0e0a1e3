#!/bin/bash
# Round-2 visit on a 2-GPU box: GPU tests (incl. two devices in one process), default bench at N=1 and
# N=2, the row-sharded 65536^2 pair and the sharded geometry workloads at N=1 and N=2.
set -u
mkdir -p gpurun_out
TAG=${1:-r2b}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_tests.log
tail -12 gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench N=1 rc=$?"
tail -3 gpurun_out/${TAG}_bench.err
for N in 2; do
  timeout 900 $TR --nproc-per-node $N --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench N=$N rc=$?"
  tail -3 gpurun_out/${TAG}_bench_${N}gpu.err
done
for N in 1 2; do
  timeout 600 $TR --nproc-per-node $N --master-port 29512 bench.py --gpus $N --workload pair65536 --steps 10 --warmup 3 > gpurun_out/${TAG}_pair65536_${N}gpu.json 2> gpurun_out/${TAG}_pair65536_${N}gpu.err; echo "pair65536 N=$N rc=$?"
  timeout 600 $TR --nproc-per-node $N --master-port 29513 bench.py --gpus $N --workload geometry --steps 3 > gpurun_out/${TAG}_geometry_${N}gpu.json 2> gpurun_out/${TAG}_geometry_${N}gpu.err; echo "geometry N=$N rc=$?"
  tail -2 gpurun_out/${TAG}_pair65536_${N}gpu.err gpurun_out/${TAG}_geometry_${N}gpu.err
done
for f in gpurun_out/${TAG}_*gpu.json; do echo "== $f"; cut -c1-1200 $f; done
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2b_bench.json".replace("r2b","'${TAG}'".strip("'"))))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], d.get("self_check"))
x=d["extra"]
print(json.dumps(x["pair_65536"])[:900])
for k in x:
    if k.startswith(("tri","res")): print(k, json.dumps(x[k])[:1200])
PY
