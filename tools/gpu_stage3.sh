#!/bin/bash
# peer push exchange: single-GPU protocol tests, then the all-pairs bench at N ranks, push vs nccl
set -u
mkdir -p gpurun_out
TAG=${1:-st3}; N=${2:-2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_matching.py -m gpu -x -q -k "staged or peer_push or two_devices" > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_tests.log
for X in push nccl; do
timeout 600 $TR --nproc-per-node $N --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --exchange $X 2> gpurun_out/${TAG}_bench_${N}gpu_$X.err | grep '^{' > gpurun_out/${TAG}_bench_${N}gpu_$X.json; echo "bench N=$N exchange=$X rc=$?"
tail -3 gpurun_out/${TAG}_bench_${N}gpu_$X.err | cut -c1-400
done
SFM_PEER_FLAGS=kernel timeout 600 $TR --nproc-per-node $N --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --exchange push --no-self-check 2> gpurun_out/${TAG}_bench_${N}gpu_pushk.err | grep '^{' > gpurun_out/${TAG}_bench_${N}gpu_pushk.json; echo "bench kernel flags rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_${N}gpu_*.json")):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("/")[-1], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],2), d.get("self_check",{}).get("all_ranks_ok"), d["roofline"]["frac"])
PY
