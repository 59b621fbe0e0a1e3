#!/bin/bash
# Scaling visit on an N-GPU box: the three workloads at N ranks (one rank per GPU, torchrun).
#   bash tools/gpu_scale.sh <tag> <N>
set -u
mkdir -p gpurun_out
TAG=${1:-scale}; N=${2:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node $N --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/${TAG}_bench_${N}gpu.err | grep '^{' > gpurun_out/${TAG}_bench_${N}gpu.json; echo "bench N=$N rc=$?"
timeout 600 $TR --nproc-per-node $N --master-port 29522 bench.py --gpus $N --workload pair65536 --steps 20 --warmup 3 2> gpurun_out/${TAG}_pair65536_${N}gpu.err | grep '^{' > gpurun_out/${TAG}_pair65536_${N}gpu.json; echo "pair65536 N=$N rc=$?"
for V in 2 8; do
timeout 600 $TR --nproc-per-node $N --master-port 29523 bench.py --gpus $N --workload geometry --views $V --steps 3 2> gpurun_out/${TAG}_geometry_v${V}_${N}gpu.err | grep '^{' > gpurun_out/${TAG}_geometry_v${V}_${N}gpu.json; echo "geometry V=$V N=$N rc=$?"
done
tail -2 gpurun_out/${TAG}_*_${N}gpu.err | cut -c1-300
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*_${N}gpu.json")):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("/")[-1], "value", d["value"], d["unit"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], {k:d[k] for k in ("sharded_list_equals_single_gpu","matches") if k in d}, d.get("self_check",{}).get("all_ranks_ok"), d.get("residuals",{}).get("obs_per_s"), d.get("residuals",{}).get("rel_diff"))
PY
