#!/bin/bash
# 8-GPU all-pairs bench: exchange / staging variants of the end-to-end step
set -u
mkdir -p gpurun_out
TAG=${1:-st8}; N=${2:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { # name, extra args
  local name=$1; shift
  timeout 400 $TR --nproc-per-node $N --master-port 29551 bench.py --gpus $N --steps 10 --warmup 3 "$@" 2> gpurun_out/${TAG}_bench_${N}gpu_$name.err | grep '^{' > gpurun_out/${TAG}_bench_${N}gpu_$name.json; echo "bench N=$N $name rc=$?"
  tail -2 gpurun_out/${TAG}_bench_${N}gpu_$name.err | cut -c1-300
}
run push12 --exchange push --stage-weights 1,2
run nccl12 --exchange nccl --stage-weights 1,2 --no-self-check
run push11 --exchange push --stage-weights 1,1 --no-self-check
run push123 --exchange push --stage-weights 1,2,3 --no-self-check
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_${N}gpu_*.json")):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("/")[-1], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],2), d.get("self_check",{}).get("all_ranks_ok"), d["roofline"]["frac"])
PY
