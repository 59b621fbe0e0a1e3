#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-st8}; N=${2:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for S in 3 2; do
timeout 600 $TR --nproc-per-node $N --master-port 2953$S bench.py --gpus $N --steps 10 --warmup 3 --stages $S 2> gpurun_out/${TAG}_bench_${N}gpu_s$S.err | grep '^{' > gpurun_out/${TAG}_bench_${N}gpu_s$S.json; echo "bench N=$N stages=$S rc=$?"
tail -2 gpurun_out/${TAG}_bench_${N}gpu_s$S.err | cut -c1-300
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_${N}gpu_s*.json")):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("/")[-1], "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],2), d.get("self_check",{}).get("all_ranks_ok"), d["roofline"]["frac"])
PY
