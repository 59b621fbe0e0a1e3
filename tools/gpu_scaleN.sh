#!/bin/bash
# bench.py at N ranks exactly as the driver launches it
set -u
mkdir -p gpurun_out
TAG=${1:-sc}; N=${2:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/${TAG}_bench_${N}gpu.err | grep '^{' > gpurun_out/${TAG}_bench_${N}gpu.json; echo "bench N=$N rc=$?"
tail -2 gpurun_out/${TAG}_bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_${N}gpu.json"))
print("N=$N value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],2), d.get("self_check",{}).get("all_ranks_ok"), d["roofline"]["frac"])
PY
