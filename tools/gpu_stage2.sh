#!/bin/bash
# staged multi-GPU upload: all-pairs bench at N ranks with 1 / 3 stages, plus the multi-GPU tests
set -u
mkdir -p gpurun_out
TAG=${1:-st2}; N=${2:-2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_matching.py -m gpu -x -q -k "staged or two_devices" > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
for S in 3 1 4; do
timeout 900 $TR --nproc-per-node $N --master-port 2952$S bench.py --gpus $N --steps 10 --warmup 3 --stages $S 2> gpurun_out/${TAG}_bench_${N}gpu_s$S.err | grep '^{' > gpurun_out/${TAG}_bench_${N}gpu_s$S.json; echo "bench N=$N stages=$S rc=$?"
tail -3 gpurun_out/${TAG}_bench_${N}gpu_s$S.err | cut -c1-400
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
