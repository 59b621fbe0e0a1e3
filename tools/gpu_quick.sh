#!/bin/bash
# Quick GPU visit: parity tests + a short bench (+ optional ubench).
set -u
mkdir -p gpurun_out
TAG=${1:-q}
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_tests.log
tail -15 gpurun_out/${TAG}_tests.log
[ "${UBENCH:-0}" = "1" ] && tools/ubench > gpurun_out/${TAG}_ubench.json 2> gpurun_out/${TAG}_ubench.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench.json"))
    r=d["roofline"]; print("value",d["value"],"e2e",d["e2e"]["value"],"tops",r["achieved"],"frac",r["frac"],"probe",r["measured_i8_probe_tops"],"clk",d["clocks"])
    print(d.get("extra"))
except Exception as e: print("no bench json",e)
PY
