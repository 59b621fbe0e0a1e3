#!/bin/bash
# full-size bench (power-capped regime) for every variant library, twice, interleaved
set -u
mkdir -p gpurun_out
for rep in 1 2; do
for lib in build/variants/lib_*.so; do
  n=$(basename $lib .so)
  SFM_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 10 --no-extras --no-cpu-baseline --no-self-check 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$n', round(d['value']), round(d['roofline']['kernel_ms_per_step'],2), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))"
done
done
