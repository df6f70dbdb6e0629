#!/bin/bash
# A/B the tracking kernel variants under gnss_sdr_b200/variants/ on ONE box (same GPU, same clocks).
# usage: tools/ab_bench.sh [steps]
STEPS=${1:-200}
for rep in 1 2; do
for v in gnss_sdr_b200/variants/*.so; do
  B200_LIB=$PWD/$v timeout 200 python bench.py --steps $STEPS --warmup 5 --no-cpu-baseline --no-acq --no-e2e 2>&1 | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), 'ms', round(d['value']/1e3,1), 'Gs/s', d['clocks'])"
done
done
