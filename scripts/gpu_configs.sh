#!/bin/bash
# extra (non-headline) bench lines for BASELINE configs 1,3,4,5 -> gpurun_out/config_N.json
mkdir -p gpurun_out
for c in ${CONFIGS:-1 3 4 5}; do
  echo "== config $c"
  timeout 900 python bench.py --config $c --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/config_$c.json
done
