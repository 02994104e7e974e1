#!/bin/bash
# quick check after a kernel change: GPU parity suite, then one bench line per config in $CONFIGS
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for c in ${CONFIGS:-3 2}; do
  echo "== config $c"
  timeout 600 python bench.py --config $c --steps ${STEPS:-16} --warmup 3 --no-cpu-baseline --no-alt-camera 2>&1 | tail -1 | python scripts/show_bench.py
done
