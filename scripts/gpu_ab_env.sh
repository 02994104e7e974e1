#!/bin/bash
# A/B of builder / traversal knobs (environment variables) over the BASELINE configs: parity subset, then one bench line per setting
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for c in ${CONFIGS:-3 2 4 5 1}; do
  for v in "" ${VARIANTS:-RT_B200_MORTON_PER_AXIS=1 RT_B200_NO_MERGED_FIRST=1}; do
    echo "== config $c ${v:-default}"
    env $v timeout 600 python bench.py --config $c --steps ${STEPS:-16} --warmup 3 --no-cpu-baseline --no-alt-camera 2>&1 | tail -1 | python scripts/show_bench.py
  done
done
