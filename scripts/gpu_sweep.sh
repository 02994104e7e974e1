#!/bin/bash
# runs the short bench once per library variant in rustracer_b200/csrc/_build/var_*.so (and a quick parity test each)
for f in rustracer_b200/csrc/_build/var_*.so; do
  echo "VARIANT $f"
  if [ "${TESTS:-0}" = "1" ]; then RT_B200_LIB=$PWD/$f timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -1; fi
  RT_B200_LIB=$PWD/$f timeout 300 python bench.py --steps ${STEPS:-48} --warmup 6 --no-cpu-baseline ${BENCH_ARGS:-} 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']; print('  value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms', round(d['ms_per_step'], 3), 'frac', round(r['frac'], 3), 'stages', {k: round(v, 3) for k, v in r['stage_ms_per_step'].items()}, 'nodes/ray', round(r['nodes_per_ray'], 2), 'tris/ray', round(r['tris_per_ray'], 2))"
done
