#!/bin/bash
# round 2: tuning sweep of the traversal kernel (variants + env knobs), short benches only
set -u
mkdir -p gpurun_out
show() { python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']; print('  value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms', round(d['ms_per_step'], 3), 'frac', round(r['frac'], 3), 'stages', {k: round(v, 3) for k, v in r['stage_ms_per_step'].items()}, 'nodes/ray', round(r['nodes_per_ray'], 2), 'tris/ray', round(r['tris_per_ray'], 2))"; }
echo "BASE"; timeout 300 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-alt-camera 2>&1 | tail -1 | show
for m in 256 1024 4096; do echo "MIN_RAYS_PER_CTA $m"; RT_B200_MIN_RAYS_PER_CTA=$m timeout 300 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-alt-camera 2>&1 | tail -1 | show; done
for b in 2 3; do echo "TRACE_BLOCKS $b"; RT_B200_TRACE_BLOCKS=$b timeout 300 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-alt-camera 2>&1 | tail -1 | show; done
for f in rustracer_b200/csrc/_build/var_*.so; do
  echo "VARIANT $f"
  RT_B200_LIB=$PWD/$f timeout 300 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-alt-camera 2>&1 | tail -1 | show
done
