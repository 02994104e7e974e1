#!/bin/bash
# quick: GPU parity tests + bench (+ optional sweep of variants)
set -u
mkdir -p gpurun_out
show() { python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']; print('  value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms', round(d['ms_per_step'], 3), 'frac', round(r['frac'], 3), 'stages', {k: round(v, 3) for k, v in r['stage_ms_per_step'].items()}, 'nodes/ray', round(r['nodes_per_ray'], 2), 'tris/ray', round(r['tris_per_ray'], 2), 'alt', d.get('alt_camera') and round(d['alt_camera']['value'],1))"; }
if [ "${TESTS:-1}" = "1" ]; then timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log; fi
echo "BASE"; timeout 300 python bench.py --steps 24 --warmup 4 --no-cpu-baseline ${BENCH_ARGS:-} 2>&1 | tail -1 | tee gpurun_out/bench_quick.log | show
echo "FIF1"; timeout 300 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-alt-camera --frames-in-flight 1 2>&1 | tail -1 | show
for f in rustracer_b200/csrc/_build/var_*.so; do
  [ -e "$f" ] || continue
  echo "VARIANT $f"
  RT_B200_LIB=$PWD/$f timeout 300 python bench.py --steps 24 --warmup 4 --no-cpu-baseline --no-alt-camera 2>&1 | tail -1 | show
done
if [ -n "${NCU:-}" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:${NCU} -s ${SKIP:-24} -c ${COUNT:-8} -f -o gpurun_out/prof_${NCU} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt-camera --frames-in-flight 1 > gpurun_out/ncu_full_${NCU}.log 2>&1
tail -2 gpurun_out/ncu_full_${NCU}.log
fi
