#!/bin/bash
# GPU tests, then the bench at 1..4 frames in flight
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for n in ${NFS:-1 2 3 4}; do
  timeout 300 python bench.py --steps 48 --warmup 6 --no-cpu-baseline --frames-in-flight $n 2>&1 | tail -1 | tee gpurun_out/bench_fif$n.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('NF', d['config'].get('frames_in_flight'), 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms', round(d['ms_per_step'], 3), 'frac', round(d['roofline']['frac'], 3))"
done
