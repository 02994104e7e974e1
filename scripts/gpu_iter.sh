#!/bin/bash
# quick iteration: GPU tests (optional), bench, optional ncu full capture of one kernel
set -u
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3; fi
timeout 600 python bench.py --steps ${STEPS:-30} --warmup 5 ${BENCH_ARGS:-} 2>&1 | tail -1 | tee gpurun_out/bench.log
if [ -n "${KERNEL:-}" ]; then bash scripts/gpu_profile.sh > /dev/null 2>&1; fi
