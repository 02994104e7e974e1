#!/bin/bash
# Evidence run for profiles/: launch list of the bench command + full captures of the two dominant kernels.
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
KERNEL=extend_kernel SKIP=${SKIP_EXT:-9} COUNT=4 bash scripts/gpu_profile.sh > /dev/null 2>&1
KERNEL=shade_kernel SKIP=8 COUNT=3 bash scripts/gpu_profile.sh > /dev/null 2>&1
ls -la gpurun_out
