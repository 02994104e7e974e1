#!/bin/bash
# Evidence run for profiles/: the driver's bench command, its ncu launch list, full captures of the two dominant kernels
# over one whole steady-state frame (8 extend launches = bounces 0..7, 8 shade launches).
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --steps 50 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
KERNEL=extend_kernel SKIP=${SKIP_EXT:-24} COUNT=8 bash scripts/gpu_profile.sh > /dev/null 2>&1
KERNEL=shade_kernel SKIP=24 COUNT=3 bash scripts/gpu_profile.sh > /dev/null 2>&1
ls -la gpurun_out
