#!/bin/bash
# ncu full capture of the dominant kernels (one GPU, few launches) + the launch list of the same command.
set -u
mkdir -p gpurun_out
K=${KERNEL:-extend_kernel}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-9} -c ${COUNT:-4} -f -o gpurun_out/prof_$K \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline ${BENCH_EXTRA:-} > gpurun_out/ncu_full_$K.log 2>&1
tail -2 gpurun_out/ncu_full_$K.log
ls -la gpurun_out/
