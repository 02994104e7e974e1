#!/bin/bash
# per-bounce diagnostic + ncu full capture of the extend kernel (launches 9..12 = first steady frame) at HEAD
set -u
mkdir -p gpurun_out
timeout 300 python scripts/gpu_diag.py 2>&1 | tail -9 | tee gpurun_out/diag.log
KERNEL=extend_kernel SKIP=${SKIP_EXT:-24} COUNT=${COUNT_EXT:-3} bash scripts/gpu_profile.sh > /dev/null 2>&1
ls -la gpurun_out
