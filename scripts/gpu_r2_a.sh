#!/bin/bash
# round 2, first GPU pass: parity tests, new-camera baseline, queue-size experiments, ncu capture of the traversal kernel
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" | tee gpurun_out/pytest_gpu.log
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -30 | tee -a gpurun_out/pytest_gpu.log
echo "== smoke" | tee gpurun_out/smoke.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee -a gpurun_out/smoke.log
echo "== bench"
timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench.log
echo "== bench fif1"
timeout 300 python bench.py --steps 20 --warmup 5 --frames-in-flight 1 --no-cpu-baseline --no-alt-camera 2>&1 | tail -1 | tee gpurun_out/bench_fif1.log
echo "== bench 4K fif1 / fif4"
timeout 300 python bench.py --steps 10 --warmup 3 --size 3840x2160 --frames-in-flight 1 --no-cpu-baseline --no-alt-camera 2>&1 | tail -1 | tee gpurun_out/bench_4k_fif1.log
timeout 300 python bench.py --steps 10 --warmup 3 --size 3840x2160 --frames-in-flight 4 --no-cpu-baseline --no-alt-camera 2>&1 | tail -1 | tee gpurun_out/bench_4k_fif4.log
for b in 4 6; do
  echo "== trace blocks $b"
  RT_B200_TRACE_BLOCKS=$b timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-alt-camera 2>&1 | tail -1 | tee gpurun_out/bench_tb$b.log
done
echo "== ncu full extend"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:extend_kernel -s 24 -c 8 -f -o gpurun_out/prof_extend_kernel \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt-camera --frames-in-flight 1 > gpurun_out/ncu_full_extend_kernel.log 2>&1
tail -2 gpurun_out/ncu_full_extend_kernel.log
ls -la gpurun_out | head -40
