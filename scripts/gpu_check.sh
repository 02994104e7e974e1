#!/bin/bash
# Runs on the GPU box through gpurun: GPU parity tests, smoke, a short bench and an ncu launch list.
# Every step has its own timeout so a hung kernel cannot hold the box until gpurun's limit.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" | tee gpurun_out/pytest_gpu.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 | tee -a gpurun_out/pytest_gpu.log
echo "== smoke" | tee gpurun_out/smoke.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20 | tee -a gpurun_out/smoke.log
echo "== bench" | tee gpurun_out/bench.log
timeout 600 python bench.py --steps ${BENCH_STEPS:-30} --warmup 5 2>&1 | tail -5 | tee -a gpurun_out/bench.log
if [ "${NCU:-1}" = "1" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  tail -3 gpurun_out/ncu_bench.log
fi
