#!/bin/bash
# what the driver runs at round end, on one GPU: GPU parity suite, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tee gpurun_out/final_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/final_smoke.txt
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.log | python scripts/show_bench.py
timeout 600 python bench.py --impl reference 2>&1 | tail -1 | tee gpurun_out/bench_reference.log | cut -c1-160
