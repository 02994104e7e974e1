#!/bin/bash
# config 3: bench per library variant, then a full ncu capture of the two-level extend kernel (bounces 0..3 of one steady-state frame)
echo BASE; timeout 600 python bench.py --config 3 --steps 16 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python scripts/show_bench.py
for f in rustracer_b200/csrc/_build/var_*.so; do [ -e "$f" ] || continue; echo "VARIANT $f"; RT_B200_LIB=$PWD/$f timeout 600 python bench.py --config 3 --steps 16 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python scripts/show_bench.py; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:extend_kernel -s 24 -c 4 -f -o gpurun_out/prof_extend_c3 \
    python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline --frames-in-flight 1 > gpurun_out/ncu_full_c3.log 2>&1
tail -1 gpurun_out/ncu_full_c3.log | cut -c1-150
