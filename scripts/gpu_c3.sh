#!/bin/bash
# config 3 (two-level traversal) A/B: parity subset + bench per library variant
timeout 900 python -m pytest tests -x -q -m gpu -k "instancing or foliage or baseline_sized or trace_ids" 2>&1 | tail -2
echo BASE; timeout 600 python bench.py --config 3 --steps 16 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python scripts/show_bench.py
for f in rustracer_b200/csrc/_build/var_*.so; do [ -e "$f" ] || continue; echo "VARIANT $f"; RT_B200_LIB=$PWD/$f timeout 600 python bench.py --config 3 --steps 16 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python scripts/show_bench.py; done
