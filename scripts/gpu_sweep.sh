#!/bin/bash
# runs the short bench once per library variant in rustracer_b200/csrc/_build/var_*.so
for f in rustracer_b200/csrc/_build/var_*.so; do
  echo "VARIANT $f"
  RT_B200_LIB=$PWD/$f timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
done
