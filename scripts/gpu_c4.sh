#!/bin/bash
# config 4 (skinned character) A/B
timeout 900 python -m pytest tests -x -q -m gpu -k "skin or shadows_glb" 2>&1 | tail -2
show4() { python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d['refit']; print('  value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'skin_ms', round(r['skin_ms'],4), 'GB/s', round(r['skinning_gbs']), 'refit_ms', round(r['refit_ms'],4), 'tlas', round(r['tlas_ms'],4))"; }
echo BASE; timeout 600 python bench.py --config 4 --steps 24 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | show4
for f in rustracer_b200/csrc/_build/var_*.so; do [ -e "$f" ] || continue; echo "VARIANT $f"; RT_B200_LIB=$PWD/$f timeout 600 python bench.py --config 4 --steps 24 --warmup 4 --no-cpu-baseline 2>&1 | tail -1 | show4; done
