#!/bin/bash
# builds library variants for A/B runs on the GPU box: scripts/build_variants.sh name1:"-DX=1 -DY=2" name2:"..."
# outputs rustracer_b200/csrc/_build/var_<name>.so (picked up by scripts/gpu_sweep.sh through RT_B200_LIB)
set -e
cd "$(dirname "$0")/../rustracer_b200/csrc"
mkdir -p _build; rm -f _build/var_*.so
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --extended-lambda --expt-relaxed-constexpr -Xcompiler -fPIC,-ffp-contract=off,-Wno-unused-function -ccbin g++"
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  ( /usr/local/cuda/bin/nvcc $FLAGS $defs -shared -o _build/var_$name.so rt_api.cu -lcudart 2>&1 | grep -E "error" || true ) &
done
wait
ls -la _build/
