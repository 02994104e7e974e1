#!/bin/bash
# N-GPU bench lines of the other BASELINE configs (sample-pass sharding; config 5 with its periodic combine every 64 steps)
set -u
N=${N:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
for c in ${CONFIGS:-5 3}; do
  steps=32; [ "$c" = "5" ] && steps=130      # config 5: 130 steps per rank -> two periodic combines (every 64) + the final one
  echo "== config $c x$N"
  timeout 900 $TR bench.py --gpus $N --config $c --steps $steps --warmup 3 2>gpurun_out/config_${c}_${N}gpu.err | tail -1 | tee gpurun_out/config_${c}_${N}gpu.json | python scripts/show_bench.py
done
