#!/bin/bash
# N-GPU pass: correctness check of the exchange step, then the bench in both partitions.  usage: N=2 bash scripts/gpu_multi.sh
set -u
N=${N:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "== multigpu_check x$N"
timeout 600 $TR tests/multigpu_check.py 2>&1 | grep -v "^W\|^\*\*\*\|Setting OMP" | tail -6 | tee gpurun_out/multigpu_check_${N}gpu.txt
echo "== bench samples x$N"
BENCH_DEBUG=1 timeout 600 $TR bench.py --gpus $N --steps ${STEPS:-48} --warmup 5 ${BENCH_ARGS:-} 2>gpurun_out/bench_${N}gpu.err | tail -1 | tee gpurun_out/bench_${N}gpu.json | python scripts/show_bench.py
grep "timed region" gpurun_out/bench_${N}gpu.err | head -8
echo "== bench tiles x$N"
timeout 600 $TR bench.py --gpus $N --steps ${STEPS:-48} --warmup 5 --partition tiles ${BENCH_ARGS:-} 2>gpurun_out/bench_${N}gpu_tiles.err | tail -1 | tee gpurun_out/bench_${N}gpu_tiles.json | python scripts/show_bench.py
if [ "${REF:-0}" = "1" ]; then
echo "== reference arm x$N"
timeout 600 $TR bench.py --impl reference --gpus $N --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('reference', round(d['value'],2), d['cpu_baseline']['cores'], 'cores')"
fi
