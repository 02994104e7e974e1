#!/bin/bash
# skinning kernel check: bit-exact skinning tests, config 4 bench line, kernel time under ncu
timeout 900 python -m pytest tests -x -q -m gpu -k "skin or shadows_glb" 2>&1 | tail -2
timeout 600 python bench.py --config 4 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/config_4.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('value', round(d['value'], 1), 'ms', round(d['ms_per_step'], 3), d['refit'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:skin_kernel -s 4 -c 6 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu-baseline --frames-in-flight 1 2>&1 | grep -E "gpu__time_duration|dram__bytes" | head -18
