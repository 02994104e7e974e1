#!/bin/bash
# Round-2 evidence run for profiles/ (one B200): the driver's bench command, its ncu launch list, full ncu captures of the
# two dominant kernels over one steady-state frame (8 extend + 8 shade launches), the other BASELINE configs, sanitizers.
# Two passes (gpurun copies back at most 64 MiB): PART=1 headline bench + launch list + extend / shade captures, PART=2 the rest.
set -u
PART=${PART:-1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
if [ "$PART" = "1" ]; then
echo "== bench (driver command)"
timeout 900 python bench.py --steps 50 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench.log | python scripts/show_bench.py
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_reference.log | cut -c1-200
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt-camera > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-120
for K in extend_kernel shade_kernel; do
  echo "== ncu full $K"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 24 -c 8 -f -o gpurun_out/prof_$K \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-alt-camera --frames-in-flight 1 > gpurun_out/ncu_full_$K.log 2>&1
  tail -1 gpurun_out/ncu_full_$K.log
done
fi
if [ "$PART" = "2" ]; then
echo "== ncu full, config 3 two-level extend kernel (bounces 0..3 of one steady-state frame)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:extend_kernel -s 24 -c 4 -f -o gpurun_out/prof_extend_c3 \
    python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline --frames-in-flight 1 > gpurun_out/ncu_full_c3.log 2>&1
tail -1 gpurun_out/ncu_full_c3.log | cut -c1-150
echo "== ncu full, config 4 skinning + refit kernels (one steady-state update)"
timeout 900 ncu --set full --clock-control none -k regex:"skin_kernel|refit_nodes_kernel" -s 6 -c 4 -f -o gpurun_out/prof_skin_c4 \
    python bench.py --config 4 --steps 2 --warmup 3 --no-cpu-baseline --frames-in-flight 1 > gpurun_out/ncu_full_c4.log 2>&1
tail -1 gpurun_out/ncu_full_c4.log | cut -c1-150
echo "== configs"
for c in 1 3 4 5; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 2>&1 | tail -1 | tee gpurun_out/config_$c.json | python scripts/show_bench.py
done
if [ "${SANITIZE:-1}" = "1" ]; then
echo "== sanitizers"
for tool in memcheck racecheck; do
  ( timeout 600 compute-sanitizer --tool $tool python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke:|ERROR SUMMARY|RACECHECK SUMMARY|=========  *(Invalid|Race|Error)" | head -10
    timeout 900 compute-sanitizer --tool $tool python -m pytest tests -x -q -m gpu -k "multi_device or skinning_refit or shadows_glb or lifecycle or frames_in_flight_bit" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|=========  *(Invalid|Race)" | head -10 ) | tee gpurun_out/sanitizer_$tool.txt
done
fi
fi
ls gpurun_out | head -50
