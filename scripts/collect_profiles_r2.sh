#!/bin/bash
# Turns the artefacts of scripts/gpu_profiles_r2.sh / gpu_multi.sh (merged into gpurun_out/) into the committed summaries under profiles/.
set -eu
R=r02
cp gpurun_out/bench.log profiles/${R}_bench_line.json
cp gpurun_out/bench_reference.log profiles/${R}_bench_reference.json
cp gpurun_out/launches.csv profiles/${R}_launches.csv
python scripts/launch_list_summary.py gpurun_out/launches.csv > profiles/${R}_launch_list.txt
python scripts/ncu_traffic.py gpurun_out/prof_extend_kernel.ncu-rep profiles/${R}_extend_traffic.json "extend_kernel<ALPHA=0,COUNT=0,SINGLE=1>" \
  "ncu --set full --clock-control none -k regex:extend_kernel -s 24 -c 8, bench.py --steps 2 --warmup 3 --frames-in-flight 1 (one steady-state frame: bounces 0..7)"
ncu -i gpurun_out/prof_extend_kernel.ncu-rep --page source --csv --print-source sass,cuda > /tmp/r2_ext_src.csv 2>/dev/null
{ echo "# ncu --set full --clock-control none --import-source on, extend_kernel<ALPHA=0,COUNT=0,SINGLE=1>, bench.py --steps 2 --warmup 3 --frames-in-flight 1"
  echo "# (config 2, reference camera, 1080p depth 8): the 8 launches of one steady-state frame (bounces 0..7)"
  python scripts/ncu_summary.py gpurun_out/prof_extend_kernel.ncu-rep 8 25; echo
  for l in 0 1 2; do
    echo "# basic-block breakdown, launch $l (bounce $l): executions, lanes active, share of warp instructions / of stall samples (blocks >= 0.8 % of instructions)"
    python scripts/ncu_sass_flow.py /tmp/r2_ext_src.csv $l > /tmp/r2_flow.txt; head -1 /tmp/r2_flow.txt; python scripts/ncu_blocks.py /tmp/r2_flow.txt 0.8 | cut -c1-190; echo
  done; } > profiles/${R}_extend_kernel_ncu.txt 2>&1
{ echo "# ncu --set full, shade_kernel<SIMPLE=1,COUNT=0> (material-sorted), same command, bounces 0..7 of one steady-state frame"
  python scripts/ncu_summary.py gpurun_out/prof_shade_kernel.ncu-rep 8 25; } > profiles/${R}_shade_kernel_ncu.txt 2>&1
for c in 1 3 4 5; do if [ -f gpurun_out/config_$c.json ]; then tail -1 gpurun_out/config_$c.json > profiles/${R}_config_$c.json; fi; done
for t in memcheck racecheck; do cp gpurun_out/sanitizer_$t.txt profiles/${R}_sanitizer_$t.txt; done
for n in 2 4 8; do
  true; [ -f gpurun_out/multigpu_check_${n}gpu.txt ] && cp gpurun_out/multigpu_check_${n}gpu.txt profiles/${R}_multigpu_check_${n}gpu.txt
  [ -f gpurun_out/bench_${n}gpu.json ] && cp gpurun_out/bench_${n}gpu.json profiles/${R}_bench_${n}gpu.json
  [ -f gpurun_out/bench_${n}gpu_tiles.json ] && cp gpurun_out/bench_${n}gpu_tiles.json profiles/${R}_bench_${n}gpu_tiles.json
  for c in 3 5; do [ -f gpurun_out/config_${c}_${n}gpu.json ] && cp gpurun_out/config_${c}_${n}gpu.json profiles/${R}_config_${c}_${n}gpu.json; done
done
python scripts/results_table.py > profiles/${R}_results_table.md
if [ -f gpurun_out/prof_extend_c3.ncu-rep ]; then
  ncu -i gpurun_out/prof_extend_c3.ncu-rep --page source --csv --print-source sass,cuda > /tmp/r2_c3_src.csv 2>/dev/null
  { echo "# config 3 (10k instances x 100k triangles, alpha MASK): extend_kernel<ALPHA=1,COUNT=0,SINGLE=0>, ncu --set full, bounces 0..3 of one frame"
    python scripts/ncu_summary.py gpurun_out/prof_extend_c3.ncu-rep 4 12; echo
    echo "# basic-block breakdown, launch 1: the instance entry (rt_traverse.h trav_enter_instance), the leaf pushes and the any-hit evaluation run at 2-3 lanes"
    python scripts/ncu_sass_flow.py /tmp/r2_c3_src.csv 1 > /tmp/r2_flow.txt; head -1 /tmp/r2_flow.txt; python scripts/ncu_blocks.py /tmp/r2_flow.txt 0.9 | cut -c1-190; } > profiles/${R}_extend_config3_ncu.txt 2>&1
fi
if [ -f gpurun_out/prof_skin_c4.ncu-rep ]; then
  { echo "# config 4 (1 M-triangle skinned character): skin_kernel and refit_nodes_kernel, ncu --set full --clock-control none, one steady-state update"
    echo "# (bench.py --config 4 --steps 2 --warmup 3 --frames-in-flight 1).  Algorithmic bytes of skinning: 256 B per vertex (128 B in + 128 B out)."
    python scripts/ncu_summary.py gpurun_out/prof_skin_c4.ncu-rep 4 0 2>&1 | grep -v "^total\|inst  " ; } > profiles/${R}_skin_refit_config4_ncu.txt 2>&1
fi
echo "profiles/${R}_* refreshed"; ls -la profiles | grep r02
