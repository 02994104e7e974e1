#!/bin/bash
# Turns the artefacts of scripts/gpu_profiles_round.sh (merged into gpurun_out/ by gpurun) into the committed summaries
# under profiles/.  Refuses to run on stale artefacts: the last gpurun call must have succeeded.
set -eu
R=${ROUND:-r01}
python - <<'PY'
import json, sys
v = json.load(open('gpurun_out/.last_call.json'))
if v.get('status') != 'ok' or v.get('rc') != 0 or 'prof_extend_kernel.ncu-rep' not in v.get('pulled_files', []):
    sys.exit(f"gpurun_out is stale: last call status={v.get('status')} rc={v.get('rc')} pulled={v.get('pulled_files')}")
PY
cp gpurun_out/bench.log profiles/${R}_bench_line.json
cp gpurun_out/launches.csv profiles/${R}_launches.csv
python scripts/launch_list_summary.py gpurun_out/launches.csv > profiles/${R}_launch_list.txt
python scripts/ncu_traffic.py gpurun_out/prof_extend_kernel.ncu-rep profiles/${R}_extend_traffic.json "extend_kernel<ALPHA=0,COUNT=0,SINGLE=1>" \
  "ncu --set full --clock-control none -k regex:extend_kernel -s 24 -c 8, bench.py --steps 2 --warmup 3 (one steady-state frame: bounces 0..7)"
{ echo "# ncu --set full --clock-control none --import-source on, extend_kernel<ALPHA=0,COUNT=0,SINGLE=1>, bench.py --steps 2 --warmup 3 (config 2, 1080p depth 8): the 8 launches of one steady-state frame (bounces 0..7)"
  python scripts/ncu_summary.py gpurun_out/prof_extend_kernel.ncu-rep 8 40; echo
  echo "# per-region breakdown, launch 0 (primary rays) and launch 1 (first diffuse bounce)"
  python scripts/ncu_regions.py gpurun_out/prof_extend_kernel.ncu-rep 0; python scripts/ncu_regions.py gpurun_out/prof_extend_kernel.ncu-rep 1; } > profiles/${R}_extend_kernel_ncu.txt 2>&1
{ echo "# ncu --set full, shade_kernel<SIMPLE=1,COUNT=0>, same command, bounces 0..2 of one steady-state frame"
  python scripts/ncu_summary.py gpurun_out/prof_shade_kernel.ncu-rep 3 40; } > profiles/${R}_shade_kernel_ncu.txt 2>&1
for c in 1 3 4 5; do if [ -f gpurun_out/config_$c.json ]; then tail -1 gpurun_out/config_$c.json > profiles/${R}_config_$c.json; fi; done
echo "profiles/${R}_* refreshed"
