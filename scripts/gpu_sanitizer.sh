#!/bin/bash
# compute-sanitizer memcheck + racecheck over smoke() and the frames-in-flight parity case (small frames)
set -u
mkdir -p gpurun_out
cat > /tmp/fif_case.py <<'PY'
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import parity_cases as pc
from rustracer_b200 import core, scenes
d = scenes.cornell_box(lucy=False)
pc.case_frames_in_flight(core.Api(), d, None, size=48, frames=6)
print("frames-in-flight case: PASS")
PY
for tool in memcheck racecheck; do
  ( timeout 600 compute-sanitizer --tool $tool python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke:|ERROR SUMMARY|RACECHECK SUMMARY|=========  *(Invalid|Race|Error)" | head -20
    timeout 600 compute-sanitizer --tool $tool python /tmp/fif_case.py 2>&1 | grep -E "PASS|ERROR SUMMARY|RACECHECK SUMMARY|Traceback|Error|=========  *(Invalid|Race)" | head -20
    # skinning kernel, cooperative BLAS / TLAS refit, multi-buffered scene, lifecycle with frames in flight
    timeout 900 compute-sanitizer --tool $tool python -m pytest tests -x -q -m gpu -k "skinning_refit or skinned_animation or lifecycle" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|=========  *(Invalid|Race)" | head -20 ) | tee gpurun_out/sanitizer_$tool.txt
done
