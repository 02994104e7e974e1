"""prints the headline fields of a bench.py JSON line (stdin)"""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); r = d.get("roofline") or {}
print("  value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 3), "gpus", d["n_gpus"], d["scaling"],
      "frac", r.get("frac") and round(r["frac"], 3), "whole", r.get("whole_frame_frac") and round(r["whole_frame_frac"], 3),
      "stages", {k: round(v, 3) for k, v in (r.get("stage_ms_per_step") or {}).items()})
bv = (d.get("engine") or {}).get("bvh") or {}
print("  nodes/ray", r.get("nodes_per_ray") and round(r["nodes_per_ray"], 2), "tris/ray", r.get("tris_per_ray") and round(r["tris_per_ray"], 2), "build ms", bv.get("build_ms_device") and round(bv["build_ms_device"], 2), "tlas ms", bv.get("tlas_ms_device") and round(bv["tlas_ms_device"], 2))
ex = (d.get("engine") or {}).get("exchange")
if ex:
    print("  exchange: fused", round(ex["fused_ms"], 4), "ms  nccl", round(ex["nccl_ms"], 4), "ms  complete image:", ex["complete_image_on_this_rank"])
if d.get("alt_camera"):
    print("  alt camera", round(d["alt_camera"]["value"], 1))
