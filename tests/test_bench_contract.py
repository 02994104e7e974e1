"""bench.py contract checks that need no GPU: the reference (CPU) arm prints ONE JSON line with the agreed keys, and the
GPU arm refuses to run without a CUDA device instead of falling back to the CPU."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mrays/s" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return      # on a GPU box the GPU tests and the driver exercise this arm
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_committed_bench_lines_follow_the_contract():
    """The evidence under profiles/ is a set of real bench.py lines: every committed line carries the contract's keys, its
    roofline fraction follows from its own achieved / peak figures, its whole-frame bytes follow from its own counters
    (SURVEY.md §8d formula), and both arms of the headline name the same workload."""
    import bench
    prof = ROOT / "profiles"
    head = json.loads((prof / "r02_bench_line.json").read_text())
    ref = json.loads((prof / "r02_bench_reference.json").read_text())
    assert ref["impl"] == "reference" and ref["config"] == head["config"] and ref["metric"] == head["metric"]
    lines = [head] + [json.loads(p.read_text()) for p in sorted(prof.glob("r02_config_*.json")) + sorted(prof.glob("r02_bench_*gpu*.json"))]
    assert len(lines) >= 12
    for d in lines:
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                  "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
            assert k in d, k
        assert d["metric"] == "Mrays/s" and d["higher_is_better"] is True and d["vs_baseline"] is None and d["warmup"] >= 3
        assert d["gpu_launches"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        r = d["roofline"]
        assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        # value = rays of the whole job (rays_per_step is one rank's share of a step) / max-over-ranks time
        assert abs(d["value"] - d["rays_per_step"] * d["n_gpus"] / d["ms_per_step"] / 1e3) / d["value"] < 0.02
        if d["n_gpus"] == 1:
            st = r["counters_per_step"]
            trav, whole = bench.algorithmic_bytes(st, spp=1)
            assert abs(whole - r["whole_frame_bytes_per_step"]) / whole < 0.01
