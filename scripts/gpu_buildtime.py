"""Scene build time (device events around rt_scene_create's BLAS + TLAS builds) for the bench scene; best of 3."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rustracer_b200 import core
bench.select_config(int(os.environ.get("CONFIG", "2")))
d = bench.build_scene_desc()
ctx = core.Context(64, 64)
best = None
for _ in range(3):
    t0 = time.perf_counter(); sc = core.Scene(ctx, d); wall = time.perf_counter() - t0
    info = sc.bvh_info()
    best = min(best, info.build_ms) if best is not None else info.build_ms
    print(f"build_ms_device {info.build_ms:.2f} (wall {wall * 1e3:.1f} ms) nodes {info.blas_nodes} depth {info.max_depth_blas}")
    del sc
print("best", round(best, 2))
