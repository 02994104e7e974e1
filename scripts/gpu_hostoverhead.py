"""Host-side cost of rt_render (enqueue only) vs GPU time, at 1080p and 512x512."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench
from rustracer_b200 import core, host
desc = bench.build_scene_desc()
for (W, H) in ((1920, 1080), (512, 512)):
    ctx = core.Context(W, H); sc = core.Scene(ctx, desc)
    cam = host.Camera(W, H).set(position=(0, 0, 14.0)); gui = host.Gui(number_of_samples=1, number_of_bounces=8)
    ubos = [bench.frame_ubo(cam, gui, f, True) for f in range(60)]
    for u in ubos[:10]: ctx.render(sc, u)
    ctx.synchronize()
    t0 = time.perf_counter()
    for u in ubos[10:]: ctx.render(sc, u)
    t1 = time.perf_counter(); ctx.synchronize(); t2 = time.perf_counter()
    print(f"{W}x{H}: enqueue {1e3*(t1-t0)/50:.3f} ms/frame, total {1e3*(t2-t0)/50:.3f} ms/frame, launches/frame {ctx.stats().n_kernel_launches}")
