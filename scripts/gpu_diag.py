"""Developer diagnostic: extend/shade time as a function of the bounce limit (per-bounce launch cost and ray counts).
env: CONFIG, W/H (resolution override), DEPTHS (comma list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from rustracer_b200 import core, host, scenes
bench.select_config(int(os.environ.get("CONFIG", "2")))
if os.environ.get("W"): bench.WIDTH, bench.HEIGHT = int(os.environ["W"]), int(os.environ["H"])
d = bench.build_scene_desc()
ctx = core.Context(bench.WIDTH, bench.HEIGHT); sc = core.Scene(ctx, d)
cam = host.Camera(bench.WIDTH, bench.HEIGHT).set(position=bench.CAM_POS)
stream = torch.cuda.Stream()
prev_t, prev_r = 0.0, 0
for depth in [int(x) for x in os.environ.get("DEPTHS", "1,2,3,4,5,6,7,8").split(",")]:
    gui = host.Gui(number_of_samples=1, number_of_bounces=depth, **bench.GUI_KW)
    if True:
        for f in range(3): ctx.render(sc, bench.frame_ubo(cam, gui, f, bool(d.fully_opaque)), flags=4, stream=stream.cuda_stream)
        ex = sh = 0.0; N = 10
        for f in range(N):
            ctx.render(sc, bench.frame_ubo(cam, gui, 3 + f, bool(d.fully_opaque)), flags=4, stream=stream.cuda_stream)
            st = ctx.stats(); ex += st.ms_extend / N; sh += st.ms_shade / N
        print(f"depth {depth}: rays {st.rays_extend} (+{st.rays_extend - prev_r}) extend {ex:.3f} ms (+{ex - prev_t:.3f}) shade {sh:.3f} ms")
        prev_t, prev_r = ex, st.rays_extend
