"""Experiment: T contexts on T streams render alternating full frames (intra-GPU sample-pass sharding) — do the
per-bounce kernel tails of one frame overlap with the other frames' work?  env: CONFIG, TS (comma list)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from rustracer_b200 import core, host
bench.select_config(int(os.environ.get("CONFIG", "2")))
d = bench.build_scene_desc()
W, H = bench.WIDTH, bench.HEIGHT
cam = host.Camera(W, H).set(position=bench.CAM_POS); gui = host.Gui(number_of_samples=1, number_of_bounces=8, **bench.GUI_KW)
ctx0 = core.Context(W, H); sc = core.Scene(ctx0, d)
for T in [int(x) for x in os.environ.get("TS", "1,2,3,4").split(",")]:
    ctxs = [ctx0] + [core.Context(W, H) for _ in range(T - 1)]
    streams = [torch.cuda.Stream() for _ in range(T)]
    def frame(f):
        u = bench.frame_ubo(cam, gui, f, bool(d.fully_opaque))
        ctxs[f % T].render(sc, u, flags=2, stream=streams[f % T].cuda_stream)
    for f in range(6 * T): frame(f)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    N = 48
    e0.record(); t0 = time.perf_counter()
    for s in streams: s.wait_event(e0)
    for f in range(N): frame(6 * T + f)
    th = time.perf_counter() - t0
    for s in streams: ev = torch.cuda.Event(); ev.record(s); torch.cuda.current_stream().wait_event(ev)
    e1.record(); torch.cuda.synchronize()
    print(f"T={T}: {e0.elapsed_time(e1) / N:.3f} ms/frame (host enqueue {1000 * th / N:.3f} ms/frame)", flush=True)
    del ctxs[1:]
