"""Developer probe (needs a -DRT_PROBE variant library in RT_B200_LIB): per-warp timeline of one extend launch.
env: DEPTH (bounce whose extend launch is probed = DEPTH-1; the frame is rendered with number_of_bounces=DEPTH)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from rustracer_b200 import core, host, _ffi as F
bench.select_config(int(os.environ.get("CONFIG", "2")))
d = bench.build_scene_desc()
ctx = core.Context(bench.WIDTH, bench.HEIGHT); sc = core.Scene(ctx, d)
cam = host.Camera(bench.WIDTH, bench.HEIGHT).set(position=bench.CAM_POS)
lib = F.load_rt(); lib.rt_debug_probe.argtypes = [C.c_void_p, C.c_uint32]
NW = 148 * 8 * 4
for depth in [int(x) for x in os.environ.get("DEPTHS", "1,2,3").split(",")]:
    gui = host.Gui(number_of_samples=1, number_of_bounces=depth, **bench.GUI_KW)
    for f in range(4):
        ctx.render(sc, bench.frame_ubo(cam, gui, f, bool(d.fully_opaque)), flags=4)
    st = ctx.stats()
    buf = np.zeros((NW, 6), np.uint64)
    assert lib.rt_debug_probe(buf.ctypes.data, NW) == 0
    t0, tx, t1, rays, iters, sm = [buf[:, k].astype(np.float64) for k in range(6)]
    start = t0.min(); end = t1.max(); dur = (end - start) / 1e3
    fin = (t1 - start) / 1e3; exh = (tx[tx > 0] - start) / 1e3
    print(f"depth {depth}: last launch = bounce {depth - 1}; kernel span {dur:.1f} us (stats: extend total {st.ms_extend * 1e3:.1f} us over {depth} launches)")
    print(f"  warp start spread {((t0 - start) / 1e3).max():.1f} us; queue exhausted seen at {np.percentile(exh, [0, 50, 100]).round(1)} us")
    print(f"  warp finish percentiles [10,50,90,99,100] = {np.percentile(fin, [10, 50, 90, 99, 100]).round(1)} us")
    print(f"  rays/warp [min,med,max] = {rays.min():.0f} {np.median(rays):.0f} {rays.max():.0f}; iterations/warp [med,90,max] = {np.median(iters):.0f} {np.percentile(iters, 90):.0f} {iters.max():.0f}")
    late = np.argsort(fin)[-8:]
    print("  latest warps: " + ", ".join(f"w{w} sm{int(sm[w])} fin {fin[w]:.0f}us exh@{(tx[w] - start) / 1e3 if tx[w] > 0 else -1:.0f}us rays {int(rays[w])} it {int(iters[w])}" for w in late))
    per_sm = np.array([fin[sm == k].max() for k in np.unique(sm)])
    print(f"  per-SM last finish [min,med,max] = {per_sm.min():.1f} {np.median(per_sm):.1f} {per_sm.max():.1f} us; mean of warp finish {fin.mean():.1f} us")
