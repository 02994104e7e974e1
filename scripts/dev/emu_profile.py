"""Developer probe (CPU only): a BASELINE config through the host-emulation build with per-level node / triangle counters.
Build: g++ -O2 -std=c++17 -fPIC -ffp-contract=off -DRT_EMU_PROFILE -w -shared -o /tmp/emuprof/librt_emu.so tests/emu/emu.cpp
usage: emu_profile.py CONFIG [W H]"""
import ctypes as C, sys, time, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from rustracer_b200 import _ffi as F, core, scenes, host
import bench

cfg = int(sys.argv[1])
bench.select_config(cfg)
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (bench.WIDTH // 8, bench.HEIGHT // 8)
lib = C.CDLL(os.environ.get("EMU_PROF_LIB", "/tmp/emuprof/librt_emu.so"))
rename = lambda n: "emu_" + n
F.bind_rt(lib, rename, optional=F.RT_CUDA_ONLY)
api = core.Api(lib, rename)
lib.emu_prof.restype = C.POINTER(C.c_ulonglong)
d = bench.build_scene_desc()
ctx = core.Context(W, H, api=api); sc = core.Scene(ctx, d)
bench.SPP, bench.BOUNCES = 1, 8
cam = host.Camera(W, H).set(position=bench.CAM_POS)
gui = bench.make_gui()
p = lib.emu_prof()
for i in range(8): p[i] = 0
nf = 2
for f in range(nf):
    ctx.render(sc, bench.frame_ubo(cam, gui, f, False))
ctx.synchronize()
rays = 0
st = ctx.stats(); rays = st.rays_extend * nf   # stats are per frame (last)
names = ["tlas nodes", "merged nodes", "object nodes", "merged tris", "object tris", "entries"]
tot = sum(p[i] for i in range(3))
print(f"config {cfg} aspect={os.environ.get('RT_B200_MORTON_ASPECT', 'default')}: nodes/ray {tot / rays:.2f} tris/ray {(p[3] + p[4]) / rays:.2f}  " + "  ".join(f"{n} {p[i] / rays:.2f}" for i, n in enumerate(names)))
