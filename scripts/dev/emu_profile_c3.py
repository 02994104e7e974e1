"""Developer probe (CPU only): config 3 through the host-emulation build with per-level node / triangle counters.
Build: g++ -O2 -std=c++17 -fPIC -ffp-contract=off -DRT_EMU_PROFILE -w -shared -o /tmp/emuprof/librt_emu.so tests/emu/emu.cpp"""
import ctypes as C, sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from rustracer_b200 import _ffi as F, core, scenes, host
import bench

W, H = int(sys.argv[1]) if len(sys.argv) > 1 else 240, int(sys.argv[2]) if len(sys.argv) > 2 else 135
import os
lib = C.CDLL(os.environ.get("EMU_PROF_LIB", "/tmp/emuprof/librt_emu.so"))
rename = lambda n: "emu_" + n
F.bind_rt(lib, rename, optional=F.RT_CUDA_ONLY)
api = core.Api(lib, rename)
lib.emu_prof.restype = C.POINTER(C.c_ulonglong)
t = time.time()
d = scenes.instanced_foliage(n_side=100, tris_per_mesh=100_000, cards=64, tex_size=1024, sky=scenes.procedural_sky(256))
print("scene desc", time.time() - t); t = time.time()
ctx = core.Context(W, H, api=api); sc = core.Scene(ctx, d)
print("build", time.time() - t); t = time.time()
bench.SPP, bench.BOUNCES = 1, 8
cam = host.Camera(W, H).set(position=(0, 1.2, 7.0))
gui = bench.make_gui()
p = lib.emu_prof()
for i in range(8): p[i] = 0
u = bench.frame_ubo(cam, gui, 0, False)
ctx.render(sc, u, flags=F.RT_RENDER_COUNTERS if hasattr(F, "RT_RENDER_COUNTERS") else 0)
ctx.synchronize()
print("render", time.time() - t)
st = ctx.stats()
rays = st.rays_extend
print({k: getattr(st, k) for k in bench.COUNTER_KEYS})
names = ["tlas nodes", "merged nodes", "object nodes", "merged tris", "object tris", "entries"]
for n, i in zip(names, range(6)): print(f"{n:14s} {p[i]:10d}  per ray {p[i] / max(rays, 1):.2f}")
bi = sc.bvh_info()
print("tlas nodes", bi.tlas_nodes, "depth", bi.max_depth_tlas, "blas nodes", bi.blas_nodes, "depth", bi.max_depth_blas)
