"""Developer probe (CPU only): decodes the TLAS of config 3 built by the host emulation (rt_scene_read_nodes(-2)) and prints
its child-box surface areas / fan-out: how the Morton normalisation was diagnosed (run with RT_B200_MORTON_ASPECT=1e30 for the
round-1 tree).  Needs the profile build of the emulation library (see emu_profile.py)."""
import ctypes as C, sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from rustracer_b200 import _ffi as F, core, scenes
lib = C.CDLL(os.environ.get("EMU_PROF_LIB", "/tmp/emuprof/librt_emu.so"))
rename = lambda n: "emu_" + n
F.bind_rt(lib, rename, optional=F.RT_CUDA_ONLY)
api = core.Api(lib, rename)
d = scenes.instanced_foliage(n_side=100, tris_per_mesh=100_000, cards=64, tex_size=1024)
ctx = core.Context(16, 16, api=api); sc = core.Scene(ctx, d)
N = sc.read_nodes(-2)
print("tlas nodes", len(N))
def bf(x): return (x.astype(np.uint32) << 16).view(np.float32)
org = N[:, 0:3].view(np.float32)
planes = N[:, 8:32].reshape(-1, 3, 8)    # axis, child
lo = org[:, :, None] + bf(planes & 0xFFFF); hi = org[:, :, None] + bf(planes >> 16)
valid = (planes & 0xFFFF) != 0x7F80
ext = np.where(valid, hi - lo, 0)
area = 2 * (ext[:, 0] * ext[:, 1] + ext[:, 1] * ext[:, 2] + ext[:, 0] * ext[:, 2])
v = valid.all(1)
meta = N[:, 6:8].copy().view(np.uint8).reshape(-1, 8)
inner = ((meta & 0x18) == 0x18) & v
print("children/node", v.sum(1).mean(), "inner/node", inner.sum(1).mean(), "leaf slots/node", (v & ~inner).sum(1).mean())
print("sum child area", area[v].sum(), "inner child area", area[inner].sum(), "leaf area", area[v & ~inner].sum())
print("root children:", v[0].sum(), "inner", inner[0].sum()); 
for j in range(8):
    if v[0, j]: print("  child", j, "inner" if inner[0, j] else "leaf", "lo", lo[0, :, j], "hi", hi[0, :, j], "prims", meta[0, j] >> 5)
