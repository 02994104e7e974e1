"""Developer probe: per-ray node counters of config 3 through the host-emulation build, merged-first on/off."""
import ctypes as C, sys, os, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from rustracer_b200 import _ffi as F, core, scenes, host
lib = C.CDLL(os.environ.get("EMU_PROF_LIB", "/tmp/emuprof/librt_emu.so"))
rename = lambda n: "emu_" + n
F.bind_rt(lib, rename, optional=F.RT_CUDA_ONLY)
api = core.Api(lib, rename)
lib.emu_prof.restype = C.POINTER(C.c_ulonglong)
d = scenes.instanced_foliage(n_side=100, tris_per_mesh=100_000, cards=64, tex_size=1024)
ctx = core.Context(16, 16, api=api); sc = core.Scene(ctx, d)
rng = np.random.default_rng(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
rays = np.zeros(n, F.RAY_DTYPE)
# primary-like rays from the bench camera towards the field
o = np.array([0, 1.2, 7.0], np.float32)
tx = rng.uniform(-4, 4, n); tz = rng.uniform(-5, 3, n)
tgt = np.stack([tx, np.zeros(n), tz], 1).astype(np.float32)
dd = tgt - o; dd /= np.linalg.norm(dd, axis=1, keepdims=True)
rays["origin"] = o; rays["direction"] = dd; rays["tmin"] = 1e-3; rays["tmax"] = 1e4
p = lib.emu_prof()
out = []
for i in range(n):
    for k in range(8): p[k] = 0
    h = sc.trace_closest(rays[i:i + 1])
    out.append([p[k] for k in range(6)] + [float(h["t"][0]), int(h["instance_id"][0]), int(h["primitive_id"][0])])
np.save(sys.argv[2], np.array(out))
print(np.array(out)[:, :6].mean(0))
