"""Generates the committed golden fixtures.  Run HERE (build container), where /root/reference exists:

    python tests/golden/make_golden.py

Outputs (tests/golden/):
  cornell_box_scene.npz   flat GPU arrays of assets/models/CornellBox/cornellBox.gltf as produced by the host
                          loader (host/gltf_host.cpp) with the reference's rules — the GPU box has no /root/reference
  cornell_box_golden.npz  oracle outputs on that scene: closest hits for a seeded ray set, any-hit results,
                          a 64x64 8-spp image (acc + rgba8), debug-mapping images, per-bounce payload traces
  rng_bsdf_golden.npz     RNG sequences, tonemap / offset_ray / BSDF sample tables from the oracle
The reference itself ships no numeric vectors, so these pin the oracle's *own* behaviour over
time (regression), and pin the host loader's output for the bundled asset.
"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from rustracer_b200 import host, _ffi as F  # noqa: E402
from oracle import orc  # noqa: E402
import util  # noqa: E402

OUT = Path(__file__).resolve().parent
ASSET = "/root/reference/assets/models/CornellBox/cornellBox.gltf"


def main():
    doc = host.load_file(ASSET)
    d = doc.scene_desc()
    util.save_scene_npz(d, OUT / "cornell_box_scene.npz")
    d2 = util.load_scene_npz(OUT / "cornell_box_scene.npz")
    s = orc.OracleScene(d2)

    rays, rng4 = util.random_rays(4096, seed=1234, extent=4.9)
    gold = {"rays": rays, "rng4": rng4}
    gold["hits_opaque"] = s.trace_closest(rays, F.RT_TRACE_OPAQUE if hasattr(F, "RT_TRACE_OPAQUE") else 1, rng4)
    gold["hits_alpha"] = s.trace_closest(rays, 0, rng4)
    srays = rays.copy(); srays["tmin"] = 0.1; srays["tmax"] = 6.0
    gold["any_alpha"] = s.trace_any(srays, 0, rng4)
    W = H = 64
    cam = host.Camera(W, H).set(position=(0, 0, 14.0))
    gui = host.Gui(number_of_samples=2, number_of_bounces=8)
    drv = host.FrameDriver(cam, gui, doc.fully_opaque())
    acc = None
    ubos = []
    for _ in range(4):
        u = drv.next_ubo(); ubos.append(bytes(u))
        acc, out, st = s.render(u, W, H, acc)
    gold["image_acc"], gold["image_out"] = acc, out
    gold["image_ubos"] = np.frombuffer(b"".join(ubos), np.uint8).reshape(4, -1)
    for name, mapping in (("albedo", 5), ("normal", 8), ("instance", 2), ("triangle", 3)):
        g2 = host.Gui(number_of_samples=1, number_of_bounces=8, mapping=mapping, antialiasing=0)
        u = host.FrameDriver(cam, g2, doc.fully_opaque()).next_ubo()
        _, o, _ = s.render(u, W, H, None)
        gold["map_" + name] = o
        gold["map_" + name + "_ubo"] = np.frombuffer(bytes(u), np.uint8)
    u = F.rt_ubo.from_buffer_copy(ubos[0])
    gold["payload_trace"] = np.stack([np.pad(s.trace_pixel(u, W, H, x, y, 8), ((0, 8), (0, 0)))[:8] for x, y in ((32, 32), (10, 50), (50, 12), (20, 40))])
    np.savez_compressed(OUT / "cornell_box_golden.npz", **gold)

    # RNG / BSDF tables
    L = orc.lib()
    t = {}
    t["tea"] = np.array([[a, b, L.orc_tea(a, b)] for a, b in ((0, 0), (1, 2), (123456, 654321), (0xFFFFFFFF, 7), (640, 480))], np.uint64)
    states = np.array([[1, 2, 3, 0], [100, 200, 7, 41], [0xFFFFFFFF, 0, 0xDEADBEEF, 5]], np.uint32)
    seqs = []
    for st4 in states:
        st = (F.c_u32 * 4)(*[int(x) for x in st4]); seqs.append([L.orc_rand(st) for _ in range(8)])
    t["pcg_states"], t["pcg_rand"] = states, np.array(seqs, np.float32)
    seed = F.c_u32(12345); t["lcg"] = np.array([L.orc_lcg_float(C.byref(seed)) for _ in range(8)], np.float32)
    rs = np.random.default_rng(5)
    pts, nrm = rs.uniform(-3, 3, (32, 3)).astype(np.float32), rs.normal(size=(32, 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True); pts[:4] *= 0.001
    off = np.zeros_like(pts)
    for i in range(32):
        L.orc_offset_ray(F.as_ptr(pts[i], F.c_f), F.as_ptr(nrm[i], F.c_f), F.as_ptr(off[i], F.c_f))
    t["offset_p"], t["offset_n"], t["offset_out"] = pts, nrm, off
    cols = rs.uniform(0, 4, (16, 3)).astype(np.float32); tm = np.zeros((5, 16, 3), np.float32)
    for m in range(5):
        for i in range(16):
            L.orc_tonemap(m, F.as_ptr(cols[i], F.c_f), F.as_ptr(tm[m, i], F.c_f))
    t["tonemap_in"], t["tonemap_out"] = cols, tm
    bs_in = np.zeros((48, 22), np.float32); bs_out = np.zeros((48, 14), np.float32)
    for i in range(48):
        n = rs.normal(size=3); n /= np.linalg.norm(n)
        v = rs.normal(size=3); v /= np.linalg.norm(v)
        if np.dot(n, v) < 0: v = -v
        bs_in[i] = [*n, *n, *v, *rs.uniform(0.05, 1, 3), rs.choice([0, 0.3, 1.0]), rs.choice([0, 0.2, 0.7, 1.0]), rs.choice([1.0, 1.5]),
                    rs.choice([0, 0.5, 1.0]), i % 2, (i // 2) % 2, 1 + i % 3, rs.uniform(), rs.uniform(), rs.choice([-1.0, 0.7])]
        L.orc_bsdf_sample(F.as_ptr(bs_in[i], F.c_f), F.as_ptr(bs_out[i], F.c_f))
    t["bsdf_in"], t["bsdf_out"] = bs_in, bs_out
    np.savez_compressed(OUT / "rng_bsdf_golden.npz", **t)
    for f in ("cornell_box_scene.npz", "cornell_box_golden.npz", "rng_bsdf_golden.npz"):
        print(f, (OUT / f).stat().st_size)


if __name__ == "__main__":
    main()
