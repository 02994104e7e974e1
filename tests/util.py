"""Shared helpers for the test-suite: fixture (de)serialisation, seeded ray sets, image metrics."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from rustracer_b200 import _ffi as F

GOLDEN = Path(__file__).resolve().parent / "golden"


def _arr(ptr, n, ctype):
    if n == 0:
        return b""
    return C.string_at(ptr, n * C.sizeof(ctype))


def save_scene_npz(d: F.rt_scene_desc, path):
    v = np.frombuffer(_arr(d.vertices, d.n_vertices, F.rt_vertex), F.VERTEX_DTYPE)
    assert (v["skin_index"] == -1).all() and d.n_images <= 1, "fixture writer only supports static untextured scenes"
    np.savez_compressed(
        path,
        position=v["position"][:, :3].copy(), normal=v["normal"][:, :3].copy(), uv0=v["uv0"].copy(), tangent=v["tangent"].copy(),
        indices=np.frombuffer(_arr(d.indices, d.n_indices, F.c_u32), np.uint32),
        prim_infos=np.frombuffer(_arr(d.prim_infos, d.n_geometries, F.rt_prim_info), np.uint32).reshape(-1, 4),
        geometries=np.frombuffer(_arr(d.geometries, d.n_geometries, F.rt_geometry), np.uint32).reshape(-1, 4),
        materials=np.frombuffer(_arr(d.materials, d.n_materials, F.rt_material), np.uint8).reshape(-1, 256),
        instances=np.frombuffer(_arr(d.instances, d.n_instances, F.rt_instance), F.INSTANCE_DTYPE),
        dlights=np.frombuffer(_arr(d.dlights, d.n_dlights, F.rt_light), F.LIGHT_DTYPE),
        plights=np.frombuffer(_arr(d.plights, d.n_plights, F.rt_light), F.LIGHT_DTYPE))


def load_scene_npz(path) -> F.rt_scene_desc:
    z = np.load(path)
    n = len(z["position"])
    v = np.zeros(n, F.VERTEX_DTYPE)
    v["position"][:, :3], v["normal"][:, :3], v["uv0"], v["tangent"] = z["position"], z["normal"], z["uv0"], z["tangent"]
    v["color"], v["skin_index"] = 1.0, -1
    keep = {"v": v, "i": np.ascontiguousarray(z["indices"]), "p": np.ascontiguousarray(z["prim_infos"]), "g": np.ascontiguousarray(z["geometries"]),
            "m": np.ascontiguousarray(z["materials"]), "inst": np.ascontiguousarray(z["instances"]), "dl": np.ascontiguousarray(z["dlights"]),
            "pl": np.ascontiguousarray(z["plights"]), "img_px": np.full((1, 1, 4), 1, np.uint8)}
    keep["img"] = (F.rt_image_desc * 1)(F.rt_image_desc(keep["img_px"].ctypes.data_as(F.c_u8p), 1, 1, 1, 0))
    keep["smp"] = (F.rt_sampler_desc * 1)(F.rt_sampler_desc(1, 1, 2, 2))
    keep["tex"] = (F.rt_texture_desc * 1)(F.rt_texture_desc(0, 0))
    d = F.rt_scene_desc()
    d.vertices, d.n_vertices = F.as_ptr(v, F.rt_vertex), n
    d.indices, d.n_indices = F.as_ptr(keep["i"], F.c_u32), len(keep["i"])
    d.prim_infos = keep["p"].ctypes.data_as(C.POINTER(F.rt_prim_info))
    d.geometries, d.n_geometries = keep["g"].ctypes.data_as(C.POINTER(F.rt_geometry)), len(keep["g"])
    d.materials, d.n_materials = keep["m"].ctypes.data_as(C.POINTER(F.rt_material)), len(keep["m"])
    d.instances, d.n_instances = F.as_ptr(keep["inst"], F.rt_instance), len(keep["inst"])
    d.images, d.n_images, d.samplers, d.n_samplers, d.textures, d.n_textures = keep["img"], 1, keep["smp"], 1, keep["tex"], 1
    d.dlights, d.n_dlights = F.as_ptr(keep["dl"], F.rt_light), len(keep["dl"])
    d.plights, d.n_plights = F.as_ptr(keep["pl"], F.rt_light), len(keep["pl"])
    d._keep = keep
    d.fully_opaque = bool((keep["m"].view(np.uint32)[:, 0] == 1).all())
    return d


def random_rays(n, seed, extent=4.5, tmin=0.001, tmax=1.0e4):
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, F.RAY_DTYPE)
    rays["origin"] = rng.uniform(-extent, extent, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["direction"] = d.astype(np.float32)
    rays["tmin"], rays["tmax"] = tmin, tmax
    rng4 = rng.integers(0, 2**32, (n, 4), dtype=np.uint32)
    return rays, rng4


def adversarial_rays(desc: F.rt_scene_desc, n, seed):
    """Rays aimed exactly at shared vertices / edge midpoints of world-space triangles, axis-parallel rays and rays
    starting on a surface (SURVEY.md §8d 'adversarial')."""
    rng = np.random.default_rng(seed)
    v = np.frombuffer(_arr(desc.vertices, desc.n_vertices, F.rt_vertex), F.VERTEX_DTYPE)
    idx = np.frombuffer(_arr(desc.indices, desc.n_indices, F.c_u32), np.uint32)
    prim = np.frombuffer(_arr(desc.prim_infos, desc.n_geometries, F.rt_prim_info), np.uint32).reshape(-1, 4)
    geo = np.frombuffer(_arr(desc.geometries, desc.n_geometries, F.rt_geometry), np.uint32).reshape(-1, 4)
    inst = np.frombuffer(_arr(desc.instances, desc.n_instances, F.rt_instance), F.INSTANCE_DTYPE)
    targets = []
    for _ in range(n):
        k = rng.integers(len(inst)); g = inst["geo_id"][k]
        M = inst["transform"][k].reshape(3, 4).astype(np.float64)
        t = rng.integers(geo[g, 1] // 3)
        tri = [v["position"][prim[g, 0] + idx[prim[g, 1] + 3 * t + j], :3].astype(np.float64) for j in range(3)]
        mode = rng.integers(3)
        p = tri[0] if mode == 0 else (0.5 * (tri[0] + tri[1]) if mode == 1 else (tri[0] + tri[1] + tri[2]) / 3.0)
        targets.append(M[:, :3] @ p + M[:, 3])
    targets = np.array(targets)
    rays = np.zeros(n, F.RAY_DTYPE)
    o = rng.uniform(-4.0, 4.0, (n, 3))
    axis = rng.integers(0, 4, n)
    for i in range(n):
        if axis[i] < 3 and i % 3 == 0:   # axis-parallel through the target
            o[i] = targets[i]; o[i, axis[i]] -= 3.0
    d = targets - o
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-12)
    rays["origin"], rays["direction"] = o.astype(np.float32), d.astype(np.float32)
    rays["tmin"], rays["tmax"] = 0.001, 1.0e4
    # a quarter of the rays start exactly on the target surface point
    q = np.arange(n) % 4 == 1
    rays["origin"][q] = targets[q].astype(np.float32)
    return rays


def hits_equal(a, b):
    return (a["instance_id"] == b["instance_id"]) & (a["primitive_id"] == b["primitive_id"]) & (a["geo_id"] == b["geo_id"]) & \
           (a["t"] == b["t"]) & (a["u"] == b["u"]) & (a["v"] == b["v"])


def psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10.0 * np.log10(255.0**2 / mse)


def mean_rel_err(acc_a, acc_b, total_samples):
    """mean_px(|L_a - L_b| / (L_b + 1e-3)) on linear luminance of acc.rgb / total samples."""
    w = np.array([0.2126, 0.7152, 0.0722])
    la = (acc_a[..., :3].astype(np.float64) / total_samples) @ w
    lb = (acc_b[..., :3].astype(np.float64) / total_samples) @ w
    return float(np.mean(np.abs(la - lb) / (lb + 1e-3)))


# ---- full scene fixtures (skinned + textured assets): every array of rt_scene_desc -------------------------------------
def save_scene_full(d: F.rt_scene_desc, path, image_files=None, **extra):
    """Writes every array of an rt_scene_desc.  image_files: {image index: encoded PNG/JPEG bytes} — those images are
    stored in their file encoding (decoded again with the host decoder by load_scene_full) instead of raw RGBA8."""
    image_files = image_files or {}
    z = dict(extra)
    z["vertices"] = np.frombuffer(_arr(d.vertices, d.n_vertices, F.rt_vertex), np.uint8)
    z["indices"] = np.frombuffer(_arr(d.indices, d.n_indices, F.c_u32), np.uint32)
    z["prim_infos"] = np.frombuffer(_arr(d.prim_infos, d.n_geometries, F.rt_prim_info), np.uint32).reshape(-1, 4)
    z["geometries"] = np.frombuffer(_arr(d.geometries, d.n_geometries, F.rt_geometry), np.uint32).reshape(-1, 4)
    z["materials"] = np.frombuffer(_arr(d.materials, d.n_materials, F.rt_material), np.uint8).reshape(-1, 256)
    z["instances"] = np.frombuffer(_arr(d.instances, d.n_instances, F.rt_instance), F.INSTANCE_DTYPE)
    z["dlights"] = np.frombuffer(_arr(d.dlights, d.n_dlights, F.rt_light), F.LIGHT_DTYPE)
    z["plights"] = np.frombuffer(_arr(d.plights, d.n_plights, F.rt_light), F.LIGHT_DTYPE)
    z["samplers"] = np.array([[s.mag_filter, s.min_filter, s.wrap_s, s.wrap_t] for s in d.samplers[:d.n_samplers]], np.uint32).reshape(-1, 4)
    z["textures"] = np.array([[t.image_index, t.sampler_index] for t in d.textures[:d.n_textures]], np.uint32).reshape(-1, 2)
    z["image_meta"] = np.array([[im.width, im.height, im.srgb, 1 if k in image_files else 0] for k, im in enumerate(d.images[:d.n_images])], np.uint32).reshape(-1, 4)
    for k, im in enumerate(d.images[:d.n_images]):
        z[f"image_{k}"] = np.frombuffer(image_files[k], np.uint8) if k in image_files else \
            np.frombuffer(C.string_at(im.rgba8, im.width * im.height * 4), np.uint8)
    if d.n_skins:
        sk = np.frombuffer(_arr(d.skins, d.n_skins * 4096, F.c_f), np.float32).reshape(d.n_skins, 256, 16)
        used = int(np.nonzero(np.abs(sk).sum(axis=(0, 2)))[0].max()) + 1
        z["skins"] = sk[:, :used].copy(); z["n_skins"] = np.uint32(d.n_skins)
    np.savez_compressed(path, **z)


def expand_skins(sk: np.ndarray) -> np.ndarray:
    """[n_skins, used_joints, 16] -> SkinRaw layout [n_skins, 256, 16] (unused joints zero, as the loader leaves them)."""
    out = np.zeros((sk.shape[0], 256, 16), np.float32)
    out[:, :sk.shape[1]] = sk
    return out


def load_scene_full(path):
    """-> (rt_scene_desc, npz).  Encoded images are decoded with the host layer's own PNG / JPEG decoder."""
    from rustracer_b200 import host
    z = np.load(path)
    keep = {k: np.ascontiguousarray(z[k]) for k in ("indices", "prim_infos", "geometries", "materials", "instances", "dlights", "plights")}
    keep["v"] = np.frombuffer(z["vertices"].tobytes(), F.VERTEX_DTYPE).copy()
    meta = z["image_meta"]; px = []
    for k, (w, h, srgb, enc) in enumerate(meta):
        a = z[f"image_{k}"]
        img = host.decode_image(a.tobytes()) if enc else a.reshape(h, w, 4)
        assert img.shape == (h, w, 4), (img.shape, w, h)
        px.append(np.ascontiguousarray(img, np.uint8))
    keep["px"] = px
    keep["img"] = (F.rt_image_desc * max(1, len(px)))(*[F.rt_image_desc(p.ctypes.data_as(F.c_u8p), int(m[0]), int(m[1]), int(m[2]), 0) for p, m in zip(px, meta)])
    keep["smp"] = (F.rt_sampler_desc * max(1, len(z["samplers"])))(*[F.rt_sampler_desc(*[int(x) for x in s]) for s in z["samplers"]])
    keep["tex"] = (F.rt_texture_desc * max(1, len(z["textures"])))(*[F.rt_texture_desc(int(t[0]), int(t[1])) for t in z["textures"]])
    d = F.rt_scene_desc()
    d.vertices, d.n_vertices = F.as_ptr(keep["v"], F.rt_vertex), len(keep["v"])
    d.indices, d.n_indices = F.as_ptr(keep["indices"], F.c_u32), len(keep["indices"])
    d.prim_infos = keep["prim_infos"].ctypes.data_as(C.POINTER(F.rt_prim_info))
    d.geometries, d.n_geometries = keep["geometries"].ctypes.data_as(C.POINTER(F.rt_geometry)), len(keep["geometries"])
    d.materials, d.n_materials = keep["materials"].ctypes.data_as(C.POINTER(F.rt_material)), len(keep["materials"])
    d.instances, d.n_instances = F.as_ptr(keep["instances"], F.rt_instance), len(keep["instances"])
    d.images, d.n_images, d.samplers, d.n_samplers = keep["img"], len(px), keep["smp"], len(z["samplers"])
    d.textures, d.n_textures = keep["tex"], len(z["textures"])
    d.dlights, d.n_dlights = F.as_ptr(keep["dlights"], F.rt_light), len(keep["dlights"])
    d.plights, d.n_plights = F.as_ptr(keep["plights"], F.rt_light), len(keep["plights"])
    if "skins" in z.files:
        keep["sk"] = expand_skins(z["skins"])
        d.skins, d.n_skins = F.as_ptr(keep["sk"], F.c_f), int(z["n_skins"])
    d._keep = keep
    d.fully_opaque = bool((keep["materials"].view(np.uint32)[:, 0] == 1).all())
    return d, z
