"""Synthetic scene generators for BASELINE.json's configs.

Everything here is host-side input construction (numpy): it produces the reference's flat arrays (128-byte
Vertex, PrimInfo, 256-byte MaterialRaw, instances, lights) exactly as `asset_loader` would hand them to the
GPU, including the reference's scene normalisation rule (aabb.rs:65-69: longest side -> 10, centred).
No reference asset is read at run time (the GPU box does not have /root/reference).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _ffi as F


# ----------------------------------------------------------------------------------------------------
# materials
# ----------------------------------------------------------------------------------------------------
def material(base_color=(1, 1, 1, 1), metallic=1.0, roughness=1.0, emissive=(0, 0, 0), alpha_mode=1, alpha_cutoff=0.5,
             ior=1.5, transmission=None, volume=None, base_color_texture=-1, unlit=False) -> F.rt_material:
    """MaterialRaw with the loader's defaults (material.rs:162-190); glTF defaults metallic = roughness = 1."""
    m = F.rt_material()
    none = F.rt_texture_info(-1, -1)
    m.alpha_mode, m.alpha_cutoff, m.double_sided, m.workflow = alpha_mode, alpha_cutoff, 1, 0
    m.base_color_texture = F.rt_texture_info(base_color_texture, 0) if base_color_texture >= 0 else none
    m.base_color[:] = base_color
    m.metallic_factor, m.roughness_factor = metallic, roughness
    m.metallic_roughness_texture = m.normal_texture = m.emissive_texture = m.occlusion_texture = none
    m.emissive_factor[:] = (*emissive, 0.0)
    m.ior, m.unlit = ior, int(unlit)
    m.transmission_texture = none
    if transmission is not None:
        m.transmission_factor, m.transmission_exist = float(transmission), 1
    m.attenuation_color[:] = (1, 1, 1)
    m.thickness_texture = none
    m.attenuation_distance = 3.4028234663852886e38
    if volume is not None:
        att_color, att_dist = volume
        m.attenuation_color[:] = att_color
        m.thickness_factor, m.attenuation_distance, m.volume_exists = 1.0, float(att_dist), 1
    m.specular_texture = m.specular_color_texture = none
    m.specular_color_factor[:] = (1, 1, 1, 1)
    m.specular_factor = 1.0
    m.sg_diffuse_texture = m.sg_specular_glossiness_texture = none
    return m


# ----------------------------------------------------------------------------------------------------
# meshes (positions, normals, uvs, triangle indices)
# ----------------------------------------------------------------------------------------------------
def box_mesh(half):
    hx, hy, hz = half
    faces = [((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((-1, 0, 0), (0, 0, 1), (0, 1, 0)), ((0, 1, 0), (0, 0, 1), (1, 0, 0)),
             ((0, -1, 0), (1, 0, 0), (0, 0, 1)), ((0, 0, 1), (1, 0, 0), (0, 1, 0)), ((0, 0, -1), (0, 1, 0), (1, 0, 0))]
    pos, nrm, uv, idx = [], [], [], []
    h = np.array([hx, hy, hz], np.float32)
    for n, a, b in faces:
        n, a, b = np.array(n, np.float32), np.array(a, np.float32), np.array(b, np.float32)
        base = len(pos)
        for (su, sv) in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
            pos.append((n + su * a + sv * b) * h)
            nrm.append(n)
            uv.append(((su + 1) / 2, (sv + 1) / 2))
        idx += [base, base + 1, base + 2, base, base + 2, base + 3]
    return np.array(pos, np.float32), np.array(nrm, np.float32), np.array(uv, np.float32), np.array(idx, np.uint32)


def grid_indices(rows, cols, wrap_cols=True):
    """Triangles of a (rows+1) x cols vertex grid, columns optionally wrapping around."""
    r = np.arange(rows, dtype=np.uint32)[:, None]
    c = np.arange(cols if wrap_cols else cols - 1, dtype=np.uint32)[None, :]
    c1 = (c + 1) % cols if wrap_cols else c + 1
    v00, v01, v10, v11 = r * cols + c, r * cols + c1, (r + 1) * cols + c, (r + 1) * cols + c1
    tris = np.stack([np.stack([v00, v10, v11], -1), np.stack([v00, v11, v01], -1)], 2)
    return tris.reshape(-1).astype(np.uint32)


def uv_sphere(radius=1.0, stacks=64, slices=65):
    th = np.linspace(0.0, math.pi, stacks + 1, dtype=np.float64)[:, None]
    ph = (np.arange(slices, dtype=np.float64) * (2 * math.pi / slices))[None, :]
    n = np.stack([np.sin(th) * np.cos(ph), np.cos(th) * np.ones_like(ph), np.sin(th) * np.sin(ph)], -1).reshape(-1, 3)
    uv = np.stack([np.broadcast_to(ph / (2 * math.pi), (stacks + 1, slices)), np.broadcast_to(th / math.pi, (stacks + 1, slices))], -1).reshape(-1, 2)
    return (n * radius).astype(np.float32), n.astype(np.float32), uv.astype(np.float32), grid_indices(stacks, slices)


def lucy_standin(rows=574, cols=391, seed=0xC0FFEE):
    """Procedural statue-like closed surface filling Lucy's local AABB (the real blob is missing;
    min (-141.4,-257.1,-866.5) max (153.5,205.8,-11.9)); 2*rows*cols triangles (default 574 x 391 = 448 868, the triangle
    count of the real scan) with scan-like high-frequency relief."""
    rng = np.random.default_rng(seed)
    lo = np.array([-141.44540405273438, -257.1471862792969, -866.5087280273438])
    hi = np.array([153.5471954345703, 205.83340454101562, -11.90410041809082])
    t = np.linspace(0.0, 1.0, rows + 1)[:, None]                  # along the long (z) axis, feet -> head
    ph = (np.arange(cols) * (2 * math.pi / cols))[None, :]
    # silhouette: pedestal, robe, waist, shoulders, head
    prof = (0.55 * np.exp(-((t - 0.04) / 0.05) ** 2) + 0.42 * np.exp(-((t - 0.30) / 0.22) ** 2) + 0.30 * np.exp(-((t - 0.62) / 0.10) ** 2)
            + 0.36 * np.exp(-((t - 0.74) / 0.05) ** 2) + 0.16 * np.exp(-((t - 0.90) / 0.045) ** 2))
    prof = prof * np.sqrt(np.clip(np.sin(np.pi * t), 0.0, 1.0)) + 1e-3
    r = np.broadcast_to(prof, (rows + 1, cols)).copy()
    # wings / arms: angular lobes in the upper body
    r *= 1.0 + 0.9 * np.exp(-((t - 0.72) / 0.12) ** 2) * (np.maximum(np.cos(ph - 0.6), 0) ** 6 + np.maximum(np.cos(ph - 2.6), 0) ** 6)
    # folds and scan noise (multi-octave)
    for k in range(1, 9):
        fa, fb = rng.integers(3, 9) * k, rng.integers(2, 30) * k
        r *= 1.0 + (0.05 / k) * np.sin(fa * ph + rng.uniform(0, 6.28) + fb * t * 6.28) * np.sin(fb * t * 3.14 + rng.uniform(0, 6.28))
    x, y, z = r * np.cos(ph), r * np.sin(ph), np.broadcast_to(t, r.shape)
    p = np.stack([x, y, z], -1).reshape(-1, 3)
    pmin, pmax = p.min(0), p.max(0)
    p = lo + (p - pmin) / (pmax - pmin) * (hi - lo)
    idx = grid_indices(rows, cols)
    # smooth vertex normals from face normals
    tri = idx.reshape(-1, 3)
    fn = np.cross(p[tri[:, 1]] - p[tri[:, 0]], p[tri[:, 2]] - p[tri[:, 0]])
    n = np.zeros_like(p)
    for k in range(3):
        np.add.at(n, tri[:, k], fn)
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    n = np.where(ln > 0, n / np.maximum(ln, 1e-30), np.array([0.0, 0.0, 1.0]))
    uv = np.stack([np.broadcast_to(ph / (2 * math.pi), r.shape), np.broadcast_to(t, r.shape)], -1).reshape(-1, 2)
    return p.astype(np.float32), n.astype(np.float32), uv.astype(np.float32), idx


# ----------------------------------------------------------------------------------------------------
# transforms (column-vector convention, 4x4 numpy float64 -> instance 3x4 float32)
# ----------------------------------------------------------------------------------------------------
def trs(translation=(0, 0, 0), rotation=(0, 0, 0, 1), scale=(1, 1, 1)) -> np.ndarray:
    x, y, z, w = rotation
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    M = np.eye(4)
    M[:3, :3] = R * np.asarray(scale, np.float64)[None, :]
    M[:3, 3] = translation
    return M


class SceneBuilder:
    """Collects geometries / materials / instances and emits an rt_scene_desc backed by numpy arrays."""

    def __init__(self):
        self.verts, self.idx, self.prims, self.geos, self.geo_bounds = [], [], [], [], []
        self.materials: list[F.rt_material] = []
        self.instances: list[tuple[int, np.ndarray]] = []
        self.images, self.samplers, self.textures = [], [], []
        self.dlights = self.plights = None
        self.skins = None
        self.sky = None
        self._nv = self._ni = 0
        # slot 0 = the loader's dummy texture / sampler / image (texture.rs:21-41, image.rs:31-43)
        self.add_image(np.full((1, 1, 4), 1, np.uint8), srgb=True)
        self.samplers.append((F.rt_sampler_desc(1, 1, 2, 2)))
        self.textures.append(F.rt_texture_desc(0, 0))

    def add_material(self, m: F.rt_material) -> int:
        self.materials.append(m)
        return len(self.materials) - 1

    def add_image(self, rgba: np.ndarray, srgb=True) -> int:
        self.images.append((np.ascontiguousarray(rgba, np.uint8), int(srgb)))
        return len(self.images) - 1

    def add_texture(self, image: int, mag=1, wrap_s=2, wrap_t=2) -> int:
        self.samplers.append(F.rt_sampler_desc(mag, mag, wrap_s, wrap_t))
        self.textures.append(F.rt_texture_desc(image, len(self.samplers) - 1))
        return len(self.textures) - 1

    def add_geometry(self, pos, nrm, uv, idx, material_id: int, color=None, weights=None, joints=None, skin_index=-1) -> int:
        n = len(pos)
        v = np.zeros(n, F.VERTEX_DTYPE)
        v["position"][:, :3] = pos
        v["normal"][:, :3] = nrm
        v["tangent"][:, 0] = 1.0
        v["color"] = 1.0 if color is None else color
        v["uv0"] = uv
        v["skin_index"] = skin_index
        if weights is not None:
            v["weights"], v["joints"] = weights, joints
        self.verts.append(v)
        self.idx.append(np.ascontiguousarray(idx, np.uint32))
        self.prims.append((self._nv, self._ni, material_id))
        self.geos.append((n, len(idx), 1 if self.materials[material_id].alpha_mode == 1 else 0))
        self.geo_bounds.append((np.asarray(pos).min(0), np.asarray(pos).max(0)))
        self._nv += n
        self._ni += len(idx)
        return len(self.geos) - 1

    def add_instance(self, geo_id: int, world: np.ndarray):
        self.instances.append((geo_id, np.asarray(world, np.float64)))

    def normalize(self):
        """The reference's load_scene rule (scene_graph.rs:277-289, aabb.rs): only the min and max corners of each
        mesh box are transformed; the union is scaled so its longest side is 10 and centred at the origin."""
        lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
        for g, M in self.instances:
            bl, bh = self.geo_bounds[g]
            a = (M @ np.array([*bl, 1.0]))[:3]
            b = (M @ np.array([*bh, 1.0]))[:3]
            lo, hi = np.minimum(lo, np.minimum(a, b)), np.maximum(hi, np.maximum(a, b))
        size = np.abs(hi - lo)
        larger = size[0] if (size[0] > size[1] and size[0] > size[2]) else (size[1] if size[1] > size[2] else size[2])
        centre = lo + (hi - lo) / 2
        T = np.eye(4)
        T[:3, 3] = -centre
        S = np.diag([10.0 / larger] * 3 + [1.0])
        A = S @ T
        self.instances = [(g, A @ M) for g, M in self.instances]
        return A

    def set_lights(self, dlights=None, plights=None):
        self.dlights, self.plights = dlights, plights

    def default_lights(self):
        """get_lights_raw (scene_graph.rs:55-81): 5 zero-intensity point lights + 1 zero-intensity directional."""
        rng = np.random.default_rng(7)
        pl = np.zeros(5, F.LIGHT_DTYPE)
        pl["color"], pl["kind"], pl["range"], pl["intensity"] = 1.0, 1, np.inf, 0.0
        pl["transform"] = (rng.random((5, 4)).astype(np.float32) - 0.5) * 20.0
        dl = np.zeros(1, F.LIGHT_DTYPE)
        dl["color"], dl["transform"], dl["kind"], dl["range"], dl["intensity"] = 1.0, 1.0, 0, np.inf, 0.0
        return dl, pl

    def build(self) -> F.rt_scene_desc:
        d = F.rt_scene_desc()
        keep = {}
        keep["v"] = np.concatenate(self.verts) if self.verts else np.zeros(0, F.VERTEX_DTYPE)
        keep["i"] = np.concatenate(self.idx) if self.idx else np.zeros(0, np.uint32)
        keep["p"] = (F.rt_prim_info * max(1, len(self.prims)))(*[F.rt_prim_info(*p, 0) for p in self.prims])
        keep["g"] = (F.rt_geometry * max(1, len(self.geos)))(*[F.rt_geometry(*g, 0) for g in self.geos])
        keep["m"] = (F.rt_material * len(self.materials))(*self.materials)
        inst = np.zeros(len(self.instances), F.INSTANCE_DTYPE)
        for k, (g, M) in enumerate(self.instances):
            inst["transform"][k] = M[:3, :].astype(np.float32).reshape(12)
            inst["geo_id"][k], inst["mask"][k], inst["flags"][k] = g, 0xFF, 1
        keep["inst"] = inst
        keep["img_px"] = [im for im, _ in self.images]
        keep["img"] = (F.rt_image_desc * len(self.images))(*[
            F.rt_image_desc(im.ctypes.data_as(F.c_u8p), im.shape[1], im.shape[0], srgb, 0) for im, srgb in self.images])
        keep["smp"] = (F.rt_sampler_desc * len(self.samplers))(*self.samplers)
        keep["tex"] = (F.rt_texture_desc * len(self.textures))(*self.textures)
        dl, pl = self.default_lights()
        keep["dl"] = np.ascontiguousarray(self.dlights if self.dlights is not None else dl, F.LIGHT_DTYPE)
        keep["pl"] = np.ascontiguousarray(self.plights if self.plights is not None else pl, F.LIGHT_DTYPE)
        d.vertices, d.n_vertices = F.as_ptr(keep["v"], F.rt_vertex), len(keep["v"])
        d.indices, d.n_indices = F.as_ptr(keep["i"], F.c_u32), len(keep["i"])
        d.prim_infos, d.geometries, d.n_geometries = keep["p"], keep["g"], len(self.geos)
        d.materials, d.n_materials = keep["m"], len(self.materials)
        d.instances, d.n_instances = F.as_ptr(inst, F.rt_instance), len(inst)
        d.images, d.n_images = keep["img"], len(self.images)
        d.samplers, d.n_samplers = keep["smp"], len(self.samplers)
        d.textures, d.n_textures = keep["tex"], len(self.textures)
        d.dlights, d.n_dlights = F.as_ptr(keep["dl"], F.rt_light), len(keep["dl"])
        d.plights, d.n_plights = F.as_ptr(keep["pl"], F.rt_light), len(keep["pl"])
        if self.skins is not None:
            keep["sk"] = np.ascontiguousarray(self.skins, np.float32)
            d.skins, d.n_skins = F.as_ptr(keep["sk"], F.c_f), keep["sk"].size // 4096
        if self.sky is not None:
            faces, srgb = self.sky
            keep["sky"] = [np.ascontiguousarray(f, np.uint8) for f in faces]
            for k in range(6):
                d.skybox_faces[k] = keep["sky"][k].ctypes.data_as(F.c_u8p)
            d.skybox_height, d.skybox_width, d.skybox_srgb = keep["sky"][0].shape[0], keep["sky"][0].shape[1], int(srgb)
        d._keep = keep
        d.fully_opaque = all(m.alpha_mode == 1 for m in self.materials)
        return d


# ----------------------------------------------------------------------------------------------------
# named scenes
# ----------------------------------------------------------------------------------------------------
def cornell_box(lucy: bool = False, blend_sphere: bool = True, lucy_rows=574, lucy_cols=391, shell=None) -> F.rt_scene_desc:
    """Cornell box with the node layout and materials of assets/models/CornellBox/cornellBox.gltf (config 1) or
    CornellBoxLucy/cornellBoxLucy.gltf (config 2, `lucy=True`: all materials OPAQUE, glass sphere and a ~448k-triangle
    procedural stand-in for the missing Lucy scan, deviation D4).
    shell: path of an .npz written by tests/util.save_scene_npz from the real cornellBox.gltf / cornellBox.bin
    (tests/golden/cornell_box_scene.npz): the nine shell meshes (walls, light, boxes, two spheres) then carry the
    asset's own vertices / normals / uvs / tangents / indices ("Cornell shell from
    cornellBox.bin"); without it the shell is procedural (same node transforms, 64x65 uv spheres)."""
    b = SceneBuilder()
    shell_z = np.load(shell) if shell is not None else None

    def shell_mesh(k, fallback):
        if shell_z is None:
            return fallback
        vo, io = int(shell_z["prim_infos"][k][0]), int(shell_z["prim_infos"][k][1])
        nv, ni = int(shell_z["geometries"][k][0]), int(shell_z["geometries"][k][1])
        return (shell_z["position"][vo:vo + nv], shell_z["normal"][vo:vo + nv], shell_z["uv0"][vo:vo + nv], shell_z["indices"][io:io + ni])

    def shell_tangents():
        if shell_z is None:
            return
        for k in range(9):
            vo, nv = int(shell_z["prim_infos"][k][0]), int(shell_z["geometries"][k][0])
            b.verts[k]["tangent"] = shell_z["tangent"][vo:vo + nv]
    white = b.add_material(material(metallic=0.0))
    white2 = b.add_material(material(metallic=0.0))
    green = b.add_material(material((0.054592281579971313, 1.0, 0.0, 1.0), metallic=0.0))
    red = b.add_material(material((1.0, 0.0, 0.00010718735575210303, 1.0), metallic=0.0))
    light = b.add_material(material(metallic=0.0, roughness=0.0, emissive=(1, 1, 1)))
    grey = b.add_material(material((0.5, 0.5, 0.5, 1.0)))
    default = b.add_material(material())
    sph_a = b.add_material(material((0.949999988079071, 0.949999988079071, 0.949999988079071, 0.050000011920928955), roughness=0.0,
                                    alpha_mode=1 if (lucy or not blend_sphere) else 3))
    if lucy:
        sph_b = b.add_material(material(metallic=0.0, roughness=0.0, ior=1.5, transmission=1.0, volume=((0.9, 0.9, 0.9), 1.8)))
        lucy_m = b.add_material(material(metallic=0.1, roughness=0.0, ior=1.5, transmission=1.0, volume=((0.9, 0.9, 0.9), 1.0)))
    else:
        sph_b = b.add_material(material(roughness=0.0, ior=1.76))
    e = 0.05000000074505806
    g_back = b.add_geometry(*shell_mesh(0, box_mesh((5, 5, e))), white)
    g_floor = b.add_geometry(*shell_mesh(1, box_mesh((5, e, 5))), white2)
    g_left = b.add_geometry(*shell_mesh(2, box_mesh((e, 5, 5))), green)
    g_right = b.add_geometry(*shell_mesh(3, box_mesh((e, 5, 5))), red)
    g_light = b.add_geometry(*shell_mesh(4, box_mesh((0.5, e, 0.5))), light)
    g_cube = b.add_geometry(*shell_mesh(5, box_mesh((0.5, 0.5, 0.5))), grey)
    g_tall = b.add_geometry(*shell_mesh(6, box_mesh((1.25, 3.0, 1.25))), default)
    g_sa = b.add_geometry(*shell_mesh(7, uv_sphere()), sph_a)
    g_sb = b.add_geometry(*shell_mesh(8, uv_sphere()), sph_b)
    shell_tangents()
    b.add_instance(g_back, trs((0, 0, -5)))
    b.add_instance(g_floor, trs((0, -5, 0)))
    b.add_instance(g_floor, trs((0, 5, 0)))          # mesh 1 is instanced twice (floor and ceiling)
    b.add_instance(g_left, trs((-5, 0, 0)))
    b.add_instance(g_right, trs((5, 0, 0)))
    b.add_instance(g_light, trs((0, 4.626189708709717, 0), scale=(3, 1, 3)))
    b.add_instance(g_cube, trs((3, -3.5, 1.6790000200271606), (0, 0.46174857020378113, 0, 0.8870108127593994), (2.5, 2.5, 2.5)))
    b.add_instance(g_tall, trs((-2.5, -2, -2), (0, 0.026176944375038147, 0, 0.9996573328971863)))
    b.add_instance(g_sa, trs((3.311300039291382, -1, 0.08760000020265579)))
    b.add_instance(g_sb, trs((-3, -4, 3)))
    if lucy:
        g_lucy = b.add_geometry(*lucy_standin(lucy_rows, lucy_cols), lucy_m)
        b.add_instance(g_lucy, trs((-0.33713316917419434, -5.119440078735352, 1.06102454662323),
                                   (0.6395068764686584, 0.30171337723731995, -0.30171340703964233, 0.6395068764686584),
                                   (0.007677134592086077,) * 3))
    b.normalize()
    return b.build()


# ----------------------------------------------------------------------------------------------------
# config 3: instanced foliage
# ----------------------------------------------------------------------------------------------------
def leaf_mask_texture(size=1024, seed=11) -> np.ndarray:
    """Procedural RGBA8 sRGB leaf-mask: green leaves with alpha 255 inside blobs, 0 outside."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:size, 0:size].astype(np.float32) / size
    a = np.zeros((size, size), np.float32)
    for _ in range(24):
        cx, cy, r, e = rng.uniform(0.1, 0.9), rng.uniform(0.1, 0.9), rng.uniform(0.05, 0.16), rng.uniform(0.3, 1.0)
        a = np.maximum(a, (((x - cx) / r) ** 2 + ((y - cy) / (r * e)) ** 2 < 1.0).astype(np.float32))
    img = np.zeros((size, size, 4), np.uint8)
    img[..., 0] = 40 + 30 * np.sin(x * 40); img[..., 1] = 140 + 60 * np.cos(y * 33); img[..., 2] = 50
    img[..., 3] = (a * 255).astype(np.uint8)
    return img


def displaced_icosphere_like(n_tris_target=100_000, seed=1):
    """One BLAS: displaced sphere with ~n_tris_target triangles (grid tessellation)."""
    rows = max(4, int(math.sqrt(n_tris_target / 2)))
    cols = max(4, n_tris_target // (2 * rows))
    rng = np.random.default_rng(seed)
    p, n, uv, idx = uv_sphere(1.0, rows, cols)
    th = np.arccos(np.clip(n[:, 1], -1, 1)); ph = np.arctan2(n[:, 2], n[:, 0])
    d = np.ones(len(p))
    for k in range(1, 6):
        d += (0.12 / k) * np.sin(k * 5 * th + rng.uniform(0, 6.28)) * np.cos(k * 7 * ph + rng.uniform(0, 6.28))
    return (p * d[:, None].astype(np.float32)), n, uv, idx


def foliage_cards(n_cards=64, seed=5):
    """Alpha-tested quads (2 triangles each) with random orientation around the unit sphere."""
    rng = np.random.default_rng(seed)
    pos, nrm, uv, idx = [], [], [], []
    for c in range(n_cards):
        centre = rng.normal(size=3); centre *= rng.uniform(1.0, 1.6) / np.linalg.norm(centre)
        t = rng.normal(size=3); t -= centre * np.dot(t, centre) / np.dot(centre, centre); t /= np.linalg.norm(t)
        b = np.cross(centre / np.linalg.norm(centre), t)
        s = rng.uniform(0.25, 0.5)
        base = len(pos)
        for (su, sv) in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
            pos.append(centre + s * (su * t + sv * b)); nrm.append(centre / np.linalg.norm(centre)); uv.append(((su + 1) / 2 * 0.5 + rng.integers(2) * 0.5, (sv + 1) / 2))
        idx += [base, base + 1, base + 2, base, base + 2, base + 3]
    return np.array(pos, np.float32), np.array(nrm, np.float32), np.array(uv, np.float32), np.array(idx, np.uint32)


def instanced_foliage(n_side=100, tris_per_mesh=100_000, cards=64, tex_size=1024, seed=2, sky=None) -> F.rt_scene_desc:
    """Config 3: n_side^2 instances of one ~tris_per_mesh BLAS on a jittered grid (random yaw, scale in [0.5,1.5]),
    a metallic-roughness material and an alpha-MASK foliage-card geometry with a procedural leaf texture."""
    rng = np.random.default_rng(seed)
    b = SceneBuilder()
    metal = b.add_material(material((0.9, 0.6, 0.3, 1.0), metallic=1.0, roughness=0.25))
    img = b.add_image(leaf_mask_texture(tex_size), srgb=True)
    tex = b.add_texture(img, mag=1, wrap_s=2, wrap_t=2)
    leaf = b.add_material(material((1, 1, 1, 1), metallic=0.0, roughness=0.8, alpha_mode=2, alpha_cutoff=0.5, base_color_texture=tex))
    ground = b.add_material(material((0.4, 0.4, 0.4, 1), metallic=0.0, roughness=0.9))
    g_body = b.add_geometry(*displaced_icosphere_like(tris_per_mesh, seed=1), metal)
    g_leaf = b.add_geometry(*foliage_cards(cards), leaf)
    g_ground = b.add_geometry(*box_mesh((n_side * 2.0, 0.1, n_side * 2.0)), ground)
    b.add_instance(g_ground, trs((0, -1.8, 0)))
    for i in range(n_side):
        for j in range(n_side):
            yaw = rng.uniform(0, 2 * math.pi); s = rng.uniform(0.5, 1.5)
            T = trs(((i - n_side / 2 + rng.uniform(-0.3, 0.3)) * 3.5, 0.0, (j - n_side / 2 + rng.uniform(-0.3, 0.3)) * 3.5),
                    (0, math.sin(yaw / 2), 0, math.cos(yaw / 2)), (s, s, s))
            b.add_instance(g_body, T); b.add_instance(g_leaf, T)
    b.normalize()
    dl = np.zeros(1, F.LIGHT_DTYPE); dl["color"], dl["transform"], dl["kind"], dl["range"], dl["intensity"] = 1.0, [[-0.4, -1.0, -0.3, 0.0]], 0, np.inf, 1.0
    b.set_lights(dlights=dl)
    if sky is not None:
        b.sky = (sky, True)
    return b.build()


def procedural_sky(size=64, seed=3):
    """Six RGBA8 faces (+x,-x,+y,-y,+z,-z): vertical gradient with a bright patch, stands in for the Yokohama skybox."""
    rng = np.random.default_rng(seed)
    faces = []
    for f in range(6):
        y = np.linspace(0, 1, size, dtype=np.float32)[:, None] * np.ones((1, size), np.float32)
        img = np.zeros((size, size, 4), np.uint8)
        base = np.array([90, 140, 220]) if f != 3 else np.array([60, 50, 40])
        img[..., :3] = np.clip(base[None, None, :] * (1.2 - 0.6 * y[..., None]) + rng.integers(0, 6, (size, size, 3)), 0, 255)
        if f == 2:
            img[size // 3: size // 2, size // 3: size // 2, :3] = 255
        img[..., 3] = 255
        faces.append(img)
    return faces


# ----------------------------------------------------------------------------------------------------
# config 4: skinned character
# ----------------------------------------------------------------------------------------------------
def skinned_character(n_tris=1_000_000, joints=256, seed=3):
    """Capsule-limb rig: one skin x `joints` joints, 4 normalised weights per vertex.  Returns (desc, pose_fn) where
    pose_fn(frame) -> (1,256,16) column-major skin matrices of a looping 60-frame animation."""
    rng = np.random.default_rng(seed)
    rows = max(8, int(math.sqrt(n_tris / 2) * 2)); cols = max(8, n_tris // (2 * rows))
    p, n, uv, idx = uv_sphere(1.0, rows, cols)
    p = p * np.array([0.6, 3.0, 0.6], np.float32)          # tall capsule, y in [-3, 3]
    t = (p[:, 1] + 3.0) / 6.0 * (joints - 1)
    j0 = np.clip(np.floor(t).astype(np.int64), 1, joints - 3)
    w = np.zeros((len(p), 4), np.float32); jj = np.zeros((len(p), 4), np.uint32)
    f = (t - j0).astype(np.float32)
    w[:, 0] = np.clip(0.5 - 0.5 * f, 0, 1) * 0.5; w[:, 1] = 1 - f * 0.5 - w[:, 0]; w[:, 2] = f * 0.5; w[:, 3] = 0.0
    w[:, 3] = rng.uniform(0, 0.1, len(p)).astype(np.float32)
    w /= w.sum(1, keepdims=True)
    jj[:, 0], jj[:, 1], jj[:, 2], jj[:, 3] = j0 - 1, j0, j0 + 1, j0 + 2
    b = SceneBuilder()
    skin_m = b.add_material(material((0.8, 0.7, 0.6, 1), metallic=0.0, roughness=0.6))
    g = b.add_geometry(p, n, uv, idx, skin_m, weights=w, joints=jj, skin_index=0)
    b.add_instance(g, np.eye(4))                            # skinned nodes get the identity instance transform
    floor = b.add_material(material((0.5, 0.5, 0.5, 1), metallic=0.0))
    gf = b.add_geometry(*box_mesh((6, 0.1, 6)), floor)
    b.add_instance(gf, trs((0, -3.6, 0)))
    lamp = b.add_material(material(metallic=0.0, roughness=0.0, emissive=(1, 1, 1)))
    gl = b.add_geometry(*box_mesh((2, 0.05, 2)), lamp)
    b.add_instance(gl, trs((0, 5.0, 0)))

    def pose(frame: int):
        ph = 2 * math.pi * (frame % 60) / 60.0
        mats = np.zeros((1, 256, 16), np.float32)
        mats[0, :, [0, 5, 10, 15]] = 1.0
        for j in range(joints):
            a = 0.25 * math.sin(ph + 0.05 * j) * (j / joints)
            M = trs((0.4 * math.sin(ph + 0.03 * j) * (j / joints), 0.0, 0.3 * math.cos(ph * 2 + 0.02 * j) * (j / joints)), (0, 0, math.sin(a / 2), math.cos(a / 2)))
            mats[0, j] = M.T.reshape(16).astype(np.float32)
        return mats
    b.skins = pose(0)
    return b.build(), pose


# ----------------------------------------------------------------------------------------------------
# config 5: transmission / volume scene
# ----------------------------------------------------------------------------------------------------
def glass_box(n_objects=64, seed=9, sphere_res=(48, 49)) -> F.rt_scene_desc:
    """Cornell-type box holding n_objects glass spheres (transmission 1, volume on, attDist in [0.5,2], random
    attenuation colour, ior 1.5, roughness 0) and one emissive quad."""
    rng = np.random.default_rng(seed)
    b = SceneBuilder()
    white = b.add_material(material(metallic=0.0)); red = b.add_material(material((0.9, 0.1, 0.1, 1), metallic=0.0))
    green = b.add_material(material((0.1, 0.9, 0.1, 1), metallic=0.0)); light = b.add_material(material(metallic=0.0, roughness=0.0, emissive=(1, 1, 1)))
    e = 0.05
    for geo, T in ((box_mesh((5, 5, e)), trs((0, 0, -5))), (box_mesh((5, e, 5)), trs((0, -5, 0))), (box_mesh((5, e, 5)), trs((0, 5, 0)))):
        b.add_instance(b.add_geometry(*geo, white), T)
    b.add_instance(b.add_geometry(*box_mesh((e, 5, 5)), green), trs((-5, 0, 0)))
    b.add_instance(b.add_geometry(*box_mesh((e, 5, 5)), red), trs((5, 0, 0)))
    b.add_instance(b.add_geometry(*box_mesh((1.5, e, 1.5)), light), trs((0, 4.9, 0)))
    side = int(math.ceil(n_objects ** (1 / 3)))
    k = 0
    for ix in range(side):
        for iy in range(side):
            for iz in range(side):
                if k >= n_objects:
                    break
                m = b.add_material(material(metallic=0.0, roughness=0.0, ior=1.5, transmission=1.0,
                                            volume=(tuple(rng.uniform(0.5, 1.0, 3)), rng.uniform(0.5, 2.0))))
                g = b.add_geometry(*uv_sphere(1.0, *sphere_res), m)
                r = 3.6 / side
                c = (np.array([ix, iy, iz]) + 0.5) / side * 8.0 - 4.0 + rng.uniform(-0.15, 0.15, 3)
                b.add_instance(g, trs(c, scale=(r, r, r)))
                k += 1
    b.normalize()
    return b.build()
