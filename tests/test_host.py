"""Host layer (host/gltf_host.cpp): glTF import rules, scene normalisation, camera / UBO construction, animation."""
import base64
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

import util
from rustracer_b200 import _ffi as F, host

REF = Path("/root/reference/assets/models")


def arr(ptr, n, ctype, dtype):
    return np.frombuffer(util._arr(ptr, n, ctype), dtype)


def write_gltf(tmp_path, with_ext=True):
    """Tiny self-contained glTF: two triangles (u16 indices, u8-normalised colours), TRS node, material extensions."""
    pos = np.array([[0, 0, 0], [2, 0, 0], [0, 2, 0], [2, 2, 0]], np.float32)
    col = np.array([[255, 0, 0, 255], [0, 255, 0, 255], [0, 0, 255, 255], [255, 255, 255, 128]], np.uint8)
    idx = np.array([0, 1, 2, 2, 1, 3], np.uint16)
    blob = pos.tobytes() + col.tobytes() + idx.tobytes()
    ext = {"KHR_materials_ior": {"ior": 1.33}, "KHR_materials_transmission": {"transmissionFactor": 0.9},
           "KHR_materials_volume": {"attenuationDistance": 2.5, "attenuationColor": [0.9, 0.8, 0.7], "thicknessFactor": 1.0}} if with_ext else {}
    g = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}],
         "nodes": [{"mesh": 0, "translation": [1, 2, 3], "scale": [2, 2, 2], "rotation": [0, 0, 0.70710678, 0.70710678]}],
         "meshes": [{"primitives": [{"attributes": {"POSITION": 0, "COLOR_0": 1}, "indices": 2, "material": 0}]}],
         "materials": [{"alphaMode": "MASK", "alphaCutoff": 0.3, "pbrMetallicRoughness": {"baseColorFactor": [0.5, 0.6, 0.7, 0.8], "metallicFactor": 0.25},
                        "emissiveFactor": [0.1, 0.2, 0.3], "extensions": ext}],
         "accessors": [{"bufferView": 0, "componentType": 5126, "count": 4, "type": "VEC3", "min": [0, 0, 0], "max": [2, 2, 0]},
                       {"bufferView": 1, "componentType": 5121, "normalized": True, "count": 4, "type": "VEC4"},
                       {"bufferView": 2, "componentType": 5123, "count": 6, "type": "SCALAR"}],
         "bufferViews": [{"buffer": 0, "byteOffset": 0, "byteLength": 48}, {"buffer": 0, "byteOffset": 48, "byteLength": 16}, {"buffer": 0, "byteOffset": 64, "byteLength": 12}],
         "buffers": [{"byteLength": len(blob), "uri": "data:application/octet-stream;base64," + base64.b64encode(blob).decode()}]}
    p = tmp_path / "tiny.gltf"
    p.write_text(json.dumps(g))
    return p


def test_tiny_gltf_import_rules(tmp_path):
    doc = host.load_file(write_gltf(tmp_path))
    d = doc.scene_desc()
    assert (d.n_vertices, d.n_indices, d.n_geometries, d.n_instances, d.n_materials) == (4, 6, 1, 1, 1)
    v = arr(d.vertices, 4, F.rt_vertex, F.VERTEX_DTYPE)
    np.testing.assert_allclose(v["color"][3], [1, 1, 1, 128 / 255], rtol=1e-6)            # normalised u8 colours
    assert (v["skin_index"] == -1).all() and (v["tangent"][:, 0] == 1).all()                # defaults geometry.rs:195,246
    np.testing.assert_allclose(v["normal"][:, :3], [[0, 0, 1]] * 4, atol=1e-6)               # create_geo_normal
    assert list(arr(d.indices, 6, F.c_u32, np.uint32)) == [0, 1, 2, 2, 1, 3]                 # u16 -> u32
    m = d.materials[0]
    assert m.alpha_mode == 2 and abs(m.alpha_cutoff - 0.3) < 1e-6 and not doc.fully_opaque()
    assert abs(m.ior - 1.33) < 1e-6 and m.transmission_exist == 1 and abs(m.transmission_factor - 0.9) < 1e-6
    assert m.volume_exists == 1 and abs(m.attenuation_distance - 2.5) < 1e-6 and abs(m.attenuation_color[1] - 0.8) < 1e-6
    assert m.base_color_texture.index == -1 and abs(m.metallic_factor - 0.25) < 1e-6 and abs(m.roughness_factor - 1.0) < 1e-6
    assert d.geometries[0].opaque == 0
    # default lights: 5 zero-intensity point lights + 1 zero-intensity directional (scene_graph.rs:69-79)
    assert d.n_plights == 5 and d.n_dlights == 1 and all(d.plights[i].intensity == 0 for i in range(5))
    # scene normalisation: only min/max corners are transformed; longest side -> 10, centred (aabb.rs:65-85)
    inst = arr(d.instances, 1, F.rt_instance, F.INSTANCE_DTYPE)
    M = inst["transform"][0].reshape(3, 4).astype(np.float64)
    # node = T(1,2,3) R(90 deg about z) S(2): local corners (0,0,0),(2,2,0) -> world (1,2,3) and (-3,6,3)
    lo, hi = np.array([-3, 2, 3.0]), np.array([1, 6, 3.0])
    s = 10.0 / 4.0
    expect = np.diag([s, s, s]) @ (np.array([[0, -2, 0], [2, 0, 0], [0, 0, 2.0]]))
    np.testing.assert_allclose(M[:, :3], expect, atol=1e-5)
    np.testing.assert_allclose(M[:, 3], s * (np.array([1, 2, 3.0]) - (lo + (hi - lo) / 2)), atol=1e-5)
    assert inst["mask"][0] == 0xFF and inst["geo_id"][0] == 0


def test_camera_and_ubo_match_reference_formulas():
    cam = host.Camera(1920, 1080).set(position=(1, 2, 5), direction=(0.1, -0.2, -1))
    V, P = cam.view_matrix().astype(np.float64), cam.projection_matrix().astype(np.float64)
    f = np.array([0.1, -0.2, -1]); f /= np.linalg.norm(f)
    s = np.cross(f, [0, 1, 0]); s /= np.linalg.norm(s); u = np.cross(s, f)
    e = np.array([1, 2, 5.0])
    np.testing.assert_allclose(V[:3, :3], np.stack([s, u, -f]), atol=1e-6)                  # look_at_rh
    np.testing.assert_allclose(V[:3, 3], [-s @ e, -u @ e, f @ e], atol=1e-5)
    t = np.tan(np.radians(60) / 2); a = 1920 / 1080; n, fa = 0.1, 10.0
    GL = np.array([[1 / (a * t), 0, 0, 0], [0, 1 / t, 0, 0], [0, 0, (fa + n) / (n - fa), 2 * fa * n / (n - fa)], [0, 0, -1, 0]])
    Cm = np.array([[1, 0, 0, 0], [0, -1, 0, 0], [0, 0, .5, .5], [0, 0, 0, 1.0]])                 # OPENGL_TO_VULKAN_RT camera.rs:116-118
    np.testing.assert_allclose(P, Cm @ GL, atol=1e-6)
    gui = host.Gui()
    drv = host.FrameDriver(cam, gui, True)
    u0 = drv.next_ubo()
    assert (u0.number_of_samples, u0.total_number_of_samples, u0.number_of_bounces, u0.frame_count) == (3, 3, 5, 0)   # Gui::new defaults
    assert u0.exposure == 5.0 and u0.random_seed == 3 and u0.has_sky == 0 and u0.antialiasing == 1 and u0.fully_opaque == 1
    MVI = np.array(u0.model_view_inverse[:], np.float64).reshape(4, 4).T
    np.testing.assert_allclose(MVI @ V, np.eye(4), atol=1e-5)
    PI = np.array(u0.projection_inverse[:], np.float64).reshape(4, 4).T
    np.testing.assert_allclose(PI @ P, np.eye(4), atol=1e-4)
    u1 = drv.next_ubo()
    assert (u1.total_number_of_samples, u1.frame_count) == (6, 1)
    # sample budget: stops at max_number_of_samples; mapping != Render disables accumulation and forces 1 bounce
    g2 = host.Gui(number_of_samples=4, max_number_of_samples=10)
    d2 = host.FrameDriver(cam, g2, True)
    assert [d2.next_ubo().number_of_samples for _ in range(4)] == [4, 4, 2, 0]
    g3 = host.Gui(mapping=5); d3 = host.FrameDriver(cam, g3, True)
    u = d3.next_ubo(); u = d3.next_ubo()
    assert u.number_of_bounces == 1 and u.total_number_of_samples == u.number_of_samples


def test_png_decoder_roundtrip():
    import io
    from PIL import Image
    rng = np.random.default_rng(0)
    lib = F.load_host()
    for mode, ch in (("RGBA", 4), ("RGB", 3), ("L", 1), ("LA", 2)):
        a = rng.integers(0, 256, (13, 17, ch), dtype=np.uint8)
        im = Image.fromarray(a[..., 0] if ch == 1 else a, mode)
        buf = io.BytesIO(); im.save(buf, "PNG")
        raw = buf.getvalue()
        out = F.c_u8p(); w, h = F.c_u32(), F.c_u32()
        data = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
        assert lib.gv_decode_png(data, len(raw), C.byref(out), C.byref(w), C.byref(h)) == 0, lib.gv_last_error()
        px = np.ctypeslib.as_array(out, shape=(h.value, w.value, 4)).copy(); lib.gv_free(out)
        assert (px == np.asarray(im.convert("RGBA"))).all(), mode


def test_load_errors_are_reported(tmp_path):
    with pytest.raises(host.HostError):
        host.load_file(tmp_path / "missing.gltf")
    (tmp_path / "bad.gltf").write_text("{ not json")
    with pytest.raises(host.HostError):
        host.load_file(tmp_path / "bad.gltf")


@pytest.mark.skipif(not REF.exists(), reason="reference assets only exist in the build container")
def test_bundled_cornell_box_matches_fixture(cornell_desc):
    d = host.load_file(REF / "CornellBox/cornellBox.gltf").scene_desc()
    assert (d.n_vertices, d.n_indices // 3, d.n_geometries, d.n_instances, d.n_materials) == (16808, 16720, 9, 10, 9)
    for name, ct, dt in (("vertices", F.rt_vertex, F.VERTEX_DTYPE), ("instances", F.rt_instance, F.INSTANCE_DTYPE)):
        a = arr(getattr(d, name), getattr(d, "n_" + name), ct, dt); b = arr(getattr(cornell_desc, name), getattr(cornell_desc, "n_" + name), ct, dt)
        assert a.tobytes() == b.tobytes(), name
    assert d.materials[7].alpha_mode == 3 and abs(d.materials[7].base_color[3] - 0.05) < 1e-6   # the BLEND sphere
    assert abs(d.materials[8].ior - 1.76) < 1e-6 and d.materials[4].emissive_factor[0] == 1.0


@pytest.mark.skipif(not REF.exists(), reason="reference assets only exist in the build container")
def test_skinned_animated_glb_end_to_end():
    """shadows.glb (CesiumMan): skin tagging, PNG texture, default-material fallback, spot light as point light,
    animation sampling -> skin matrices; the emulated CUDA core must agree with the oracle after animating."""
    from emu_lib import emu_api
    from oracle import orc
    from rustracer_b200 import core
    doc = host.load_file(REF / "shadows.glb")
    d = doc.scene_desc()
    v = arr(d.vertices, d.n_vertices, F.rt_vertex, F.VERTEX_DTYPE)
    assert (v["skin_index"] >= 0).sum() == 5697 and v["joints"].max() == 18 and d.n_skins == 1
    assert d.n_images == 2 and d.images[1].width == 1024 and d.images[1].srgb == 1 and d.n_plights == 1 and d.plights[0].intensity == 20.0
    assert not doc.static_scene() and doc.need_compute()
    inst = doc.get_instances()
    assert any(np.allclose(t.reshape(3, 4), np.eye(4)[:3]) for t in inst["transform"])           # skinned node -> identity instance
    o = orc.OracleScene(d)
    ctx = core.Context(64, 48, api=emu_api()); sc = core.Scene(ctx, d)
    k0 = doc.get_skins().copy()
    doc.animate(0.7)
    k1 = doc.get_skins()
    assert np.abs(k1 - k0).max() > 1e-3
    o.update_skins(k1); sc.update_skins(k1)
    np.testing.assert_allclose(sc.read_vertices()["position"], o.read_vertices(d.n_vertices)["position"], rtol=1e-6, atol=1e-6)
    rays, _ = util.random_rays(3000, seed=2, extent=3.0)
    a, b = sc.trace_closest(rays, 1), o.trace_closest(rays, 1)
    assert util.hits_equal(a, b).all()


SHADOWS = util.GOLDEN / "shadows_glb_scene.npz"


@pytest.mark.skipif(not REF.exists(), reason="reference assets only exist in the build container")
def test_shadows_glb_fixture_is_what_the_loader_produces():
    """tests/golden/shadows_glb_scene.npz (committed, drives the GPU-box tests) == the host loader's output for the
    reference's shadows.glb, including the loader-driven animation frames (make_shadows_fixture.py)."""
    d, z = util.load_scene_full(SHADOWS)
    doc = host.load_file(REF / "shadows.glb"); ld = doc.scene_desc()
    for name, ct in (("vertices", F.rt_vertex), ("instances", F.rt_instance), ("materials", F.rt_material), ("plights", F.rt_light), ("dlights", F.rt_light)):
        assert util._arr(getattr(d, name), getattr(d, "n_" + name), ct) == util._arr(getattr(ld, name), getattr(ld, "n_" + name), ct), name
    assert util._arr(d.indices, d.n_indices, F.c_u32) == util._arr(ld.indices, ld.n_indices, F.c_u32)
    assert d.n_images == ld.n_images == 2 and C.string_at(d.images[1].rgba8, 1024 * 1024 * 4) == C.string_at(ld.images[1].rgba8, 1024 * 1024 * 4)
    for f in (0, 13, 59):
        doc.animate(float(z["anim_times"][f]))
        assert (util.expand_skins(z["anim_skins"][f]) == doc.get_skins()).all(), f
        assert z["anim_instances"][f].tobytes() == doc.get_instances().tobytes(), f


def test_shadows_glb_fixture_animation_oracle_vs_emulation():
    """The committed shadows.glb fixture (CesiumMan, 19 joints, textured, point light of intensity 20) animated by its
    loader-produced skin matrices: skinned vertices bit-exact and closest hits bit-exact, CUDA sources (host emulation) vs
    the oracle.  The same fixture drives rt_scene_update_skins on the B200 in tests/test_gpu_parity.py."""
    from emu_lib import emu_api
    from oracle import orc
    from rustracer_b200 import core
    import parity_cases as pc
    d, z = util.load_scene_full(SHADOWS)
    assert d.n_skins == 1 and d.n_plights == 1 and d.plights[0].intensity == 20.0 and z["anim_skins"].shape[:3] == (60, 1, 19)
    pc.case_shadows_glb(emu_api(), d, z, frames=(7, 41), size=40, n_rays=3000)


def _test_image(w, h, seed=1):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    a = np.stack([(np.sin(x / 7.0) + np.cos(y / 5.0)) * 60 + 128, (x * 3 + y * 2) % 256, rng.integers(0, 255, (h, w)) * 0.3 + 100], -1)
    return np.clip(a, 0, 255).astype(np.uint8)


def test_jpeg_decoder_matches_libjpeg():
    """Baseline + progressive, 4:4:4 / 4:2:2 / 4:2:0 / grey, odd sizes, restart markers: bit-identical to libjpeg-turbo
    (same integer IDCT, triangle upsampling and YCbCr conversion).  The reference's decoder (jpeg-decoder 0.3.0) is not
    available; T.81 allows decoders to differ by an LSB or two, see host/gltf_host.cpp."""
    Image = pytest.importorskip("PIL.Image")
    import io
    cases = 0
    for (w, h) in [(64, 48), (67, 35), (16, 16), (200, 133)]:
        for sub in (0, 1, 2):
            for prog in (False, True):
                for grey in (False, True):
                    im = Image.fromarray(_test_image(w, h))
                    kw = dict(quality=85, progressive=prog, optimize=prog)
                    if grey:
                        im = im.convert("L")
                    else:
                        kw["subsampling"] = sub
                    b = io.BytesIO(); im.save(b, "JPEG", **kw)
                    ref = np.asarray(Image.open(io.BytesIO(b.getvalue())).convert("RGBA"))
                    got = host.decode_image(b.getvalue())
                    assert got.shape == ref.shape and np.array_equal(got, ref), (w, h, sub, prog, grey)
                    cases += 1
    b = io.BytesIO(); Image.fromarray(_test_image(128, 96)).save(b, "JPEG", quality=90, restart_marker_blocks=3)
    assert np.array_equal(host.decode_image(b.getvalue()), np.asarray(Image.open(io.BytesIO(b.getvalue())).convert("RGBA")))
    assert cases == 48
    with pytest.raises(host.HostError):
        host.decode_image(b.getvalue()[:200])          # truncated stream: error, not garbage


def test_skybox_directory_loader(tmp_path):
    """cubumap.rs:29-50,86-106: six .png/.jpg files, sorted by face name or alias; anything else is an error."""
    Image = pytest.importorskip("PIL.Image")
    names = ["right.png", "negx.jpg", "top.png", "negy.png", "posz.jpg", "back.png"]          # mixed aliases / formats
    want = []
    for i, n in enumerate(names):
        a = _test_image(32, 32, seed=i)
        Image.fromarray(a).save(tmp_path / n, quality=95)
        want.append(np.asarray(Image.open(tmp_path / n).convert("RGBA")))
    (tmp_path / "readme.txt").write_text("ignored")
    faces = host.load_skybox_dir(tmp_path)
    assert len(faces) == 6 and all(np.array_equal(f, w) for f, w in zip(faces, want))
    (tmp_path / "extra.png").write_bytes((tmp_path / "top.png").read_bytes())
    with pytest.raises(host.HostError):
        host.load_skybox_dir(tmp_path)                  # 7 images: resource_manager::load_cubemap asserts len == 6
    ref_dir = Path("/root/reference/assets/skyboxs/Yokohama")
    if ref_dir.is_dir():                                # the reference's own skybox (absent on the GPU box)
        faces = host.load_skybox_dir(ref_dir)
        for f, n in zip(faces, ("posx", "negx", "posy", "negy", "posz", "negz")):
            assert np.array_equal(f, np.asarray(Image.open(ref_dir / f"{n}.jpg").convert("RGBA"))), n


# ---- MikkTSpace restatement (host/mikktspace_gen.h; geometry.rs:192-212,296-350) ---------------------------------
def _gen_tangents(pos, nrm, uv, idx):
    v = np.zeros(len(pos), F.VERTEX_DTYPE)
    v["position"][:, :3] = pos; v["normal"][:, :3] = nrm; v["uv0"] = uv
    v["tangent"] = [1, 0, 0, 0]                                     # the reference's pre-fill (geometry.rs:195)
    idx = np.ascontiguousarray(idx, np.uint32).reshape(-1)
    F.load_host().gv_generate_tangents(F.as_ptr(v, F.rt_vertex), len(v), F.as_ptr(idx, F.c_u32), len(idx))
    return v["tangent"].copy()


def _grid(n, fn):
    """(n+1)^2 vertices over (u, v) in [0,1]^2, two CCW triangles per cell; fn(u, v) -> position, normal."""
    us, vs = np.meshgrid(np.linspace(0, 1, n + 1), np.linspace(0, 1, n + 1), indexing="xy")
    uv = np.stack([us.ravel(), vs.ravel()], 1).astype(np.float32)
    pos, nrm = fn(uv[:, 0].astype(np.float64), uv[:, 1].astype(np.float64))
    idx = []
    for j in range(n):
        for i in range(n):
            a = j * (n + 1) + i; b = a + 1; c = a + n + 1; d = c + 1
            idx += [a, b, d, a, d, c]
    return pos.astype(np.float32), nrm.astype(np.float32), uv, np.array(idx, np.uint32)


def test_mikktspace_plane_and_mirrored_uv():
    plane = lambda u, v: (np.stack([2 * u, 3 * v, 0 * u], 1), np.tile([0.0, 0.0, 1.0], (len(u), 1)))
    pos, nrm, uv, idx = _grid(6, plane)
    t = _gen_tangents(pos, nrm, uv, idx)
    # tangent = d(position)/du direction, unit length; UV orientation preserved -> w = -1 (the reference's inverted sign)
    np.testing.assert_allclose(t[:, :3], np.tile([1, 0, 0], (len(t), 1)), atol=1e-6)
    assert (t[:, 3] == -1).all()
    # mirror the u coordinate: dP/du flips and the orientation flag with it
    t2 = _gen_tangents(pos, nrm, np.stack([1 - uv[:, 0], uv[:, 1]], 1), idx)
    np.testing.assert_allclose(t2[:, :3], np.tile([-1, 0, 0], (len(t), 1)), atol=1e-6)
    assert (t2[:, 3] == 1).all()
    # a constant UV mapping has no tangent: such triangles cannot found a group, vertices keep the default frame
    t3 = _gen_tangents(pos, nrm, np.zeros_like(uv), idx)
    assert (t3 == [1, 0, 0, 1]).all()


def test_mikktspace_sphere_matches_analytic_tangent():
    def sphere(u, v):
        th, ph = 2 * np.pi * u, np.pi * (0.1 + 0.8 * v)              # stay away from the poles
        p = np.stack([np.sin(ph) * np.cos(th), np.cos(ph), np.sin(ph) * np.sin(th)], 1)
        return p, p
    pos, nrm, uv, idx = _grid(48, sphere)
    t = _gen_tangents(pos, nrm, uv, idx)
    th = 2 * np.pi * uv[:, 0].astype(np.float64)
    analytic = np.stack([-np.sin(th), 0 * th, np.cos(th)], 1)        # dP/du, normalised
    inner = np.ones(len(t), bool)
    inner[(uv[:, 0] == 0) | (uv[:, 0] == 1) | (uv[:, 1] == 0) | (uv[:, 1] == 1)] = False   # one-sided at the borders
    assert np.abs(np.linalg.norm(t[:, :3], axis=1) - 1).max() < 1e-5
    assert np.abs((t[:, :3] * nrm).sum(1)).max() < 1e-5              # in the tangent plane of the vertex normal
    assert (t[inner, :3] * analytic[inner]).sum(1).min() > 0.999
    assert len(set(t[:, 3])) == 1                                    # one orientation over the whole surface


def test_mikktspace_degenerate_and_duplicate_vertices():
    plane = lambda u, v: (np.stack([u, v, 0 * u], 1), np.tile([0.0, 0.0, 1.0], (len(u), 1)))
    pos, nrm, uv, idx = _grid(3, plane)
    base = _gen_tangents(pos, nrm, uv, idx)
    # un-index the mesh: every corner its own vertex; welding must give the same per-vertex result
    pos2, nrm2, uv2 = pos[idx], nrm[idx], uv[idx]
    t2 = _gen_tangents(pos2, nrm2, uv2, np.arange(len(idx), dtype=np.uint32))
    np.testing.assert_allclose(t2, base[idx], atol=1e-6)
    # a degenerate triangle (two equal positions) that shares vertex 5 with good triangles copies their frame;
    # one built from otherwise unused vertices keeps the default frame with w = +1
    extra_p = np.array([[9, 9, 0], [9, 9, 0], [8, 9, 0]], np.float32)
    pos3 = np.concatenate([pos, extra_p]); nrm3 = np.concatenate([nrm, np.tile([0, 0, 1], (3, 1)).astype(np.float32)])
    uv3 = np.concatenate([uv, np.array([[0, 0], [1, 0], [0, 1]], np.float32)])
    n0 = len(pos)
    idx3 = np.concatenate([idx, np.array([5, 5, 6, n0, n0 + 1, n0 + 2], np.uint32)])
    t3 = _gen_tangents(pos3, nrm3, uv3, idx3)
    np.testing.assert_allclose(t3[:n0], base, atol=1e-6)
    assert (t3[n0:] == [1, 0, 0, 1]).all()


VIEWER = Path(__file__).resolve().parents[1] / "host" / "_build" / "gltf_viewer"


def test_cpp_gltf_viewer_cli_usage_and_errors(tmp_path):
    """The C++ headless gltf_viewer (host/gltf_viewer.cpp) mirrors the reference CLI `-f <file>` (args.rs:4-10)."""
    import subprocess
    assert VIEWER.exists(), "run __graft_entry__.build() (host/Makefile links the viewer once the CUDA core is built)"
    r = subprocess.run([str(VIEWER)], capture_output=True, text=True)
    assert r.returncode == 2 and "usage: gltf_viewer -f <file>" in r.stderr
    r = subprocess.run([str(VIEWER), "-f", str(tmp_path / "missing.gltf")], capture_output=True, text=True)
    assert r.returncode == 1 and "load_file" in r.stderr               # import errors are reported, not swallowed


@pytest.mark.gpu
def test_cpp_gltf_viewer_matches_python_viewer(tmp_path):
    """Same file, same options: the compiled host (2 frames in flight, C ABI only) and the ctypes host give the same PNG."""
    import subprocess
    from PIL import Image
    from rustracer_b200 import gltf_viewer
    f = write_gltf(tmp_path)
    common = ["-f", str(f), "--width", "96", "--height", "64", "--spp", "12", "--samples-per-frame", "3", "--bounces", "4", "--camera", "0", "0", "14"]
    r = subprocess.run([str(VIEWER), *common, "-o", str(tmp_path / "cpp.png")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "12 spp in 4 frames" in r.stdout
    gltf_viewer.main([*common, "-o", str(tmp_path / "py.png")])
    a, b = np.asarray(Image.open(tmp_path / "cpp.png")), np.asarray(Image.open(tmp_path / "py.png"))
    assert a.shape == (64, 96, 3) and (a == b).all() and a.std() > 0
    # the same frames split over three replicas in ONE process (rt_multi, tile partition; the box has one GPU, so all replicas
    # sit on device 0): gathered image identical to the single-GPU one
    r = subprocess.run([str(VIEWER), *common, "--gpus", "3", "--same-device", "-o", str(tmp_path / "multi.png")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "on 3 GPU(s)" in r.stdout
    c = np.asarray(Image.open(tmp_path / "multi.png"))
    assert (c == a).all()


def test_malformed_gltf_is_an_error_not_a_crash(tmp_path):
    """The gltf crate validates indices at import and the reference panics on the rest; the C++ loader must turn every
    dangling index / wrong type into an error (found by fuzzing the tiny file: these used to read out of bounds)."""
    base = json.loads(write_gltf(tmp_path).read_text())

    def broken(mutate):
        g = json.loads(json.dumps(base)); mutate(g)
        p = tmp_path / "bad.gltf"; p.write_text(json.dumps(g))
        with pytest.raises(RuntimeError):
            host.load_file(p).scene_desc()

    broken(lambda g: g["scenes"][0].__setitem__("nodes", [7]))                  # scene -> missing node
    broken(lambda g: g["nodes"][0].__setitem__("mesh", 1))                      # node -> missing mesh
    broken(lambda g: g.__setitem__("meshes", []))
    broken(lambda g: g.__setitem__("nodes", 0))                                 # wrong JSON type
    broken(lambda g: g["nodes"][0].__setitem__("children", [0]))                # cycle
    broken(lambda g: g["nodes"][0].__setitem__("skin", 0))                      # node -> missing skin
    broken(lambda g: g["accessors"][0].__setitem__("bufferView", 9))            # accessor -> missing bufferView
    broken(lambda g: g["bufferViews"][0].__setitem__("buffer", 3))              # bufferView -> missing buffer
    broken(lambda g: g["accessors"][2].__setitem__("count", 10 ** 9))           # accessor beyond the buffer
    broken(lambda g: g.pop("bufferViews"))
    broken(lambda g: g["meshes"][0]["primitives"][0]["attributes"].__setitem__("POSITION", 5))
    broken(lambda g: g.__setitem__("animations", [{"channels": [{"sampler": 0, "target": {"node": 4, "path": "translation"}}],
                                                   "samplers": [{"input": 0, "output": 0}]}]))   # channel -> missing node
    g1 = json.loads(json.dumps(base)); g1["accessors"][0].update(type="SCALAR", componentType=5123)   # POSITION with one component:
    p1 = tmp_path / "narrow.gltf"; p1.write_text(json.dumps(g1))                                # missing components read as 0, never outside the buffer
    try:
        host.load_file(p1).scene_desc()
    except RuntimeError:
        pass
    broken(lambda g: g.__setitem__("animations", [{"channels": [{"sampler": 0, "target": {"node": 0, "path": "rotation"}}],
                                                   "samplers": [{"input": 0, "output": 2}]}]))   # sampler output shorter / narrower than its input
    idx_bad = np.array([0, 1, 2, 2, 1, 9], np.uint16).tobytes()

    def bad_index(g):
        raw = bytearray(base64.b64decode(g["buffers"][0]["uri"].split(",", 1)[1])); raw[64:76] = idx_bad
        g["buffers"][0]["uri"] = "data:application/octet-stream;base64," + base64.b64encode(bytes(raw)).decode()
    broken(bad_index)                                                                           # vertex index beyond the vertex count


def test_unorm8_sequence_equals_division():
    """rt_surface.h unorm8(): x * rn(1/255) plus one fused residual correction is byte / 255.0f for every byte (the product
    alone differs for 126 of them).  The fused multiply-adds are evaluated exactly in float64 and rounded once."""
    x = np.arange(256, dtype=np.float32)
    c = np.float32(0.0039215688593685626984)
    assert c == np.float32(1) / np.float32(255)
    q0 = (x * c).astype(np.float32)
    r = (x.astype(np.float64) - q0.astype(np.float64) * 255.0).astype(np.float32)
    q = (r.astype(np.float64) * np.float64(c) + q0.astype(np.float64)).astype(np.float32)
    assert np.array_equal(q, x / np.float32(255))
    assert np.count_nonzero(q0 != x / np.float32(255)) > 0
