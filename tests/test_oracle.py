"""Pins the CPU oracle: independent known answers for the RNGs, committed golden vectors, BVH-vs-brute-force."""
import ctypes as C

import numpy as np
import pytest

import util
from oracle import orc
from rustracer_b200 import _ffi as F, host

M32 = 0xFFFFFFFF


def py_tea(v0, v1):   # lib/Random.glsl:12-25 restated independently in Python
    s0 = 0
    for _ in range(16):
        s0 = (s0 + 0x9E3779B9) & M32
        v0 = (v0 + ((((v1 << 4) & M32) + 0xA341316C) ^ (v1 + s0) ^ ((v1 >> 5) + 0xC8013EA4))) & M32
        v1 = (v1 + ((((v0 << 4) & M32) + 0xAD90777D) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7E95761E))) & M32
    return v0


def py_pcg4d(v):   # lib/Random.glsl:81-98
    v = [(x * 1664525 + 1013904223) & M32 for x in v]
    def rnd(v):
        v[0] = (v[0] + v[1] * v[3]) & M32; v[1] = (v[1] + v[2] * v[0]) & M32
        v[2] = (v[2] + v[0] * v[1]) & M32; v[3] = (v[3] + v[1] * v[2]) & M32
    rnd(v)
    v = [x ^ (x >> 16) for x in v]
    rnd(v)
    return v


def test_tea_and_pcg_known_answers():
    L = orc.lib()
    for a, b in ((0, 0), (1, 2), (123456, 654321), (M32, 7), (1919, 1079)):
        assert L.orc_tea(a, b) == py_tea(a, b)
    for st in ([1, 2, 3, 4], [0, 0, 0, 0], [M32, 12345, 99, 7]):
        i, o = (F.c_u32 * 4)(*st), (F.c_u32 * 4)()
        L.orc_pcg4d(i, o)
        assert list(o) == py_pcg4d(list(st))
    # rand(): w++ then 23-bit mantissa trick
    st = (F.c_u32 * 4)(5, 6, 7, 0)
    r = L.orc_rand(st)
    x = py_pcg4d([5, 6, 7, 1])[0]
    assert st[3] == 1 and r == np.float32(np.uint32(0x3F800000 | (x >> 9)).view(np.float32) - np.float32(1.0))
    assert 0.0 <= r < 1.0
    seed = F.c_u32(42)
    f = L.orc_lcg_float(C.byref(seed))
    s = (1664525 * 42 + 1013904223) & M32
    assert seed.value == s and f == np.float32((s & 0xFFFFFF) / 16777216.0)


def test_rng_bsdf_tables_match_golden():
    g = np.load(util.GOLDEN / "rng_bsdf_golden.npz")
    L = orc.lib()
    for a, b, r in g["tea"]:
        assert L.orc_tea(int(a), int(b)) == int(r)
    for st4, seq in zip(g["pcg_states"], g["pcg_rand"]):
        st = (F.c_u32 * 4)(*[int(x) for x in st4])
        assert [L.orc_rand(st) for _ in range(8)] == list(seq)
    out = np.zeros(3, np.float32)
    for p, n, o in zip(g["offset_p"], g["offset_n"], g["offset_out"]):
        L.orc_offset_ray(F.as_ptr(np.ascontiguousarray(p), F.c_f), F.as_ptr(np.ascontiguousarray(n), F.c_f), F.as_ptr(out, F.c_f))
        assert (out == o).all()
    for m in range(5):
        for c, o in zip(g["tonemap_in"], g["tonemap_out"][m]):
            L.orc_tonemap(m, F.as_ptr(np.ascontiguousarray(c), F.c_f), F.as_ptr(out, F.c_f))
            np.testing.assert_allclose(out, o, rtol=1e-6, atol=1e-7)
    bo = np.zeros(14, np.float32)
    for i, o in zip(g["bsdf_in"], g["bsdf_out"]):
        L.orc_bsdf_sample(F.as_ptr(np.ascontiguousarray(i), F.c_f), F.as_ptr(bo, F.c_f))
        np.testing.assert_allclose(bo, o, rtol=2e-5, atol=1e-6)


def test_bsdf_energy_and_direction_sanity():
    """Size-independent properties: sampled directions are unit length and on the expected side; lobe probabilities sum to 1."""
    g = np.load(util.GOLDEN / "rng_bsdf_golden.npz")
    for i, o in zip(g["bsdf_in"], g["bsdf_out"]):
        assert abs(o[8] + o[9] + o[10] - 1.0) < 1e-5
        if o[0]:
            assert abs(np.linalg.norm(o[1:4]) - 1.0) < 1e-4
            n, lobe = i[0:3], int(i[18])
            if lobe in (1, 2):
                assert np.dot(o[1:4], n) > -1e-4
            assert (o[4:7] >= 0).all()


def test_closest_hits_match_golden(cornell_oracle, golden):
    h = cornell_oracle.trace_closest(golden["rays"], 1, golden["rng4"])
    assert util.hits_equal(h, golden["hits_opaque"]).all()
    h = cornell_oracle.trace_closest(golden["rays"], 0, golden["rng4"])
    assert util.hits_equal(h, golden["hits_alpha"]).all()
    # the BLEND sphere (alpha 0.05) lets most alpha-tested rays through: the two sets must differ
    assert (~util.hits_equal(golden["hits_alpha"], golden["hits_opaque"])).sum() > 10
    srays = golden["rays"].copy(); srays["tmin"] = 0.1; srays["tmax"] = 6.0
    assert (cornell_oracle.trace_any(srays, 0, golden["rng4"]) == golden["any_alpha"]).all()


def test_bvh_traversal_equals_brute_force(cornell_desc, cornell_oracle):
    rays, rng4 = util.random_rays(1500, seed=99)
    adv = util.adversarial_rays(cornell_desc, 1500, seed=3)
    for r in (rays, adv):
        for flags in (0, 1):
            a = cornell_oracle.trace_closest(r, flags, rng4[: len(r)])
            b = cornell_oracle.trace_closest(r, flags, rng4[: len(r)], brute=True)
            assert util.hits_equal(a, b).all()


def test_render_matches_golden(cornell_oracle, golden):
    W = H = 64
    acc = None
    for raw in golden["image_ubos"]:
        u = F.rt_ubo.from_buffer_copy(raw.tobytes())
        acc, out, _ = cornell_oracle.render(u, W, H, acc)
    np.testing.assert_allclose(acc, golden["image_acc"], rtol=1e-5, atol=1e-6)
    assert util.psnr(out, golden["image_out"]) > 60
    for name in ("albedo", "normal", "instance", "triangle"):
        u = F.rt_ubo.from_buffer_copy(golden["map_" + name + "_ubo"].tobytes())
        _, o, _ = cornell_oracle.render(u, W, H, None)
        assert np.abs(o.astype(int) - golden["map_" + name].astype(int)).max() <= 1
    u = F.rt_ubo.from_buffer_copy(golden["image_ubos"][0].tobytes())
    for k, (x, y) in enumerate(((32, 32), (10, 50), (50, 12), (20, 40))):
        rec = np.pad(cornell_oracle.trace_pixel(u, W, H, x, y, 8), ((0, 8), (0, 0)))[:8]
        np.testing.assert_allclose(rec, golden["payload_trace"][k], rtol=1e-5, atol=1e-6)


def test_empty_and_degenerate_inputs():
    """Edge cases: zero rays, rays that miss everything, zero-area triangles, a scene with an empty geometry."""
    from rustracer_b200 import scenes
    b = scenes.SceneBuilder()
    m = b.add_material(scenes.material(metallic=0.0))
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [2, 2, 2], [2, 2, 2], [2, 2, 2]], np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (6, 1)); uv = np.zeros((6, 2), np.float32)
    g0 = b.add_geometry(pos, nrm, uv, np.array([0, 1, 2, 3, 4, 5], np.uint32), m)   # second triangle is degenerate
    g1 = b.add_geometry(pos[:3], nrm[:3], uv[:3], np.zeros(0, np.uint32), m)        # no triangles at all
    b.add_instance(g0, np.eye(4)); b.add_instance(g1, np.eye(4))
    s = orc.OracleScene(b.build())
    assert len(s.trace_closest(np.zeros(0, F.RAY_DTYPE))) == 0
    rays = np.zeros(3, F.RAY_DTYPE)
    rays["origin"] = [[0.25, 0.25, 1], [5, 5, 5], [2, 2, 3]]; rays["direction"] = [[0, 0, -1], [0, 0, 1], [0, 0, -1]]
    rays["tmin"], rays["tmax"] = 0.001, 100
    h = s.trace_closest(rays, 1)
    assert h["t"][0] == 1.0 and h["primitive_id"][0] == 0 and h["instance_id"][0] == 0
    assert h["t"][1] == -1.0 and h["instance_id"][1] == M32 and h["t"][2] == -1.0
    # open interval: a hit exactly at tmax or tmin is rejected
    rays["tmax"][0] = 1.0
    assert s.trace_closest(rays[:1], 1)["t"][0] == -1.0
    rays["tmax"][0] = 100; rays["tmin"][0] = 1.0
    assert s.trace_closest(rays[:1], 1)["t"][0] == -1.0


def test_tonemap_operators_match_published_curves():
    """lib/Tonemapping.glsl:9-53 against independent float64 restatements of the published operators (Hable's
    Uncharted 2 filmic curve, Hejl / Burgess-Dawson, Narkowicz' ACES fit, Reinhard) plus known answers: each curve's
    value at the white point / at zero."""
    l = orc.lib()
    rng = np.random.default_rng(9)
    x = np.concatenate([rng.uniform(0, 4, (200, 3)), rng.uniform(0, 0.05, (50, 3)), [[0, 0, 0], [11.2 / 2] * 3, [1, 1, 1]]]).astype(np.float32)

    def run(mode):
        out = np.zeros_like(x)
        for i in range(len(x)):
            a = np.ascontiguousarray(x[i]); b = np.zeros(3, np.float32)
            l.orc_tonemap(mode, F.as_ptr(a, F.c_f), F.as_ptr(b, F.c_f)); out[i] = b
        return out.astype(np.float64)

    X = x.astype(np.float64)
    g = lambda c: np.power(c, 1 / 2.2)

    def hable(c):
        A, B, Cc, D, E, Fv = 0.15, 0.50, 0.10, 0.20, 0.02, 0.30
        return (c * (A * c + Cc * B) + D * E) / (c * (A * c + B) + D * Fv) - E / Fv

    np.testing.assert_allclose(run(0), g(X / (X + 1)), rtol=2e-5, atol=2e-6)                               # Reinhard + gamma
    nz = X.min(1) > 0            # at exactly 0 the curve is a float32 rounding residue (~1e-9) that the 1/2.2 power lifts to ~2e-4
    np.testing.assert_allclose(run(1)[nz], g(hable(2 * X[nz]) / hable(11.2)), rtol=5e-5, atol=5e-6)       # Uncharted 2, exposure bias 2, W = 11.2
    h = np.maximum(0, X - 0.004)
    np.testing.assert_allclose(run(2), (h * (6.2 * h + 0.5)) / (h * (6.2 * h + 1.7) + 0.06), rtol=2e-5, atol=2e-6)   # Hejl-Richard (gamma baked in)
    np.testing.assert_allclose(run(3), g(np.clip((X * (2.51 * X + 0.03)) / (X * (2.43 * X + 0.59) + 0.14), 0, 1)), rtol=2e-5, atol=2e-6)   # ACES (Narkowicz)
    np.testing.assert_allclose(run(7), g(X), rtol=2e-5, atol=2e-6)                                          # any other mode: gamma only
    # known answers: the Uncharted curve maps its white point (input W/2 after the x2 exposure bias) to exactly 1, zero to zero
    np.testing.assert_allclose(run(1)[-2], [1, 1, 1], atol=2e-6)
    assert np.abs(run(1)[-3]).max() < 1e-3
    np.testing.assert_allclose(run(0)[-1], [0.5 ** (1 / 2.2)] * 3, rtol=1e-6)
