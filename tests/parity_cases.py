"""Parity cases shared by the GPU tests (product library through the C ABI, `-m gpu`) and by the pre-GPU
host-emulation checks (`-m "not gpu"`).  Every case compares against the CPU oracle on the same seeded inputs.
Tolerances (BASELINE.json north_star): IDs / t / barycentrics bit-exact; images mean relative error < 1 %,
PSNR >= 40 dB; debug-mapping channels within 1 LSB."""
import numpy as np

import util
from oracle import orc
from rustracer_b200 import _ffi as F, core, host, scenes

MRE_TOL = 0.01
PSNR_TOL = 40.0


def make(api, desc, w=64, h=64):
    ctx = core.Context(w, h, api=api)
    return ctx, core.Scene(ctx, desc)


SCALAR = 2   # RT_TRACE_SCALAR: one thread per ray; default = the persistent wavefront traversal rt_render runs


def check_ids(sc, o, rays, flags=0, rng4=None, what=""):
    """Closest-hit ids / t / u / v bit-exact vs the oracle through BOTH GPU paths: the persistent wavefront traversal of the
    frame kernels (default of rt_trace_closest) and the scalar per-thread loop."""
    ho = o.trace_closest(rays, flags, rng4)
    for mode in (0, SCALAR):
        hg = sc.trace_closest(rays, flags | mode, rng4)
        bad = ~util.hits_equal(hg, ho)
        assert not bad.any(), (what, "scalar" if mode else "wavefront", int(bad.sum()), hg[bad][:3], ho[bad][:3])
    return ho


def check_any(sc, o, rays, flags=0, rng4=None, what=""):
    oo = o.trace_any(rays, flags, rng4)
    for mode in (0, SCALAR):
        og = sc.trace_any(rays, flags | mode, rng4)
        assert (og == oo).all(), (what, "scalar" if mode else "wavefront", int((og != oo).sum()))
    return oo


def bounce_ray_sets(o, ubo, W, H, stride=1, max_rays=1 << 19, max_shadow=1 << 17):
    """SURVEY.md §8d ray set (iv): the oracle's own bounce >= 1 segments (origin on a surface, tMin = 0.001) and shadow rays."""
    return o.record_bounce_rays(ubo, W, H, stride=stride, min_bounce=1, max_rays=max_rays, max_shadow=max_shadow)


def case_trace_golden(api, cornell_desc, cornell_oracle, golden, n_adv=3000):
    ctx, sc = make(api, cornell_desc)
    for flags, key in ((1, "hits_opaque"), (0, "hits_alpha")):
        for mode in (0, SCALAR):
            h = sc.trace_closest(golden["rays"], flags | mode, golden["rng4"])
            assert util.hits_equal(h, golden[key]).all(), (key, mode)
    adv = util.adversarial_rays(cornell_desc, n_adv, seed=11)
    for flags in (0, 1):
        check_ids(sc, cornell_oracle, adv, flags, what="adversarial")
    srays = golden["rays"].copy(); srays["tmin"] = 0.1; srays["tmax"] = 6.0
    for mode in (0, SCALAR):
        assert (sc.trace_any(srays, mode, golden["rng4"]) == golden["any_alpha"]).all()
    assert len(sc.trace_closest(np.zeros(0, F.RAY_DTYPE))) == 0   # empty input
    # ray set (iv): bounce >= 1 segments and shadow rays of the oracle's own path tracer (BLEND sphere: alpha test with the payload RNG)
    W = H = 96
    cam = host.Camera(W, H); gui = host.Gui(number_of_samples=1, number_of_bounces=8)
    ubo = host.FrameDriver(cam, gui, cornell_desc.fully_opaque).next_ubo()
    rays, rng4, srays, srng4 = bounce_ray_sets(cornell_oracle, ubo, W, H)
    assert len(rays) > 5000
    check_ids(sc, cornell_oracle, rays, 0, rng4, what="bounce rays")
    check_ids(sc, cornell_oracle, rays, 1, None, what="bounce rays opaque")


def case_render_golden(api, cornell_desc, golden):
    ctx, sc = make(api, cornell_desc)
    for raw in golden["image_ubos"]:
        ctx.render(sc, F.rt_ubo.from_buffer_copy(raw.tobytes()))
    acc, out = ctx.readback()
    assert util.mean_rel_err(acc, golden["image_acc"], 8) < MRE_TOL
    assert util.psnr(out[..., :3], golden["image_out"][..., :3]) >= PSNR_TOL
    for name in ("albedo", "normal", "instance", "triangle"):
        ctx.resize(64, 64)
        ctx.render(sc, F.rt_ubo.from_buffer_copy(golden["map_" + name + "_ubo"].tobytes()))
        _, o = ctx.readback()
        assert np.abs(o.astype(int) - golden["map_" + name].astype(int)).max() <= 1, name


def case_tile_partition(api, cornell_desc, golden):
    ctx, sc = make(api, cornell_desc)
    u = F.rt_ubo.from_buffer_copy(golden["image_ubos"][0].tobytes())
    ctx.render(sc, u)
    acc_full, out_full = ctx.readback()
    ctx.resize(64, 64)
    for part in range(3):
        ctx.render(sc, u, strip_rows=8, n_parts=3, part=part)
    acc_p, out_p = ctx.readback()
    assert (acc_full == acc_p).all() and (out_full == out_p).all()


def case_frames_in_flight(api, cornell_desc, golden, size=64, frames=7):
    """InFlightFrames (app/src/lib.rs:34): n frames overlap on the GPU, accumulation stays in submission order, so the
    accumulation image and every presented RGBA8 frame are bit-identical to the strictly ordered n = 1 run."""
    ctx, sc = make(api, cornell_desc, size, size)
    cam = host.Camera(size, size).set(position=(0, 0, 14.0)); gui = host.Gui(number_of_samples=1, number_of_bounces=6)
    drv = host.FrameDriver(cam, gui, cornell_desc.fully_opaque)
    ubos = [drv.next_ubo() for _ in range(frames)]
    ref_out = []
    for u in ubos:
        ctx.render(sc, u)
        ref_out.append(ctx.readback(want_acc=False)[1].copy())
    ref_acc, _ = ctx.readback()
    ref_stats = ctx.stats()
    for n in (2, 3, 4):
        ctx.set_frames_in_flight(n)
        ctx.resize(size, size)
        bufs = [np.zeros((size, size, 4), np.uint8) for _ in ubos]
        tickets = []
        for u, b in zip(ubos, bufs):
            ctx.render(sc, u)
            tickets.append(ctx.readback_async(b))
        assert tickets == list(range(tickets[0], tickets[0] + len(ubos)))
        for t in tickets:
            ctx.frame_wait(t)
        for f, (b, r) in enumerate(zip(bufs, ref_out)):
            assert (b == r).all(), (n, f)
        acc, out = ctx.readback()
        assert (acc == ref_acc).all() and (out == ref_out[-1]).all(), n
        st = ctx.stats()
        assert st.rays_extend == ref_stats.rays_extend and st.rays_shadow == ref_stats.rays_shadow
        # a consumer of the accumulation image joins the frames in flight: tonemap-only pass == last frame's image
        ctx.tonemap(ubos[-1]); _, out2 = ctx.readback()
        assert (out2 == ref_out[-1]).all()
    ctx.set_frames_in_flight(1)
    acc, out = ctx.readback()
    assert (acc == ref_acc).all() and (out == ref_out[-1]).all()       # shrinking keeps the presented image
    import pytest
    with pytest.raises(core.RtError):
        ctx.set_frames_in_flight(5)
    with pytest.raises(core.RtError):
        ctx.frame_wait(10 ** 9)


def case_lifecycle_in_flight(api, cornell_desc, size=48):
    """Resize / destroy / scene updates while frames are in flight must synchronise internally: no stale frame may land
    in a re-allocated image, nothing may crash, and the images equal those of a fresh, strictly ordered context."""
    cam = host.Camera(size, size).set(position=(0, 0, 14.0)); gui = host.Gui(number_of_samples=1, number_of_bounces=5)
    ubos = [host.FrameDriver(cam, gui, cornell_desc.fully_opaque).next_ubo() for _ in range(1)]
    drv = host.FrameDriver(cam, gui, cornell_desc.fully_opaque)
    ubos = [drv.next_ubo() for _ in range(6)]
    ref_ctx, ref_sc = make(api, cornell_desc, size, size)
    for u in ubos[:3]:
        ref_ctx.render(ref_sc, u)
    ref_acc, ref_out = ref_ctx.readback()
    ctx, sc = make(api, cornell_desc, size * 2, size)
    ctx.set_frames_in_flight(4)
    big = host.FrameDriver(host.Camera(size * 2, size).set(position=(0, 0, 14.0)), gui, cornell_desc.fully_opaque)
    for _ in range(5):
        ctx.render(sc, big.next_ubo())
    ctx.resize(size, size)                        # frames of the old size are still in flight here
    for u in ubos[:3]:
        ctx.render(sc, u)
    acc, out = ctx.readback()
    assert (acc == ref_acc).all() and (out == ref_out).all()
    dl, pl = bright_lights()
    for u in ubos[3:]:
        ctx.render(sc, u)
    sc.update_lights(dl, pl)                      # waits for the frames that still read the old light arrays
    ctx.render(sc, ubos[0])
    assert np.isfinite(ctx.readback()[0]).all()
    for u in ubos:
        ctx.render(sc, u)
    sc.close(); ctx.close()                       # destroy with frames in flight
    ref_sc.close(); ref_ctx.close()


def case_instancing(api):
    b = scenes.SceneBuilder()
    m = b.add_material(scenes.material((0.8, 0.3, 0.2, 1), metallic=0.0))
    g = b.add_geometry(*scenes.uv_sphere(1.0, 12, 16), m)
    rng = np.random.default_rng(4)
    for _ in range(40):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        b.add_instance(g, scenes.trs(rng.uniform(-8, 8, 3), q, rng.uniform(0.3, 1.5, 3)))
    # baked instances next to the instanced ones: a single-use floor and a tiny (<= 256 triangles) mesh used twice
    gf = b.add_geometry(*scenes.box_mesh((9, 0.1, 9)), m)
    b.add_instance(gf, scenes.trs((0, -9, 0)))
    gs = b.add_geometry(*scenes.uv_sphere(1.0, 6, 8), m)
    b.add_instance(gs, scenes.trs((0, 5, 0), scale=(3, 3, 3))); b.add_instance(gs, scenes.trs((4, 5, 1), scale=(2, 1, 2)))
    d = b.build()
    o = orc.OracleScene(d)
    ctx, sc = make(api, d, 32, 32)
    rays, _ = util.random_rays(4000, seed=8, extent=9.0)
    check_ids(sc, o, rays, 1)
    assert sc.bvh_info().tlas_nodes >= 2
    inst = np.frombuffer(util._arr(d.instances, d.n_instances, F.rt_instance), F.INSTANCE_DTYPE).copy()
    inst["transform"][:, 3] += 1.5
    sc.update_instances(inst); o.update_instances(inst)
    check_ids(sc, o, rays, 1)


def bright_lights():
    pl = np.zeros(2, F.LIGHT_DTYPE); pl["color"] = [[1, .9, .8, 0], [.2, .4, 1, 0]]; pl["transform"] = [[0, 3, 2, 1], [-3, -2, 3, 1]]
    pl["kind"], pl["range"], pl["intensity"] = 1, 1e30, [20, 10]
    dl = np.zeros(1, F.LIGHT_DTYPE); dl["color"], dl["transform"], dl["intensity"] = 1, [[-.5, -1, -.3, 0]], 2.0
    return dl, pl


def case_lights(api, cornell_desc, size=48, frames=2):
    o = orc.OracleScene(cornell_desc)
    ctx, sc = make(api, cornell_desc, size, size)
    dl, pl = bright_lights()
    o.update_lights(dl, pl); sc.update_lights(dl, pl)
    cam = host.Camera(size, size).set(position=(0, 0, 14)); gui = host.Gui(number_of_samples=2, number_of_bounces=6)
    d1, d2 = host.FrameDriver(cam, gui, False), host.FrameDriver(cam, gui, False)
    acc = None
    for _ in range(frames):
        acc, out, st = o.render(d1.next_ubo(), size, size, acc); ctx.render(sc, d2.next_ubo())
    acc_e, out_e = ctx.readback()
    s = ctx.stats()
    assert st.rays_shadow > 0 and abs(int(s.rays_shadow) - int(st.rays_shadow)) <= 0.001 * st.rays_shadow
    assert abs(int(s.shaded_hits) - int(st.shaded_hits)) <= 0.002 * st.shaded_hits and s.shaded_hits < s.rays_extend   # real hits, not shade invocations
    assert util.mean_rel_err(acc_e, acc, 2 * frames) < MRE_TOL and util.psnr(out_e[..., :3], out[..., :3]) >= PSNR_TOL
    # counted frame: texture taps / light candidates feed bench.py's algorithmic bytes and must follow the oracle's
    u = d2.next_ubo(); ctx.render(sc, u, flags=1); sg = ctx.stats(); _, _, so = o.render(d1.next_ubo(), size, size, acc)
    assert so.light_cands > 0 and abs(int(sg.light_cands) - int(so.light_cands)) <= 0.002 * so.light_cands and sg.tex_taps == so.tex_taps == 0
    # ray set (iv): the oracle's shadow rays (tMin 0.1, any-hit with the BLEND sphere) and bounce rays with NEE on
    rays, rng4, srays, srng4 = bounce_ray_sets(o, u, size, size)
    assert len(srays) > 100
    check_any(sc, o, srays, 0, srng4, what="shadow rays")
    check_ids(sc, o, rays, 0, rng4, what="bounce rays (lights)")


def skinned_scene(seed=3, rows=24, cols=16, joints=8):
    """Small capsule skinned to a chain of joints (config-4 shape in miniature)."""
    rng = np.random.default_rng(seed)
    b = scenes.SceneBuilder()
    m = b.add_material(scenes.material((0.7, 0.7, 0.9, 1), metallic=0.0))
    pos, nrm, uv, idx = scenes.uv_sphere(1.0, rows, cols)
    pos = pos * np.array([0.5, 2.0, 0.5], np.float32)
    t = (pos[:, 1] + 2.0) / 4.0 * (joints - 1)
    j0 = np.clip(np.floor(t).astype(np.uint32), 0, joints - 2)
    w1 = (t - j0).astype(np.float32)
    weights = np.zeros((len(pos), 4), np.float32); jidx = np.zeros((len(pos), 4), np.uint32)
    weights[:, 0], weights[:, 1] = 1 - w1, w1
    jidx[:, 0], jidx[:, 1] = j0, j0 + 1
    g = b.add_geometry(pos, nrm, uv, idx, m, weights=weights, joints=jidx, skin_index=0)
    b.add_instance(g, np.eye(4))
    floor_m = b.add_material(scenes.material(metallic=0.0))
    gf = b.add_geometry(*scenes.box_mesh((4, 0.1, 4)), floor_m)
    b.add_instance(gf, scenes.trs((0, -2.6, 0)))

    def pose(phase):
        mats = np.zeros((1, 256, 16), np.float32)
        mats[0, :, [0, 5, 10, 15]] = 1.0
        for j in range(joints):
            a = 0.35 * np.sin(phase + 0.7 * j)
            M = scenes.trs((0.3 * np.sin(phase * 1.3 + j), 0.05 * j * np.cos(phase), 0), (0, 0, np.sin(a / 2), np.cos(a / 2)))
            mats[0, j] = M.T.reshape(16).astype(np.float32)   # column-major
        return mats
    b.skins = pose(0.0)
    return b.build(), pose


def case_skinning(api):
    d, pose = skinned_scene()
    o = orc.OracleScene(d)
    ctx, sc = make(api, d, 32, 32)
    rays, _ = util.random_rays(3000, seed=21, extent=3.0)
    n = d.n_vertices
    # refit is idempotent: refitting with the pose the BVH was built from reproduces the built nodes byte for byte
    # (child boxes are exact min / max unions either way), and so does a rebuild (deterministic builder)
    built = sc.read_nodes(-1)
    assert built.shape[0] >= 2 and built.shape[1] == 32
    sc.update_skins(pose(0.0), rebuild=False)
    assert (sc.read_nodes(-1) == built).all()
    # a rebuild allocates child / primitive ranges with atomics, so nodes may be numbered differently from run to run:
    # compare the nodes as a multiset, ignoring the two base-index words
    def canon(nodes):
        k = np.delete(nodes, [4, 5], axis=1)
        return k[np.lexsort(k.T[::-1])]
    sc.update_skins(pose(0.0), rebuild=True)
    rebuilt = sc.read_nodes(-1)
    assert rebuilt.shape == built.shape and (canon(rebuilt) == canon(built)).all()
    for k, phase in enumerate((0.0, 0.9, 2.3)):
        if k:
            mats = pose(phase)
            sc.update_skins(mats, rebuild=(k == 2)); o.update_skins(mats)
        vg, vo = sc.read_vertices(n), o.read_vertices(n)
        # AnimationCompute.comp: explicitly rounded operations in the oracle's order on both sides -> positions, normals and
        # tangents bit-exact (SURVEY.md §8d "Skinning kernel: bit-exact vs oracle under the same fma policy")
        for f in ("position", "normal", "tangent"):
            assert (vg[f].view(np.uint32) == vo[f].view(np.uint32)).all(), (f, np.abs(vg[f] - vo[f]).max())
        assert (vg["joints"] == vo["joints"]).all() and (vg["skin_index"] == vo["skin_index"]).all()
        # identical vertices -> identical triangles: closest hits bit-exact through refit and rebuild alike
        check_ids(sc, o, rays, 1, what=f"skinned pose {k}")


def case_shadows_glb(api, d, z, frames=(0, 7, 23, 41, 59), size=128, n_rays=100000, spp=2):
    """The reference's shadows.glb through its loader-produced animation (fixture tests/golden/shadows_glb_scene.npz):
    per frame rt_scene_update_skins (skinning kernel + BLAS refit + TLAS refit, main.rs:384-395) then vertices, closest hits,
    shadow-ray occlusion and a rendered image with NEE (point light on, textured base colour) against the oracle."""
    o = orc.OracleScene(d)
    ctx, sc = make(api, d, size, size)
    rays, rng4 = util.random_rays(n_rays, seed=17, extent=5.0)
    cam = host.Camera(size, size).set(position=(0.0, 0.0, 9.0)); gui = host.Gui(number_of_samples=spp, number_of_bounces=5, animation=1)
    d1, d2 = host.FrameDriver(cam, gui, d.fully_opaque), host.FrameDriver(cam, gui, d.fully_opaque)
    moved = 0.0
    v0 = sc.read_vertices()
    for f in frames:
        mats = util.expand_skins(z["anim_skins"][f])
        sc.update_skins(mats); o.update_skins(mats)
        vg, vo = sc.read_vertices(), o.read_vertices(d.n_vertices)
        for name in ("position", "normal", "tangent"):
            assert (vg[name].view(np.uint32) == vo[name].view(np.uint32)).all(), (f, name)
        moved = max(moved, float(np.abs(vg["position"] - v0["position"]).max()))
        check_ids(sc, o, rays, 1, what=f"shadows.glb frame {f}")
        u = d1.next_ubo(); ctx.render(sc, u); acc, out, st = o.render(d2.next_ubo(), size, size, None)
        acc_g, out_g = ctx.readback()
        mre, ps = util.mean_rel_err(acc_g, acc, spp), util.psnr(out_g[..., :3], out[..., :3])
        assert mre < MRE_TOL and ps >= PSNR_TOL, (f, mre, ps)
        sg = ctx.stats()
        assert st.rays_shadow > 0 and abs(int(sg.rays_shadow) - int(st.rays_shadow)) <= 0.003 * st.rays_shadow + 2     # NEE is live
        brays, brng, srays, srng = bounce_ray_sets(o, u, size, size, max_rays=20000, max_shadow=20000)
        check_ids(sc, o, brays, 0, brng, what=f"shadows.glb bounce rays frame {f}")
        check_any(sc, o, srays, 0, srng, what=f"shadows.glb shadow rays frame {f}")
    assert moved > 0.05          # the animation really deforms the character


def case_lucy_ids(api, n_rays=100000, rows=120, cols=121):
    d = scenes.cornell_box(lucy=True, lucy_rows=rows, lucy_cols=cols, shell=util.GOLDEN / "cornell_box_scene.npz")   # §8d config 2: real shell + stand-in
    o = orc.OracleScene(d)
    ctx, sc = make(api, d, 32, 32)
    rays, _ = util.random_rays(n_rays, seed=5)
    check_ids(sc, o, rays, 1, what="lucy random")
    check_ids(sc, o, util.adversarial_rays(d, min(20000, n_rays), seed=13), 1, what="lucy adversarial")
    # ray set (iv): the oracle's own bounce >= 1 rays at the reference camera (inside the box, incoherent, origin on surfaces)
    W, H = 480, 270
    ubo = host.FrameDriver(host.Camera(W, H), host.Gui(number_of_samples=1, number_of_bounces=8), True).next_ubo()
    brays, brng, _, _ = bounce_ray_sets(o, ubo, W, H, max_rays=min(n_rays, 1 << 19))
    assert len(brays) > min(n_rays, 1 << 19) // 4
    check_ids(sc, o, brays, 1, what="lucy bounce rays")
    return d, o, ctx, sc


def render_compare(api, d, o, W, H, gui_kw, frames, cam_pos=(0, 0, 14.0), opaque=None, tol_mre=MRE_TOL, tol_psnr=PSNR_TOL):
    opaque = d.fully_opaque if opaque is None else opaque
    ctx, sc = make(api, d, W, H)
    cam = host.Camera(W, H).set(position=cam_pos); gui = host.Gui(**gui_kw)
    d1, d2 = host.FrameDriver(cam, gui, opaque), host.FrameDriver(cam, gui, opaque)
    acc = None
    for _ in range(frames):
        ctx.render(sc, d1.next_ubo()); acc, out, st = o.render(d2.next_ubo(), W, H, acc)
    acc_g, out_g = ctx.readback()
    total = d1.total.value
    mre, ps = util.mean_rel_err(acc_g, acc, total), util.psnr(out_g[..., :3], out[..., :3])
    assert mre < tol_mre and ps >= tol_psnr, (mre, ps)
    return ctx, sc, st


def case_foliage(api, n_side=6, tris=2000, size=64, n_rays=20000):
    """Config-3 shape in miniature: instanced BLAS + alpha-MASK textured cards + sky + directional light."""
    d = scenes.instanced_foliage(n_side=n_side, tris_per_mesh=tris, cards=16, tex_size=64, sky=scenes.procedural_sky(16))
    o = orc.OracleScene(d)
    ctx, sc = make(api, d, 16, 16)
    rays, rng4 = util.random_rays(n_rays, seed=31, extent=4.0)
    for flags in (0, 1):
        check_ids(sc, o, rays, flags, rng4)
    srays = rays.copy(); srays["tmin"] = 0.1; srays["tmax"] = 3.0
    check_any(sc, o, srays, 0, rng4)
    assert not d.fully_opaque
    render_compare(api, d, o, size, size, dict(number_of_samples=2, number_of_bounces=6, sky=1), 3, cam_pos=(0, 1.0, 6.0))


def case_glass(api, n_objects=8, size=64, res=(16, 17)):
    """Config-5 shape in miniature: transmission + volume attenuation + refraction."""
    d = scenes.glass_box(n_objects=n_objects, sphere_res=res)
    o = orc.OracleScene(d)
    render_compare(api, d, o, size, size, dict(number_of_samples=2, number_of_bounces=8), 4)


def case_skinned_character(api, n_tris=20000, joints=64, size=48, frames=3):
    """Config-4 shape in miniature: per frame skinning kernel + BLAS refit + TLAS rebuild + render, accumulation off."""
    d, pose = scenes.skinned_character(n_tris=n_tris, joints=joints)
    o = orc.OracleScene(d)
    ctx, sc = make(api, d, size, size)
    cam = host.Camera(size, size).set(position=(0, 0, 9.0)); gui = host.Gui(number_of_samples=2, number_of_bounces=5, animation=1)
    d1, d2 = host.FrameDriver(cam, gui, True), host.FrameDriver(cam, gui, True)
    for f in range(1, frames + 1):
        mats = pose(f * 7)
        sc.update_skins(mats); o.update_skins(mats)
        ctx.render(sc, d1.next_ubo()); acc, out, st = o.render(d2.next_ubo(), size, size, None)
        acc_g, out_g = ctx.readback()
        # vertices differ by rounding (FMA contraction in the skinning kernel) -> silhouettes may move by a pixel
        assert util.mean_rel_err(acc_g, acc, 2) < 0.02 and util.psnr(out_g[..., :3], out[..., :3]) >= 35.0


def case_skinned_in_flight(api, n_tris=20000, joints=64, size=48, frames=7):
    """Animated scene with frames in flight: the skin update of frame f+1 is queued while earlier frames still render
    (multi-buffered skinned vertices / triangles / BVH nodes).  Every presented frame must be bit-identical to the
    strictly serial run (1 scene copy, 1 frame in flight)."""
    d, pose = scenes.skinned_character(n_tris=n_tris, joints=joints)
    ctx, sc = make(api, d, size, size)
    cam = host.Camera(size, size).set(position=(0, 0, 9.0)); gui = host.Gui(number_of_samples=1, number_of_bounces=5, animation=1)
    drv = host.FrameDriver(cam, gui, True)
    ubos = [drv.next_ubo() for _ in range(frames)]
    poses = [pose(3 + 5 * f) for f in range(frames)]
    sc.set_versions(1)
    ref = []
    for u, m in zip(ubos, poses):
        sc.update_skins(m); ctx.render(sc, u)
        ref.append(ctx.readback(want_acc=False)[1].copy())
    assert any((ref[0] != r).any() for r in ref[1:])                 # the animation does move pixels
    for n_ver, nf in ((2, 2), (3, 4), (2, 4), (4, 3)):
        sc.set_versions(n_ver); ctx.set_frames_in_flight(nf); ctx.resize(size, size)
        bufs = [np.zeros((size, size, 4), np.uint8) for _ in range(frames)]
        tickets = []
        for u, m, b in zip(ubos, poses, bufs):
            sc.update_skins(m); ctx.render(sc, u)
            tickets.append(ctx.readback_async(b))
        for t in tickets:
            ctx.frame_wait(t)
        for f, (b, r) in enumerate(zip(bufs, ref)):
            assert (b == r).all(), (n_ver, nf, f)
    # a rebuild is synchronous and replicates the new topology into every copy
    sc.update_skins(poses[0], rebuild=True); ctx.render(sc, ubos[0])
    first = ctx.readback(want_acc=False)[1].copy()
    sc.update_skins(poses[1]); ctx.render(sc, ubos[1]); sc.update_skins(poses[0]); ctx.render(sc, ubos[0])
    assert (ctx.readback(want_acc=False)[1] == first).all()
    with __import__("pytest").raises(core.RtError):
        sc.set_versions(5)


def textured_scene(seed=17):
    """Nine spheres / quads whose materials walk through every texture slot of MaterialRaw (Material.glsl:51-73):
    base colour (sRGB), normal map + tangents, metallic-roughness, emissive, transmission, KHR_materials_specular
    factors + textures, the specular-glossiness workflow with both of its textures, vertex colours, the second UV set,
    nearest / linear filtering and clamp / mirror / repeat wrapping."""
    rng = np.random.default_rng(seed)
    b = scenes.SceneBuilder()

    def noise(n, srgb, lo=0, hi=256):
        return b.add_image(rng.integers(lo, hi, (n, n, 4), dtype=np.uint8), srgb=srgb)

    def normal_map(n):
        xy = rng.uniform(-0.5, 0.5, (n, n, 2)); z = np.sqrt(1 - (xy ** 2).sum(-1))
        v = np.concatenate([xy, z[..., None], np.ones((n, n, 1))], -1)
        return b.add_image(np.round((v * 0.5 + 0.5) * 255).astype(np.uint8), srgb=False)

    TI = F.rt_texture_info
    base = b.add_texture(noise(16, True), mag=1, wrap_s=2, wrap_t=2)
    base_near = b.add_texture(noise(8, True), mag=0, wrap_s=0, wrap_t=1)
    nrm = b.add_texture(normal_map(16), mag=1, wrap_s=1, wrap_t=2)
    mr = b.add_texture(noise(16, False), mag=1, wrap_s=2, wrap_t=0)
    emis = b.add_texture(noise(8, True, 0, 40), mag=1, wrap_s=2, wrap_t=2)
    trans = b.add_texture(noise(8, False), mag=0, wrap_s=2, wrap_t=2)
    spec = b.add_texture(noise(8, False), mag=1, wrap_s=2, wrap_t=2)
    spec_col = b.add_texture(noise(8, True), mag=1, wrap_s=1, wrap_t=1)
    sg_diff = b.add_texture(noise(16, True), mag=1, wrap_s=2, wrap_t=2)
    sg_spec = b.add_texture(noise(16, True), mag=1, wrap_s=0, wrap_t=0)
    mats = []
    m = scenes.material((0.9, 0.8, 0.7, 1), metallic=0.1, roughness=0.6, base_color_texture=base); mats.append(m)
    m = scenes.material((1, 1, 1, 1), metallic=0.0, roughness=0.8, base_color_texture=base_near); m.base_color_texture = TI(base_near, 1); mats.append(m)   # uv set 1
    m = scenes.material((0.8, 0.8, 0.8, 1), metallic=0.0, roughness=0.5); m.normal_texture = TI(nrm, 0); mats.append(m)
    m = scenes.material((0.9, 0.6, 0.3, 1), metallic=1.0, roughness=1.0); m.metallic_roughness_texture = TI(mr, 0); mats.append(m)
    m = scenes.material((0.5, 0.5, 0.5, 1), metallic=0.0, roughness=0.9, emissive=(1.0, 0.8, 0.6)); m.emissive_texture = TI(emis, 0); mats.append(m)
    m = scenes.material((0.9, 0.95, 1.0, 1), metallic=0.0, roughness=0.1, ior=1.45, transmission=0.9, volume=((0.8, 0.9, 0.7), 1.5)); m.transmission_texture = TI(trans, 0); mats.append(m)
    m = scenes.material((0.6, 0.2, 0.2, 1), metallic=0.0, roughness=0.4); m.specular_exist, m.specular_factor = 1, 0.7
    m.specular_color_factor[:] = (0.9, 0.7, 0.5, 1); m.specular_texture, m.specular_color_texture = TI(spec, 0), TI(spec_col, 0); mats.append(m)
    m = scenes.material((1, 1, 1, 1), metallic=0.3, roughness=0.3); m.workflow = 1
    m.sg_diffuse_factor[:] = (0.8, 0.7, 0.6, 1); m.sg_specular_glossiness_factor[:] = (0.6, 0.5, 0.4, 0.7)
    m.sg_diffuse_texture, m.sg_specular_glossiness_texture = TI(sg_diff, 0), TI(sg_spec, 0); mats.append(m)
    m = scenes.material((1, 1, 1, 1), metallic=0.0, roughness=1.0, unlit=True, base_color_texture=base); mats.append(m)
    ids = [b.add_material(m) for m in mats]
    for k, mid in enumerate(ids):
        pos, nrmv, uv, idx = scenes.uv_sphere(0.8, 10, 14)
        col = np.ones((len(pos), 4), np.float32); col[:, :3] = rng.uniform(0.6, 1.0, (len(pos), 3))
        g = b.add_geometry(pos, nrmv, uv * 3.0 - 0.7, idx, mid, color=col)           # uvs outside [0,1]: the wrap modes matter
        v = b.verts[g]
        t = np.cross(np.array([0, 1.0, 0]), nrmv); t /= np.maximum(np.linalg.norm(t, axis=1, keepdims=True), 1e-6)
        v["tangent"][:, :3] = t; v["tangent"][:, 3] = np.where(rng.random(len(pos)) < 0.5, -1.0, 1.0)
        v["uv1"] = uv[:, ::-1] * 2.0
        b.add_instance(g, scenes.trs(((k % 3 - 1) * 2.0, (k // 3 - 1) * 2.0, 0.0)))
    wall = b.add_material(scenes.material((0.8, 0.8, 0.8, 1), metallic=0.0, roughness=1.0))
    gw = b.add_geometry(*scenes.box_mesh((4.0, 4.0, 0.1)), wall); b.add_instance(gw, scenes.trs((0, 0, -1.5)))
    lamp = b.add_material(scenes.material((1, 1, 1, 1), metallic=0.0, emissive=(4, 4, 4)))
    gl = b.add_geometry(*scenes.box_mesh((1.5, 0.05, 1.5)), lamp); b.add_instance(gl, scenes.trs((0, 3.9, 2.0)))
    b.dlights, b.plights = bright_lights()          # NEE + shadow rays through the textured BSDFs
    return b.build()


def case_textured_materials(api, size=96, frames=3, n_rays=20000):
    d = textured_scene()
    o = orc.OracleScene(d)
    # closest hits bit-exact, then the rendered image and the texture-dependent debug channels
    ctx, sc = make(api, d, 32, 32)
    rays, _ = util.random_rays(n_rays, seed=23)
    check_ids(sc, o, rays, 1)
    del ctx, sc
    o2 = o
    ctx, sc, st = render_compare(api, d, o2, size, size, dict(number_of_samples=4, number_of_bounces=5), frames, cam_pos=(0, 0, 14.0))
    # counted frame: texture taps and light candidates (bench.py's algorithmic bytes) follow the oracle's counts
    cam = host.Camera(size, size).set(position=(0, 0, 14.0)); gui = host.Gui(number_of_samples=1, number_of_bounces=5)
    u = host.FrameDriver(cam, gui, d.fully_opaque).next_ubo()
    ctx.resize(size, size); ctx.render(sc, u, flags=1); sg = ctx.stats(); _, _, so = o2.render(u, size, size, None)
    assert so.tex_taps > 0 and abs(int(sg.tex_taps) - int(so.tex_taps)) <= 0.003 * so.tex_taps, (sg.tex_taps, so.tex_taps)
    assert abs(int(sg.light_cands) - int(so.light_cands)) <= 0.003 * so.light_cands and abs(int(sg.shaded_hits) - int(so.shaded_hits)) <= 0.003 * so.shaded_hits
    for mapping in (6, 7, 9, 10, 11):          # albedo, normal, metallic, roughness, transmission style channels (RayTracing.rchit:258-286)
        ctx.resize(size, size)
        cam = host.Camera(size, size).set(position=(0, 0, 14.0)); gui = host.Gui(number_of_samples=1, number_of_bounces=1, mapping=mapping, antialiasing=0)
        d1, d2 = host.FrameDriver(cam, gui, d.fully_opaque), host.FrameDriver(cam, gui, d.fully_opaque)
        ctx.render(sc, d1.next_ubo()); _, out, _ = o2.render(d2.next_ubo(), size, size, None)
        _, out_g = ctx.readback()
        diff = np.abs(out_g.astype(int) - out.astype(int))
        assert (diff > 1).mean() < 0.002, (mapping, diff.max(), (diff > 1).mean())    # texture-filtered channels: 1 LSB (silhouette pixels aside)


def case_frame_options(api, cornell_desc, cornell_oracle, size=48):
    """UBO variants of RayTracing.rgen / Tonemapping.glsl / rchit debug path: thin lens (LCG stream), orthographic
    camera, every tone-map mode, DISTANCE and HEAT mappings, debug == 1, no anti-aliasing, spp > 1, bounce limits."""
    variants = [
        dict(aperture=0.6, focus_distance=12.0), dict(orthographic_fov_dis=4.0), dict(selected_tone_map_mode=1), dict(selected_tone_map_mode=2),
        dict(selected_tone_map_mode=3), dict(selected_tone_map_mode=7), dict(mapping=4, map_scale=25.0), dict(mapping=1, map_scale=1.0),
        dict(debug=1), dict(antialiasing=0), dict(number_of_samples=5), dict(number_of_bounces=1), dict(number_of_bounces=2), dict(exposure=0.5, scale=2.0),
        dict(mapping=6), dict(mapping=7), dict(mapping=9), dict(mapping=10), dict(mapping=11),
    ]
    dl, pl = bright_lights()
    ctx, sc = make(api, cornell_desc, size, size)
    cornell_oracle.update_lights(dl, pl); sc.update_lights(dl, pl)
    try:
        cam = host.Camera(size, size).set(position=(0.5, 0.3, 13.0), direction=(-0.05, -0.02, -1))
        for kw in variants:
            base = dict(number_of_samples=2, number_of_bounces=5)
            base.update(kw)
            gui = host.Gui(**base)
            d1, d2 = host.FrameDriver(cam, gui, False), host.FrameDriver(cam, gui, False)
            ctx.resize(size, size)
            acc = None
            for _ in range(2):
                ctx.render(sc, d1.next_ubo()); acc, out, st = cornell_oracle.render(d2.next_ubo(), size, size, acc)
            acc_g, out_g = ctx.readback()
            mre, ps = util.mean_rel_err(acc_g, acc, max(1, d1.total.value)), util.psnr(out_g[..., :3], out[..., :3])
            assert mre < MRE_TOL and ps >= PSNR_TOL, (kw, mre, ps)
    finally:
        d0 = np.frombuffer(util._arr(cornell_desc.dlights, cornell_desc.n_dlights, F.rt_light), F.LIGHT_DTYPE)
        p0 = np.frombuffer(util._arr(cornell_desc.plights, cornell_desc.n_plights, F.rt_light), F.LIGHT_DTYPE)
        cornell_oracle.update_lights(d0, p0)
