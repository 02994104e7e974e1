"""Pre-GPU logic checks: the CUDA core sources compiled as a host emulation (tests/emu) against the oracle.
These do NOT establish GPU parity (tests/test_gpu_parity.py does, on the B200 box); they keep the builder,
traversal and wavefront shading logic honest on the GPU-less build machine."""
import parity_cases as pc
from emu_lib import emu_api


def test_emu_trace_ids_bit_exact(cornell_desc, cornell_oracle, golden):
    pc.case_trace_golden(emu_api(), cornell_desc, cornell_oracle, golden)


def test_emu_render_matches_golden(cornell_desc, golden):
    pc.case_render_golden(emu_api(), cornell_desc, golden)


def test_emu_tile_partition_is_bit_identical(cornell_desc, golden):
    pc.case_tile_partition(emu_api(), cornell_desc, golden)


def test_emu_frames_in_flight_bookkeeping(cornell_desc, golden):
    pc.case_frames_in_flight(emu_api(), cornell_desc, golden, size=24, frames=4)


def test_emu_instancing_and_tlas_update():
    pc.case_instancing(emu_api())


def test_emu_lights_and_shadow_rays(cornell_desc):
    pc.case_lights(emu_api(), cornell_desc)


def test_emu_skinning_refit_rebuild():
    pc.case_skinning(emu_api())


def test_emu_lucy_ids():
    pc.case_lucy_ids(emu_api(), n_rays=20000, rows=60, cols=61)


def test_emu_foliage_alpha_mask_sky():
    pc.case_foliage(emu_api(), n_side=4, tris=800, size=40, n_rays=8000)


def test_emu_glass_volume():
    pc.case_glass(emu_api(), n_objects=6, size=40, res=(12, 13))


def test_emu_skinned_character():
    pc.case_skinned_character(emu_api(), n_tris=4000, joints=32, size=32, frames=2)


def test_emu_frame_options(cornell_desc, cornell_oracle):
    pc.case_frame_options(emu_api(), cornell_desc, cornell_oracle, size=32)


def test_emu_skinned_in_flight_bookkeeping():
    pc.case_skinned_in_flight(emu_api(), n_tris=1500, joints=16, size=20, frames=4)


def test_emu_textured_materials():
    pc.case_textured_materials(emu_api(), size=40, frames=2, n_rays=4000)


def test_emu_lifecycle_in_flight(cornell_desc):
    pc.case_lifecycle_in_flight(emu_api(), cornell_desc, size=16)


def test_flat_scene_morton_cells_stay_cubic(monkeypatch):
    """A flat field of instances + baked foliage cards (config 3 in small): normalising the Morton codes per axis stretches the
    thin axis until its bits are noise (round 1).  The builder clamps the per-axis stretch (RT_MORTON_MAX_ASPECT): the same
    frame must need clearly fewer node visits than with the round-1 normalisation, and the two images must be identical."""
    import numpy as np
    from rustracer_b200 import core, host, scenes, _ffi as F
    d = scenes.instanced_foliage(n_side=48, tris_per_mesh=1000, cards=32, tex_size=64, sky=scenes.procedural_sky(16))
    cam = host.Camera(96, 54).set(position=(0, 1.2, 7.0))
    gui = host.Gui(number_of_samples=1, number_of_bounces=4, sky=1)
    import ctypes as C
    u = F.rt_ubo(); total = F.c_u32(0)
    F.load_host().gv_build_ubo(C.byref(cam.c), C.byref(gui.g), C.byref(total), 0, 0, 3, C.byref(u))
    out = {}
    for name, env in (("cubic", None), ("per_axis", "1e30")):
        if env: monkeypatch.setenv("RT_B200_MORTON_ASPECT", env)
        ctx = core.Context(96, 54, api=emu_api()); sc = core.Scene(ctx, d)
        ctx.render(sc, u, flags=1)          # RT_RENDER_COUNTERS
        ctx.synchronize()
        st = ctx.stats()
        out[name] = (st.nodes / max(st.rays_extend, 1), ctx.readback()[0].copy())
    assert np.array_equal(out["cubic"][1], out["per_axis"][1])
    assert out["cubic"][0] < 0.9 * out["per_axis"][0], (out["cubic"][0], out["per_axis"][0])


def test_tlas_leaves_hold_one_instance():
    """rt_scene_read_nodes(geo = -2) returns the TLAS; every TLAS leaf slot holds exactly one instance (rt_build.h
    WideOut::leaf_max = 1: an instance is only entered after its own box test), BLAS leaf slots hold up to three triangles."""
    import numpy as np
    from rustracer_b200 import core, scenes
    d = scenes.instanced_foliage(n_side=12, tris_per_mesh=600, cards=200, tex_size=16)    # 200 cards = 400 triangles: not baked
    ctx = core.Context(8, 8, api=emu_api()); sc = core.Scene(ctx, d)
    def leaf_counts(nodes):
        meta = nodes[:, 6:8].copy().view(np.uint8).reshape(-1, 8)
        planes = nodes[:, 8:32].reshape(-1, 3, 8)
        valid = ((planes & 0xFFFF) != 0x7F80).all(1)
        inner = (meta & 0x18) == 0x18
        leaf = valid & ~inner
        return np.array([bin(int(m) >> 5).count("1") for m in meta[leaf]])
    tl = leaf_counts(sc.read_nodes(-2))
    assert len(tl) == 288 and (tl == 1).all()                     # 144 bodies + 144 card sets (+ the ground is baked)
    bl = leaf_counts(sc.read_nodes(0))
    assert bl.max() == 3 and bl.min() >= 1
