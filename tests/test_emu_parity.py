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
