"""ctypes mirror of include/rt_b200.h and include/gltf_host.h.

Only struct layouts and library loading live here.  The product library (librt_b200.so, CUDA) is loaded
by `load_rt()`, which fails loudly when the extension is missing — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
RT_LIB = Path(os.environ.get("RT_B200_LIB") or (Path(__file__).resolve().parent / "csrc" / "_build" / "librt_b200.so"))
HOST_LIB = ROOT / "host" / "_build" / "libgltf_host.so"

c_f = C.c_float
c_u32 = C.c_uint32
c_i32 = C.c_int32
c_u64 = C.c_uint64
c_u8p = C.POINTER(C.c_uint8)


class rt_vertex(C.Structure):
    _fields_ = [("position", c_f * 4), ("normal", c_f * 4), ("tangent", c_f * 4), ("color", c_f * 4),
                ("weights", c_f * 4), ("joints", c_u32 * 4), ("uv0", c_f * 2), ("uv1", c_f * 2),
                ("skin_index", c_i32), ("_pad", c_u32 * 3)]


class rt_prim_info(C.Structure):
    _fields_ = [("v_offset", c_u32), ("i_offset", c_u32), ("material_id", c_u32), ("_pad", c_u32)]


class rt_texture_info(C.Structure):
    _fields_ = [("index", c_i32), ("coord", c_i32)]


class rt_material(C.Structure):
    _fields_ = [
        ("alpha_mode", c_u32), ("alpha_cutoff", c_f), ("double_sided", c_u32), ("workflow", c_u32),
        ("_pad0", c_f * 2), ("base_color_texture", rt_texture_info), ("base_color", c_f * 4),
        ("metallic_factor", c_f), ("roughness_factor", c_f), ("metallic_roughness_texture", rt_texture_info),
        ("normal_texture", rt_texture_info), ("emissive_texture", rt_texture_info), ("emissive_factor", c_f * 4),
        ("occlusion_texture", rt_texture_info), ("ior", c_f), ("unlit", c_u32),
        ("transmission_texture", rt_texture_info), ("transmission_factor", c_f), ("transmission_exist", c_u32),
        ("attenuation_color", c_f * 3), ("thickness_factor", c_f), ("thickness_texture", rt_texture_info),
        ("attenuation_distance", c_f), ("volume_exists", c_u32),
        ("specular_texture", rt_texture_info), ("specular_color_texture", rt_texture_info),
        ("specular_color_factor", c_f * 4), ("specular_factor", c_f), ("specular_exist", c_u32), ("_pad1", c_f * 2),
        ("sg_diffuse_factor", c_f * 4), ("sg_specular_glossiness_factor", c_f * 4),
        ("sg_diffuse_texture", rt_texture_info), ("sg_specular_glossiness_texture", rt_texture_info),
    ]


class rt_light(C.Structure):
    _fields_ = [("color", c_f * 4), ("transform", c_f * 4), ("kind", c_u32), ("range", c_f),
                ("intensity", c_f), ("_pad", c_u32)]


class rt_ubo(C.Structure):
    _fields_ = [
        ("model_view", c_f * 16), ("projection", c_f * 16), ("model_view_inverse", c_f * 16),
        ("projection_inverse", c_f * 16),
        ("aperture", c_f), ("focus_distance", c_f), ("fov_angle", c_f), ("orthographic_fov_dis", c_f),
        ("heatmap_scale", c_f), ("total_number_of_samples", c_u32), ("number_of_samples", c_u32),
        ("number_of_bounces", c_u32), ("random_seed", c_u32), ("has_sky", c_u32), ("antialiasing", c_u32),
        ("mapping", c_u32), ("frame_count", c_u32), ("debug", c_u32), ("fully_opaque", c_u32),
        ("exposure", c_f), ("tone_mapping_mode", c_u32),
    ]


class rt_geometry(C.Structure):
    _fields_ = [("v_len", c_u32), ("i_len", c_u32), ("opaque", c_u32), ("_pad", c_u32)]


class rt_instance(C.Structure):
    _fields_ = [("transform", c_f * 12), ("geo_id", c_u32), ("mask", c_u32), ("flags", c_u32), ("_pad", c_u32)]


class rt_image_desc(C.Structure):
    _fields_ = [("rgba8", c_u8p), ("width", c_u32), ("height", c_u32), ("srgb", c_u32), ("_pad", c_u32)]


class rt_sampler_desc(C.Structure):
    _fields_ = [("mag_filter", c_u32), ("min_filter", c_u32), ("wrap_s", c_u32), ("wrap_t", c_u32)]


class rt_texture_desc(C.Structure):
    _fields_ = [("image_index", c_u32), ("sampler_index", c_u32)]


class rt_scene_desc(C.Structure):
    _fields_ = [
        ("vertices", C.POINTER(rt_vertex)), ("n_vertices", c_u32),
        ("indices", C.POINTER(c_u32)), ("n_indices", c_u32),
        ("prim_infos", C.POINTER(rt_prim_info)),
        ("geometries", C.POINTER(rt_geometry)), ("n_geometries", c_u32),
        ("materials", C.POINTER(rt_material)), ("n_materials", c_u32),
        ("instances", C.POINTER(rt_instance)), ("n_instances", c_u32),
        ("images", C.POINTER(rt_image_desc)), ("n_images", c_u32),
        ("samplers", C.POINTER(rt_sampler_desc)), ("n_samplers", c_u32),
        ("textures", C.POINTER(rt_texture_desc)), ("n_textures", c_u32),
        ("dlights", C.POINTER(rt_light)), ("n_dlights", c_u32),
        ("plights", C.POINTER(rt_light)), ("n_plights", c_u32),
        ("skins", C.POINTER(c_f)), ("n_skins", c_u32),
        ("skybox_faces", c_u8p * 6),
        ("skybox_width", c_u32), ("skybox_height", c_u32), ("skybox_srgb", c_u32), ("_pad", c_u32),
    ]


class rt_combine_desc(C.Structure):
    _fields_ = [("peer_blocks", C.POINTER(C.c_void_p)), ("n_peers", c_u32), ("epoch", c_u32), ("row0", c_u32), ("row1", c_u32),
                ("strip_rows", c_u32), ("n_parts", c_u32), ("part", c_u32), ("gather_to", C.c_int32), ("n_senders", c_u32),
                ("gather_acc", c_u32)]


RT_GATHER_ALL, RT_GATHER_NONE = -1, -2
RT_PARTITION_TILES, RT_PARTITION_SAMPLE_PASSES = 0, 1


class rt_ray(C.Structure):
    _fields_ = [("origin", c_f * 3), ("tmin", c_f), ("direction", c_f * 3), ("tmax", c_f)]


class rt_hit(C.Structure):
    _fields_ = [("t", c_f), ("u", c_f), ("v", c_f), ("instance_id", c_u32), ("primitive_id", c_u32),
                ("geo_id", c_u32)]


class rt_stats(C.Structure):
    _fields_ = [("ms_total", c_f), ("ms_raygen", c_f), ("ms_extend", c_f), ("ms_shade", c_f), ("ms_shadow", c_f),
                ("ms_accum", c_f), ("n_extend_launches", c_u32), ("n_kernel_launches", c_u32),
                ("rays_extend", c_u64), ("rays_shadow", c_u64), ("shaded_hits", c_u64), ("pixel_samples", c_u64),
                ("nodes", c_u64), ("tris", c_u64), ("insts", c_u64), ("anyhits", c_u64), ("tex_taps", c_u64),
                ("light_cands", c_u64)]


class rt_render_opts(C.Structure):
    _fields_ = [("flags", c_u32), ("strip_rows", c_u32), ("n_parts", c_u32), ("part", c_u32)]


class rt_bvh_info(C.Structure):
    _fields_ = [("blas_nodes", c_u64), ("blas_tris", c_u64), ("tlas_nodes", c_u64), ("bytes", c_u64),
                ("max_depth_blas", c_u32), ("max_depth_tlas", c_u32),
                ("build_ms", c_f), ("refit_ms", c_f), ("skin_ms", c_f), ("tlas_ms", c_f)]


class gv_camera(C.Structure):
    _fields_ = [("position", c_f * 3), ("direction", c_f * 3), ("fov", c_f), ("aspect_ratio", c_f),
                ("z_near", c_f), ("z_far", c_f)]


class gv_gui(C.Structure):
    _fields_ = [("aperture", c_f), ("focus_distance", c_f), ("number_of_samples", c_u32),
                ("number_of_bounces", c_u32), ("max_number_of_samples", c_u32), ("acc", c_u32), ("sky", c_u32),
                ("antialiasing", c_u32), ("debug", c_u32), ("mapping", c_u32), ("animation", c_u32),
                ("map_scale", c_f), ("scale", c_f), ("orthographic_fov_dis", c_f), ("exposure", c_f),
                ("selected_tone_map_mode", c_u32)]


# numpy views of the same layouts (used to build scenes without per-field ctypes traffic)
VERTEX_DTYPE = np.dtype([("position", "<f4", 4), ("normal", "<f4", 4), ("tangent", "<f4", 4), ("color", "<f4", 4),
                         ("weights", "<f4", 4), ("joints", "<u4", 4), ("uv0", "<f4", 2), ("uv1", "<f4", 2),
                         ("skin_index", "<i4"), ("_pad", "<u4", 3)])
RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("tmin", "<f4"), ("direction", "<f4", 3), ("tmax", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("instance_id", "<u4"), ("primitive_id", "<u4"),
                      ("geo_id", "<u4")])
INSTANCE_DTYPE = np.dtype([("transform", "<f4", 12), ("geo_id", "<u4"), ("mask", "<u4"), ("flags", "<u4"),
                           ("_pad", "<u4")])
LIGHT_DTYPE = np.dtype([("color", "<f4", 4), ("transform", "<f4", 4), ("kind", "<u4"), ("range", "<f4"),
                        ("intensity", "<f4"), ("_pad", "<u4")])
assert VERTEX_DTYPE.itemsize == C.sizeof(rt_vertex) == 128
assert C.sizeof(rt_material) == 256 and C.sizeof(rt_ubo) == 324 and C.sizeof(rt_light) == 48
assert C.sizeof(rt_instance) == 64 == INSTANCE_DTYPE.itemsize and C.sizeof(rt_prim_info) == 16
assert C.sizeof(rt_ray) == 32 == RAY_DTYPE.itemsize and C.sizeof(rt_hit) == 24 == HIT_DTYPE.itemsize

RT_EXPORTS = [
    "rt_last_error", "rt_version", "rt_context_create", "rt_context_destroy", "rt_frame_resize",
    "rt_scene_create", "rt_scene_destroy", "rt_scene_update_instances", "rt_scene_update_skins",
    "rt_scene_update_lights", "rt_scene_set_skybox", "rt_render", "rt_tonemap", "rt_synchronize",
    "rt_readback", "rt_upload_accumulation", "rt_device_ptrs", "rt_last_frame_stats", "rt_trace_closest",
    "rt_trace_any", "rt_scene_read_vertices", "rt_scene_bvh_info", "rt_ipc_export", "rt_ipc_open",
    "rt_ipc_close", "rt_context_set_frames_in_flight", "rt_join", "rt_readback_async",
    "rt_frame_wait", "rt_scene_read_nodes", "rt_scene_set_versions",
    "rt_combine", "rt_readback_display", "rt_combine_ptrs",
    "rt_multi_create", "rt_multi_destroy", "rt_multi_scene_create", "rt_multi_render", "rt_multi_combine",
    "rt_multi_readback", "rt_multi_synchronize", "rt_multi_replica",
]
# multi-GPU entry points: CUDA-only (the test-only host emulation has no peers to talk to)
RT_CUDA_ONLY = ("rt_ipc_export", "rt_ipc_open", "rt_ipc_close", "rt_combine", "rt_readback_display", "rt_combine_ptrs",
                "rt_multi_create", "rt_multi_destroy", "rt_multi_scene_create", "rt_multi_render", "rt_multi_combine",
                "rt_multi_readback", "rt_multi_synchronize", "rt_multi_replica")
GV_EXPORTS = [
    "gv_last_error", "gv_load_file", "gv_doc_free", "gv_doc_scene_desc", "gv_doc_fully_opaque",
    "gv_doc_static_scene", "gv_doc_need_compute", "gv_doc_aabb_trans", "gv_doc_animate", "gv_doc_get_skins",
    "gv_doc_get_instances", "gv_doc_set_skybox", "gv_camera_default", "gv_camera_view_matrix",
    "gv_camera_projection_matrix", "gv_mat4_inverse", "gv_gui_default", "gv_build_ubo", "gv_decode_png", "gv_decode_image",
    "gv_load_skybox_dir", "gv_free", "gv_generate_tangents",
]

_rt = None
_host = None


def load_host() -> C.CDLL:
    global _host
    if _host is None:
        if not HOST_LIB.exists():
            raise RuntimeError(f"host library missing: {HOST_LIB} (run `python -c 'import __graft_entry__ as g; g.build()'`)")
        lib = C.CDLL(str(HOST_LIB))
        lib.gv_last_error.restype = C.c_char_p
        lib.gv_load_file.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        lib.gv_doc_free.argtypes = [C.c_void_p]
        lib.gv_doc_free.restype = None
        lib.gv_doc_scene_desc.argtypes = [C.c_void_p, C.POINTER(rt_scene_desc)]
        for f in ("gv_doc_fully_opaque", "gv_doc_static_scene", "gv_doc_need_compute"):
            getattr(lib, f).argtypes = [C.c_void_p]
        lib.gv_doc_aabb_trans.argtypes = [C.c_void_p, C.POINTER(c_f)]
        lib.gv_doc_aabb_trans.restype = None
        lib.gv_doc_animate.argtypes = [C.c_void_p, c_f]
        lib.gv_doc_get_skins.argtypes = [C.c_void_p, C.POINTER(C.POINTER(c_f)), C.POINTER(c_u32)]
        lib.gv_doc_get_instances.argtypes = [C.c_void_p, C.POINTER(C.POINTER(rt_instance)), C.POINTER(c_u32)]
        lib.gv_doc_set_skybox.argtypes = [C.c_void_p, c_u8p * 6, c_u32, c_u32, c_u32]
        lib.gv_camera_default.argtypes = [C.POINTER(gv_camera), c_u32, c_u32]
        lib.gv_camera_default.restype = None
        lib.gv_camera_view_matrix.argtypes = [C.POINTER(gv_camera), C.POINTER(c_f)]
        lib.gv_camera_view_matrix.restype = None
        lib.gv_camera_projection_matrix.argtypes = [C.POINTER(gv_camera), C.POINTER(c_f)]
        lib.gv_camera_projection_matrix.restype = None
        lib.gv_mat4_inverse.argtypes = [C.POINTER(c_f), C.POINTER(c_f)]
        lib.gv_gui_default.argtypes = [C.POINTER(gv_gui)]
        lib.gv_gui_default.restype = None
        lib.gv_build_ubo.argtypes = [C.POINTER(gv_camera), C.POINTER(gv_gui), C.POINTER(c_u32), c_u32, c_u32, c_u32,
                                     C.POINTER(rt_ubo)]
        lib.gv_build_ubo.restype = None
        lib.gv_decode_png.argtypes = [c_u8p, C.c_size_t, C.POINTER(c_u8p), C.POINTER(c_u32), C.POINTER(c_u32)]
        lib.gv_decode_image.argtypes = [c_u8p, C.c_size_t, C.POINTER(c_u8p), C.POINTER(c_u32), C.POINTER(c_u32)]
        lib.gv_load_skybox_dir.argtypes = [C.c_char_p, c_u8p * 6, C.POINTER(c_u32), C.POINTER(c_u32)]
        lib.gv_generate_tangents.argtypes = [C.POINTER(rt_vertex), c_u32, C.POINTER(c_u32), c_u32]
        lib.gv_generate_tangents.restype = None
        lib.gv_free.argtypes = [C.c_void_p]
        lib.gv_free.restype = None
        _host = lib
    return _host


def bind_rt(lib: C.CDLL, rename=lambda n: n, optional=()) -> C.CDLL:
    """Attach argument/return types for the rt_* entry points of include/rt_b200.h to `lib`."""
    vp, vpp = C.c_void_p, C.POINTER(C.c_void_p)
    sigs = {
        "rt_last_error": ([], C.c_char_p), "rt_version": ([], C.c_char_p),
        "rt_context_create": ([C.c_int, c_u32, c_u32, vpp], C.c_int),
        "rt_context_destroy": ([vp], None),
        "rt_frame_resize": ([vp, c_u32, c_u32], C.c_int),
        "rt_scene_create": ([vp, C.POINTER(rt_scene_desc), vpp], C.c_int),
        "rt_scene_destroy": ([vp], None),
        "rt_scene_update_instances": ([vp, C.POINTER(rt_instance), c_u32], C.c_int),
        "rt_scene_update_skins": ([vp, C.POINTER(c_f), c_u32, C.c_int], C.c_int),
        "rt_scene_update_lights": ([vp, C.POINTER(rt_light), c_u32, C.POINTER(rt_light), c_u32], C.c_int),
        "rt_scene_set_skybox": ([vp, c_u8p * 6, c_u32, c_u32, c_u32], C.c_int),
        "rt_render": ([vp, vp, C.POINTER(rt_ubo), C.POINTER(rt_render_opts), vp], C.c_int),
        "rt_tonemap": ([vp, C.POINTER(rt_ubo), vp], C.c_int),
        "rt_synchronize": ([vp], C.c_int),
        "rt_readback": ([vp, C.POINTER(c_f), c_u8p], C.c_int),
        "rt_upload_accumulation": ([vp, C.POINTER(c_f)], C.c_int),
        "rt_device_ptrs": ([vp, vpp, vpp], C.c_int),
        "rt_last_frame_stats": ([vp, C.POINTER(rt_stats)], C.c_int),
        "rt_trace_closest": ([vp, C.POINTER(rt_ray), c_u32, c_u32, C.POINTER(c_u32), C.POINTER(rt_hit)], C.c_int),
        "rt_trace_any": ([vp, C.POINTER(rt_ray), c_u32, c_u32, C.POINTER(c_u32), c_u8p], C.c_int),
        "rt_scene_read_vertices": ([vp, C.POINTER(rt_vertex), c_u32], C.c_int),
        "rt_scene_bvh_info": ([vp, C.POINTER(rt_bvh_info)], C.c_int),
        "rt_ipc_export": ([vp, vp], C.c_int),
        "rt_ipc_open": ([vp, vp, vpp], C.c_int),
        "rt_ipc_close": ([vp, vp], C.c_int),
        "rt_context_set_frames_in_flight": ([vp, c_u32], C.c_int),
        "rt_join": ([vp, vp], C.c_int),
        "rt_readback_async": ([vp, c_u8p, C.POINTER(C.c_uint64)], C.c_int),
        "rt_frame_wait": ([vp, C.c_uint64], C.c_int),
        "rt_scene_set_versions": ([vp, c_u32], C.c_int),
        "rt_scene_read_nodes": ([vp, C.c_int, C.POINTER(c_f), c_u32, C.POINTER(c_u32)], C.c_int),
        "rt_combine": ([vp, C.POINTER(rt_combine_desc), C.POINTER(rt_ubo), vp], C.c_int),
        "rt_readback_display": ([vp, c_u8p, C.POINTER(c_f)], C.c_int),
        "rt_combine_ptrs": ([vp, vpp, vpp, C.POINTER(C.c_uint64)], C.c_int),
        "rt_multi_create": ([C.POINTER(C.c_int), c_u32, c_u32, c_u32, c_u32, vpp], C.c_int),
        "rt_multi_destroy": ([vp], None),
        "rt_multi_scene_create": ([vp, C.POINTER(rt_scene_desc)], C.c_int),
        "rt_multi_render": ([vp, C.POINTER(rt_ubo)], C.c_int),
        "rt_multi_combine": ([vp, C.POINTER(rt_ubo)], C.c_int),
        "rt_multi_readback": ([vp, C.POINTER(rt_ubo), C.POINTER(c_f), c_u8p], C.c_int),
        "rt_multi_synchronize": ([vp], C.c_int),
        "rt_multi_replica": ([vp, c_u32, vpp, vpp], C.c_int),
    }
    assert set(sigs) == set(RT_EXPORTS)
    for name, (args, res) in sigs.items():
        try:
            fn = getattr(lib, rename(name))
        except AttributeError:
            if name in optional:
                continue
            raise
        fn.argtypes, fn.restype = args, res
    return lib


def load_rt() -> C.CDLL:
    """Load the CUDA core.  Raises (never falls back) when the extension has not been built."""
    global _rt
    if _rt is None:
        if not RT_LIB.exists():
            raise RuntimeError(f"CUDA extension missing: {RT_LIB} — build it with __graft_entry__.build(); "
                               "rustracer_b200 has no CPU fallback")
        _rt = bind_rt(C.CDLL(str(RT_LIB)))
    return _rt


def as_ptr(arr: np.ndarray, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))
