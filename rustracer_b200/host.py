"""Python view of the C++ host layer (host/gltf_host.cpp, include/gltf_host.h).

Mirrors the reference's host-side names: `load_file` -> Doc (asset_loader/src/scene_graph.rs:438),
`Doc.animate`, `Doc.get_skins`, `Camera` (app/src/camera.rs), `Gui` defaults (gltf_viewer/src/gui_state.rs:303)
and the per-frame UBO fill of `GltfViewer::update` (gltf_viewer/src/main.rs:189-242).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi as F


class HostError(RuntimeError):
    pass


class Doc:
    """asset_loader::Doc — owns the flat GPU arrays built from a glTF file."""

    def __init__(self, handle):
        self._h = handle
        self._lib = F.load_host()

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.gv_doc_free(self._h)
            self._h = None

    def scene_desc(self) -> F.rt_scene_desc:
        d = F.rt_scene_desc()
        if self._lib.gv_doc_scene_desc(self._h, C.byref(d)):
            raise HostError(self._lib.gv_last_error().decode())
        d._owner = self  # keep the doc alive while the desc is referenced
        return d

    def fully_opaque(self) -> bool:
        return bool(self._lib.gv_doc_fully_opaque(self._h))

    def static_scene(self) -> bool:
        return bool(self._lib.gv_doc_static_scene(self._h))

    def need_compute(self) -> bool:
        return bool(self._lib.gv_doc_need_compute(self._h))

    def aabb_trans(self) -> np.ndarray:
        out = (F.c_f * 16)()
        self._lib.gv_doc_aabb_trans(self._h, out)
        return np.array(out, dtype=np.float32).reshape(4, 4).T  # column-major storage -> math layout

    def animate(self, t: float) -> None:
        if self._lib.gv_doc_animate(self._h, float(t)):
            raise HostError(self._lib.gv_last_error().decode())

    def get_skins(self) -> np.ndarray:
        p = C.POINTER(F.c_f)()
        n = F.c_u32()
        self._lib.gv_doc_get_skins(self._h, C.byref(p), C.byref(n))
        if n.value == 0:
            return np.zeros((0, 256, 16), np.float32)
        return np.ctypeslib.as_array(p, shape=(n.value, 256, 16)).copy()

    def get_instances(self) -> np.ndarray:
        p = C.POINTER(F.rt_instance)()
        n = F.c_u32()
        self._lib.gv_doc_get_instances(self._h, C.byref(p), C.byref(n))
        buf = C.string_at(p, n.value * C.sizeof(F.rt_instance))
        return np.frombuffer(buf, dtype=F.INSTANCE_DTYPE).copy()

    def set_skybox(self, faces, srgb: bool = True) -> None:
        """faces: six HxWx4 uint8 arrays ordered +x,-x,+y,-y,+z,-z (asset_loader/src/cubumap.rs:20-50)."""
        faces = [np.ascontiguousarray(f, dtype=np.uint8) for f in faces]
        h, w = faces[0].shape[:2]
        arr = (F.c_u8p * 6)(*[f.ctypes.data_as(F.c_u8p) for f in faces])
        if self._lib.gv_doc_set_skybox(self._h, arr, w, h, int(srgb)):
            raise HostError(self._lib.gv_last_error().decode())


def decode_image(data: bytes) -> np.ndarray:
    """PNG / JPEG bytes -> HxWx4 uint8 (asset_loader::image::Image::load_image, image.rs:60-83)."""
    lib = F.load_host()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    out = F.c_u8p(); w = F.c_u32(); h = F.c_u32()
    if lib.gv_decode_image(buf, len(data), C.byref(out), C.byref(w), C.byref(h)):
        raise HostError(lib.gv_last_error().decode())
    try:
        return np.ctypeslib.as_array(out, shape=(h.value, w.value, 4)).copy()
    finally:
        lib.gv_free(out)


def load_skybox_dir(path: str) -> list:
    """asset_loader::cubumap::SkyBox::new (cubumap.rs:86-106): six faces, +x,-x,+y,-y,+z,-z, each HxWx4 uint8."""
    lib = F.load_host()
    faces = (F.c_u8p * 6)(); w = F.c_u32(); h = F.c_u32()
    if lib.gv_load_skybox_dir(str(path).encode(), faces, C.byref(w), C.byref(h)):
        raise HostError(lib.gv_last_error().decode())
    out = []
    for f in range(6):
        out.append(np.ctypeslib.as_array(faces[f], shape=(h.value, w.value, 4)).copy())
        lib.gv_free(faces[f])
    return out


def load_file(path: str) -> Doc:
    lib = F.load_host()
    h = C.c_void_p()
    if lib.gv_load_file(str(path).encode(), C.byref(h)):
        raise HostError(lib.gv_last_error().decode())
    return Doc(h)


class Camera:
    """app::camera::Camera (crates/libs/app/src/camera.rs:15-118)."""

    def __init__(self, width: int, height: int):
        self.c = F.gv_camera()
        F.load_host().gv_camera_default(C.byref(self.c), width, height)

    def set(self, position=None, direction=None, fov=None):
        if position is not None:
            self.c.position[:] = [float(x) for x in position]
        if direction is not None:
            d = np.asarray(direction, np.float32)
            d = d / np.linalg.norm(d)
            self.c.direction[:] = [float(x) for x in d]
        if fov is not None:
            self.c.fov = float(fov)
        return self

    def view_matrix(self) -> np.ndarray:
        o = (F.c_f * 16)()
        F.load_host().gv_camera_view_matrix(C.byref(self.c), o)
        return np.array(o, np.float32).reshape(4, 4).T

    def projection_matrix(self) -> np.ndarray:
        o = (F.c_f * 16)()
        F.load_host().gv_camera_projection_matrix(C.byref(self.c), o)
        return np.array(o, np.float32).reshape(4, 4).T


class Gui:
    """gltf_viewer gui_state.rs Gui with Gui::new defaults."""

    def __init__(self, **overrides):
        self.g = F.gv_gui()
        F.load_host().gv_gui_default(C.byref(self.g))
        for k, v in overrides.items():
            if not hasattr(self.g, k):
                raise AttributeError(k)
            setattr(self.g, k, v)


class FrameDriver:
    """Per-frame bookkeeping of GltfViewer::update (main.rs:189-242): sample budgeting, UBO fill."""

    def __init__(self, camera: Camera, gui: Gui, fully_opaque: bool, random_seed: int = 3):
        self.camera, self.gui = camera, gui
        self.fully_opaque = int(fully_opaque)
        self.random_seed = random_seed
        self.total = F.c_u32(0)
        self.frame_count = 0  # monotonic (deviation D1: the reference resets it every wall-clock second)

    def reset_samples(self):
        self.total = F.c_u32(0)

    def next_ubo(self) -> F.rt_ubo:
        u = F.rt_ubo()
        F.load_host().gv_build_ubo(C.byref(self.camera.c), C.byref(self.gui.g), C.byref(self.total), self.frame_count,
                                   self.fully_opaque, self.random_seed, C.byref(u))
        self.frame_count += 1
        return u
