"""Python view of the C ABI (include/rt_b200.h): Context / Scene / render / readback / trace.

This is the call surface a host uses in place of the reference's Vulkan calls.  It binds the
CUDA library built from rustracer_b200/csrc; if that library is missing, construction raises — there is no
CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi as F


class RtError(RuntimeError):
    pass


class Api:
    """Function table over a loaded library exporting the rt_* entry points."""

    def __init__(self, lib: C.CDLL | None = None, rename=lambda n: n):
        self.lib = lib if lib is not None else F.load_rt()
        self._rename = rename

    def __getattr__(self, name):
        if name.startswith("rt_"):
            return getattr(self.lib, self._rename(name))
        raise AttributeError(name)

    def check(self, rc: int):
        if rc:
            raise RtError(self.rt_last_error().decode())


class Context:
    """Replaces vulkan::Context + the storage / accumulation images of BaseApp."""

    def __init__(self, width: int, height: int, device: int = 0, api: Api | None = None):
        self.api = api or Api()
        self.width, self.height = width, height
        self._h = C.c_void_p()
        self.api.check(self.api.rt_context_create(device, width, height, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self.api.rt_context_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def resize(self, width: int, height: int):
        self.api.check(self.api.rt_frame_resize(self._h, width, height))
        self.width, self.height = width, height

    def render(self, scene: "Scene", ubo: F.rt_ubo, flags: int = 0, strip_rows: int = 0, n_parts: int = 0, part: int = 0,
               stream=None):
        opts = F.rt_render_opts(flags, strip_rows, n_parts, part)
        self.api.check(self.api.rt_render(self._h, scene._h, C.byref(ubo), C.byref(opts), stream))

    def set_frames_in_flight(self, n: int):
        """InFlightFrames of the reference (app/src/lib.rs:34): up to n rt_render calls overlap on the GPU; the
        accumulation stays in submission order, so images are bit-identical to n = 1."""
        self.api.check(self.api.rt_context_set_frames_in_flight(self._h, n))

    def join(self, stream=None):
        self.api.check(self.api.rt_join(self._h, stream))

    def readback_async(self, out_buf: np.ndarray) -> int:
        """Queues the device->host copy of the last submitted frame's RGBA8 image; returns its ticket."""
        assert out_buf.dtype == np.uint8 and out_buf.size == self.width * self.height * 4 and out_buf.flags.c_contiguous
        t = C.c_uint64()
        self.api.check(self.api.rt_readback_async(self._h, out_buf.ctypes.data_as(F.c_u8p), C.byref(t)))
        return int(t.value)

    def frame_wait(self, ticket: int):
        self.api.check(self.api.rt_frame_wait(self._h, ticket))

    def tonemap(self, ubo: F.rt_ubo, stream=None):
        self.api.check(self.api.rt_tonemap(self._h, C.byref(ubo), stream))

    def synchronize(self):
        self.api.check(self.api.rt_synchronize(self._h))

    def readback(self, want_acc: bool = True, want_out: bool = True, acc_buf=None, out_buf=None):
        acc = out = None
        if want_acc:
            acc = acc_buf if acc_buf is not None else np.empty((self.height, self.width, 4), np.float32)
        if want_out:
            out = out_buf if out_buf is not None else np.empty((self.height, self.width, 4), np.uint8)
        self.api.check(self.api.rt_readback(self._h, F.as_ptr(acc, F.c_f) if acc is not None else None,
                                            out.ctypes.data_as(F.c_u8p) if out is not None else None))
        return acc, out

    def upload_accumulation(self, acc: np.ndarray):
        acc = np.ascontiguousarray(acc, np.float32)
        assert acc.shape == (self.height, self.width, 4)
        self.api.check(self.api.rt_upload_accumulation(self._h, F.as_ptr(acc, F.c_f)))

    def device_ptrs(self):
        a, o = C.c_void_p(), C.c_void_p()
        self.api.check(self.api.rt_device_ptrs(self._h, C.byref(a), C.byref(o)))
        return a.value, o.value

    # ---- multi-GPU combine (include/rt_b200.h "Multi-GPU combine, device-synchronised") ----
    def ipc_handle(self) -> bytes:
        h = (C.c_uint8 * 64)()
        self.api.check(self.api.rt_ipc_export(self._h, h))
        return bytes(h)

    def ipc_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        self.api.check(self.api.rt_ipc_open(self._h, (C.c_uint8 * 64).from_buffer_copy(handle), C.byref(p)))
        return p.value

    def combine_ptrs(self):
        b, d, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        self.api.check(self.api.rt_combine_ptrs(self._h, C.byref(b), C.byref(d), C.byref(n)))
        return b.value, d.value, int(n.value)

    def combine(self, peers, ubo: F.rt_ubo, epoch: int, rows=(0, 0), tiles=None, gather_to=F.RT_GATHER_ALL, n_senders=None,
                gather_acc=False, stream=None):
        """rt_combine: fused peer-memory reduce + tonemap + gather of this rank's band.  peers: accumulation-block pointers of
        the other ranks.  tiles = (strip_rows, n_parts, part) for the tile partition."""
        arr = (C.c_void_p * max(1, len(peers)))(*peers)
        d = F.rt_combine_desc(arr, len(peers), epoch, rows[0], rows[1], 0, 0, 0, gather_to,
                              (len(peers) if gather_to == F.RT_GATHER_ALL else 0) if n_senders is None else n_senders, int(gather_acc))
        if tiles:
            d.strip_rows, d.n_parts, d.part = tiles
        self.api.check(self.api.rt_combine(self._h, C.byref(d), C.byref(ubo), stream))

    def readback_display(self, want_sum: bool = False):
        out = np.empty((self.height, self.width, 4), np.uint8)
        acc = np.empty((self.height, self.width, 4), np.float32) if want_sum else None
        self.api.check(self.api.rt_readback_display(self._h, out.ctypes.data_as(F.c_u8p), F.as_ptr(acc, F.c_f) if want_sum else None))
        return out, acc

    def stats(self) -> F.rt_stats:
        st = F.rt_stats()
        self.api.check(self.api.rt_last_frame_stats(self._h, C.byref(st)))
        return st


class Scene:
    """Replaces create_global + Buffers::new + create_as + pipeline/descriptor setup."""

    def __init__(self, ctx: Context, desc: F.rt_scene_desc):
        self.ctx, self.api = ctx, ctx.api
        self._h = C.c_void_p()
        self.n_vertices = desc.n_vertices
        self.api.check(self.api.rt_scene_create(ctx._h, C.byref(desc), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self.api.rt_scene_destroy(self._h)
        self._h = None

    def __del__(self):
        self.close()

    def update_instances(self, inst: np.ndarray):
        inst = np.ascontiguousarray(inst, F.INSTANCE_DTYPE)
        self.api.check(self.api.rt_scene_update_instances(self._h, F.as_ptr(inst, F.rt_instance), len(inst)))

    def update_skins(self, mats: np.ndarray, rebuild: bool = False):
        mats = np.ascontiguousarray(mats, np.float32)
        self.api.check(self.api.rt_scene_update_skins(self._h, F.as_ptr(mats, F.c_f), mats.size // 4096, int(rebuild)))

    def set_versions(self, n: int):
        """Copies of the buffers a skin update rewrites (default 2 for skinned scenes): updates overlap frames in flight."""
        self.api.check(self.api.rt_scene_set_versions(self._h, n))

    def update_lights(self, dlights: np.ndarray, plights: np.ndarray):
        d = np.ascontiguousarray(dlights, F.LIGHT_DTYPE)
        p = np.ascontiguousarray(plights, F.LIGHT_DTYPE)
        self.api.check(self.api.rt_scene_update_lights(self._h, F.as_ptr(d, F.rt_light), len(d), F.as_ptr(p, F.rt_light), len(p)))

    def set_skybox(self, faces, srgb: bool = True):
        faces = [np.ascontiguousarray(f, np.uint8) for f in faces]
        h, w = faces[0].shape[:2]
        arr = (F.c_u8p * 6)(*[f.ctypes.data_as(F.c_u8p) for f in faces])
        self.api.check(self.api.rt_scene_set_skybox(self._h, arr, w, h, int(srgb)))

    def _rng(self, rng4, n):
        if rng4 is None:
            return None, None
        a = np.ascontiguousarray(rng4, np.uint32).reshape(n, 4)
        return a, F.as_ptr(a, F.c_u32)

    def trace_closest(self, rays: np.ndarray, flags: int = 0, rng4=None) -> np.ndarray:
        rays = np.ascontiguousarray(rays, F.RAY_DTYPE)
        hits = np.zeros(len(rays), F.HIT_DTYPE)
        keep, rp = self._rng(rng4, len(rays))
        self.api.check(self.api.rt_trace_closest(self._h, F.as_ptr(rays, F.rt_ray), len(rays), flags, rp, F.as_ptr(hits, F.rt_hit)))
        return hits

    def trace_any(self, rays: np.ndarray, flags: int = 0, rng4=None) -> np.ndarray:
        rays = np.ascontiguousarray(rays, F.RAY_DTYPE)
        occ = np.zeros(len(rays), np.uint8)
        keep, rp = self._rng(rng4, len(rays))
        self.api.check(self.api.rt_trace_any(self._h, F.as_ptr(rays, F.rt_ray), len(rays), flags, rp, occ.ctypes.data_as(F.c_u8p)))
        return occ

    def read_vertices(self, n: int | None = None) -> np.ndarray:
        n = self.n_vertices if n is None else n
        v = np.zeros(n, F.VERTEX_DTYPE)
        self.api.check(self.api.rt_scene_read_vertices(self._h, F.as_ptr(v, F.rt_vertex), n))
        return v

    def read_nodes(self, geo: int = -1) -> np.ndarray:
        """128-byte wide nodes as (n, 32) uint32: geo >= 0 that geometry's BLAS, -1 the merged world-space BLAS, -2 the TLAS."""
        n = F.c_u32()
        self.api.check(self.api.rt_scene_read_nodes(self._h, geo, None, 0, C.byref(n)))
        out = np.zeros((n.value, 32), np.float32)
        if n.value:
            self.api.check(self.api.rt_scene_read_nodes(self._h, geo, F.as_ptr(out, F.c_f), n.value, C.byref(n)))
        return out.view(np.uint32)

    def bvh_info(self) -> F.rt_bvh_info:
        info = F.rt_bvh_info()
        self.api.check(self.api.rt_scene_bvh_info(self._h, C.byref(info)))
        return info


class Multi:
    """rt_multi: one process, several GPUs (SURVEY.md §8b last row).  mode: F.RT_PARTITION_TILES (every device renders its
    strips of each frame; bit-identical to one GPU) or F.RT_PARTITION_SAMPLE_PASSES (whole frames round-robin)."""

    def __init__(self, devices, width: int, height: int, mode: int, api: Api | None = None):
        self.api = api or Api()
        self.width, self.height, self.n = width, height, len(devices)
        self._h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        self.api.check(self.api.rt_multi_create(arr, len(devices), width, height, mode, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None):
            self.api.rt_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def scene(self, desc: F.rt_scene_desc):
        self.api.check(self.api.rt_multi_scene_create(self._h, C.byref(desc)))

    def render(self, ubo: F.rt_ubo):
        self.api.check(self.api.rt_multi_render(self._h, C.byref(ubo)))

    def combine(self, ubo: F.rt_ubo):
        self.api.check(self.api.rt_multi_combine(self._h, C.byref(ubo)))

    def synchronize(self):
        self.api.check(self.api.rt_multi_synchronize(self._h))

    def readback(self, ubo: F.rt_ubo):
        acc = np.empty((self.height, self.width, 4), np.float32); out = np.empty((self.height, self.width, 4), np.uint8)
        self.api.check(self.api.rt_multi_readback(self._h, C.byref(ubo), F.as_ptr(acc, F.c_f), out.ctypes.data_as(F.c_u8p)))
        return acc, out
