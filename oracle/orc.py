"""TEST INFRASTRUCTURE — ctypes view of the CPU oracle (oracle/oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import this.
The product package (rustracer_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from rustracer_b200 import _ffi as F

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "liborc.so"
_lib = None


def build(force: bool = False) -> None:
    if force or not LIB.exists() or LIB.stat().st_mtime < max((HERE / "oracle.cpp").stat().st_mtime,
                                                                (HERE / "orc_math.h").stat().st_mtime):
        subprocess.check_call(["make", "-C", str(HERE)], stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB.exists():
            build()
        l = C.CDLL(str(LIB))
        vp, vpp = C.c_void_p, C.POINTER(C.c_void_p)
        l.orc_last_error.restype = C.c_char_p
        l.orc_scene_create.argtypes = [C.POINTER(F.rt_scene_desc), vpp]
        l.orc_scene_destroy.argtypes = [vp]
        l.orc_scene_destroy.restype = None
        l.orc_scene_update_instances.argtypes = [vp, C.POINTER(F.rt_instance), F.c_u32]
        l.orc_scene_update_skins.argtypes = [vp, C.POINTER(F.c_f), F.c_u32]
        l.orc_scene_update_lights.argtypes = [vp, C.POINTER(F.rt_light), F.c_u32, C.POINTER(F.rt_light), F.c_u32]
        l.orc_scene_read_vertices.argtypes = [vp, C.POINTER(F.rt_vertex), F.c_u32]
        for f in ("orc_trace_closest", "orc_trace_closest_brute"):
            getattr(l, f).argtypes = [vp, C.POINTER(F.rt_ray), F.c_u32, F.c_u32, C.POINTER(F.c_u32), C.POINTER(F.rt_hit)]
        l.orc_trace_any.argtypes = [vp, C.POINTER(F.rt_ray), F.c_u32, F.c_u32, C.POINTER(F.c_u32), F.c_u8p]
        l.orc_render.argtypes = [vp, C.POINTER(F.rt_ubo), F.c_u32, F.c_u32, C.POINTER(F.c_f), F.c_u8p, F.c_u32, F.c_u32,
                                 C.POINTER(F.rt_stats)]
        l.orc_record_bounce_rays.argtypes = [vp, C.POINTER(F.rt_ubo), F.c_u32, F.c_u32, F.c_u32, F.c_u32, C.POINTER(F.rt_ray), C.POINTER(F.c_u32), F.c_u32,
                                             C.POINTER(F.rt_ray), C.POINTER(F.c_u32), F.c_u32, C.POINTER(F.c_u32), C.POINTER(F.c_u32)]
        l.orc_trace_pixel.argtypes = [vp, C.POINTER(F.rt_ubo), F.c_u32, F.c_u32, F.c_u32, F.c_u32, C.POINTER(F.c_f), F.c_u32]
        l.orc_tea.argtypes = [F.c_u32, F.c_u32]
        l.orc_tea.restype = F.c_u32
        l.orc_pcg4d.argtypes = [C.POINTER(F.c_u32), C.POINTER(F.c_u32)]
        l.orc_pcg4d.restype = None
        l.orc_rand.argtypes = [C.POINTER(F.c_u32)]
        l.orc_rand.restype = F.c_f
        l.orc_lcg_float.argtypes = [C.POINTER(F.c_u32)]
        l.orc_lcg_float.restype = F.c_f
        l.orc_offset_ray.argtypes = [C.POINTER(F.c_f)] * 3
        l.orc_offset_ray.restype = None
        l.orc_tonemap.argtypes = [F.c_u32, C.POINTER(F.c_f), C.POINTER(F.c_f)]
        l.orc_tonemap.restype = None
        l.orc_bsdf_sample.argtypes = [C.POINTER(F.c_f), C.POINTER(F.c_f)]
        l.orc_bsdf_sample.restype = None
        l.orc_set_num_threads.argtypes = [C.c_int]
        l.orc_set_num_threads.restype = None
        _lib = l
    return _lib


class OracleError(RuntimeError):
    pass


class OracleScene:
    def __init__(self, desc: F.rt_scene_desc):
        self._l = lib()
        self._h = C.c_void_p()
        if self._l.orc_scene_create(C.byref(desc), C.byref(self._h)):
            raise OracleError(self._l.orc_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.orc_scene_destroy(self._h)
            self._h = None

    def _rng(self, rng4, n):
        if rng4 is None:
            return None, None
        a = np.ascontiguousarray(rng4, np.uint32).reshape(n, 4)
        return a, F.as_ptr(a, F.c_u32)

    def trace_closest(self, rays: np.ndarray, flags: int = 0, rng4=None, brute: bool = False) -> np.ndarray:
        rays = np.ascontiguousarray(rays, F.RAY_DTYPE)
        hits = np.zeros(len(rays), F.HIT_DTYPE)
        keep, rp = self._rng(rng4, len(rays))
        fn = self._l.orc_trace_closest_brute if brute else self._l.orc_trace_closest
        if fn(self._h, F.as_ptr(rays, F.rt_ray), len(rays), flags, rp, F.as_ptr(hits, F.rt_hit)):
            raise OracleError(self._l.orc_last_error().decode())
        return hits

    def trace_any(self, rays: np.ndarray, flags: int = 0, rng4=None) -> np.ndarray:
        rays = np.ascontiguousarray(rays, F.RAY_DTYPE)
        occ = np.zeros(len(rays), np.uint8)
        keep, rp = self._rng(rng4, len(rays))
        if self._l.orc_trace_any(self._h, F.as_ptr(rays, F.rt_ray), len(rays), flags, rp, occ.ctypes.data_as(F.c_u8p)):
            raise OracleError(self._l.orc_last_error().decode())
        return occ

    def render(self, ubo: F.rt_ubo, width: int, height: int, acc: np.ndarray | None = None, rows=None):
        """One frame.  Returns (acc RGBA32F HxWx4, out RGBA8 HxWx4, stats)."""
        if acc is None:
            acc = np.zeros((height, width, 4), np.float32)
        out = np.zeros((height, width, 4), np.uint8)
        st = F.rt_stats()
        r0, r1 = rows if rows else (0, 0)
        if self._l.orc_render(self._h, C.byref(ubo), width, height, F.as_ptr(acc, F.c_f), out.ctypes.data_as(F.c_u8p),
                              r0, r1, C.byref(st)):
            raise OracleError(self._l.orc_last_error().decode())
        return acc, out, st

    def record_bounce_rays(self, ubo: F.rt_ubo, width: int, height: int, stride: int = 1, min_bounce: int = 1, max_rays: int = 1 << 20,
                           max_shadow: int = 1 << 18):
        """SURVEY.md §8d ray set (iv): (rays, rng4, shadow_rays, shadow_rng4) the oracle's own path tracer generates for this frame."""
        rays = np.zeros(max_rays, F.RAY_DTYPE); rng4 = np.zeros((max_rays, 4), np.uint32)
        srays = np.zeros(max_shadow, F.RAY_DTYPE); srng4 = np.zeros((max_shadow, 4), np.uint32)
        n, ns = F.c_u32(), F.c_u32()
        if self._l.orc_record_bounce_rays(self._h, C.byref(ubo), width, height, stride, min_bounce, F.as_ptr(rays, F.rt_ray), F.as_ptr(rng4, F.c_u32), max_rays,
                                          F.as_ptr(srays, F.rt_ray), F.as_ptr(srng4, F.c_u32), max_shadow, C.byref(n), C.byref(ns)):
            raise OracleError(self._l.orc_last_error().decode())
        return rays[:n.value].copy(), rng4[:n.value].copy(), srays[:ns.value].copy(), srng4[:ns.value].copy()

    def trace_pixel(self, ubo: F.rt_ubo, width: int, height: int, x: int, y: int, max_records: int = 16) -> np.ndarray:
        rec = np.zeros((max_records, 16), np.float32)
        n = self._l.orc_trace_pixel(self._h, C.byref(ubo), width, height, x, y, F.as_ptr(rec, F.c_f), max_records)
        return rec[:n]

    def update_instances(self, inst: np.ndarray):
        inst = np.ascontiguousarray(inst, F.INSTANCE_DTYPE)
        if self._l.orc_scene_update_instances(self._h, F.as_ptr(inst, F.rt_instance), len(inst)):
            raise OracleError(self._l.orc_last_error().decode())

    def update_skins(self, mats: np.ndarray):
        mats = np.ascontiguousarray(mats, np.float32)
        n = mats.size // 4096
        if self._l.orc_scene_update_skins(self._h, F.as_ptr(mats, F.c_f), n):
            raise OracleError(self._l.orc_last_error().decode())

    def update_lights(self, dlights: np.ndarray, plights: np.ndarray):
        d = np.ascontiguousarray(dlights, F.LIGHT_DTYPE)
        p = np.ascontiguousarray(plights, F.LIGHT_DTYPE)
        self._l.orc_scene_update_lights(self._h, F.as_ptr(d, F.rt_light), len(d), F.as_ptr(p, F.rt_light), len(p))

    def read_vertices(self, n: int) -> np.ndarray:
        v = np.zeros(n, F.VERTEX_DTYPE)
        if self._l.orc_scene_read_vertices(self._h, F.as_ptr(v, F.rt_vertex), n):
            raise OracleError(self._l.orc_last_error().decode())
        return v
