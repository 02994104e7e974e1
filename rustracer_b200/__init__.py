"""rustracer_b200 — B200-native path-tracing core behind the reference's render-path surface.

Layers:
  csrc/        hand-written CUDA for sm_100a + the C ABI (include/rt_b200.h)     -> librt_b200.so
  host.py      ctypes view of the C++ host layer (host/gltf_host.cpp): glTF import with the reference's
               rules, animation, camera, per-frame UBO
  core.py      ctypes view of the C ABI: Context / Scene / render / readback / trace
  scenes.py    synthetic scene generators for BASELINE.json's configs
There is no CPU fallback: importing `core` objects without the built CUDA library raises.
"""
from . import _ffi  # noqa: F401

__all__ = ["_ffi"]
__version__ = "0.1.0"
