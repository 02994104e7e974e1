"""TEST INFRASTRUCTURE — loads the host-emulation build of the CUDA core (tests/emu/emu.cpp).

Used only by `-m "not gpu"` tests to check builder / traversal / shading logic on the GPU-less build box.
"""
import ctypes as C
import subprocess
from pathlib import Path

from rustracer_b200 import _ffi as F
from rustracer_b200.core import Api

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "librt_emu.so"
IPC = F.RT_CUDA_ONLY      # multi-GPU entry points exist in the CUDA library only
_api = None


def emu_api() -> Api:
    global _api
    if _api is None:
        subprocess.check_call(["make", "-C", str(HERE)], stdout=subprocess.DEVNULL)
        lib = C.CDLL(str(LIB))
        rename = lambda n: "emu_" + n  # noqa: E731
        F.bind_rt(lib, rename, optional=IPC)
        _api = Api(lib, rename)
    return _api
