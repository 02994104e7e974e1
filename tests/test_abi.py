"""The C-ABI library loads and exports every symbol include/*.h declares (no compute without a GPU)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from rustracer_b200 import _ffi as F

ROOT = Path(__file__).resolve().parents[1]


def declared(header: str, prefix: str):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"_[a-z0-9_]+)\s*\(", text)))


def test_rt_exports_match_header():
    names = declared("rt_b200.h", "rt")
    assert set(names) == set(F.RT_EXPORTS), set(names) ^ set(F.RT_EXPORTS)
    lib = F.load_rt()
    for n in names:
        assert hasattr(lib, n), n
    assert b"sm_100a" in lib.rt_version()


def test_host_exports_match_header():
    names = declared("gltf_host.h", "gv")
    assert set(names) == set(F.GV_EXPORTS), set(names) ^ set(F.GV_EXPORTS)
    lib = F.load_host()
    for n in names:
        assert hasattr(lib, n), n


def test_struct_layouts():
    # SURVEY.md §8a row a12-T
    assert C.sizeof(F.rt_vertex) == 128 and F.rt_vertex.skin_index.offset == 112 and F.rt_vertex.uv0.offset == 96
    m = F.rt_material
    assert C.sizeof(m) == 256
    offs = dict(alpha_mode=0, alpha_cutoff=4, workflow=12, base_color_texture=24, base_color=32, metallic_factor=48, roughness_factor=52,
                metallic_roughness_texture=56, normal_texture=64, emissive_texture=72, emissive_factor=80, occlusion_texture=96, ior=104,
                unlit=108, transmission_texture=112, transmission_factor=120, transmission_exist=124, attenuation_color=128,
                thickness_factor=140, thickness_texture=144, attenuation_distance=152, volume_exists=156, specular_texture=160,
                specular_color_texture=168, specular_color_factor=176, specular_factor=192, specular_exist=196, sg_diffuse_factor=208,
                sg_specular_glossiness_factor=224, sg_diffuse_texture=240, sg_specular_glossiness_texture=248)
    for k, v in offs.items():
        assert getattr(m, k).offset == v, k
    assert C.sizeof(F.rt_ubo) == 324 and F.rt_ubo.aperture.offset == 256 and F.rt_ubo.tone_mapping_mode.offset == 320
    assert C.sizeof(F.rt_light) == 48 and C.sizeof(F.rt_prim_info) == 16 and C.sizeof(F.rt_instance) == 64


def test_no_device_means_error_not_fallback():
    """Without a CUDA device context creation must fail loudly (the product has no CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = F.load_rt()
    h = C.c_void_p()
    assert lib.rt_context_create(0, 64, 64, C.byref(h)) != 0
    assert b"no CUDA device" in lib.rt_last_error() or b"CUDA" in lib.rt_last_error()


def test_product_never_imports_oracle_or_emu():
    """The product package must not import, link or execute the oracle or the host-emulation build."""
    for p in (ROOT / "rustracer_b200").rglob("*.py"):
        text = p.read_text()
        for needle in ("import orc", "from oracle", "oracle.orc", "emu_lib", "librt_emu", "liborc"):
            assert needle not in text, (p, needle)
    for p in list((ROOT / "rustracer_b200" / "csrc").glob("*.h")) + list((ROOT / "rustracer_b200" / "csrc").glob("*.cu")):
        text = p.read_text()
        assert "#include \"../../oracle" not in text and "oracle/" not in text.replace("oracle/oracle.cpp", ""), p
    import subprocess
    out = subprocess.run(["ldd", str(F.RT_LIB)], capture_output=True, text=True).stdout
    assert "liborc" not in out and "librt_emu" not in out
