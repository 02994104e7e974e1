"""Generates tests/golden/shadows_glb_scene.npz from the reference's assets/models/shadows.glb (CesiumMan: 1 skin x 19
joints, 57 LINEAR TRS channels over 0..2 s, 1024^2 PNG base-colour texture, spot light on a bone).
Run HERE (build container, where /root/reference exists):

    python tests/golden/make_shadows_fixture.py

The fixture holds what the host loader (host/gltf_host.cpp, the reference's asset_loader rules) produces for the file
— every rt_scene_desc array, the texture in its original PNG encoding — plus the loader-driven animation: for 60 frames
at 1/30 s the skin matrices (Doc::animate -> get_skins, animation.rs:77-146, skinning.rs:39-50) and the instance list.
The GPU box has no /root/reference: the -m gpu tests drive rt_scene_update_skins from this file."""
import json
import struct
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from rustracer_b200 import host  # noqa: E402
import util  # noqa: E402

ASSET = Path("/root/reference/assets/models/shadows.glb")
OUT = Path(__file__).resolve().parent / "shadows_glb_scene.npz"
FRAMES, DT = 60, 1.0 / 30.0


def glb_images(path):
    """encoded bytes of the GLB's embedded images, by glTF image index"""
    raw = path.read_bytes()
    assert raw[:4] == b"glTF"
    jlen = struct.unpack_from("<I", raw, 12)[0]
    js = json.loads(raw[20:20 + jlen])
    boff = 20 + jlen + 8
    out = {}
    for k, im in enumerate(js.get("images", [])):
        bv = js["bufferViews"][im["bufferView"]]
        out[k] = raw[boff + bv.get("byteOffset", 0): boff + bv.get("byteOffset", 0) + bv["byteLength"]]
    return out


def main():
    doc = host.load_file(ASSET)
    d = doc.scene_desc()
    files = {k + 1: b for k, b in glb_images(ASSET).items()}      # loader image 0 is the 1x1 dummy (image.rs:31-43)
    for k, b in files.items():                                     # the stored encoding must decode to what the loader produced
        im = d.images[k]
        import ctypes as C
        assert (host.decode_image(b) == np.frombuffer(C.string_at(im.rgba8, im.width * im.height * 4), np.uint8).reshape(im.height, im.width, 4)).all()
    skins, insts, times = [], [], []
    sk0 = doc.get_skins()
    used = int(np.nonzero(np.abs(sk0).sum(axis=(0, 2)))[0].max()) + 1
    for f in range(FRAMES):
        t = f * DT
        doc.animate(t)
        skins.append(doc.get_skins()[:, :used].copy()); insts.append(doc.get_instances().copy()); times.append(t)
    doc2 = host.load_file(ASSET)                                    # un-animated copy for the static arrays
    util.save_scene_full(doc2.scene_desc(), OUT, image_files=files, anim_times=np.array(times, np.float32),
                         anim_skins=np.stack(skins), anim_instances=np.stack(insts))
    d3, z = util.load_scene_full(OUT)
    print(OUT, OUT.stat().st_size, "bytes;", d3.n_vertices, "vertices,", d3.n_indices // 3, "triangles,", d3.n_skins, "skin,", used, "joints,", FRAMES, "animation frames")


if __name__ == "__main__":
    main()
