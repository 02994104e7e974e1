"""Headless counterpart of the reference's `gltf_viewer` example (crates/examples/gltf_viewer/src/main.rs).

    python -m rustracer_b200.gltf_viewer -f scene.gltf [-o out.png] [--width 1920 --height 1080] [--spp 64]
                                         [--bounces 5] [--skybox DIR] [--mapping 0] [--animate T]

`-f/--file` mirrors the reference CLI (args.rs:4-10).  The window, swapchain and imgui panel are out of scope; GUI
defaults (gui_state.rs:303-332) are used for everything not given on the command line.  Needs a CUDA device.
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np

from . import core, host

def main(argv=None):
    ap = argparse.ArgumentParser(prog="gltf_viewer")
    ap.add_argument("-f", "--file", required=True, help="Path of the glTF file")
    ap.add_argument("-o", "--output", default="render.png")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=64, help="total samples per pixel (accumulated)")
    ap.add_argument("--samples-per-frame", type=int, default=3)
    ap.add_argument("--bounces", type=int, default=5)
    ap.add_argument("--skybox", default=None)
    ap.add_argument("--mapping", type=int, default=0)
    ap.add_argument("--tone-map", type=int, default=0)
    ap.add_argument("--animate", type=float, default=None, help="evaluate the glTF animation at time T (seconds)")
    ap.add_argument("--camera", type=float, nargs=3, default=None, metavar=("X", "Y", "Z"))
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)

    doc = host.load_file(a.file)
    sky = False
    if a.skybox:
        doc.set_skybox(host.load_skybox_dir(a.skybox))     # SkyBox::new, cubumap.rs:86-106
        sky = True
    desc = doc.scene_desc()
    ctx = core.Context(a.width, a.height, device=a.device)
    scene = core.Scene(ctx, desc)
    if a.animate is not None:                      # GltfViewer::state_change animation branch (main.rs:377-410)
        doc.animate(a.animate)
        if doc.need_compute():
            scene.update_skins(doc.get_skins())
        scene.update_instances(doc.get_instances())
    cam = host.Camera(a.width, a.height)
    if a.camera:
        cam.set(position=a.camera)
    gui = host.Gui(number_of_samples=a.samples_per_frame, number_of_bounces=a.bounces, max_number_of_samples=a.spp, sky=int(sky),
                   mapping=a.mapping, selected_tone_map_mode=a.tone_map)
    drv = host.FrameDriver(cam, gui, doc.fully_opaque())
    frames = 0
    while True:
        ubo = drv.next_ubo()
        if ubo.number_of_samples == 0:
            break
        ctx.render(scene, ubo)
        frames += 1
        if a.mapping != 0:
            break
    _, out = ctx.readback(want_acc=False)
    from PIL import Image
    Image.fromarray(out[..., :3]).save(a.output)
    st = ctx.stats()
    print(f"{a.output}: {a.width}x{a.height}, {drv.total.value} spp in {frames} frames, last frame {st.ms_total:.2f} ms")


if __name__ == "__main__":
    sys.exit(main())
