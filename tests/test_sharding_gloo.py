"""N>1 path on CPU: world_size-2 `gloo` run of the sample-pass sharding logic (frame assignment, private sums,
all-reduce combine, tonemap of the reduced image) with the host-emulation renderer standing in for the GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
from rustracer_b200 import _ffi as F, core, host, sharding

W = H = 40
N_FRAMES = 6


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _frame_ubo(cam, gui, g, opaque):
    import ctypes as C
    u = F.rt_ubo(); total = F.c_u32(g)
    F.load_host().gv_build_ubo(C.byref(cam.c), C.byref(gui.g), C.byref(total), g, int(opaque), 3, C.byref(u))
    return u


def _worker(rank, world, port, out_path):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu")); sys.path.insert(0, os.path.dirname(__file__))
    from emu_lib import emu_api
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    d = util.load_scene_npz(util.GOLDEN / "cornell_box_scene.npz")
    ctx = core.Context(W, H, api=emu_api()); sc = core.Scene(ctx, d)
    cam = host.Camera(W, H).set(position=(0, 0, 14.0)); gui = host.Gui(number_of_samples=1, number_of_bounces=5)
    for g in sharding.frames_of_rank(N_FRAMES, rank, world):
        ctx.render(sc, _frame_ubo(cam, gui, g, d.fully_opaque))
    acc, _ = ctx.readback()
    t = torch.from_numpy(acc.copy())
    dist.all_reduce(t)                                   # the exchange step (NCCL / peer-memory reduce on the GPUs)
    ctx.upload_accumulation(t.numpy())
    ctx.tonemap(_frame_ubo(cam, gui, N_FRAMES - 1, d.fully_opaque))
    acc2, out = ctx.readback()
    r0, r1 = sharding.reduce_rows(rank, world, H)
    rows = torch.zeros((H, W, 4), dtype=torch.uint8); rows[r0:r1] = torch.from_numpy(out[r0:r1].copy())
    dist.all_reduce(rows)                                # all-gather of the RGBA8 row bands
    if rank == 0:
        np.savez(out_path, acc=acc2, out=rows.numpy())
    dist.barrier(); dist.destroy_process_group()


def test_sample_pass_sharding_world2(tmp_path, cornell_desc):
    out_path = str(tmp_path / "w2.npz")
    mp.spawn(_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    z = np.load(out_path)
    # single-process reference: all frames on one (emulated) device
    from emu_lib import emu_api
    ctx = core.Context(W, H, api=emu_api()); sc = core.Scene(ctx, cornell_desc)
    cam = host.Camera(W, H).set(position=(0, 0, 14.0)); gui = host.Gui(number_of_samples=1, number_of_bounces=5)
    for g in range(N_FRAMES):
        ctx.render(sc, _frame_ubo(cam, gui, g, cornell_desc.fully_opaque))
    acc, out = ctx.readback()
    np.testing.assert_allclose(z["acc"], acc, rtol=1e-5, atol=1e-6)     # fp32 sum order differs
    assert np.abs(z["out"].astype(int) - out.astype(int)).max() <= 1


def test_partition_arithmetic():
    for world in (1, 2, 3, 4, 8):
        frames = sorted(g for r in range(world) for g in sharding.frames_of_rank(17, r, world))
        assert frames == list(range(17))
        for r in range(world):
            assert [sharding.global_frame(s, r, world) for s in range(3)] == [r, r + world, r + 2 * world]
        rows = sorted(y for r in range(world) for y in sharding.owned_rows(1080, 8, world, r))
        assert rows == list(range(1080))
        bands = [sharding.reduce_rows(r, world, 1080) for r in range(world)]
        assert bands[0][0] == 0 and bands[-1][1] == 1080 and all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
