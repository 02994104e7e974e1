"""Multi-GPU correctness check (run with torchrun on >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py

Sample-pass sharding: every rank renders its frames into a private sum; rt_combine (fused peer-memory
reduce + tonemap + all-gather of the RGBA8 bands, device-side flags, no host barrier) runs once mid-way (periodic display
refresh, acc untouched) and once at the end; EVERY rank must then hold the complete image, equal (fp32 reassociation aside)
to rank 0 rendering all frames alone.
Tile sharding (§8e A): every rank renders its strips of each frame; after rt_combine (rooted gather incl. the RGBA32F
strips) rank 0 holds an image bit-identical to the single-GPU frame."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from rustracer_b200 import _ffi as F, core, host, scenes, sharding  # noqa: E402
import ctypes as C  # noqa: E402

W, H, K = 640, 360, 8


def ubo_of(cam, gui, g, opaque):
    u = F.rt_ubo(); total = F.c_u32(g)
    F.load_host().gv_build_ubo(C.byref(cam.c), C.byref(gui.g), C.byref(total), g, int(opaque), 3, C.byref(u))
    return u


def open_peers(ctx, rank, world):
    gathered = [None] * world
    dist.all_gather_object(gathered, ctx.ipc_handle())
    return [ctx.ipc_open(hb) for r, hb in enumerate(gathered) if r != rank]


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = scenes.cornell_box(lucy=True, lucy_rows=120, lucy_cols=121)
    ctx = core.Context(W, H, device=local); sc = core.Scene(ctx, d)
    cam = host.Camera(W, H); gui = host.Gui(number_of_samples=1, number_of_bounces=8)
    peers = open_peers(ctx, rank, world)
    n_frames = K * world
    rows = sharding.reduce_rows(rank, world, H)
    ctx.set_frames_in_flight(3)     # the sharded passes overlap on each GPU; rt_combine joins them on the device
    mine = sharding.frames_of_rank(n_frames, rank, world)
    for k, g in enumerate(mine):
        ctx.render(sc, ubo_of(cam, gui, g, True))
        if k == len(mine) // 2 - 1:  # periodic refresh: every rank has rendered K/2 frames
            ctx.combine(peers, ubo_of(cam, gui, (K // 2) * world - 1, True), epoch=1, rows=rows)
    final = ubo_of(cam, gui, n_frames - 1, True)
    ctx.combine(peers, final, epoch=2, rows=rows)              # no host barrier before it: the ranks synchronise on the device
    out, band_sum = ctx.readback_display(want_sum=True)
    acc_private, _ = ctx.readback()
    ok = True
    # every rank holds the complete image: compare the RGBA8 images across ranks bit for bit
    t = torch.from_numpy(out.astype(np.int32)).cuda(); t0 = t.clone(); dist.broadcast(t0, 0)
    same_everywhere = bool((t == t0).all().item())
    flag = torch.tensor([int(same_everywhere)], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        ref = core.Context(W, H, device=local); rsc = core.Scene(ref, d)
        for g in range(n_frames):
            ref.render(rsc, ubo_of(cam, gui, g, True))
        racc, rout = ref.readback()
        r0, r1 = rows
        err = np.abs(band_sum[r0:r1] - racc[r0:r1]).max() / max(1e-6, np.abs(racc[r0:r1]).max())
        lsb = int(np.abs(out.astype(int) - rout.astype(int)).max())
        print(f"sample-pass combine: world {world}: rank-0 band max rel sum err {err:.2e}, full image max RGBA8 diff vs 1 GPU {lsb}, identical on every rank: {bool(flag.item())}")
        ok &= err < 1e-5 and lsb <= 1 and bool(flag.item())
    # tile sharding: strips of each frame, rooted gather to rank 0 (peer 0 of every other rank is rank 0)
    dist.barrier()
    ctx.set_frames_in_flight(1)
    ctx.resize(W, H)
    ubos = [ubo_of(cam, gui, g, True) for g in range(3)]
    for u in ubos:
        ctx.render(sc, u, strip_rows=8, n_parts=world, part=rank)
    ctx.combine(peers, ubos[-1], epoch=3, tiles=(8, world, rank), gather_to=(F.RT_GATHER_NONE if rank == 0 else 0),
                n_senders=(world - 1 if rank == 0 else 0), gather_acc=True)
    out_t, acc_t = ctx.readback_display(want_sum=True)
    dist.barrier()
    if rank == 0:
        ref = core.Context(W, H, device=local); rsc = core.Scene(ref, d)
        for u in ubos:
            ref.render(rsc, u)
        full_acc, full = ref.readback()
        same = bool((out_t == full).all()) and bool((acc_t == full_acc).all())
        print(f"tile partition x{world}: image and accumulation gathered on rank 0 bit-identical to the 1-GPU frames: {same}")
        ok &= same
        print("MULTIGPU_CHECK", "PASS" if ok else "FAIL")
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
