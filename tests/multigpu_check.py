"""Multi-GPU correctness check (run with torchrun on >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_check.py

Sample-pass sharding: every rank renders its frames, the fused peer-memory reduce + tonemap (rt_reduce_peers) combines
them; rank 0 then renders ALL frames alone and compares its row band (fp32 sum order differs -> tolerance).
Tile sharding: every rank renders its strips of one frame into its own image; the union must be bit-identical to the
full frame rendered by rank 0."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from rustracer_b200 import _ffi as F, core, host, scenes, sharding  # noqa: E402

W, H, K = 640, 360, 8


def ubo_of(cam, gui, g, opaque):
    u = F.rt_ubo(); total = F.c_u32(g)
    F.load_host().gv_build_ubo(C.byref(cam.c), C.byref(gui.g), C.byref(total), g, int(opaque), 3, C.byref(u))
    return u


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = scenes.cornell_box(lucy=True, lucy_rows=120, lucy_cols=121)
    ctx = core.Context(W, H, device=local); sc = core.Scene(ctx, d)
    cam = host.Camera(W, H).set(position=(0, 0, 14.0)); gui = host.Gui(number_of_samples=1, number_of_bounces=8)
    n_frames = K * world
    ctx.set_frames_in_flight(3)     # the sharded passes overlap on each GPU; rt_reduce_peers joins them
    for g in sharding.frames_of_rank(n_frames, rank, world):
        ctx.render(sc, ubo_of(cam, gui, g, True))
    ctx.synchronize()
    handle = (C.c_uint8 * 64)(); ctx.api.check(ctx.api.rt_ipc_export(ctx._h, handle))
    gathered = [None] * world; dist.all_gather_object(gathered, bytes(handle))
    peers = []
    for r, hb in enumerate(gathered):
        if r != rank:
            p = C.c_void_p(); ctx.api.check(ctx.api.rt_ipc_open(ctx._h, (C.c_uint8 * 64).from_buffer_copy(hb), C.byref(p))); peers.append(p.value)
    dist.barrier()
    r0, r1 = sharding.reduce_rows(rank, world, H)
    final = ubo_of(cam, gui, n_frames - 1, True)
    ctx.api.check(ctx.api.rt_reduce_peers(ctx._h, (C.c_void_p * len(peers))(*peers), len(peers), C.byref(final), r0, r1, None))
    acc, out = ctx.readback()
    dist.barrier()
    ok = True
    if rank == 0:
        ref = core.Context(W, H, device=local)
        rsc = core.Scene(ref, d)
        for g in range(n_frames):
            ref.render(rsc, ubo_of(cam, gui, g, True))
        racc, rout = ref.readback()
        err = np.abs(acc[r0:r1] - racc[r0:r1]).max() / max(1e-6, np.abs(racc[r0:r1]).max())
        lsb = np.abs(out[r0:r1].astype(int) - rout[r0:r1].astype(int)).max()
        print(f"sample-pass reduce: world {world}, rows {r0}..{r1}: max rel acc err {err:.2e}, max RGBA8 diff {lsb}")
        ok &= err < 1e-5 and lsb <= 1
    # tile sharding
    ctx.resize(W, H)
    u0 = ubo_of(cam, gui, 0, True)
    ctx.render(sc, u0, strip_rows=8, n_parts=world, part=rank)
    acc_t, out_t = ctx.readback()
    t = torch.from_numpy(out_t.astype(np.int32)).cuda()
    dist.all_reduce(t)                                   # rows are disjoint: the sum is the assembled image
    if rank == 0:
        ref = core.Context(W, H, device=local); rsc = core.Scene(ref, d)
        ref.render(rsc, u0); _, full = ref.readback()
        same = bool((t.cpu().numpy() == full.astype(np.int32)).all())
        print(f"tile partition x{world}: assembled image bit-identical to the 1-GPU frame: {same}")
        ok &= same
        print("MULTIGPU_CHECK", "PASS" if ok else "FAIL")
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
