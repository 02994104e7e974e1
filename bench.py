#!/usr/bin/env python
"""bench.py — headline benchmark of the path-tracing hot path (BASELINE.json: Mrays/s at 1080p depth 8).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port on the host cores
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU (sample-pass sharding, §8e B)

A step is one frame: 1 spp per pixel at max depth 8 over the Lucy-in-Cornell stand-in scene (config 2: Cornell shell
from the bundled cornellBox asset (tests/golden/cornell_box_scene.npz), the real Lucy scan is not shipped),
seen from the reference's default camera (app/src/lib.rs:331-338: position (0,0,1) looking down -z, fov 60), accumulated
like the reference does (RayTracing.rgen:132-166).
Mrays/s counts every traced segment (path segments + shadow rays), SURVEY.md §8d.
Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

WIDTH, HEIGHT, BOUNCES, SPP = 1920, 1080, 8, 1
CONFIG = 2          # BASELINE.json configs[1] is the headline workload; --config selects the others (extra, not the driver's line)
POSE = None         # config 4: per-frame skin matrices
GUI_KW = {}
CAM_POS = (0, 0, 1.0)    # the reference's default camera (app/src/lib.rs:331-338): inside the box, looking down -z
CAM_ALT = (0, 0, 14.0)   # round-1 headline camera (outside the box: 2.4 rays/pixel, 30 % misses); kept as a second, labelled figure
SHELL = ROOT / "tests" / "golden" / "cornell_box_scene.npz"   # the real cornellBox.gltf shell, exported by tests/golden/make_golden.py
NAMES = {1: "procedural Cornell box (configs[0]) 512x512, 1 spp/frame, max depth 8",
         2: "Lucy-in-Cornell stand-in (BASELINE configs[1]) 1920x1080, 1 spp/frame accumulated, max depth 8",
         3: "10k instances of a 100k-triangle BLAS + alpha-MASK foliage cards (configs[2]) 1920x1080, sky + directional light, max depth 8",
         4: "skinned ~1M-triangle character, 256 joints (configs[3]): per frame skinning + BLAS refit + TLAS + 1 spp render 1920x1080",
         5: "64 glass/volume objects in a Cornell-type box (configs[4]) 3840x2160, 1 spp/frame, max depth 8"}


def select_config(c: int):
    global WIDTH, HEIGHT, CONFIG, GUI_KW, CAM_POS
    CONFIG = c
    if c == 1:
        WIDTH, HEIGHT = 512, 512
    elif c == 2:
        pass
    elif c == 3:
        GUI_KW = {"sky": 1}; CAM_POS = (0, 1.2, 7.0)
    elif c == 4:
        GUI_KW = {"animation": 1}; CAM_POS = (0, 0, 9.0)
    elif c == 5:
        WIDTH, HEIGHT = 3840, 2160


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled in-process every 5 ms (the same
    counters `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*` prints; a 100 ms nvidia-smi loop
    would miss a sub-100 ms timed region), nvidia-smi subprocess as the fallback."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device_index: int):
        self.dev, self.proc, self.lines = device_index, None, []
        self.nvml, self.handle, self.samples, self.stop_flag, self.source = None, None, [], False, None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        self.nvml = pynvml
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.dev).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            return pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            return pynvml.nvmlDeviceGetHandleByIndex(self.dev)

    def _poll(self):
        n = self.nvml
        bits = {"hw_slowdown": n.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksEventReasonSwPowerCap}
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((sm, {k for k, b in bits.items() if r & b}))
            except Exception:
                break
            time.sleep(0.005)

    def prepare(self):
        """NVML initialisation takes milliseconds and varies from process to process: do it before the barrier that
        precedes the timed region, so that the ranks enter the region together."""
        try:
            self.handle = self._nvml_handle()
            self.max_sm = self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)
            self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None
        return self

    def start(self):
        if self.nvml is None and self.handle is None:
            self.prepare()
        if self.nvml is not None:
            try:
                self.t = threading.Thread(target=self._poll, daemon=True)
                self.t.start()
                self.source = "nvml"
                return
            except Exception:
                self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.dev)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
            self.source = "nvidia-smi"
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            sm = [s for s, _ in self.samples]
            reasons = set().union(*[r for _, r in self.samples]) if self.samples else set()
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_sm, "reasons": sorted(reasons), "samples": len(sm), "source": "nvml 5 ms poll"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def build_scene_desc():
    global POSE
    from rustracer_b200 import scenes
    shell = SHELL if SHELL.exists() else None
    if CONFIG == 1:
        return scenes.cornell_box(lucy=False, shell=shell)
    if CONFIG == 3:
        return scenes.instanced_foliage(n_side=100, tris_per_mesh=100_000, cards=64, tex_size=1024, sky=scenes.procedural_sky(256))
    if CONFIG == 4:
        d, POSE = scenes.skinned_character(n_tris=1_000_000, joints=256)
        return d
    if CONFIG == 5:
        return scenes.glass_box(n_objects=64)
    return scenes.cornell_box(lucy=True, shell=shell)


def make_gui():
    from rustracer_b200 import host
    # max_number_of_samples: Gui::new's 5000 would silently end the sample budget (number_of_samples -> 0) in long or
    # many-GPU runs; the bench accumulates without a budget
    return host.Gui(number_of_samples=SPP, number_of_bounces=BOUNCES, max_number_of_samples=0x7FFFFFFF, **GUI_KW)


def frame_ubo(cam, gui, frame_index: int, fully_opaque: bool):
    """UBO of global frame `frame_index` (0-based): 1 spp, total = frame_index + 1 (GltfViewer::update, main.rs:207-237)."""
    from rustracer_b200 import _ffi as F
    u = F.rt_ubo()
    total = F.c_u32(frame_index)
    F.load_host().gv_build_ubo(C.byref(cam.c), C.byref(gui.g), C.byref(total), frame_index, int(fully_opaque), 3, C.byref(u))
    assert u.number_of_samples == SPP and u.total_number_of_samples == frame_index + 1 or gui.g.animation, "sample budget exhausted"
    return u


COUNTER_KEYS = ("rays_extend", "rays_shadow", "shaded_hits", "pixel_samples", "nodes", "tris", "insts", "anyhits", "tex_taps", "light_cands")


def algorithmic_bytes(st, spp=SPP):
    """SURVEY.md §8d formula, every count taken from in-kernel counters of the same frames.  Returns (traversal bytes of
    extend + shadow rays, whole-frame bytes).  shaded_hits are closest hits only (misses run the miss stage: no vertex /
    material gather); texture taps of shading and of the any-hit stage share one counter (16 B per bilinear tap)."""
    rays = st["rays_extend"] + st["rays_shadow"]
    trav = 48 * rays + 80 * st["nodes"] + 48 * st["tris"] + 64 * st["insts"] + (12 + 96 + 32) * st["anyhits"]
    shade = (12 + 384 + 16 + 256 + 48 + 128) * st["shaded_hits"] + 16 * st["tex_taps"] + 48 * st["light_cands"]
    pix = 36 * st["pixel_samples"] / spp
    return trav, trav + shade + pix


def stats_dict(st):
    return {k: int(getattr(st, k)) for k in COUNTER_KEYS}


# ----------------------------------------------------------------------------------------------------
# CPU arm (oracle port)
# ----------------------------------------------------------------------------------------------------
def cpu_sample(desc, frames: int, rows, warm: int = 0):
    """Renders `frames` frames of rows [rows) at 1080p with the oracle on all host threads.  Returns (Mrays/s, rays, seconds)."""
    from oracle import orc
    from rustracer_b200 import host
    s = orc.OracleScene(desc)
    cam = host.Camera(WIDTH, HEIGHT).set(position=CAM_POS)
    gui = make_gui()
    acc = np.zeros((HEIGHT, WIDTH, 4), np.float32)
    rays, secs = 0, 0.0
    for f in range(warm + frames):
        u = frame_ubo(cam, gui, f, desc.fully_opaque)
        t0 = time.perf_counter()
        acc, out, st = s.render(u, WIDTH, HEIGHT, acc, rows=rows)
        dt = time.perf_counter() - t0
        if f >= warm:
            rays += st.rays_extend + st.rays_shadow; secs += dt
    return rays / secs / 1e6, rays, secs


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import orc
    desc = build_scene_desc()
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the CPU arm uses every host core regardless
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    orc.lib().orc_set_num_threads(int(ncpu))
    cores = orc.lib().orc_num_threads() if hasattr(orc.lib(), "orc_num_threads") else ncpu
    # calibrate a row band so that one step takes ~1 s
    m, rays, secs = cpu_sample(desc, 1, (HEIGHT // 2 - 20, HEIGHT // 2 + 20))
    band = int(max(8, min(HEIGHT, 40 * (1.0 / max(secs, 1e-3)))))
    r0 = (HEIGHT - band) // 2
    from rustracer_b200 import host
    s = orc.OracleScene(desc)
    cam = host.Camera(WIDTH, HEIGHT).set(position=CAM_POS); gui = make_gui()
    acc = np.zeros((HEIGHT, WIDTH, 4), np.float32)
    total_rays, t_total, samples = 0, 0.0, 0
    for f in range(args.warmup + args.steps):
        u = frame_ubo(cam, gui, f, desc.fully_opaque)
        t0 = time.perf_counter(); acc, out, st = s.render(u, WIDTH, HEIGHT, acc, rows=(r0, r0 + band)); dt = time.perf_counter() - t0
        if f >= args.warmup:
            total_rays += st.rays_extend + st.rays_shadow; t_total += dt; samples += st.pixel_samples
    v = total_rays / t_total / 1e6
    sample = f"rows {r0}..{r0 + band} of each {WIDTH}x{HEIGHT} frame, {args.steps} frames x 1 spp, depth 8, all host threads"
    line = {"impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(desc),
            "note": "reference cannot be built/run here (Rust+Vulkan RT, no toolchain): CPU oracle port stands in (BASELINE.md §3)",
            "samples_per_s": samples / t_total,
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": int(cores), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(desc):
    return {"workload": NAMES[CONFIG],
            "camera": {"position": list(CAM_POS), "direction": [0, 0, -1], "fov_deg": 60,
                       "note": "reference default camera (app/src/lib.rs:331-338)" if CAM_POS == (0, 0, 1.0) else "scene-specific camera"},
            "shell": "cornellBox.gltf asset (tests/golden/cornell_box_scene.npz)" if (CONFIG in (1, 2) and SHELL.exists()) else "procedural",
            "triangles": int(desc.n_indices // 3), "instances": int(desc.n_instances), "width": WIDTH, "height": HEIGHT, "spp_per_step": SPP, "max_depth": BOUNCES,
            "l2": "per-frame path-state/hit streams (~480 MB at 1080p) exceed the 126 MB L2; the ~28 MB BVH is L2-resident by design (SURVEY.md App. G)"}


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from rustracer_b200 import _ffi as F, core, host
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # a dedicated (non-default) torch stream: its handle is what the C ABI launches on, so torch.cuda.Event timing
    # brackets exactly our kernels (handle 0 would make rt_render fall back to the context's private stream)
    tstream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    desc = build_scene_desc()
    ctx = core.Context(WIDTH, HEIGHT, device=local)
    t0 = time.perf_counter(); scene = core.Scene(ctx, desc); build_s = time.perf_counter() - t0
    first_build_ms = float(scene.bvh_info().build_ms)
    # the first build of a process also pays one-time costs (module loading, the first large allocations of a fresh GPU):
    # 3 ms to > 100 ms from run to run.  The scene is built a second time and that (warm) build is the reported figure.
    scene.close()
    t0 = time.perf_counter(); scene = core.Scene(ctx, desc); build_s = time.perf_counter() - t0
    info = scene.bvh_info()
    cam = host.Camera(WIDTH, HEIGHT).set(position=CAM_POS)
    gui = make_gui()
    K, Wm = args.steps, args.warmup
    # sample-pass sharding (§8e B): global frame g = step * world + rank; every rank starts from a zero accumulation
    from rustracer_b200 import sharding
    tiles = args.partition == "tiles" and world > 1
    if tiles:      # strong scaling (§8e A): every rank renders its interleaved 8-row strips of the SAME global frame s
        ubos = [frame_ubo(cam, gui, s, desc.fully_opaque) for s in range(Wm + K)]
        final_ubo = ubos[-1]
    else:          # weak scaling (§8e B): rank r renders whole global frames s * world + r into a private sum
        ubos = [frame_ubo(cam, gui, sharding.global_frame(s, rank, world), desc.fully_opaque) for s in range(Wm + K)]
        final_ubo = frame_ubo(cam, gui, (Wm + K) * world - 1, desc.fully_opaque)
    combine_every = args.combine_every if args.combine_every >= 0 else (64 if CONFIG == 5 else 0)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    poses = [np.ascontiguousarray(POSE(s), np.float32) for s in range(Wm + K)] if POSE is not None else None   # host animation evaluated up front

    def pre_step(s):
        if poses is not None:     # config 4: skin upload + skinning kernel + BLAS refit + TLAS refit belong to the step
            scene.update_skins(poses[s])

    epoch = [0]
    side = torch.cuda.Stream(device=local)      # periodic display refreshes run beside the frames (§8d config 5)

    def combine(ubo, on_side=False):
        """the exchange step, device-synchronised (rt_combine): no host-side wait, no barrier"""
        if world == 1:
            return
        epoch[0] += 1
        if tiles:      # rooted gather of the strips' RGBA8 (the frame is presented on rank 0)
            ctx.combine(peers, ubo, epoch[0], tiles=(8, world, rank), gather_to=(F.RT_GATHER_NONE if rank == 0 else 0),
                        n_senders=(world - 1 if rank == 0 else 0), stream=stream)
        else:          # fused reduce + tonemap of this rank's band + all-gather: every rank ends up with the complete image
            ctx.combine(peers, ubo, epoch[0], rows=(row0, row1), stream=(side.cuda_stream if on_side else stream))

    def render_step(s, flags=0, exchange=True):
        pre_step(s)
        if tiles:
            ctx.render(scene, ubos[s], flags=flags, strip_rows=8, n_parts=world, part=rank, stream=stream)
            if exchange:
                combine(ubos[s])
        else:
            ctx.render(scene, ubos[s], flags=flags, stream=stream)
            if exchange and combine_every and (s + 1) % combine_every == 0 and s + 1 < Wm + K:
                combine(frame_ubo(cam, gui, (s + 1) * world - 1, desc.fully_opaque), on_side=True)

    def frames(lo, hi, flags=0):
        for s in range(lo, hi):
            render_step(s, flags, exchange=False)

    # ---- device-timed run: inputs resident, K frames (+ final cross-GPU reduce) ----
    # frames in flight (the reference's InFlightFrames, app/src/lib.rs:34): the path tracing of NF consecutive frames
    # overlaps on the GPU, the accumulation stays in submission order (bit-identical images, tests/parity_cases.py)
    NF = max(1, min(8, args.frames_in_flight))
    ctx.set_frames_in_flight(NF)
    if POSE is not None:
        scene.set_versions(max(2, min(4, args.scene_versions)))   # skin updates write the next copy while frames read the previous ones
    ctx.resize(WIDTH, HEIGHT)
    # peers map the accumulation image only now, after the last (re)allocation: rt_frame_resize with unchanged dimensions
    # clears in place and a size change is refused while a handle is exported
    peers = []
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, ctx.ipc_handle())
        peers = [ctx.ipc_open(hb) for r, hb in enumerate(gathered) if r != rank]     # accumulation blocks of the other ranks, in rank order
    row0, row1 = sharding.reduce_rows(rank, world, HEIGHT)

    frames(0, Wm)
    ctx.synchronize()
    sampler = ClockSampler(local).prepare()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    e0.record()
    t_loop0 = time.perf_counter()
    for s in range(Wm, Wm + K):
        render_step(s)
    if not tiles:
        combine(final_ubo)                     # all ranks' sums -> the complete tonemapped image on every rank; synchronised on the device
    t_enq = time.perf_counter() - t_loop0
    ctx.join(stream)                           # the timing stream waits (on the device) for every frame in flight and for the combine
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if world > 1 and os.environ.get("BENCH_DEBUG"):
        print(f"[rank {rank}] timed region {ms:.3f} ms; host enqueue {1e3 * t_enq:.3f} ms; clocks {clocks}", file=sys.stderr, flush=True)
    tmax = torch.tensor([ms], device="cuda")
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_max = float(tmax.item())

    # ---- counted pass (outside the timed region): exact rays / nodes / triangles for the same K frames ----
    ctx.set_frames_in_flight(1)                # per-frame statistics and per-kernel event timing: one frame at a time
    ctx.resize(WIDTH, HEIGHT)
    frames(0, Wm)
    tot = dict.fromkeys(COUNTER_KEYS, 0)
    ext_ms, launches, n_ext = 0.0, 0, 0
    for s in range(Wm, Wm + K):
        render_step(s, flags=1 | 4, exchange=False)
        st = ctx.stats()
        for k, v in stats_dict(st).items():
            tot[k] += v
        launches += st.n_kernel_launches
    # per-stage times of an uncounted frame sequence (counters slow the kernels down)
    ctx.resize(WIDTH, HEIGHT)
    frames(0, Wm)
    stage = dict(raygen=0.0, extend=0.0, shade=0.0, shadow=0.0, accum=0.0)
    upd = dict(skin_ms=0.0, refit_ms=0.0, tlas_ms=0.0)
    for s in range(Wm, Wm + K):
        render_step(s, flags=4, exchange=False)
        st = ctx.stats()
        if poses is not None:      # update stages event-timed one frame at a time (nothing else runs on the GPU)
            bi = scene.bvh_info()
            upd["skin_ms"] += bi.skin_ms / K; upd["refit_ms"] += bi.refit_ms / K; upd["tlas_ms"] += bi.tlas_ms / K
        stage["raygen"] += st.ms_raygen; stage["extend"] += st.ms_extend; stage["shade"] += st.ms_shade; stage["shadow"] += st.ms_shadow; stage["accum"] += st.ms_accum
        n_ext += st.n_extend_launches
    rays_rank = tot["rays_extend"] + tot["rays_shadow"]
    t_rays = torch.tensor([float(rays_rank), float(tot["pixel_samples"])], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t_rays)
    rays_all, samples_all = float(t_rays[0].item()), float(t_rays[1].item())
    value = rays_all / (ms_max * 1e-3) / 1e6

    # ---- end-to-end through the C ABI with host buffers: UBO from host each frame, RGBA8 image read back each frame ----
    # Like the reference's draw loop (app/src/lib.rs:400-401) the host submits frame s (UBO by value + the frame's image
    # copy to pinned host memory queued behind it) and waits on the fence of an earlier frame before it reuses that
    # frame's host buffer; with 2*NF host buffers the device never runs out of submitted frames.
    ctx.set_frames_in_flight(NF)
    HB = 2 * NF     # host staging buffers: the host runs up to HB frames ahead, the device keeps NF in flight
    pinned = [torch.empty((HEIGHT, WIDTH, 4), dtype=torch.uint8, pin_memory=True) for _ in range(HB)]
    out_np = [p.numpy() for p in pinned]
    ctx.resize(WIDTH, HEIGHT)
    for s in range(0, Wm):
        render_step(s, exchange=False); ctx.frame_wait(ctx.readback_async(out_np[s % HB]))
    ctx.synchronize()
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    tickets = []
    for s in range(Wm, Wm + K):
        if len(tickets) >= HB:
            ctx.frame_wait(tickets[-HB])                   # the image of frame s-HB is in host memory; its buffer is reused now
        render_step(s, exchange=False)                     # config 4: 256 mat4 (16 KB) host -> device per step
        tickets.append(ctx.readback_async(out_np[s % HB]))  # device -> pinned host, 4 B/pixel, behind the frame
    for t in tickets[-HB:]:
        ctx.frame_wait(t)                                  # every step's result has reached the host inside the timed region
    ctx.join(stream)
    e3.record(); torch.cuda.synchronize()
    e2e_ms = max(e2.elapsed_time(e3), (time.perf_counter() - t0) * 1e3)
    ctx.set_frames_in_flight(1)
    t_e2e = torch.tensor([e2e_ms], device="cuda")
    if dist is not None:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = rays_all / (float(t_e2e.item()) * 1e-3) / 1e6

    # ---- the exchange step alone, two ways (outside the headline region): the fused peer-memory kernel vs NCCL ----
    exchange = None
    if world > 1 and not tiles:
        ctx.set_frames_in_flight(NF); ctx.synchronize()
        reps = 10

        class _Cai:     # torch view of the RGBA32F accumulation image (zero copy) for the NCCL arm
            def __init__(self, ptr, shape):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": "<f4", "data": (ptr, False), "version": 2}
        acc_ptr, _ = ctx.device_ptrs()
        acc_t = torch.as_tensor(_Cai(acc_ptr, (HEIGHT, WIDTH, 4)), device=f"cuda:{local}")
        scratch = acc_t.clone()
        for arm in ("fused", "nccl"):
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(reps):
                if arm == "fused":
                    combine(final_ubo)
                else:       # "NCCL accumulation reduce" (BASELINE.json): all-reduce of the sums, then every rank tonemaps the whole image
                    dist.all_reduce(scratch)
                    ctx.tonemap(final_ubo, stream=stream)
            ctx.join(stream); c1.record(); torch.cuda.synchronize()
            t = torch.tensor([c0.elapsed_time(c1) / reps], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
            exchange = (exchange or {}) | {arm + "_ms": float(t.item())}
        disp, _ = ctx.readback_display()
        exchange |= {"fused": "rt_combine: snapshot + ONE kernel (peer loads of the other ranks' sums for this rank's band, tonemap, peer stores of the RGBA8 band into every rank's display) + device-side flags",
                     "nccl": "torch.distributed all_reduce (NCCL) of the RGBA32F sums + rt_tonemap of the whole image on every rank",
                     "bytes_rgba32f": WIDTH * HEIGHT * 16, "complete_image_on_this_rank": bool((disp[..., 3] == 255).all()),
                     "periodic_every_steps": combine_every}

    # ---- second, labelled figure: the round-1 camera outside the box (not the headline; kept for continuity) ----
    alt = None
    if CONFIG == 2 and world == 1 and not args.no_alt_camera:
        cam2 = host.Camera(WIDTH, HEIGHT).set(position=CAM_ALT)
        ubos2 = [frame_ubo(cam2, gui, s, desc.fully_opaque) for s in range(Wm + K)]
        ctx.set_frames_in_flight(NF); ctx.resize(WIDTH, HEIGHT)
        for s in range(Wm):
            ctx.render(scene, ubos2[s], stream=stream)
        ctx.synchronize(); torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for s in range(Wm, Wm + K):
            ctx.render(scene, ubos2[s], stream=stream)
        ctx.join(stream); a1.record(); torch.cuda.synchronize()
        alt_ms = a0.elapsed_time(a1)
        ctx.set_frames_in_flight(1); ctx.resize(WIDTH, HEIGHT)
        alt_rays = 0
        for s in range(Wm + K):
            ctx.render(scene, ubos2[s], stream=stream)
            if s >= Wm:
                st2 = ctx.stats(); alt_rays += st2.rays_extend + st2.rays_shadow
        alt = {"camera": {"position": list(CAM_ALT), "direction": [0, 0, -1], "note": "round-1 headline camera, outside the box (about half the rays per pixel, ~30 % misses)"},
               "value": alt_rays / (alt_ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": alt_ms / K, "rays_per_step": alt_rays / K}

    if rank == 0:
        peak, peak_src = measured_peak()
        trav_bytes, frame_bytes = algorithmic_bytes(tot)
        # dominant kernel = the persistent traversal (extend_kernel; shadow_kernel is the same code in any-hit mode and is
        # included when the scene casts shadow rays): algorithmic bytes of every traced ray / sum of the launch durations
        trav_ms = stage["extend"] + stage["shadow"]
        achieved = trav_bytes / (trav_ms * 1e-3) / 1e9 if trav_ms > 0 else None
        traffic, issue = None, None
        tj = ROOT / "profiles" / "r02_extend_traffic.json"     # dram bytes per launch + issue-slot figures from the committed ncu capture
        if tj.exists():
            try:
                tjd = json.loads(tj.read_text())
                traffic = float(tjd["mean_dram_bytes_per_launch"]); issue = tjd.get("issue_slots")
            except Exception:
                traffic = None
        n_trav = n_ext * (2 if tot["rays_shadow"] else 1)
        roof = {"bound": "hbm", "kernel": "extend_kernel (persistent BVH8 traversal + watertight test)" + (" + shadow_kernel" if tot["rays_shadow"] else ""),
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write, profiles/r02_extend_traffic.json)",
                "algorithmic_bytes_per_launch": trav_bytes / max(1, n_trav), "peak_source": peak_src,
                "algorithmic_bytes_per_ray": trav_bytes / max(1, rays_rank), "avg_launch_ms": trav_ms / max(1, n_trav), "launches": n_trav,
                "nodes_per_ray": tot["nodes"] / max(1, rays_rank), "tris_per_ray": tot["tris"] / max(1, rays_rank),
                "regimes": {"lone": "frac/achieved: CUDA events around each launch, one frame at a time (frames in flight = 1) over the same K frames",
                            "overlapped_upper_bound_frac": trav_bytes / (ms * 1e-3) / 1e9 / peak,
                            "overlapped": "traversal bytes / the whole timed region of `value` (frames in flight overlap traversal with shading): an upper bound for the kernel's share"},
                "issue_slots": issue,
                "note": "HBM is the roofline BASELINE.json names; the BVH is L2-resident, DRAM traffic is a small fraction of the algorithmic bytes and the kernel is bound by instruction issue (see issue_slots and DESIGN.md §4)",
                "whole_frame_algorithmic_gbs": frame_bytes / (ms * 1e-3) / 1e9, "whole_frame_frac": frame_bytes / (ms * 1e-3) / 1e9 / peak,
                "whole_frame_bytes_per_step": frame_bytes / K, "counters_per_step": {k: v / K for k, v in tot.items()},
                "stage_ms_per_step": {k: v / K for k, v in stage.items()}}
        cpu = None
        if not args.no_cpu_baseline and world == 1:     # reported at N=1 only (torchrun pins OMP_NUM_THREADS=1)
            try:
                from oracle import orc
                m, rays, secs = cpu_sample(desc, 1, (0, HEIGHT))
                nfr = int(max(1, min(64, round(15.0 / max(secs, 1e-3)))))   # ~15 s of CPU work
                if nfr > 1:
                    m, rays, secs = cpu_sample(desc, nfr, (0, HEIGHT), warm=0)
                cpu = {"value": m, "unit": "Mrays/s", "cores": int(orc.lib().orc_num_threads()), "kind": "port",
                       "sample": f"{nfr} full {WIDTH}x{HEIGHT} frames x 1 spp, depth 8 ({rays} rays, {secs:.1f} s), oracle/oracle.cpp, OpenMP all host threads"}
            except Exception as ex:   # the oracle is test infrastructure; its absence must not break the GPU arm
                cpu = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "port", "sample": f"unavailable: {ex}"}
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_max / K,
                "higher_is_better": True, "scaling": "strong" if tiles else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(desc),      # identical keys / values in both arms
                "engine": {"parallelism": (f"tile partition x{world}: every frame split into interleaved 8-row strips, RGBA8 strips gathered on rank 0 every step (rt_combine, rooted)" if tiles else
                                           f"sample-pass sharding x{world} (frames g = step*{world}+rank); rt_combine at the end: fused peer-memory reduce + tonemap + all-gather, device-side flags, no host barrier"),
                           "exchange": exchange,
                           "frames_in_flight": NF,
                           "bvh": {"nodes": int(info.blas_nodes), "depth": int(info.max_depth_blas), "bytes": int(info.bytes), "build_s": build_s,
                                   "build_ms_device": float(info.build_ms), "build_ms_device_first": first_build_ms, "tlas_ms_device": float(info.tlas_ms)}},
                "samples_per_s": samples_all / (ms_max * 1e-3),
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 324 + (16384 if POSE is not None else 0), "d2h_bytes_per_step": WIDTH * HEIGHT * 4},
                "refit": ((upd | {"skinning_gbs": 256.0 * desc.n_vertices / (upd["skin_ms"] * 1e-3) / 1e9 if upd["skin_ms"] > 0 else None,
                                  "refit_algorithmic_gbs": (160.0 * info.blas_nodes + 36.0 * info.blas_tris) / (upd["refit_ms"] * 1e-3) / 1e9 if upd["refit_ms"] > 0 else None,
                                  "vertices": int(desc.n_vertices), "timing": "CUDA events per update stage, one frame at a time"}) if POSE is not None else None),
                "gpu_launches": int(launches),
                "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "alt_camera": alt,
                "rays_per_step": rays_rank / K}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scene-versions", type=int, default=3, help="config 4: copies of the buffers a skin update rewrites (2..4)")
    ap.add_argument("--partition", default="samples", choices=["samples", "tiles"],
                    help="N > 1: samples = whole frames per rank, private sums, fused combine (weak scaling, default); tiles = every frame split into interleaved 8-row strips (strong scaling)")
    ap.add_argument("--combine-every", type=int, default=-1, help="sample passes: periodic display combine every N steps on a side stream (default: 64 for config 5, else off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-alt-camera", action="store_true", help="skip the second (round-1 camera) figure of config 2")
    ap.add_argument("--size", default=None, help="WxH override (experiments only; the headline size is 1920x1080)")
    ap.add_argument("--camera-z", type=float, default=None, help="camera z override (experiments only)")
    ap.add_argument("--frames-in-flight", type=int, default=4, help="frames whose path tracing may overlap on one GPU (1..4; reference: IN_FLIGHT_FRAMES = 2)")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE.json config (default 2 = the headline workload)")
    args = ap.parse_args()
    select_config(args.config)
    global WIDTH, HEIGHT, CAM_POS
    if args.size:
        WIDTH, HEIGHT = (int(x) for x in args.size.lower().split("x"))
    if args.camera_z is not None:
        CAM_POS = (CAM_POS[0], CAM_POS[1], float(args.camera_z))
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
