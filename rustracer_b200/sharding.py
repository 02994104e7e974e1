"""Frame sharding over the GPUs of one box: pure index arithmetic shared by bench.py and the tests.

Partition A — tiles: rows are cut into strips of `strip_rows`; rank r owns strips with strip % world == r
(rt_render_opts).  Bit-identical to the single-GPU image.
Partition B — sample passes: rank r renders global frames g with g % world == r into a private RGBA32F sum; the sums
are combined by rt_combine (fused peer-memory reduce + tonemap + all-gather, device-synchronised) — fp32 sum order differs from 1 GPU, so the
result is compared within tolerance.
"""
from __future__ import annotations


def global_frame(step: int, rank: int, world: int) -> int:
    """Global frame index rendered by `rank` at its local `step` (sample-pass sharding)."""
    return step * world + rank


def frames_of_rank(n_frames: int, rank: int, world: int) -> list[int]:
    return [g for g in range(n_frames) if g % world == rank]


def reduce_rows(rank: int, world: int, height: int) -> tuple[int, int]:
    """Row range [row0, row1) of the image that `rank` reduces and tonemaps."""
    per = (height + world - 1) // world
    return min(height, rank * per), min(height, (rank + 1) * per)


def owned_rows(height: int, strip_rows: int, world: int, rank: int) -> list[int]:
    """Rows rendered by `rank` under the tile partition (must match rt_core.h::owned_rows / local_to_pixel)."""
    rows = []
    n_strips = (height + strip_rows - 1) // strip_rows
    for k in range(rank, n_strips, world):
        rows.extend(range(k * strip_rows, min(height, (k + 1) * strip_rows)))
    return rows
