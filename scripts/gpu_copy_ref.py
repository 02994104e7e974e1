"""Reference point for the skinning kernel's roofline: a plain device copy of the same footprint (499 543 vertices x 128 B in,
the same out) timed like MEASURED_PEAKS.json times its 2 GiB copy (torch b.copy_(a), CUDA events, best of N), plus the 2 GiB
copy itself on this box."""
import torch
def best(nbytes, reps=30):
    a = torch.empty(nbytes // 2, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); b.copy_(a); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    t = min(ts)
    return t * 1e3, 2 * nbytes / t / 1e6
for name, nb in (("skinning footprint (63.9 MB in + 63.9 MB out)", 499543 * 128), ("2 GiB (MEASURED_PEAKS recipe)", 2 << 30)):
    us, gbs = best(nb)
    print(f"{name}: {us:.1f} us, {gbs:.0f} GB/s read+write")
