"""Read-bandwidth probe: what a plain streaming read of the SpMV's byte count achieves on this GPU (context for
roofline fractions: MEASURED_PEAKS.json's figure is a copy, read+write)."""
import torch
n = 350_000_000 // 8
x = torch.randn(n, dtype=torch.float64, device="cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
y = torch.empty_like(x)
def t(fn, reps=10):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for name, fn, bytes_ in (("sum (read)", lambda: x.sum(), n * 8), ("copy (read+write)", lambda: y.copy_(x), 2 * n * 8),
                         ("dot (2 reads)", lambda: torch.dot(x, y), 2 * n * 8), ("abs max", lambda: x.abs().max(), 0)):
    ms = t(fn)
    print(f"{name:20s} {ms*1e3:8.1f} us  {bytes_/ms/1e6:8.1f} GB/s")
