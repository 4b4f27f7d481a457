"""Ceilings on this box for the store-bound / load-bound kernels: write-only, read-only and copy bandwidth over a
157 MB fp32 volume (cfg2's vol32), CUDA events, best of 20."""
import torch
n = 8 * 64 * 120 * 160 * 4
a = torch.empty(n, device="cuda"); b = torch.empty(n, device="cuda")
def best(fn, iters=20):
    t = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); t.append(e0.elapsed_time(e1))
    return min(t)
for name, fn, nbytes in (("fill (write-only)", lambda: a.fill_(1.0), 4 * n), ("sum (read-only)", lambda: a.sum(), 4 * n),
                         ("copy (read+write)", lambda: b.copy_(a), 8 * n)):
    ms = best(fn)
    print("%-20s %.1f us  %.0f GB/s" % (name, ms * 1e3, nbytes / ms / 1e6))
