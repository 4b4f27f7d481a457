"""Isolated timing of the CTA-pair plane-ring kernel (precision 3xf16r2d) on a 32->32 layer at cfg2 size (64 x 120 x 160) with
0 / 1 / 2 residual volumes in the epilogue (pre1 of the second source, pre2: model.py `_cost_volume`).
Run on a B200:  python profiles/bench_ring_residual.py [--lib other/libestdepth_b200.so]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from estdepth_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=None)
args = ap.parse_args()
if args.lib:
    _lib.LIB_PATH = os.path.abspath(args.lib)
from estdepth_b200 import ops, packing  # noqa: E402

dev = "cuda"


def timeit(fn, n=40, rounds=5):
    for i in range(5):
        fn(i)
    best = 1e30
    for _ in range(rounds):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


g = torch.Generator().manual_seed(0)
D, H, W = 64, 120, 160
w = torch.randn(32, 32, 3, 3, 3, generator=g) / (32 * 27) ** 0.5
pc = packing.attach_tc(ops.PackedConv(packing.pack_weight(w, list(range(32)), list(range(32))), torch.ones(32), torch.zeros(32), 8, 32, 8, 32, "relu", "relu")).to(dev)
# several distinct volumes so that consecutive launches do not find their operands in L2 (the model's do not)
xs = [torch.randn(8, D, H, W, 4, device=dev) for _ in range(3)]
r0 = [torch.randn(8, D, H, W, 4, device=dev) for _ in range(3)]
r1 = [torch.randn(8, D, H, W, 4, device=dev) for _ in range(3)]
outs = [torch.empty(8, D, H, W, 4, device=dev) for _ in range(3)]
for name, kw in (("no residual", lambda i: {}), ("one residual", lambda i: dict(res0=r0[i % 3])), ("two residuals", lambda i: dict(res0=r0[i % 3], res1=r1[i % 3]))):
    t = timeit(lambda i: ops.conv3d(pc, xs[i % 3], outs[i % 3], precision="3xf16r2d", **kw(i)))
    print("ring2d 32->32 %-14s %7.1f us" % (name, t))

# the 16-channel layers (logit-head convolutions 16->16, ConvGRU gates 32->16): two M tiles per epilogue group
for cin in (16, 32):
    w = torch.randn(16, cin, 3, 3, 3, generator=g) / (cin * 27) ** 0.5
    pc = packing.attach_tc(ops.PackedConv(packing.pack_weight(w, list(range(cin)), list(range(16))), torch.ones(16), torch.zeros(16), cin // 4, 16, 4, 16, "relu", "relu")).to(dev)
    xin = [torch.randn(cin // 4, D, H, W, 4, device=dev) for _ in range(3)]
    t = timeit(lambda i: ops.conv3d(pc, xin[i % 3], outs[i % 3][:4], precision="3xf16r2d"))
    print("ring2d %2d->16               %7.1f us" % (cin, t))
