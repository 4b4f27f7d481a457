"""Isolated timing of the planar tensor-core kernel on the layer shapes of the context branch (ResNet-50 on the 3 target frames
of a 480x640 window) and of the matching-feature net (5 frames), pre-split tensors on both sides as in the model.
Run on a B200:  python profiles/bench_planar_layers.py [--lib path/to/other/libestdepth_b200.so] [--only 2,12] [--launches 3]
(--lib times another build of the library -- A/B of kernel changes on one box; --only / --launches: a few launches of selected
layers for an ncu capture)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from estdepth_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=None)
ap.add_argument("--only", default=None)
ap.add_argument("--launches", type=int, default=0)
args = ap.parse_args()
if args.lib:
    _lib.LIB_PATH = os.path.abspath(args.lib)
from estdepth_b200 import ops, packing  # noqa: E402

dev = "cuda"


def timeit(fn, n=50, rounds=5):
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


# (name, cin, cout, taps, N, H, W, residual)
LAYERS = [
    ("r50 layer1 1x1 256->64", 256, 64, 1, 3, 120, 160, False),
    ("r50 layer1 3x3 64->64", 64, 64, 9, 3, 120, 160, False),
    ("r50 layer1 1x1 64->256 +res", 64, 256, 1, 3, 120, 160, True),
    ("r50 layer1 1x1 64->256 (downsample)", 64, 256, 1, 3, 120, 160, False),
    ("r50 layer2 1x1 512->128", 512, 128, 1, 3, 60, 80, False),
    ("r50 layer2 3x3 128->128", 128, 128, 9, 3, 60, 80, False),
    ("r50 layer2 1x1 128->512 +res", 128, 512, 1, 3, 60, 80, True),
    ("r50 layer3 1x1 1024->256", 1024, 256, 1, 3, 30, 40, False),
    ("r50 layer3 3x3 256->256", 256, 256, 9, 3, 30, 40, False),
    ("r50 layer3 1x1 256->1024 +res", 256, 1024, 1, 3, 30, 40, True),
    ("r50 layer4 1x1 2048->512", 2048, 512, 1, 3, 15, 20, False),
    ("r50 layer4 3x3 512->512", 512, 512, 9, 3, 15, 20, False),
    ("r50 layer4 1x1 512->2048 +res", 512, 2048, 1, 3, 15, 20, True),
    ("psm layer2 3x3 64->64 +res", 64, 64, 9, 5, 120, 160, True),
    ("psm layer3 3x3 128->128 +res", 128, 128, 9, 5, 120, 160, True),
]

g = torch.Generator().manual_seed(0)
total = 0.0
only = [int(v) for v in args.only.split(",")] if args.only else range(len(LAYERS))
for name, cin, cout, taps, n, h, w, with_res in [LAYERS[i] for i in only]:
    k = 3 if taps == 9 else 1
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * taps) ** 0.5
    pc = packing.pack_conv2d(wt, torch.ones(cout), torch.zeros(cout), "relu", dev)[0]
    xs = ops.to_split(torch.randn(cin // 4, n, h, w, 4, device=dev))
    outs = [torch.empty(cout // 4, n, h, w, 4, device=dev) for _ in range(2)]
    res = ops.to_split(torch.randn(cout // 4, n, h, w, 4, device=dev)) if with_res else None
    run = lambda i: ops.conv_planar(pc, xs, outs[i % 2], res0=res, taps=taps, in_split=(True, False), res_split=with_res, out_split=True)  # noqa: E731
    if args.launches:
        for i in range(args.launches):
            run(i)
        torch.cuda.synchronize()
        continue
    t = timeit(run)
    flop = 2.0 * taps * cin * cout * n * h * w
    mb = 4.0 * n * h * w * (cin + cout * (2 if with_res else 1)) / 1e6
    total += t
    print("%-38s %7.1f us  %6.1f TF/s  %6.0f GB/s" % (name, t, flop / t * 1e-6, mb / t * 1e3))
print("sum %.1f us" % total)
