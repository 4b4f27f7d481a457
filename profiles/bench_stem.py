"""Isolated timing of the two stem kernels and the vol4 max-pool at cfg2 size, next to cuDNN's fp32 convolution / torch's max-pool
on the same tensors.  Run on a B200:  python profiles/bench_stem.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from estdepth_b200 import ops  # noqa: E402

dev = "cuda"
torch.backends.cudnn.allow_tf32 = False


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


g = torch.Generator().manual_seed(0)
img3 = torch.rand(3, 3, 480, 640, generator=g).to(dev)
img5 = torch.rand(5, 3, 480, 640, generator=g).to(dev)
w7 = (torch.randn(64, 3, 7, 7, generator=g) / 12).to(dev)
w7t = w7.permute(1, 2, 3, 0).contiguous()
b7 = torch.randn(64, generator=g).to(dev)
w3 = (torch.randn(32, 3, 3, 3, generator=g) / 5).to(dev)
b3 = torch.randn(32, generator=g).to(dev)
out7 = torch.empty(16, 3, 240, 320, 4, device=dev)
out3 = torch.empty(8, 5, 240, 320, 4, device=dev)
pool = torch.empty(16, 3, 120, 160, 4, device=dev)
print("stem7 (7x7/2, 3->64, 3 frames)   %7.1f us   cuDNN fp32 conv + bias + relu %7.1f us" % (
    timeit(lambda i: ops.stem7_conv(img3, w7t, b7, out7, out_split=True)),
    timeit(lambda i: F.relu(F.conv2d(img3, w7, b7, stride=2, padding=3)))))
print("stem  (3x3/2, 3->32, 5 frames)   %7.1f us   cuDNN fp32 conv + bias + relu %7.1f us" % (
    timeit(lambda i: ops.stem_conv(img5, w3, b3, out3, out_split=True)),
    timeit(lambda i: F.relu(F.conv2d(img5, w3, b3, stride=2, padding=1)))))
nchw = torch.randn(3, 64, 240, 320, device=dev)
print("max-pool 3x3/2 on vol4s          %7.1f us   torch max_pool2d (NCHW)       %7.1f us" % (
    timeit(lambda i: ops.maxpool3x3s2_vol4(out7, pool, in_split=True, out_split=True)),
    timeit(lambda i: F.max_pool2d(nchw, 3, stride=2, padding=1))))
