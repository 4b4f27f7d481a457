"""Isolated timing of the 3-D convolution kernels at cfg2 size (64 x 120 x 160): back-to-back launches between two CUDA
events, outputs rotating over buffers larger than L2.  python profiles/bench_conv3d.py [precision ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from estdepth_b200 import ops, packing  # noqa: E402

D, H, W = 64, 120, 160
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
precisions = sys.argv[1:] or ["3xf16r", "3xf16"]


def layer(cin_chunks, cout_pad, out_chunks):
    cin = 4 * cin_chunks
    w = torch.randn(cout_pad, cin, 3, 3, 3, generator=g) / (27 * cin) ** 0.5
    return packing.attach_tc(ops.PackedConv(packing.pack_weight(w, list(range(cin)), list(range(cout_pad))).to(dev),
                                            torch.ones(cout_pad, device=dev), torch.zeros(cout_pad, device=dev), cin_chunks, cout_pad,
                                            out_chunks, cout_pad, "relu", "relu"))


def timed(fn, n=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name, cin_chunks, cout_pad, out_chunks in (("32to32", 8, 32, 8), ("36to32", 9, 32, 8), ("16to16", 4, 16, 4), ("32to16", 8, 16, 4), ("36to40", 9, 40, 9)):
    pc = layer(cin_chunks, cout_pad, out_chunks)
    x = torch.randn(cin_chunks, D, H, W, 4, generator=g).to(dev)
    outs = [torch.empty(out_chunks, D, H, W, 4, device=dev) for _ in range(3)]
    res = torch.randn(out_chunks, D, H, W, 4, generator=g).to(dev)
    for prec in precisions:
        us = timed(lambda i: ops.conv3d(pc, x, outs[i % 3], precision=prec))
        us_res = timed(lambda i: ops.conv3d(pc, x, outs[i % 3], res0=res, precision=prec))
        gf = 54.0 * 4 * cin_chunks * cout_pad * D * H * W / 1e9
        print("%-7s %-7s %8.1f us  %6.1f TF/s   (+residual: %8.1f us)" % (name, prec, us, gf / us * 1e3, us_res))
