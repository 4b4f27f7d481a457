"""Isolated timing of the plane-ring / planar kernels with fp32 vs PRE-SPLIT (vol4s) tensors on either side, (cfg2 shapes).  Run on a B200:  python profiles/bench_split.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from estdepth_b200 import ops, packing, synth  # noqa: E402

dev = "cuda"


def timeit(fn, n=40, rounds=4):
    """Best of `rounds` batches of `n` back-to-back launches (a fresh box ramps its clocks and hits its power cap at its own pace)."""
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
for cin, cout, pad in ((32, 32, 32), (16, 16, 16), (36, 33, 48)):
    chunks = (cin + 3) // 4
    w = torch.randn(cout, cin, 3, 3, 3, generator=g) / (cin * 27) ** 0.5
    order = list(range(cout)) + [-1] * (pad - cout)
    oc = (cout + 3) // 4
    pc = packing.attach_tc(ops.PackedConv(packing.pack_weight(w, list(range(cin)), order), torch.ones(pad), torch.zeros(pad), chunks, pad, oc, pad, "relu", "relu")).to(dev)
    x = torch.randn(chunks, D, H, W, 4, device=dev)
    xs = ops.to_split(x)
    outs = [torch.empty((oc + 1) // 2 * 2, D, H, W, 4, device=dev) for _ in range(2)]
    res = torch.randn(oc, D, H, W, 4, device=dev)
    for name, kw, xin, nout in (("fp32 -> fp32", {}, x, oc), ("split -> fp32", dict(in_split=(True, False)), xs, oc),
                                ("fp32 -> split", dict(out_split=True), x, (oc + 1) // 2 * 2), ("split -> split", dict(in_split=(True, False), out_split=True), xs, (oc + 1) // 2 * 2)):
        t = timeit(lambda i: ops.conv3d(pc, xin, outs[i % 2][:nout], precision="3xf16r2", **kw))
        print("ring2 %2d->%2d %-15s %7.1f us" % (cin, cout, name, t))
    t = timeit(lambda i: ops.conv3d(pc, x, outs[i % 2][:oc], res0=res, precision="3xf16r2"))
    print("ring2 %2d->%2d fp32 + residual    %7.1f us" % (cin, cout, t))
    for name, kw, xin, nout in (("fp32 -> fp32", {}, x, oc), ("split -> fp32", dict(in_split=(True, False)), xs, oc),
                                ("fp32 -> split", dict(out_split=True), x, (oc + 1) // 2 * 2), ("split -> split", dict(in_split=(True, False), out_split=True), xs, (oc + 1) // 2 * 2)):
        t = timeit(lambda i: ops.conv3d(pc, xin, outs[i % 2][:nout], precision="3xf16r2d", **kw))
        print("ring2 DUAL %2d->%2d %-15s %7.1f us" % (cin, cout, name, t))
    t = timeit(lambda i: ops.conv3d(pc, x, outs[i % 2][:oc], res0=res, precision="3xf16r2d"))
    print("ring2 DUAL %2d->%2d fp32 + residual    %7.1f us" % (cin, cout, t))

N, H2, W2 = 5, 120, 160
for cin, cout in ((64, 64), (128, 128)):
    w = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    pc = packing.pack_conv2d(w, torch.ones(cout), torch.zeros(cout), "relu", dev)[0]
    x = torch.randn(cin // 4, N, H2, W2, 4, device=dev)
    xs = ops.to_split(x)
    outs = [torch.empty(cout // 4, N, H2, W2, 4, device=dev) for _ in range(2)]
    for name, kw, xin in (("fp32 -> fp32", {}, x), ("split -> fp32", dict(in_split=(True, False)), xs), ("fp32 -> split", dict(out_split=True), x),
                          ("split -> split", dict(in_split=(True, False), out_split=True), xs)):
        t = timeit(lambda i: ops.conv_planar(pc, xin, outs[i % 2], **kw))
        print("planar %3d->%3d %-15s %7.1f us" % (cin, cout, name, t))

