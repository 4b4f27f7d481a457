"""Isolated timing of the planar tcgen05 3x3 convolution (conv2d_tc.cu) on the 2-D feeder shapes, next to cuDNN strict fp32."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from estdepth_b200 import ops, packing  # noqa: E402

torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


def timed(fn, n=20):
    for i in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


cases = [  # maps, cin, cout, H, W, dilation
    (5, 32, 32, 240, 320, 1), (5, 64, 64, 120, 160, 1), (5, 128, 128, 120, 160, 1), (5, 128, 128, 120, 160, 2), (5, 320, 128, 120, 160, 1),
    (3, 64, 64, 120, 160, 1), (3, 128, 128, 60, 80, 1), (3, 256, 256, 30, 40, 1), (3, 512, 512, 15, 20, 1),
    (3, 2048, 256, 15, 20, 1), (3, 1280, 256, 30, 40, 1), (3, 256, 128, 30, 40, 1), (3, 640, 128, 60, 80, 1), (3, 128, 64, 60, 80, 1),
    (3, 320, 64, 120, 160, 1), (3, 128, 32, 120, 160, 1), (3, 96, 32, 240, 320, 1)]
for N, cin, cout, H, W, dil in cases:
    x = torch.randn(N, cin, H, W, generator=g).to(dev)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).to(dev)
    b = torch.zeros(cout, device=dev)
    pcs = packing.pack_conv2d(w, torch.ones(cout), torch.zeros(cout), "relu", dev, cout_slice=64 if cout >= 64 else 32)
    x4 = ops.nchw_to_vol4(x)
    out4 = torch.empty(cout // 4, N, H, W, 4, device=dev)
    step = pcs[0].cout_pad // 4

    def run_tc():
        for i, pc in enumerate(pcs):
            ops.conv_planar(pc, x4, out4[step * i:step * i + pc.out_chunks], dilation=dil)

    t_tc = timed(run_tc)
    t_cudnn = timed(lambda: torch.cudnn_convolution_relu(x, w, b, (1, 1), (dil, dil), (dil, dil), 1))
    t_in = timed(lambda: ops.nchw_to_vol4(x, x4))
    t_out = timed(lambda: ops.vol4_to_nchw(out4))
    gf = 2.0 * N * cout * cin * 9 * H * W / 1e9
    print("N%d %4d->%3d %3dx%3d d%d: tcgen05 %7.1f us (%5.1f TF/s, %d launches) | cuDNN fp32 %7.1f us | nchw->vol4 %5.1f us, vol4->nchw %5.1f us"
          % (N, cin, cout, H, W, dil, t_tc, gf / t_tc * 1e3, len(pcs), t_cudnn, t_in, t_out))
