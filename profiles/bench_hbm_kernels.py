"""Isolated timing of the HBM-bound kernels at cfg2 size: K1 warp->cost, K3 EST attention, K4 head+soft-argmin, K5 GRU glue.
Back-to-back launches between two CUDA events, outputs rotating over buffers larger than L2."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from estdepth_b200 import ops, synth  # noqa: E402

D, H, W = (64, 120, 160) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1:4])
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
peak = 6551.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, n=30):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


def report(name, t, nbytes):
    print("%-16s %8.1f us  %7.1f MB  %7.0f GB/s  %5.1f %% of %.0f GB/s" % (name, t * 1e6, nbytes / 1e6, nbytes / t / 1e9, 100 * nbytes / t / 1e9 / peak, peak))


poses = synth.camera_track(5).to(dev)
K4 = synth.intrinsics(4 * H, 4 * W).clone()
K4[:2] *= 0.25
K4 = K4.to(dev)
dmin, dmax = 0.1, 10.0
interval = (dmax - dmin) / (D - 1)
dv = (torch.arange(D, dtype=torch.float32) * interval + dmin).to(dev)

maps = [torch.randn(8, H, W, 4, generator=g).to(dev) for _ in range(2)]
homo = ops.homography_setup(poses[1].contiguous(), poses[0].contiguous(), K4)
vols = [torch.empty(8, D, H, W, 4, device=dev) for _ in range(3)]
report("K1 warp_cost", timed(lambda i: ops.warp_cost(maps[0], maps[1], homo, dv, vols[i % 3])), 4.0 * 32 * H * W * (D + 2))

keys = [torch.rand(4, D, H, W, 4, generator=g).to(dev) for _ in range(4)]
vals = [torch.rand(4, D, H, W, 4, generator=g).to(dev) for _ in range(4)]
outs = [torch.empty(4, D, H, W, 4, device=dev) for _ in range(3)]
for n in (1, 2, 3):
    w30 = torch.stack([ops.volume_warp_setup(poses[1].contiguous(), poses[j].contiguous(), K4) for j in (0, 2, 3)[:n]])
    report("K3 est N=%d" % n, timed(lambda i: ops.est_attend(keys[0], keys[1:1 + n], vals[1:1 + n], w30, dv, dmin, interval, out=outs[i % 3])),
           4.0 * 16 * D * H * W * (2 + 2 * n))

hw, hb = torch.randn(16, generator=g).to(dev), torch.zeros(1, device=dev)
logits = [torch.empty(D, H, W, device=dev) for _ in range(3)]
depth = [torch.empty(4 * H, 4 * W, device=dev) for _ in range(3)]
prob = [torch.empty(4 * H, 4 * W, device=dev) for _ in range(3)]
report("K4 head+argmin", timed(lambda i: ops.head_softargmin(dv, hidden=keys[i % 4], head_w=hw, head_b=hb, logits_out=logits[i % 3],
                                                             depth_out=depth[i % 3], prob_out=prob[i % 3], up=4)),
       4.0 * (16 * D * H * W + D * H * W + 2 * 16 * H * W))

# premix: both halves of pre0 on every frame of a window, one launch (5 x [32,H,W] -> 5 x [16 chunks,H,W,4])
fea = [torch.randn(5, 32, H, W, generator=g).to(dev) for _ in range(3)]
w64, b64 = (torch.randn(64, 32, generator=g) / 5).to(dev), torch.randn(64, generator=g).to(dev)
mix = [torch.empty(5, 16, H, W, 4, device=dev) for _ in range(3)]
report("premix x5", timed(lambda i: ops.premix_batch(fea[i % 3], w64, b64, out=mix[i % 3])), 4.0 * 5 * (32 + 64) * H * W)
