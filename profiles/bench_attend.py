"""K3 (EST attention gather) isolated at cfg2 size, N = 1..3:  python profiles/bench_attend.py [path/to/other/libestdepth_b200.so]
(the optional argument times another build of the library: A/B of kernel variants on one box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from estdepth_b200 import _lib  # noqa: E402

if len(sys.argv) > 1:
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])
from estdepth_b200 import ops, synth  # noqa: E402

dev = "cuda"
D, H, W = 64, 120, 160
g = torch.Generator().manual_seed(5)
poses = synth.camera_track(5).to(dev)
K4 = synth.intrinsics(480, 640).clone()
K4[:2] *= 0.25
K4 = K4.to(dev)
kv = [torch.randn(4, D, H, W, 4, generator=g).to(dev) for _ in range(8)]
hs = [torch.empty(4, D, H, W, 4, device=dev) for _ in range(2)]
tabs = ops.volume_warp_tables_torch([poses[2], poses[1], poses[3], poses[0]], 1, K4)[0].contiguous()
dv = (torch.arange(D, dtype=torch.float32) * (9.9 / (D - 1)) + 0.1).to(dev)
tag = os.path.basename(os.path.dirname(_lib.LIB_PATH))
for n in (1, 2, 3):
    def run(i):
        ops.est_attend(kv[0], [kv[1 + 2 * j] for j in range(n)], [kv[2 + 2 * j] for j in range(n)], tabs[:n].contiguous(), dv, 0.1, 9.9 / (D - 1), out=hs[i % 2])
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    nbytes = 4.0 * 16 * D * H * W * (2 + 2 * n)
    print("%-16s N=%d %7.1f us  %6.0f GB/s  (%.3f of 6552)" % (tag, n, us, nbytes / us / 1e3, nbytes / us / 1e3 / 6552))
