"""Experiment (library built with `make EXTRA=-DESTD_RING_TIMING`): where does the MMA issuer of conv3d_ring.cu wait?"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from estdepth_b200 import ops, packing, _lib  # noqa: E402

D, H, W = 64, 120, 160
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
w = torch.randn(32, 32, 3, 3, 3, generator=g) / 30
pc = packing.attach_tc(ops.PackedConv(packing.pack_weight(w, list(range(32)), list(range(32))).to(dev), torch.ones(32, device=dev),
                                      torch.zeros(32, device=dev), 8, 32, 8, 32, "relu", "relu"))
x = torch.randn(8, D, H, W, 4, generator=g).to(dev)
y = torch.empty_like(x)
for _ in range(3):
    ops.conv3d(pc, x, y, precision="3xf16r")
torch.cuda.synchronize()
lib = _lib.get()
buf = (ctypes.c_longlong * (148 * 4))()
lib.estd_ring_timing.argtypes = [ctypes.c_void_p]
assert lib.estd_ring_timing(buf) == 0
t = torch.tensor(list(buf), dtype=torch.float64).reshape(148, 4)
print("per CTA mean: wait ready %.0f clk, wait acc_empty %.0f clk, issuer total %.0f clk, stages %.1f" % tuple(t.mean(0).tolist()))
print("per stage: ready %.0f, acc_empty %.0f, total %.0f" % tuple((t[:, :3].sum(0) / t[:, 3].sum()).tolist()))
print("max total %.0f min total %.0f" % (t[:, 2].max().item(), t[:, 2].min().item()))
