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
prec = sys.argv[1] if len(sys.argv) > 1 else "3xf16r"
for _ in range(3):
    ops.conv3d(pc, x, y, precision=prec)
torch.cuda.synchronize()
lib = _lib.get()
buf = (ctypes.c_longlong * (148 * 4))()
fn = lib.estd_ring2_timing if prec == "3xf16r2" else lib.estd_ring_timing
fn.argtypes = [ctypes.c_void_p]
assert fn(buf) == 0
t = torch.tensor(list(buf), dtype=torch.float64).reshape(148, 4)
t = t[t[:, 3] > 0]
print("per CTA mean: wait ready %.0f clk, wait acc_empty %.0f clk, issuer total %.0f clk, stages %.1f" % tuple(t.mean(0).tolist()))
print("per stage: ready %.0f, acc_empty %.0f, total %.0f" % tuple((t[:, :3].sum(0) / t[:, 3].sum()).tolist()))
print("max total %.0f min total %.0f" % (t[:, 2].max().item(), t[:, 2].min().item()))
if prec == "3xf16r2":
    buf8 = (ctypes.c_longlong * (148 * 8))()
    lib.estd_ring2_epi_timing.argtypes = [ctypes.c_void_p]
    assert lib.estd_ring2_epi_timing(buf8) == 0
    e = torch.tensor(list(buf8), dtype=torch.float64).reshape(148, 8)
    n = e[:, 2].sum()
    print("epilogue warp 4, per hand-over (clk): wait acc_full %.0f | full->arrive %.0f = 2 x (tcgen05.ld %.0f + math/stores %.0f) + wait::st %.0f + barrier %.0f"
          % (e[:, 3].sum() / n, e[:, 1].sum() / n, e[:, 4].sum() / n / 2, e[:, 5].sum() / n / 2, e[:, 0].sum() / n, e[:, 6].sum() / n))
