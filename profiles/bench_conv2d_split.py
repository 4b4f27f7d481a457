"""Is 'three-term TF32 split through cuDNN' faster than cuDNN's strict-fp32 path for the 2-D feeder layers?
y = conv(cat[x_hi, x_hi, x_lo], cat[w_hi, w_lo, w_hi]) with cudnn.allow_tf32=True (inputs pre-truncated to TF32, so the
tensor-core products are exact) vs y = conv(x, w) with allow_tf32=False."""
import torch, torch.nn.functional as F
torch.backends.cudnn.benchmark = True
M = -8192
def split(t):
    hi = (t.view(torch.int32) & M).view(torch.float32)
    lo = ((t - hi).view(torch.int32) & M).view(torch.float32)
    return hi, lo
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
cases = [  # N, Cin, Cout, H, W, k, stride, dil
    (5, 3, 32, 480, 640, 3, 2, 1), (5, 32, 32, 240, 320, 3, 1, 1), (5, 32, 64, 240, 320, 3, 2, 1), (5, 64, 64, 120, 160, 3, 1, 1),
    (5, 128, 128, 120, 160, 3, 1, 1), (5, 128, 128, 120, 160, 3, 1, 2), (5, 320, 128, 120, 160, 3, 1, 1),
    (3, 64, 64, 120, 160, 1, 1, 1), (3, 64, 64, 120, 160, 3, 1, 1), (3, 128, 128, 60, 80, 3, 1, 1), (3, 256, 256, 30, 40, 3, 1, 1),
    (3, 512, 512, 15, 20, 3, 1, 1), (3, 2048, 256, 15, 20, 3, 1, 1), (3, 1280, 256, 30, 40, 3, 1, 1), (3, 320, 64, 120, 160, 3, 1, 1)]
tot_a = tot_b = 0
for (N, ci, co, H, W, k, s, d) in cases:
    x = torch.randn(N, ci, H, W, device="cuda"); w = torch.randn(co, ci, k, k, device="cuda") / (ci * k * k) ** 0.5
    pad = d * (k // 2)
    for fmt in (torch.contiguous_format, torch.channels_last):
        xf, wf = x.contiguous(memory_format=fmt), w.contiguous(memory_format=fmt)
        torch.backends.cudnn.allow_tf32 = False
        ta = timeit(lambda: F.conv2d(xf, wf, None, s, pad, d))
        ref = F.conv2d(xf, wf, None, s, pad, d)
        xh, xl = split(x); wh, wl = split(w)
        x3 = torch.cat([xh, xh, xl], 1).contiguous(memory_format=fmt); w3 = torch.cat([wh, wl, wh], 1).contiguous(memory_format=fmt)
        torch.backends.cudnn.allow_tf32 = True
        tb = timeit(lambda: F.conv2d(x3, w3, None, s, pad, d))
        y = F.conv2d(x3, w3, None, s, pad, d)
        tsplit = timeit(lambda: torch.cat([split(x)[0], split(x)[0], split(x)[1]], 1))
        ref64 = F.conv2d(x.double().cpu(), w.double().cpu(), None, s, pad, d) if N * ci * H * W < 3e7 and co * ci < 70000 else None
        e_ref = (ref.cpu().double() - ref64).abs().max().item() if ref64 is not None else float("nan")
        e_spl = (y.cpu().double() - ref64).abs().max().item() if ref64 is not None else float("nan")
        gf = 2.0 * N * co * ci * k * k * (H // s) * (W // s) / 1e9
        print("N%d %4d->%4d %3dx%3d k%d s%d d%d %-13s fp32 %.3f ms (%.0f TF/s) | split-tf32 %.3f ms (+split %.3f) | err fp32 %.1e split %.1e"
              % (N, ci, co, H, W, k, s, d, str(fmt).split(".")[-1], ta, gf / ta, tb, tsplit, e_ref, e_spl))
