"""Where does the tensor-core kernels' error come from, and is it a BIAS that can be compensated?

(1) One 32->32 3x3x3 layer (and planar 3x3 layers of several widths) on random data against an fp64 convolution on the GPU:
    signed error e = y - y_ref regressed on y_ref (slope = multiplicative bias), rms before / after removing the slope, for
    ReLU'd (non-negative) and signed inputs.
(2) The whole cfg2 model against the benchmark-size reference fixture with parts switched between the tensor-core and the
    exact / cuDNN fp32 implementations: which stage contributes how much of the depth error.
Run on a B200:  python profiles/trunc_probe.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from estdepth_b200 import DepthNetHybrid, ops, packing, synth  # noqa: E402
from tests.helpers import to_vol4, from_vol4  # noqa: E402

dev = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def stats(name, y, ref):
    e = (y.double() - ref).flatten()
    r = ref.flatten()
    slope = float((e * r).sum() / (r * r).sum())
    res = e - slope * r
    print("  %-34s rms err %.3e  max %.3e | slope %+.3e (x 2^-24 = %+.1f) | rms after slope removal %.3e | ref rms %.3f"
          % (name, float(e.pow(2).mean().sqrt()), float(e.abs().max()), slope, slope / 2.0 ** -24, float(res.pow(2).mean().sqrt()),
             float(r.pow(2).mean().sqrt())))
    return slope


def conv3d_probe():
    print("== 3-D 3x3x3 layers vs fp64 (one layer; 'relu' = non-negative inputs, as after a ReLU)")
    for cin, cout in ((32, 32), (16, 16)):
        for kind in ("relu", "signed"):
            g = torch.Generator().manual_seed(1)
            D, H, W = 8, 32, 64
            x = torch.randn(cin, D, H, W, generator=g)
            if kind == "relu":
                x = x.relu()
            w = torch.randn(cout, cin, 3, 3, 3, generator=g) / (cin * 27) ** 0.5
            ref = F.conv3d(x.double().to(dev).unsqueeze(0), w.double().to(dev), None, 1, 1)[0]
            pw = packing.pack_weight(w, list(range(cin)), list(range(cout)))
            pc = packing.attach_tc(ops.PackedConv(pw, torch.ones(cout), torch.zeros(cout), cin // 4, cout, cout // 4, cout, "none", "none")).to(dev)
            xin = to_vol4(x).to(dev)
            print(" %d->%d %s" % (cin, cout, kind))
            for prec in ("fp32", "3xf16", "3xf16r", "3xf16r2", "3xf16r2d"):
                out = torch.empty(cout // 4, D, H, W, 4, device=dev)
                ops.conv3d(pc, xin, out, precision=prec)
                stats(prec, from_vol4(out), ref)


def conv2d_probe():
    print("== planar 3x3 layers vs fp64")
    for cin, cout in ((64, 64), (128, 128), (320, 128), (1280, 256)):
        g = torch.Generator().manual_seed(2)
        N, H, W = 2, 60, 80
        x = torch.randn(N, cin, H, W, generator=g).relu()
        w = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
        ref = F.conv2d(x.double().to(dev), w.double().to(dev), None, 1, 1)
        pcs = packing.pack_conv2d(w, torch.ones(cout), torch.zeros(cout), "none", dev, cout_slice=64 if cout > 32 else 32)
        x4 = ops.nchw_to_vol4(x.to(dev))
        out = torch.empty(cout // 4, N, H, W, 4, device=dev)
        ops.conv_planar(pcs[0], x4, out)
        y = ops.vol4_to_nchw(out)
        print(" %d->%d relu inputs (%d accumulating MMAs per accumulator)" % (cin, cout, 9 * cin // 16))
        stats("planar 3xf16", y, ref)
        stats("cuDNN fp32", F.conv2d(x.to(dev), w.to(dev), None, 1, 1), ref)


def attribution():
    print("== cfg2 (480x640, D=64, R50) vs the reference fixture: worst |depth diff| by configuration")
    from oracle.make_golden import subsample
    gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "joint_r50_d64_480x640_g3.npz"))
    for prec, psm_tc, ctx_tc in (("3xf16r2d", True, True), ("3xf16r2d", False, False), ("3xf16r2", True, True), ("3xf16r2", False, True), ("3xf16r2", True, False), ("3xf16r2", False, False),
                                 ("3xf16", True, True), ("3xf16", False, False), ("fp32", True, True), ("fp32", False, False)):
        model = DepthNetHybrid(ndepths=64, depth_min=0.1, depth_max=10.0, resnet=50, precision=prec)
        model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0))
        model.eval().to(dev)
        model.matchingFeature.tensor_cores = psm_tc
        model.semanticFeature.tensor_cores = model.CostRegNet.tensor_cores = ctx_tc
        state = pstate = None
        worst = {}
        for w, start in enumerate((0, 3)):
            imgs, poses, K, sample = synth.synth_inputs(5, 480, 640, seed=0, start=start)
            out, state, pstate = model(imgs.to(dev), poses, K, sample, state, pstate, mode="val")
            for key, val in out.items():
                gk = gold["w%d/%s" % (w, "_".join(str(k) for k in key))]
                d = float(np.abs(subsample(key, val, 4).cpu().numpy() - gk).max())
                tag = "depth%d" % key[2] if key[0] == "depth" else key[0]
                worst[tag] = max(worst.get(tag, 0.0), d)
        print("  conv3d %-8s psm %-5s context %-5s: %s" % (prec, "tc" if psm_tc else "cudnn", "tc" if ctx_tc else "cudnn",
                                                             {k: "%.2e" % v for k, v in sorted(worst.items())}))
        del model
        torch.cuda.empty_cache()


if __name__ == "__main__":
    with torch.no_grad():
        conv3d_probe()
        conv2d_probe()
        attribution()
