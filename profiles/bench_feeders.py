"""Experiment: the cuDNN-side 2-D feeders (ResNet-50 trunk, context decoder, refinement) in NCHW vs channels_last,
strict fp32 (allow_tf32=False), cudnn.benchmark=True.  python profiles/bench_feeders.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from estdepth_b200 import DepthNetHybrid, synth, encoders  # noqa: E402

torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
model = DepthNetHybrid(ndepths=64, depth_min=0.1, depth_max=10.0, resnet=50)
model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0))
model.eval().to(dev)
imgs = torch.randn(3, 3, 480, 640, device=dev)
imgs5 = torch.randn(5, 3, 480, 640, device=dev)
logits = torch.randn(3, 64, 120, 160, device=dev)


def timed(fn, n=5):
    with torch.no_grad():
        for _ in range(3):
            out = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


def run(cl):
    encoders._folded.cache.clear()
    encoders._folded.channels_last = cl
    x = imgs.contiguous(memory_format=torch.channels_last) if cl else imgs
    t_enc, maps = timed(lambda: model.semanticFeature(x))
    t_dec, sem = timed(lambda: model.CostRegNet.context(maps))
    t_ref, _ = timed(lambda: model.CostRegNet.refine(sem, logits.contiguous(memory_format=torch.channels_last) if cl else logits, maps[0]))
    encoders._folded.channels_last = False
    t_psm, feats = timed(lambda: model.matchingFeature(imgs5))
    return (t_enc, t_dec, t_ref, t_psm), [m.float().contiguous() for m in maps] + [sem.contiguous(), feats.contiguous()]


a, ma = run(False)
print("NCHW          : resnet50 %.2f ms, context decoder %.2f ms, refine %.2f ms, psm(tc) %.2f ms" % a)
if hasattr(encoders._folded, "channels_last"):
    b, mb = run(True)
    print("channels_last : resnet50 %.2f ms, context decoder %.2f ms, refine %.2f ms, psm(tc) %.2f ms" % b)
    print("max |diff|:", [float((p - q).abs().max()) for p, q in zip(ma, mb)])
