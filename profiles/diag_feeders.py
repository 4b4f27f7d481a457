"""Diagnostic: 2-D feeders on the planar tensor-core kernel vs cuDNN strict fp32, per stage, at a given size."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from estdepth_b200 import DepthNetHybrid, synth, ops  # noqa: E402

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (480, 640)
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
model = DepthNetHybrid(ndepths=64, depth_min=0.1, depth_max=10.0, resnet=50)
model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0))
model.eval().to(dev)
imgs = synth.synth_inputs(5, H, W, seed=0, start=3)[0].to(dev)
x = 2 * (imgs[0] / 255.) - 1.


def rel(a, b):
    return float((a - b).abs().max()), float(b.abs().max())


with torch.no_grad():
    enc, dec, psm = model.semanticFeature, model.CostRegNet, model.matchingFeature
    enc.tensor_cores = False
    maps_ref = enc(x[1:4])
    enc.tensor_cores = True
    maps_tc = enc(x[1:4])
    for i, (a, b) in enumerate(zip(maps_tc, maps_ref)):
        print("resnet map %d %s: max|diff| %.2e, max|ref| %.2e" % ((i, tuple(b.shape)) + rel(a, b)))
    dec.tensor_cores = False
    sem_ref = dec.context(maps_ref)
    dec.tensor_cores = True
    sem_tc_same_in = dec.context([m.clone() for m in maps_ref])
    sem_tc = dec.context(maps_tc)
    print("context decoder (same inputs) : max|diff| %.2e, max|ref| %.2e" % rel(sem_tc_same_in, sem_ref))
    print("context decoder (tc encoder)  : max|diff| %.2e, max|ref| %.2e" % rel(sem_tc, sem_ref))
    logits = torch.randn(3, 64, H // 4, W // 4, device=dev)
    dec.tensor_cores = False
    h_ref, f_ref = dec.refine(sem_ref, logits, maps_ref[0])
    dec.tensor_cores = True
    h_tc, f_tc = dec.refine(sem_ref, logits, maps_ref[0])
    print("refine (same inputs) half: max|diff| %.2e; full: max|diff| %.2e (max %.2f)" % (rel(h_tc, h_ref)[0], rel(f_tc, f_ref)[0], float(f_ref.max())))
    psm.tensor_cores = False
    p_ref = psm(x)
    psm.tensor_cores = True
    p_tc = psm(x)
    print("psm features: max|diff| %.2e, max|ref| %.2e" % rel(p_tc, p_ref))
    ops.check_status(dev)
