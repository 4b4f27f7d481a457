"""cProfile of the host side of one steady-state cfg2 window (which Python calls take the 15 ms of issue time)."""
import cProfile, pstats, os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from estdepth_b200 import DepthNetHybrid, synth
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
model = DepthNetHybrid(ndepths=64, depth_min=0.1, depth_max=10.0, resnet=50, geometry=(sys.argv[1] if len(sys.argv) > 1 else "torch"))
model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0))
model.eval().to(dev)
w1 = [t.to(dev) for t in synth.synth_inputs(5, 480, 640, seed=0, start=0)[:3]]
w2 = [t.to(dev) for t in synth.synth_inputs(5, 480, 640, seed=0, start=3)[:3]]
_, state, pstate = model(w1[0], w1[1], w1[2], None, mode="val")
for _ in range(3):
    model(w2[0], w2[1], w2[2], None, state, pstate, mode="val")
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    model(w2[0], w2[1], w2[2], None, state, pstate, mode="val")
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
