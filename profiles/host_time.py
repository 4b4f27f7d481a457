"""Is the step host-bound?  Host time to ISSUE one steady-state cfg2 window vs. the device time to execute it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from estdepth_b200 import DepthNetHybrid, synth
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
geometry = sys.argv[1] if len(sys.argv) > 1 else "torch"
model = DepthNetHybrid(ndepths=64, depth_min=0.1, depth_max=10.0, resnet=50, geometry=geometry)
model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0))
model.eval().to(dev)
w1 = [t.to(dev) for t in synth.synth_inputs(5, 480, 640, seed=0, start=0)[:3]]
w2 = [t.to(dev) for t in synth.synth_inputs(5, 480, 640, seed=0, start=3)[:3]]
_, state, pstate = model(w1[0], w1[1], w1[2], None, mode="val")
for _ in range(3):
    model(w2[0], w2[1], w2[2], None, state, pstate, mode="val")
torch.cuda.synchronize()
n = 20
t0 = time.perf_counter()
for _ in range(n):
    model(w2[0], w2[1], w2[2], None, state, pstate, mode="val")
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("geometry=%s: host issue time %.2f ms/step, wall incl. device %.2f ms/step" % (geometry, (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3))
