"""One steady-state cfg2 step (5x480x640, D=64, R50, EST window) between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches_rNN.csv python profiles/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv3d -c 2 \
      -o gpurun_out/conv3d_rNN python profiles/profile_step.py
Numbers printed by a run under ncu are never bench values.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from estdepth_b200 import DepthNetHybrid, synth  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
V, H, W, D, resnet = {"cfg2": (5, 480, 640, 64, 50), "cfg1": (5, 128, 160, 32, 18), "cfg5": (5, 640, 960, 128, 50)}[workload]
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
model = DepthNetHybrid(ndepths=D, depth_min=0.1, depth_max=10.0, resnet=resnet, precision=(sys.argv[2] if len(sys.argv) > 2 else "3xf16r2d"))
model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0))
model.eval().to(dev)
w1 = synth.synth_inputs(V, H, W, seed=0, start=0)
w2 = list(synth.synth_inputs(V, H, W, seed=0, start=V - 2)[:3])
w2[0] = w2[0].to(dev)                       # images resident, camera parameters on the host (as bench.py)
_, state, pstate = model(w1[0].to(dev), w1[1], w1[2], None, mode="val")
for _ in range(2):
    model(w2[0], w2[1], w2[2], None, state, pstate, mode="val")
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
model(w2[0], w2[1], w2[2], None, state, pstate, mode="val")
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one %s step" % workload)
