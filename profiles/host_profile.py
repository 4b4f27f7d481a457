"""Where the host spends the issue time of a step: wall-clock issue time (GPU idle at the start, nothing waited for) and a
cProfile of 10 steps, for a steady-state cfg2 Joint window and for an ESTM step (3 frames, 2 memory volumes) with and without
frame ids.      python profiles/host_profile.py [joint|estm|estm_ids]"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from estdepth_b200 import DepthNetHybrid, sharding, synth  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "joint"
dev = torch.device("cuda:0")
model = DepthNetHybrid(ndepths=64, depth_min=0.1, depth_max=10.0, resnet=50)
model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0))
model.eval().to(dev)
if mode == "joint":
    w1 = synth.synth_inputs(5, 480, 640, seed=0, start=0)[:3]
    w2 = synth.synth_inputs(5, 480, 640, seed=0, start=3)[:3]
    _, state, pstate = model(w1[0].to(dev), w1[1], w1[2], None, mode="val")
    img = w2[0].to(dev)

    def step(i):
        model(img, w2[1], w2[2], None, state, pstate, mode="val")
else:
    wins = [synth.synth_inputs(3, 480, 640, seed=0, start=s)[:3] for s in range(16)]
    wins = [(w[0].to(dev), w[1], w[2]) for w in wins]
    memory = []
    for s in range(3):
        pre = sharding._flatten_memory(memory)
        _, c, p = model(*wins[s], None, pre[0], pre[1], mode="val")
        memory = (memory + [(c, p)])[-2:]
    pre = sharding._flatten_memory(memory)

    def step(i):
        s = 3 + i % 13
        model(*wins[s], None, pre[0], pre[1], mode="val", frame_ids=[s, s + 1, s + 2] if mode == "estm_ids" else None)

for i in range(4):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(2):
    step(4 + i)
issue = (time.perf_counter() - t0) / 2 * 1e3
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10):
    step(i)
torch.cuda.synchronize()
total = (time.perf_counter() - t0) / 10 * 1e3
print("%s: host issue %.2f ms per step (GPU idle at start), %.2f ms per step end to end over 10 steps" % (mode, issue, total))
pr = cProfile.Profile()
pr.enable()
for i in range(10):
    step(i)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print(s.getvalue()[:5000])
