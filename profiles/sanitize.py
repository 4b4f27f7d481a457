"""What compute-sanitizer wraps: one small window pair (no-EST, then EST) through every kernel family of the library, without
the CPU oracle (128x160, D=32, ResNet-18 -- the smoke configuration), plus the exact-fp32 and output-stationary 3-D kernels.

    compute-sanitizer --tool memcheck  --log-file gpurun_out/memcheck.log  python profiles/sanitize.py
    compute-sanitizer --tool racecheck --log-file gpurun_out/racecheck.log python profiles/sanitize.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from estdepth_b200 import DepthNetHybrid, synth, _lib  # noqa: E402

dev = torch.device("cuda:0")
precisions = sys.argv[1:] or ["3xf16r2", "3xf16r", "3xf16", "fp32"]
for precision in precisions:
    model = DepthNetHybrid(ndepths=32, depth_min=0.1, depth_max=10.0, resnet=18, precision=precision)
    model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0))
    model.eval().to(dev)
    state = pstate = None
    n0 = _lib.launch_count()
    for start in (0, 3):
        imgs, poses, K, sample = synth.synth_inputs(5, 128, 160, seed=0, start=start)
        out, state, pstate = model(imgs.to(dev), poses.to(dev), K.to(dev), sample, state, pstate, mode="val")
    model.check()
    torch.cuda.synchronize()
    print("precision %-8s: %d library launches, depth range %.3f .. %.3f" % (
        precision, _lib.launch_count() - n0, float(out[("depth", 0, 0)].min()), float(out[("depth", 0, 0)].max())), flush=True)
    del model
