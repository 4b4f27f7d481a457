"""GPU check of the opt-in cross-window feature cache (not collected by pytest; run on a B200):

    python tests/run_frame_cache_gpu.py

Drives the ESTM protocol (3-frame sliding windows, 2-deep memory, eval_hybrid_seq.py:169-193) twice over the same
synthetic clip at 480x640 / D=64 / R50 -- without and with ``frame_ids`` -- and reports the largest difference of the
depth maps (the cached features come from a differently composed batch: cuDNN may pick another stem algorithm, the
in-house kernels are batch-invariant) and the time per step of both."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from estdepth_b200 import DepthNetHybrid, sharding, synth  # noqa: E402

torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
H, W, D, N_FRAMES = 480, 640, 64, 12
model = DepthNetHybrid(ndepths=D, depth_min=0.1, depth_max=10.0, resnet=50)
model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0))
model.eval().to(dev)
windows = [synth.synth_inputs(3, H, W, seed=0, start=s) for s in range(N_FRAMES - 2)]
windows = [(w[0].to(dev), w[1], w[2]) for w in windows]


def run(with_ids):
    mem, maps = [], []
    model._feat_cache.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s, (imgs, poses, K) in enumerate(windows):
        pre = sharding._flatten_memory(mem)
        out, costs, cposes = model(imgs, poses, K, None, pre[0], pre[1], mode="val",
                                   frame_ids=[s, s + 1, s + 2] if with_ids else None)
        mem.append((costs, cposes))
        if len(mem) > 2:
            mem.pop(0)
        maps.append(out[("depth", 0, 0)])
    torch.cuda.synchronize()
    return torch.cat(maps), (time.perf_counter() - t0) / len(windows) * 1e3


run(False)                                         # warm-up (cuDNN autotuning, workspace)
plain, ms_plain = run(False)
cached, ms_cached = run(True)
diff = (plain - cached).abs().max().item()
print("ESTM %d steps at %dx%d D=%d: %.2f ms/step without ids, %.2f ms/step with frame_ids; max |depth diff| = %.3e"
      % (len(windows), H, W, D, ms_plain, ms_cached, diff))
assert diff < 1e-4, diff
print("OK")
