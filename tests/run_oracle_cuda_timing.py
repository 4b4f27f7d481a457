"""Not collected by pytest.  How fast is the REFERENCE ALGORITHM as plain PyTorch ops on the same GPU?

The reference itself (xxlong0/ESTDepth) cannot travel to the GPU box, so this runs its CPU restatement
(oracle/estdepth_oracle.py: the same sequence of torch ops, bit-identical to the reference on CPU) with every tensor on
cuda:0 -- i.e. the reference's PyTorch-CUDA eval path (cuDNN convolutions, ATen grid_sample) -- on the bench workload
(cfg2 steady-state Joint window), once as shipped (cudnn.allow_tf32=True, cudnn.benchmark=True, eval_hybrid.py:13) and once
in strict fp32, next to this repository's model on the same inputs.  north_star asks for >= 10x.

    python tests/run_oracle_cuda_timing.py [cfg2|cfg1]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from estdepth_b200 import DepthNetHybrid, synth  # noqa: E402
from oracle import estdepth_oracle as orc  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
V, H, W, D, resnet = {"cfg2": (5, 480, 640, 64, 50), "cfg1": (5, 128, 160, 32, 18)}[workload]
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
model = DepthNetHybrid(ndepths=D, depth_min=0.1, depth_max=10.0, resnet=resnet)
sd_cpu = synth.synth_state_dict(model.state_dict(), seed=0)
model.load_state_dict(sd_cpu)
model.eval().to(dev)

sd = {k: v.to(dev) for k, v in sd_cpu.items()}
cfg = dict(ndepths=D, depth_min=0.1, depth_max=10.0, resnet=resnet, est=True)
w1 = [t.to(dev) for t in synth.synth_inputs(V, H, W, seed=0, start=0)[:3]]
w2 = [t.to(dev) for t in synth.synth_inputs(V, H, W, seed=0, start=V - 2)[:3]]
torch.set_default_device(dev)        # the oracle's factory calls (pixel grids, plane depths) follow the inputs to the GPU


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


results = {}
with torch.no_grad():
    for name, tf32 in (("as shipped (cudnn TF32 allowed)", True), ("strict fp32", False)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        _, ostate, opose = orc.forward(sd, cfg, w1[0], w1[1], w1[2], None, None)
        ms = timed(lambda: orc.forward(sd, cfg, w2[0], w2[1], w2[2], ostate, opose))
        results[name] = ms
        print("reference algorithm as PyTorch-CUDA ops, %s: %.1f ms per window = %.1f depth frames/s" % (name, ms, (V - 2) / ms * 1e3))
    torch.backends.cudnn.allow_tf32 = False
    _, state, pstate = model(w1[0], w1[1], w1[2], None, mode="val")
    ms = timed(lambda: model(w2[0], w2[1], w2[2], None, state, pstate, mode="val"), n=10)
    print("estdepth_b200: %.1f ms per window = %.1f depth frames/s  (%.1fx as shipped, %.1fx strict fp32)"
          % (ms, (V - 2) / ms * 1e3, results["as shipped (cudnn TF32 allowed)"] / ms, results["strict fp32"] / ms))
    want, _, _ = orc.forward(sd, cfg, w2[0], w2[1], w2[2], ostate, opose)
    out, _, _ = model(w2[0], w2[1], w2[2], None, state, pstate, mode="val")
    print("max |depth - PyTorch-CUDA strict fp32| over the window: %.2e"
          % max(float((out[k] - want[k]).abs().max()) for k in out if k[0] == "depth"))

    # which of the two GPU paths is closer to the reference's CPU arithmetic (the pinned oracle) at this size?
    torch.set_default_device("cpu")
    torch.set_num_threads(os.cpu_count() or 1)
    c1 = [t.cpu() for t in w1]
    c2 = [t.cpu() for t in w2]
    _, cstate, cpose = orc.forward(sd_cpu, cfg, c1[0], c1[1], c1[2], None, None)
    cpu, _, _ = orc.forward(sd_cpu, cfg, c2[0], c2[1], c2[2], cstate, cpose)
    for name, res in (("estdepth_b200", out), ("PyTorch-CUDA strict fp32", want)):
        worst = {}
        for k in res:
            d = (res[k].cpu() - cpu[k]).abs()
            tag = "depth%d" % k[2] if k[0] == "depth" else k[0]
            worst[tag] = max(worst.get(tag, 0.0), float(d.max()))
        print("%s vs CPU oracle (window 2, %dx%d D=%d): %s" % (name, H, W, D, {k: "%.1e" % v for k, v in sorted(worst.items())}))
