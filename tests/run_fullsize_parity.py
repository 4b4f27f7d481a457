"""Not collected by pytest (the CPU oracle needs ~1.5 min at this size; run on the GPU box):

    python tests/run_fullsize_parity.py [H W D]

Full-size parity (BASELINE config 2: 5 x 480 x 640, D = 64, ResNet-50): both Joint windows (no-EST, then EST with the
returned state) through this repository's model in several arithmetic configurations against the CPU oracle (bit-identical
to the reference on CPU, tests/test_oracle_vs_reference.py) on the same synthetic inputs and weights.  Gate: 1e-3 abs on
every depth map (north_star)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from estdepth_b200 import DepthNetHybrid, synth  # noqa: E402
from oracle import estdepth_oracle as orc  # noqa: E402

_a = [a for a in sys.argv[1:] if not a.startswith("--")]
H, W, D = (int(_a[0]), int(_a[1]), int(_a[2])) if len(_a) >= 3 else (480, 640, 64)
V, resnet = 5, 50
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
tmpl = DepthNetHybrid(ndepths=D, depth_min=0.1, depth_max=10.0, resnet=resnet)
sd = synth.synth_state_dict(tmpl.state_dict(), seed=0)
cfg = dict(ndepths=D, depth_min=0.1, depth_max=10.0, resnet=resnet, est=True)
w1 = synth.synth_inputs(V, H, W, seed=0, start=0)[:3]
w2 = synth.synth_inputs(V, H, W, seed=0, start=V - 2)[:3]
torch.set_num_threads(os.cpu_count() or 1)
with torch.no_grad():
    ref1, cstate, cpose = orc.forward(sd, cfg, w1[0], w1[1], w1[2], None, None)
    ref2, cstate2, _ = orc.forward(sd, cfg, w2[0], w2[1], w2[2], cstate, cpose)
print("CPU oracle done (%dx%d, D=%d)" % (H, W, D))

configs = [("3xf16r2", "3xf16"), ("fp32", "fp32")] if "--all" not in sys.argv else [("3xf16r2", "3xf16"), ("3xf16r", "3xf16"), ("3xf16", "3xf16"), ("fp32", "3xf16"), ("fp32", "fp32"), ("3xf16r2", "fp32")]
worst_default = 0.0
for prec, feat in configs:
    model = DepthNetHybrid(ndepths=D, depth_min=0.1, depth_max=10.0, resnet=resnet, precision=prec, feature_precision=feat)
    model.load_state_dict(sd)
    model.eval().to(dev)
    with torch.no_grad():
        # camera parameters stay on the host: the model then derives the warps' matrices with the reference's torch ops on
        # the host, bit-identical to the oracle's (with CUDA poses the reference's own GPU LU decides, see --cuda-poses)
        mv = (lambda t: t.to(dev)) if "--cuda-poses" in sys.argv else (lambda t: t)
        out1, state, pstate = model(w1[0].to(dev), mv(w1[1]), mv(w1[2]), None, mode="val")
        out2, state2, _ = model(w2[0].to(dev), mv(w2[1]), mv(w2[2]), None, state, pstate, mode="val")
    line = {}
    for wname, out, ref in (("w1", out1, ref1), ("w2", out2, ref2)):
        for k in out:
            tag = "%s/%s" % (wname, "depth%d" % k[2] if k[0] == "depth" else k[0])
            line[tag] = max(line.get(tag, 0.0), float((out[k].cpu() - ref[k]).abs().max()))
    line["w2/state_value"] = float((state2["values"][0].cpu() - cstate2["values"][0]).abs().max())
    line["w2/state_key_rel"] = float((state2["keys"][0].cpu() - cstate2["keys"][0]).abs().max() / cstate2["keys"][0].abs().max())
    depth_worst = max(v for k, v in line.items() if "depth" in k)
    print("precision=%-8s feature_precision=%-6s worst depth %.1e | %s" % (prec, feat, depth_worst, {k: "%.1e" % v for k, v in sorted(line.items())}))
    if (prec, feat) == configs[0]:
        worst_default = depth_worst
        # where do the differences sit?  (coordinates within an ulp of the [-1, 1] sampling range flip between "sampled" and
        # "zero-filled", quirk Q10: such voxels are isolated; anything else would be a region)
        ev = (state2["values"][0].cpu() - cstate2["values"][0]).abs().amax(dim=(0, 1))          # [D, H, W]
        for thr in (1e-3, 1e-2):
            idx = (ev > thr).nonzero()
            print("  state_value: %d of %d voxels differ by > %.0e" % (idx.shape[0], ev.numel(), thr), end="")
            if idx.shape[0]:
                print("; d in [%d, %d], h in [%d, %d], w in [%d, %d]; first: %s" % (
                    idx[:, 0].min(), idx[:, 0].max(), idx[:, 1].min(), idx[:, 1].max(), idx[:, 2].min(), idx[:, 2].max(), idx[:8].tolist()))
            else:
                print()
        for key in (("depth", 0, 0), ("depth", 2, 0), ("depth", 0, 2)):
            ed = (out2[key].cpu() - ref2[key]).abs()[0, 0]
            idx = (ed > 1e-3).nonzero()
            print("  %s: %d of %d pixels differ by > 1e-3%s" % (key, idx.shape[0], ed.numel(),
                  ("; h in [%d, %d], w in [%d, %d]" % (idx[:, 0].min(), idx[:, 0].max(), idx[:, 1].min(), idx[:, 1].max())) if idx.shape[0] else ""))
    del model
print("default configuration: worst |depth - oracle| = %.2e (%s the 1e-3 gate)" % (worst_default, "within" if worst_default < 1e-3 else "ABOVE"))
