"""Which 3-D layers' arithmetic decides the depth error?  cfg2 against the benchmark-size reference fixtures (head gain 3 and
10) with groups of layers switched from the plane-ring schedule (162 truncating accumulates per accumulator) to the
output-stationary kernel (54).  Run on a B200:  python profiles/layer_precision_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from estdepth_b200 import DepthNetHybrid, synth  # noqa: E402
from oracle.make_golden import subsample  # noqa: E402

dev = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
GROUPS = {
    "pre": ["pre1", "pre2", "pre2_pair"], "dres01": ["dres0.0", "dres0.1", "dres1.0", "dres1.1"], "dres2": ["dres2"],
    "value_key": ["value_key"], "heads": ["head0", "head1"], "gru": ["gate", "output"],
}
VARIANTS = [[], ["heads"], ["heads", "gru"], ["heads", "gru", "value_key"], ["pre"], ["dres01"], ["dres2", "value_key"],
            ["pre", "dres01", "dres2", "value_key", "heads", "gru"]]
with torch.no_grad():
    for gain in (3.0, 10.0):
        gold = np.load(os.path.join(GOLD, "joint_r50_d64_480x640_g%d.npz" % gain))
        for variant in VARIANTS:
            model = DepthNetHybrid(ndepths=64, depth_min=0.1, depth_max=10.0, resnet=50)
            model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0, head_gain=gain))
            model.eval().to(dev)
            L = model._layers(torch.device(dev, 0))
            for grp in variant:
                for name in GROUPS[grp]:
                    L[name].precision = "3xf16"
            state = pstate = None
            worst = {}
            for w, start in enumerate((0, 3)):
                imgs, poses, K, sample = synth.synth_inputs(5, 480, 640, seed=0, start=start)
                out, state, pstate = model(imgs.to(dev), poses, K, sample, state, pstate, mode="val")
                for key, val in out.items():
                    gk = gold["w%d/%s" % (w, "_".join(str(k) for k in key))]
                    d = float(np.abs(subsample(key, val, 4).cpu().numpy() - gk).max())
                    tag = "depth%d" % key[2] if key[0] == "depth" else key[0]
                    worst[tag] = max(worst.get(tag, 0.0), d)
            print("gain %4.1f  output-stationary: %-48s %s" % (gain, "+".join(variant) or "(none: all ring2)", {k: "%.2e" % v for k, v in sorted(worst.items()) if k.startswith("depth")}), flush=True)
            del model
            torch.cuda.empty_cache()
