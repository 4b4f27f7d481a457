"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE.  Usage (from the repo root, where /root/reference is mounted):

    python -m oracle.make_golden

The reference ships no golden vectors, so these fixtures ARE the reference's
outputs: its own ``DepthNetHybrid.forward(..., mode='val')``, ``homo_warping``, ``warp_volume``,
``EpipolarTransformer`` and ``depthlayer`` executed on CPU fp32 (torch 2.11) on seeded synthetic inputs
and seeded synthetic weights (``estdepth_b200.synth``).  Inputs and weights are NOT stored -- tests
regenerate them from the same seeds -- only the reference's outputs are.

    python -m oracle.make_golden --fullsize     # the benchmark-size fixtures below (about 10 min of CPU)

Fixtures:
  joint_r18_d32_128x160.npz   cfg1: 5-frame Joint windows 1 (no EST, quirk Q3) and 2 (EST, pre_num=1)
  joint_r50_d64_128x128.npz   same protocol, ResNet-50 / D=64 (the architecture of cfg2, small image)
  estm_r18_d32_128x160.npz    ESTM protocol: 5 sliding 3-frame calls, memory_size=2 (quirks Q4, Q5)
  ops_small.npz               op-level outputs on small tensors
Benchmark-size fixtures (``--fullsize``; outputs stored SUB-SAMPLED, see ``subsample``):
  joint_r50_d64_480x640_g3.npz    BASELINE cfg2: both Joint windows at 5 x 480 x 640, D=64, ResNet-50, head gain 3 (synth default)
  joint_r50_d64_480x640_g10.npz   the same with the logit heads scaled by 10 (logit sigma ~ 3)
  estm_r50_d64_480x640.npz        BASELINE cfg3 (first 4 steps): ESTM protocol, 3-frame windows, memory 2, 480 x 640
  joint_r50_d128_640x960.npz      BASELINE cfg5: both Joint windows at 5 x 640 x 960, D=128
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference  # noqa: E402
from estdepth_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
STATE_STRIDE = 4     # hidden-state tensors are stored sub-sampled by this stride in h and w


def _np(t):
    return t.detach().cpu().numpy()


def build_reference_model(ref, resnet, ndepths, seed=0):
    m = ref.model_hybrid.DepthNetHybrid(ndepths=ndepths, depth_min=0.1, depth_max=10.0, resnet=resnet,
                                        IF_EST_transformer=True)
    m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=seed))
    m.eval()
    return m


def joint_fixture(ref, resnet, ndepths, height, width, name):
    m = build_reference_model(ref, resnet, ndepths)
    out = {}
    state, poses_state = None, None
    with torch.no_grad():
        for w, start in enumerate((0, 3)):                 # windows of 5 frames, stride 3 (eval_hybrid.py:195-196)
            imgs, poses, K, sample = synth.synth_inputs(5, height, width, seed=0, start=start)
            outputs, state, poses_state = m(imgs, poses, K, sample, state, poses_state, mode="val")
            for key, val in outputs.items():
                out["w%d/%s" % (w, "_".join(str(k) for k in key))] = _np(val)
            out["w%d/state_key" % w] = _np(state["keys"][0][..., ::STATE_STRIDE, ::STATE_STRIDE])
            out["w%d/state_value" % w] = _np(state["values"][0][..., ::STATE_STRIDE, ::STATE_STRIDE])
            out["w%d/state_pose" % w] = _np(poses_state[0])
    np.savez_compressed(os.path.join(GOLDEN_DIR, name), **out)
    return out


def estm_fixture(ref, resnet, ndepths, height, width, name, frames=7, memory=2):
    """The sliding-window protocol of eval_hybrid_seq.py:169-193 (lwindow=3, memory_size=2)."""
    m = build_reference_model(ref, resnet, ndepths)
    out = {}
    mem_costs, mem_poses = [], []
    with torch.no_grad():
        for step in range(frames - 2):
            imgs, poses, K, sample = synth.synth_inputs(3, height, width, seed=0, start=step)
            if mem_poses:
                pre_costs = {"keys": [c["keys"][0] for c in mem_costs], "values": [c["values"][0] for c in mem_costs]}
                pre_poses = [p[0] for p in mem_poses]
            else:
                pre_costs, pre_poses = None, None
            outputs, costs, cposes = m(imgs, poses, K, sample, pre_costs, pre_poses, mode="val")
            mem_costs.append(costs)
            mem_poses.append(cposes)
            if len(mem_costs) > memory:
                mem_costs.pop(0)
                mem_poses.pop(0)
            out["s%d/depth_0_2" % step] = _np(outputs[("depth", 0, 2)])
            out["s%d/depth_0_0" % step] = _np(outputs[("depth", 0, 0)])
            out["s%d/depth_0_3" % step] = _np(outputs[("depth", 0, 3)])
            out["s%d/state_pose" % step] = _np(cposes[0])
    np.savez_compressed(os.path.join(GOLDEN_DIR, name), **out)
    return out


FULL_STRIDE = 4          # benchmark-size fixtures keep every 4th row / column of a map ...
FULL_STATE_STRIDE = 16   # ... and every 16th of the hidden-state volumes


def subsample(key, arr, stride=FULL_STRIDE):
    """Which pixels of an output map a benchmark-size fixture keeps (shared by the generator and tests/test_gpu_fullsize_golden.py).

    Works on torch tensors and numpy arrays [..., H, W].  ("depth", t, 3|2) and the probability maps are nearest x4
    replications of quarter-resolution maps (hybrid_depth_decoder.py:202-209, 259-260): offset (0, 0) with stride 4 keeps
    every quarter-resolution pixel once, i.e. those maps are stored WITHOUT loss.  ("depth", t, 1) is a x2 replication of a
    half-resolution map: offset (2, 0) alternates between its even and odd rows' sources.  ("depth", t, 0) is a genuine
    full-resolution map: offset (1, 2)."""
    oy, ox = {1: (2, 0), 0: (1, 2)}.get(key[2], (0, 0)) if key[0] == "depth" else (0, 0)
    return arr[..., oy::stride, ox::stride]


def fullsize_joint_fixture(ref, resnet, ndepths, height, width, name, head_gain=synth.HEAD_GAIN, stride=FULL_STRIDE):
    m = ref.model_hybrid.DepthNetHybrid(ndepths=ndepths, depth_min=0.1, depth_max=10.0, resnet=resnet, IF_EST_transformer=True)
    m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=0, head_gain=head_gain))
    m.eval()
    out = {"meta": np.array([height, width, ndepths, resnet, stride, FULL_STATE_STRIDE, head_gain], dtype=np.float64)}
    state, poses_state = None, None
    with torch.no_grad():
        for w, start in enumerate((0, 3)):
            imgs, poses, K, sample = synth.synth_inputs(5, height, width, seed=0, start=start)
            outputs, state, poses_state = m(imgs, poses, K, sample, state, poses_state, mode="val")
            for key, val in outputs.items():
                out["w%d/%s" % (w, "_".join(str(k) for k in key))] = _np(subsample(key, val, stride)).copy()
            out["w%d/state_key" % w] = _np(state["keys"][0][..., ::FULL_STATE_STRIDE, ::FULL_STATE_STRIDE]).copy()
            out["w%d/state_value" % w] = _np(state["values"][0][..., ::FULL_STATE_STRIDE, ::FULL_STATE_STRIDE]).copy()
            out["w%d/state_pose" % w] = _np(poses_state[0])
            print(name, "window", w, "done", flush=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name), **out)


def fullsize_estm_fixture(ref, resnet, ndepths, height, width, name, steps=4, memory=2, stride=FULL_STRIDE):
    """First ``steps`` calls of BASELINE cfg3 (eval_hybrid_seq.py:169-193: 3-frame windows, memory_size 2)."""
    m = build_reference_model(ref, resnet, ndepths)
    out = {"meta": np.array([height, width, ndepths, resnet, stride, FULL_STATE_STRIDE, synth.HEAD_GAIN], dtype=np.float64)}
    mem_costs, mem_poses = [], []
    with torch.no_grad():
        for step in range(steps):
            imgs, poses, K, sample = synth.synth_inputs(3, height, width, seed=0, start=step)
            if mem_poses:
                pre_costs = {"keys": [c["keys"][0] for c in mem_costs], "values": [c["values"][0] for c in mem_costs]}
                pre_poses = [p[0] for p in mem_poses]
            else:
                pre_costs, pre_poses = None, None
            outputs, costs, cposes = m(imgs, poses, K, sample, pre_costs, pre_poses, mode="val")
            mem_costs.append(costs)
            mem_poses.append(cposes)
            if len(mem_costs) > memory:
                mem_costs.pop(0)
                mem_poses.pop(0)
            for key, val in outputs.items():
                out["s%d/%s" % (step, "_".join(str(k) for k in key))] = _np(subsample(key, val, stride)).copy()
            out["s%d/state_value" % step] = _np(costs["values"][0][..., ::FULL_STATE_STRIDE, ::FULL_STATE_STRIDE]).copy()
            out["s%d/state_pose" % step] = _np(cposes[0])
            print(name, "step", step, "done", flush=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name), **out)


def fullsize_main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = load_reference()
    fullsize_joint_fixture(ref, 50, 64, 480, 640, "joint_r50_d64_480x640_g3.npz")
    fullsize_joint_fixture(ref, 50, 64, 480, 640, "joint_r50_d64_480x640_g10.npz", head_gain=10.0)
    fullsize_estm_fixture(ref, 50, 64, 480, 640, "estm_r50_d64_480x640.npz")
    fullsize_joint_fixture(ref, 50, 128, 640, 960, "joint_r50_d128_640x960.npz", stride=8)
    for f in sorted(os.listdir(GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(GOLDEN_DIR, f)) // 1024, "KiB")


def ops_inputs(seed=0):
    """Seeded small inputs shared by make_golden and the tests (pure function)."""
    g = torch.Generator().manual_seed(1234 + seed)
    C, D, H, W = 8, 6, 20, 24
    fea = torch.randn(1, C, H, W, generator=g)
    vol = torch.randn(1, 4, D, H, W, generator=g)
    poses = synth.camera_track(3)
    K4 = synth.intrinsics(4 * H, 4 * W).unsqueeze(0).clone()
    K4[:, :2] *= 0.25
    depth_min, depth_max = 0.5, 4.0
    interval = (depth_max - depth_min) / (D - 1)
    depth_values = torch.arange(D, dtype=torch.float32) * interval + depth_min
    logits = 3.0 * torch.randn(2, D, 9, 11, generator=g)
    key_t = torch.relu(torch.randn(1, 16, D, 10, 12, generator=g))
    val_t = torch.tanh(torch.randn(1, 16, D, 10, 12, generator=g))
    wkeys = [torch.relu(torch.randn(1, 16, D, 10, 12, generator=g)) for _ in range(3)]
    wvals = [torch.tanh(torch.randn(1, 16, D, 10, 12, generator=g)) for _ in range(3)]
    return dict(fea=fea, vol=vol, poses=poses, K4=K4, depth_min=depth_min, interval=interval,
                depth_values=depth_values, logits=logits, key_t=key_t, val_t=val_t, wkeys=wkeys, wvals=wvals)


def ops_fixture(ref, name):
    x = ops_inputs()
    out = {}
    D = x["depth_values"].numel()
    _, C, H, W = x["fea"].shape
    with torch.no_grad():
        ext = torch.inverse(x["poses"]).unsqueeze(0)                       # [1,3,4,4]
        for s in (0, 2):
            sp, rp = ext[:, s].clone(), ext[:, 1].clone()
            sp[:, :3, :4] = x["K4"] @ ext[:, s, :3, :4]
            rp[:, :3, :4] = x["K4"] @ ext[:, 1, :3, :4]
            out["homo_warp_%d" % s] = _np(ref.homo.homo_warping(x["fea"], sp, rp, x["depth_values"].view(1, D, 1, 1)))
        grid = ref.homo.set_id_grid(H, W).view(1, 3, 1, H * W).repeat(1, 1, D, 1)
        dv = x["depth_values"].view(1, D, 1, 1).repeat(1, 1, H, W).view(1, 1, D, H * W)
        for j in (0, 2):
            rel = torch.matmul(x["poses"][j:j + 1], torch.inverse(x["poses"][1:2]))
            out["warp_volume_%d" % j] = _np(ref.homo.warp_volume(x["vol"], dv, rel, x["K4"], grid,
                                                                   x["depth_min"], x["interval"]))
        d, p = ref.decoder.depthlayer(x["logits"], x["depth_values"].view(1, D, 1, 1).repeat(1, 1, 9, 11))
        out["depthlayer_depth"], out["depthlayer_prob"] = _np(d), _np(p)
        est = ref.est.EpipolarTransformer(16, 16, 3)
        sd = synth.synth_state_dict(est.state_dict(), seed=3)
        est.load_state_dict(sd)
        est.eval()
        for n in (1, 2, 3):
            out["est_n%d" % n] = _np(est(target_key=x["key_t"], warped_keys=x["wkeys"][:n],
                                         target_value=x["val_t"], warped_values=x["wvals"][:n]))
    np.savez_compressed(os.path.join(GOLDEN_DIR, name), **out)
    return out


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = load_reference()
    ops_fixture(ref, "ops_small.npz")
    joint_fixture(ref, 18, 32, 128, 160, "joint_r18_d32_128x160.npz")
    joint_fixture(ref, 50, 64, 128, 128, "joint_r50_d64_128x128.npz")
    estm_fixture(ref, 18, 32, 128, 160, "estm_r18_d32_128x160.npz")
    for f in sorted(os.listdir(GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(GOLDEN_DIR, f)) // 1024, "KiB")


if __name__ == "__main__":
    if "--fullsize" in sys.argv:
        fullsize_main()
    else:
        main()
