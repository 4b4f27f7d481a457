"""The oracle against the reference's own outputs (tests/golden/, produced by oracle/make_golden.py from the
UNMODIFIED reference).  This is what pins parity: the reference ships no tests or vectors of its own."""
import os

import numpy as np
import pytest
import torch

from estdepth_b200 import synth
from oracle import estdepth_oracle as orc
from oracle.make_golden import ops_inputs
from tests.helpers import cfg_of, state_template

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
# generated on this image's CPU; another host CPU may take other oneDNN code paths -> allow fp32 round-off
TOL = 2e-4


def _sd(resnet, ndepths):
    _, tmpl = state_template(resnet, ndepths)
    return synth.synth_state_dict(tmpl, seed=0)


@pytest.mark.parametrize("resnet,ndepths,height,width,name", [
    (18, 32, 128, 160, "joint_r18_d32_128x160.npz"),
    (50, 64, 128, 128, "joint_r50_d64_128x128.npz"),
])
def test_joint_windows(resnet, ndepths, height, width, name):
    gold = np.load(os.path.join(GOLDEN, name))
    sd, cfg = _sd(resnet, ndepths), cfg_of(resnet, ndepths)
    state = pstate = None
    with torch.no_grad():
        for w, start in enumerate((0, 3)):
            imgs, poses, K, _ = synth.synth_inputs(5, height, width, seed=0, start=start)
            outputs, state, pstate = orc.forward(sd, cfg, imgs, poses, K, state, pstate)
            n = 0
            for key, val in outputs.items():
                name_ = "w%d/%s" % (w, "_".join(str(k) for k in key))
                if name_ not in gold:
                    continue            # argmax maps are an oracle extra (the reference discards the index)
                assert np.abs(val.numpy() - gold[name_]).max() < TOL, name_
                n += 1
            assert n == 18
            assert np.abs(state["keys"][0][..., ::4, ::4].numpy() - gold["w%d/state_key" % w]).max() < TOL
            assert np.abs(state["values"][0][..., ::4, ::4].numpy() - gold["w%d/state_value" % w]).max() < TOL
            assert np.array_equal(pstate[0].numpy(), gold["w%d/state_pose" % w])
    # quirk Q4: window 2 hands back window 1's pose
    assert np.array_equal(gold["w0/state_pose"], gold["w1/state_pose"])


def test_estm_protocol():
    gold = np.load(os.path.join(GOLDEN, "estm_r18_d32_128x160.npz"))
    sd, cfg = _sd(18, 32), cfg_of(18, 32)
    mem = []
    with torch.no_grad():
        for step in range(5):
            imgs, poses, K, _ = synth.synth_inputs(3, 128, 160, seed=0, start=step)
            pre = ({"keys": [c["keys"][0] for c, _ in mem], "values": [c["values"][0] for c, _ in mem]},
                   [p[0] for _, p in mem]) if mem else (None, None)
            outputs, costs, cposes = orc.forward(sd, cfg, imgs, poses, K, pre[0], pre[1])
            mem.append((costs, cposes))
            if len(mem) > 2:
                mem.pop(0)
            for s in (2, 0, 3):
                assert np.abs(outputs[("depth", 0, s)].numpy() - gold["s%d/depth_0_%d" % (step, s)]).max() < TOL
            assert np.array_equal(cposes[0].numpy(), gold["s%d/state_pose" % step])
    # Q4: the memory pose never advances past the first target's pose
    assert all(np.array_equal(gold["s0/state_pose"], gold["s%d/state_pose" % s]) for s in range(5))


@pytest.mark.parametrize("sampler", ["aten", "explicit"])
def test_ops(sampler):
    gold = np.load(os.path.join(GOLDEN, "ops_small.npz"))
    x = ops_inputs()
    tol = 1e-6 if sampler == "aten" else 2e-4          # explicit = closed-form tap arithmetic (coordinate round-off)
    ext = torch.inverse(x["poses"]).unsqueeze(0)
    for s in (0, 2):
        sp, rp = ext[:, s].clone(), ext[:, 1].clone()
        sp[:, :3, :4] = x["K4"] @ ext[:, s, :3, :4]
        rp[:, :3, :4] = x["K4"] @ ext[:, 1, :3, :4]
        got = orc.homo_warp(x["fea"], sp, rp, x["depth_values"], sampler)
        assert np.abs(got.numpy() - gold["homo_warp_%d" % s]).max() < tol
    for j in (0, 2):
        rel = torch.matmul(x["poses"][j:j + 1], torch.inverse(x["poses"][1:2]))
        got = orc.warp_volume(x["vol"], rel, x["K4"], x["depth_values"], x["depth_min"], x["interval"], sampler)
        assert np.abs(got.numpy() - gold["warp_volume_%d" % j]).max() < tol
    d, p, idx = orc.soft_argmin(x["logits"], x["depth_values"], up=1)
    assert np.abs(d.numpy() - gold["depthlayer_depth"]).max() < 1e-6
    assert np.abs(p.numpy() - gold["depthlayer_prob"]).max() < 1e-6
    assert torch.equal(idx, torch.softmax(x["logits"], 1).argmax(1, keepdim=True))
    tmpl = {k: torch.empty(s) for k, s in (("gate_conv.weight", (32, 32, 3, 3, 3)), ("gate_conv.bias", (32,)),
            ("reset_gate_norm.weight", (16,)), ("reset_gate_norm.bias", (16,)), ("update_gate_norm.weight", (16,)),
            ("update_gate_norm.bias", (16,)), ("output_conv.weight", (16, 32, 3, 3, 3)), ("output_conv.bias", (16,)),
            ("output_norm.weight", (16,)), ("output_norm.bias", (16,)))}
    sd = {"CostRegNet.epipolar_transformer." + k: v for k, v in synth.synth_state_dict(tmpl, seed=3).items()}
    for n in (1, 2, 3):
        got = orc.est_fuse(sd, x["key_t"], x["wkeys"][:n], x["val_t"], x["wvals"][:n])
        assert np.abs(got.numpy() - gold["est_n%d" % n]).max() < 1e-5


def test_attention_is_mean_not_sum():
    """Quirk Q6: N identical sources get softmax weight 1/N each and torch.mean divides by N again... no -- the
    weighted values are averaged, so identical sources give value * (1/N) * N / N = value / N."""
    x = ops_inputs()
    v = x["wvals"][0]
    for n in (1, 2, 3):
        h = orc.est_attention(x["key_t"], [x["wkeys"][0]] * n, [v] * n)
        assert torch.allclose(h, v / n, atol=1e-6)


def test_benchmark_size_first_window_cfg2():
    """The oracle against the reference's outputs AT THE BENCHMARKED SIZE (5 x 480 x 640, D=64, ResNet-50; fixture made by
    ``python -m oracle.make_golden --fullsize``): window 1 (the no-EST path, quirk Q3), about 15 s of CPU.  The second window and
    the other benchmark-size fixtures (cfg3, cfg5, head gain 10) are checked against the CUDA path in tests/test_gpu_fullsize_golden.py."""
    from oracle.make_golden import FULL_STATE_STRIDE, subsample
    gold = np.load(os.path.join(GOLDEN, "joint_r50_d64_480x640_g3.npz"))
    stride = int(gold["meta"][4])
    sd, cfg = _sd(50, 64), cfg_of(50, 64)
    imgs, poses, K, _ = synth.synth_inputs(5, 480, 640, seed=0, start=0)
    with torch.no_grad():
        outputs, state, pstate = orc.forward(sd, cfg, imgs, poses, K)
    n = 0
    for key, val in outputs.items():
        name_ = "w0/%s" % "_".join(str(k) for k in key)
        if name_ not in gold:
            continue
        assert np.abs(subsample(key, val, stride).numpy() - gold[name_]).max() < TOL, name_
        n += 1
    assert n == 18
    s = FULL_STATE_STRIDE
    assert np.abs(state["values"][0][..., ::s, ::s].numpy() - gold["w0/state_value"]).max() < TOL
    assert np.array_equal(pstate[0].numpy(), gold["w0/state_pose"])
