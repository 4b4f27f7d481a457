"""state_dict compatibility with checkpoints written by the reference (eval_hybrid.py:328-333 loads strictly)."""
import json
import os

import pytest
import torch

from estdepth_b200 import DepthNetHybrid, synth

FIXTURE = os.path.join(os.path.dirname(__file__), "golden", "state_dict_keys.json")


@pytest.mark.parametrize("tag,resnet,ndepths", [("r18_d32", 18, 32), ("r50_d64", 50, 64)])
def test_keys_and_shapes_equal_the_reference_fixture(tag, resnet, ndepths):
    want = json.load(open(FIXTURE))[tag]                  # dumped from the reference's own module
    sd = DepthNetHybrid(ndepths=ndepths, depth_min=0.1, depth_max=10.0, resnet=resnet).state_dict()
    assert [k for k, _ in want] == list(sd.keys())
    assert all(list(sd[k].shape) == shape for k, shape in want)


def test_strict_load_and_no_est_variant():
    m = DepthNetHybrid(ndepths=32, depth_min=0.1, depth_max=10.0, resnet=18)
    sd = synth.synth_state_dict(m.state_dict(), seed=0)
    m.load_state_dict(sd, strict=True)
    assert torch.equal(m.pre0[0].weight, sd["pre0.0.weight"])
    plain = DepthNetHybrid(ndepths=32, depth_min=0.1, depth_max=10.0, resnet=18, IF_EST_transformer=False)
    assert not any("epipolar_transformer" in k for k in plain.state_dict())


def test_synthetic_weights_are_deterministic():
    m = DepthNetHybrid(ndepths=32, depth_min=0.1, depth_max=10.0, resnet=18)
    a = synth.synth_state_dict(m.state_dict(), seed=0)
    b = synth.synth_state_dict(m.state_dict(), seed=0)
    c = synth.synth_state_dict(m.state_dict(), seed=1)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert not torch.equal(a["pre1.0.weight"], c["pre1.0.weight"])
