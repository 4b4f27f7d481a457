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


def test_frame_id_feature_cache_computes_each_frame_once():
    """Host logic of DepthNetHybrid._matching_features, on the CPU path of the matching-feature net:
    with frame ids, the frames two consecutive windows share are not recomputed, the features equal the uncached ones, and
    the cache is bounded and dropped when the parameters change."""
    import torch
    from tests.helpers import synth_model_and_state
    model, sd = synth_model_and_state(18, 32)
    g = torch.Generator().manual_seed(0)
    frames = torch.rand(8, 3, 128, 160, generator=g) * 2 - 1
    batches = []
    inner = model.matchingFeature.forward
    model.matchingFeature.forward = lambda x: (batches.append(x.shape[0]), inner(x))[1]
    with torch.no_grad():
        w1 = model._matching_features(frames[0:5].unsqueeze(0), frame_ids=[0, 1, 2, 3, 4])
        w2 = model._matching_features(frames[3:8].unsqueeze(0), frame_ids=[3, 4, 5, 6, 7])           # Joint stride: 2 frames shared
        assert batches == [5, 3]
        plain = model._matching_features(frames[3:8].unsqueeze(0))
        assert batches == [5, 3, 5] and tuple(w2.shape) == tuple(plain.shape) == (1, 5, 32, 32, 40)
        assert float((w2 - plain).abs().max()) < 1e-5 * float(plain.abs().max())
        assert torch.equal(w1[0, 3:5], w2[0, 0:2])
        # a frame shown twice in one call is computed once; B sequences of ids; bounded size
        model.feature_cache_size = 4
        two = model._matching_features(torch.stack([frames[0:3], frames[[0, 0, 2]]]), frame_ids=[["a", "b", "c"], ["a", "a", "c"]])
        assert batches[-1] == 5 and torch.equal(two[1, 0], two[1, 1]) and len(model._feat_cache) == 6
        model._matching_features(frames[0:3].unsqueeze(0), frame_ids=[10, 11, 12])
        assert len(model._feat_cache) == 4
        import pytest
        with pytest.raises(ValueError):
            model._matching_features(frames[0:3].unsqueeze(0), frame_ids=[1, 2])
        model.load_state_dict(sd)
        assert len(model._feat_cache) == 0


def test_forward_and_constructor_signatures_are_drop_in():
    """The boundary of SURVEY.md 8b: positional order and names of the reference's constructor
    (hybrid_models/model_hybrid.py:15-16) and forward (:110); extensions may only follow them, with defaults."""
    import inspect
    from estdepth_b200 import DepthNetHybrid
    ctor = list(inspect.signature(DepthNetHybrid.__init__).parameters.values())[1:]
    assert [(p.name, p.default) for p in ctor[:5]] == [("ndepths", 64), ("depth_min", 0.01), ("depth_max", 10.0), ("resnet", 50),
                                                       ("IF_EST_transformer", True)]
    assert all(p.default is not inspect.Parameter.empty for p in ctor[5:])
    fwd = list(inspect.signature(DepthNetHybrid.forward).parameters.values())[1:]
    assert [p.name for p in fwd[:7]] == ["imgs", "cam_poses", "cam_intr", "sample", "pre_costs", "pre_cam_poses", "mode"]
    assert fwd[4].default is None and fwd[5].default is None and fwd[6].default == "train"
    assert all(p.default is not inspect.Parameter.empty for p in fwd[7:])
    # the shim the eval drivers import (eval_hybrid.py:11, eval_hybrid_seq.py:10)
    from hybrid_models.model_hybrid import DepthNetHybrid as Shim
    assert Shim is DepthNetHybrid


def test_direct_stem_kernels_are_used_for_torchvision_stems_only():
    """encoders.ContextEncoder._plain_stem: the library's 7x7/2 convolution and 3x3/2 max-pool kernels implement exactly
    torchvision's stem (resnet_encoder.py:40-51 runs encoder.conv1 / bn1 / relu / maxpool); anything else keeps the cuDNN path."""
    from estdepth_b200.encoders import ContextEncoder
    for layers in (18, 50):
        enc = ContextEncoder(layers)
        assert ContextEncoder._plain_stem(enc.encoder)
    enc = ContextEncoder(18)
    enc.encoder.maxpool = torch.nn.MaxPool2d(3, stride=2, padding=1, ceil_mode=True)
    assert not ContextEncoder._plain_stem(enc.encoder)
    enc = ContextEncoder(18)
    enc.encoder.conv1 = torch.nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=True)
    assert not ContextEncoder._plain_stem(enc.encoder)
    enc = ContextEncoder(18)
    enc.encoder.conv1 = torch.nn.Conv2d(3, 64, 7, stride=1, padding=3, bias=False)
    assert not ContextEncoder._plain_stem(enc.encoder)
