"""Shared test helpers (seeded models / inputs; no reference access)."""
import torch

from estdepth_b200 import synth


def state_template(resnet, ndepths, est=True):
    """Key names/shapes of the reference state dict, taken from the product module (identical by construction;
    tests/test_state_dict.py pins the key list against a committed fixture generated from the reference)."""
    from estdepth_b200.model import DepthNetHybrid
    m = DepthNetHybrid(ndepths=ndepths, depth_min=0.1, depth_max=10.0, resnet=resnet, IF_EST_transformer=est)
    return m, m.state_dict()


def synth_model_and_state(resnet, ndepths, seed=0):
    m, template = state_template(resnet, ndepths)
    sd = synth.synth_state_dict(template, seed=seed)
    m.load_state_dict(sd)
    m.eval()
    return m, sd


def cfg_of(resnet, ndepths):
    return dict(ndepths=ndepths, depth_min=0.1, depth_max=10.0, resnet=resnet, est=True)


def to_vol4(x):
    """[C,D,H,W] -> vol4 [C/4,D,H,W,4] (pure torch, for building test inputs)."""
    C, D, H, W = x.shape
    return x.reshape(C // 4, 4, D, H, W).permute(0, 2, 3, 4, 1).contiguous()


def from_vol4(v):
    ch, D, H, W, _ = v.shape
    return v.permute(0, 4, 1, 2, 3).reshape(ch * 4, D, H, W).contiguous()


def to_map4(x):
    C, H, W = x.shape
    return x.reshape(C // 4, 4, H, W).permute(0, 2, 3, 1).contiguous()
