"""Live oracle-vs-reference comparison; runs only where /root/reference is mounted (the build container)."""
import pytest
import torch

from estdepth_b200 import synth
from oracle import estdepth_oracle as orc
from oracle.ref_loader import load_reference, reference_available
from tests.helpers import cfg_of

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not mounted")


def test_forward_is_bit_identical_to_the_reference():
    ref = load_reference()
    m = ref.model_hybrid.DepthNetHybrid(ndepths=32, depth_min=0.1, depth_max=10.0, resnet=18)
    sd = synth.synth_state_dict(m.state_dict(), seed=2)
    m.load_state_dict(sd)
    m.eval()
    cfg = cfg_of(18, 32)
    state = pstate = ostate = opstate = None
    with torch.no_grad():
        for start in (0, 3):
            imgs, poses, K, sample = synth.synth_inputs(5, 128, 160, seed=9, start=start)
            r_out, state, pstate = m(imgs, poses, K, sample, state, pstate, mode="val")
            o_out, ostate, opstate = orc.forward(sd, cfg, imgs, poses, K, ostate, opstate)
            for k, v in r_out.items():
                assert torch.equal(v, o_out[k]), (start, k)
            assert torch.equal(state["keys"][0], ostate["keys"][0]) and torch.equal(state["values"][0], ostate["values"][0])
            assert torch.equal(pstate[0], opstate[0])


def test_state_dict_names_match_the_reference():
    from estdepth_b200 import DepthNetHybrid
    ref = load_reference()
    for resnet, d in ((18, 32), (50, 64)):
        r = ref.model_hybrid.DepthNetHybrid(ndepths=d, depth_min=0.1, depth_max=10.0, resnet=resnet).state_dict()
        o = DepthNetHybrid(ndepths=d, depth_min=0.1, depth_max=10.0, resnet=resnet).state_dict()
        assert list(r.keys()) == list(o.keys())
        assert all(r[k].shape == o[k].shape and r[k].dtype == o[k].dtype for k in r)
