"""Full-size (BASELINE config 2: 120x160 quarter-res, D=64) checks through size-independent properties -- the CPU oracle
needs ~10 s per step at this size, so these tests use identities that hold for any input instead:

  K1  identity homography (same pose, corner-aligned sampling -- with the default align_corners=False the reference's
      normalisation makes even the identity pose resample at x*W/(W-1)-0.5, quirk Q1): x0 == ref_mix + src_mix;
      linearity in the source map
  K2  delta filter = shift with zero fill (bit-exact); linearity; tensor-core vs exact-fp32 kernel agreement
  K3  N identical sources under the identity warp: h == value / N (quirk Q6)
  K4  constant logits: depth == mean of the plane depths, prob == 1/D, argmax == 0
"""
import pytest
import torch

from estdepth_b200 import ops, packing, synth

pytestmark = pytest.mark.gpu
D, H, W = 64, 120, 160
DEV = "cuda"


def _vol(chunks, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(chunks, D, H, W, 4, generator=g).to(DEV)


def test_k1_identity_pose_and_linearity():
    g = torch.Generator().manual_seed(0)
    ref = torch.randn(8, H, W, 4, generator=g).to(DEV)
    src = torch.randn(8, H, W, 4, generator=g).to(DEV)
    src2 = torch.randn(8, H, W, 4, generator=g).to(DEV)
    K4 = synth.intrinsics(4 * H, 4 * W).clone()
    K4[:2] *= 0.25
    pose = synth.camera_track(3)
    dv = (torch.arange(D, dtype=torch.float32) * (9.9 / (D - 1)) + 0.1).to(DEV)
    same = ops.homography_setup(pose[1].to(DEV), pose[1].to(DEV), K4.to(DEV))
    x0 = ops.warp_cost(ref, src, same, dv, align_corners=True)
    # identity homography samples pixel (w,h) itself: ix = w only up to fp32 round-off of the normalise /
    # un-normalise pair, so the bilinear weights are (1-e, e) with e ~ 1e-5: compare with that slack
    want = (ref + src).unsqueeze(1).expand(-1, D, -1, -1, -1)
    # border pixels can fall outside [-1,1] by one ulp and are then zero-filled, exactly like the reference (Q10)
    assert (x0 - want)[:, :, 1:-1, 1:-1].abs().max().item() < 2e-3
    moved = ops.homography_setup(pose[1].to(DEV), pose[0].to(DEV), K4.to(DEV))
    zero = torch.zeros_like(ref)
    a, b = ops.warp_cost(zero, src, moved, dv), ops.warp_cost(zero, src2, moved, dv)
    ab = ops.warp_cost(zero, 2.0 * src - 3.0 * src2, moved, dv)
    assert (ab - (2.0 * a - 3.0 * b)).abs().max().item() < 1e-4
    frac_zero = (a.abs().sum(dim=(0, 4)) == 0).float().mean().item()
    assert 0.01 < frac_zero < 0.5          # near planes leave the source image (quirk Q10), far planes do not


@pytest.mark.parametrize("precision", ["fp32", "3xtf32", "3xf16", "3xf16r", "3xf16r2", "3xf16r2d"])
def test_k2_delta_filter_is_a_zero_filled_shift(precision):
    x = _vol(8, 1)
    w = torch.zeros(32, 32, 3, 3, 3)
    for c in range(32):
        w[c, c, 0, 2, 1] = 1.0                     # tap (kd=0, kh=2, kw=1): out[d,h,w] = in[d-1, h+1, w]
    pc = packing.attach_tc(ops.PackedConv(packing.pack_weight(w, list(range(32)), list(range(32))).to(DEV),
                                          torch.ones(32, device=DEV), torch.zeros(32, device=DEV), 8, 32, 8, 32, "none", "none"))
    y = ops.conv3d(pc, x, torch.empty_like(x), precision=precision)
    want = torch.zeros_like(x)
    want[:, 1:, :-1] = x[:, :-1, 1:]
    if precision == "fp32":
        assert torch.equal(y, want)
    else:
        # two-term splits reproduce fp32 inputs to 2^-21 relative.  The truncation-bias compensation (csrc/common.cuh) assumes
        # that every MMA of the filter's footprint adds something; a one-tap filter adds zeros in all the others, so the output is
        # over-corrected by at most n_mma * 0.272 * 2^-24 (n_mma = 162 on the ring schedules, 54 on the output-stationary one)
        n_mma = 162 if precision in ("3xf16r", "3xf16r2") else 54
        assert (y - want).abs().max().item() <= (2.0 ** -20 + n_mma * 0.272 * 2.0 ** -24) * x.abs().max().item()


def test_k2_linearity_and_kernel_agreement():
    g = torch.Generator().manual_seed(2)
    x, y = _vol(8, 3), _vol(8, 4)
    w = torch.randn(32, 32, 3, 3, 3, generator=g) / 30
    pc = packing.attach_tc(ops.PackedConv(packing.pack_weight(w, list(range(32)), list(range(32))).to(DEV),
                                          torch.ones(32, device=DEV), torch.zeros(32, device=DEV), 8, 32, 8, 32, "none", "none"))
    exact = ops.conv3d(pc, x, torch.empty_like(x), precision="fp32")
    for precision in ("3xtf32", "3xf16", "3xf16r", "3xf16r2", "3xf16r2d"):
        got = ops.conv3d(pc, x, torch.empty_like(x), precision=precision)
        assert (got - exact).abs().max().item() < (4e-5 if precision in ("3xf16r", "3xf16r2") else 2e-5), precision   # outputs are O(1)
    fx = exact
    fy = ops.conv3d(pc, y, torch.empty_like(x), precision="3xf16")
    fxy = ops.conv3d(pc, 0.5 * x + 2.0 * y, torch.empty_like(x), precision="3xf16")
    assert (fxy - (0.5 * fx + 2.0 * fy)).abs().max().item() < 5e-5


@pytest.mark.parametrize("n_src", [1, 3])
def test_k3_identical_sources_identity_warp(n_src):
    key, val = _vol(4, 5).abs(), torch.tanh(_vol(4, 6))
    K4 = synth.intrinsics(4 * H, 4 * W).clone()
    K4[:2] *= 0.25
    eye = torch.eye(4, device=DEV)
    w30 = torch.stack([ops.volume_warp_setup(eye, eye, K4.to(DEV)) for _ in range(n_src)])
    dmin, dmax = 0.1, 10.0
    interval = (dmax - dmin) / (D - 1)
    dv = (torch.arange(D, dtype=torch.float32) * interval + dmin).to(DEV)
    h = ops.est_attend(key, [key] * n_src, [val] * n_src, w30, dv, dmin, interval, align_corners=True)
    # the identity warp lands on voxel centres up to ~1e-5 voxels of round-off
    # ... and border voxels may be forced outside by one ulp (zero-filled, quirk Q10): compare the interior
    assert (h - val / n_src)[:, 1:-1, 1:-1, 1:-1].abs().max().item() < 5e-4


def test_k4_constant_logits():
    dv = (torch.arange(D, dtype=torch.float32) * (9.9 / (D - 1)) + 0.1).to(DEV)
    logits = torch.full((D, H, W), 0.37, device=DEV)
    depth = torch.empty(4 * H, 4 * W, device=DEV)
    prob = torch.empty_like(depth)
    idx = torch.empty(4 * H, 4 * W, device=DEV, dtype=torch.int32)
    ops.head_softargmin(dv, logits_in=logits, depth_out=depth, prob_out=prob, argmax_out=idx, up=4)
    assert (depth - dv.mean()).abs().max().item() < 1e-5
    assert (prob - 1.0 / D).abs().max().item() < 1e-7
    assert int(idx.abs().max()) == 0
