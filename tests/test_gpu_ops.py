"""GPU parity of every kernel behind the C ABI against the CPU oracle / the reference's golden outputs.

All calls go through ``estdepth_b200.ops`` -> ctypes -> libestdepth_b200.so (the C ABI).  Tolerances are written
next to each check; they are fp32 round-off scale, far inside the 1e-3 depth gate of BASELINE.json.
"""
import os
import zlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from estdepth_b200 import ops, packing, synth
from oracle import estdepth_oracle as orc
from oracle.make_golden import ops_inputs
from tests.helpers import from_vol4, to_map4, to_vol4

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def maxdiff(a, b):
    return (a.detach().cpu().double() - torch.as_tensor(b).double()).abs().max().item()


def test_library_loads_and_counts_launches():
    from estdepth_b200 import _lib
    lib = _lib.get()
    assert lib.estd_version() >= 100
    before = _lib.launch_count()
    ops.scalar_to_vol4(torch.zeros(2, 3, 4, device=DEV))
    assert _lib.launch_count() == before + 1


def test_errors_are_reported_not_swallowed():
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        ops.scalar_to_vol4(torch.zeros(2, 3, 4))
    pc = ops.PackedConv(torch.zeros(27, 12, 24, device=DEV), torch.zeros(24, device=DEV), torch.zeros(24, device=DEV),
                        3, 24, 6, 24, "none", "none")
    x = torch.zeros(3, 4, 8, 8, 4, device=DEV)
    y = torch.zeros(6, 4, 8, 8, 4, device=DEV)
    with pytest.raises(RuntimeError, match="no kernel"):
        ops.conv3d(pc, x, y)


def test_layout_roundtrip():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(16, 5, 7, 9, generator=g)
    v = ops.ncdhw_to_vol4(x.to(DEV))
    assert maxdiff(v, to_vol4(x)) == 0.0
    assert maxdiff(ops.vol4_to_ncdhw(v), x) == 0.0
    s = torch.randn(5, 7, 9, generator=g)
    sv = ops.scalar_to_vol4(s.to(DEV)).cpu()
    assert torch.equal(sv[0, ..., 0], s) and sv[0, ..., 1:].abs().max() == 0


def test_batched_geometry_tables_equal_the_per_pair_kernels_and_the_reference_ops():
    """estd_homography_table / estd_volume_warp_table (one launch per window / fusion step) == the per-pair kernels bit for
    bit, and agree with the reference's fp32 torch op sequence (model_hybrid.py:74-88, homo_utils.py:469-471, :51, :258,
    hybrid_depth_decoder.py:235) to fp32 round-off."""
    poses = synth.camera_track(5).to(DEV)
    K4 = synth.intrinsics(480, 640).clone()
    K4[:2] *= 0.25
    K4 = K4.to(DEV)
    pairs = [(t + 1, s) for t in range(3) for s in (t, t + 2)]
    table = ops.homography_table(poses, K4, pairs)
    for i, (r, s) in enumerate(pairs):
        assert torch.equal(table[i], ops.homography_setup(poses[r].contiguous(), poses[s].contiguous(), K4))
    want = ops.homography_table_torch(poses.cpu(), K4.cpu(), pairs)
    assert (table.cpu() - want).abs().max().item() < 2e-5 * want.abs().max().item()
    memory = [synth.camera_track(1, start=7)[0].to(DEV), synth.camera_track(1, start=9)[0].to(DEV)]
    all_poses = [poses[t + 1] for t in range(3)] + memory
    tabs = ops.volume_warp_tables(all_poses, 3, K4)
    ref = ops.volume_warp_tables_torch([p.cpu() for p in all_poses], 3, K4.cpu())
    for i in range(3):
        others = [j for j in range(5) if j != i]
        assert tabs[i].shape == (4, 30)
        for n, j in enumerate(others):
            assert torch.equal(tabs[i][n], ops.volume_warp_setup(all_poses[i].contiguous(), all_poses[j].contiguous(), K4))
        assert (tabs[i].cpu() - ref[i]).abs().max().item() < 2e-5 * ref[i].abs().max().item()
    with pytest.raises(RuntimeError, match="out of range"):
        ops.homography_table(poses, K4, [(1, 7)])


def test_premix_matches_matmul():
    g = torch.Generator().manual_seed(2)
    fea = torch.randn(32, 30, 40, generator=g)
    w = torch.randn(32, 32, generator=g) / 5
    b = torch.randn(32, generator=g)
    out = ops.premix(fea.to(DEV), w.to(DEV), b.to(DEV))
    want = to_map4(torch.einsum("oc,chw->ohw", w.double(), fea.double()).float() + b.view(-1, 1, 1))
    assert maxdiff(out, want) < 2e-5            # 32-term fp32 dot products of O(1) values


def test_premix_batch_is_the_stacked_single_map_premix():
    """One launch for all frames and both halves of pre0 (what the model issues) == per-map, per-half launches, bit for bit
    (ragged pixel count: 33 * 37 is not a multiple of the 32-pixel warp tile)."""
    g = torch.Generator().manual_seed(3)
    fea = torch.randn(5, 32, 33, 37, generator=g).to(DEV)
    w_ref, w_src = (torch.randn(32, 32, generator=g) / 5).to(DEV), (torch.randn(32, 32, generator=g) / 5).to(DEV)
    b = torch.randn(32, generator=g).to(DEV)
    both = ops.premix_batch(fea, torch.cat([w_ref, w_src], 0).contiguous(), torch.cat([b, torch.zeros_like(b)]).contiguous())
    assert tuple(both.shape) == (5, 16, 33, 37, 4)
    for v in range(5):
        assert torch.equal(both[v, :8], ops.premix(fea[v], w_ref, b))
        assert torch.equal(both[v, 8:], ops.premix(fea[v], w_src, None))
    want = to_map4((torch.einsum("oc,chw->ohw", w_src.double().cpu(), fea[4].double().cpu())).float())
    assert maxdiff(both[4, 8:], want) < 2e-5


@pytest.mark.parametrize("h,w,H,W", [(3, 5, 120, 160), (7, 10, 120, 160), (15, 20, 120, 160), (30, 40, 120, 160), (1, 1, 9, 7), (5, 4, 13, 11)])
def test_upsample_bilinear_vol4_matches_interpolate(h, w, H, W):
    """SPP branch tail (networks/psm_submodule.py:104-114): ReLU(conv + b) -> F.upsample(bilinear) -> cat, as one kernel
    writing a chunk slice of a wider vol4 buffer; arithmetic of ATen's upsample_bilinear2d (align_corners=False)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(h * 100 + w)
    src = torch.randn(3, 8, h, w, generator=g).to(DEV)
    bias = torch.randn(8, generator=g).to(DEV)
    wide = torch.full((5, 3, H, W, 4), 7.0, device=DEV)
    ops.upsample_bilinear_vol4(src, wide[2:4], bias=bias, relu=True)
    want = F.interpolate(F.relu(src + bias.view(1, -1, 1, 1)), size=(H, W), mode="bilinear", align_corners=False)
    assert maxdiff(wide[2:4], ops.nchw_to_vol4(want.contiguous()).cpu()) < 2e-6
    assert float((wide[:2] - 7.0).abs().max()) == 0.0 and float((wide[4:] - 7.0).abs().max()) == 0.0      # neighbours untouched
    plain = ops.upsample_bilinear_vol4(src, torch.empty(2, 3, H, W, 4, device=DEV))
    assert maxdiff(plain, ops.nchw_to_vol4(F.interpolate(src, size=(H, W), mode="bilinear", align_corners=False).contiguous()).cpu()) < 2e-6


@pytest.mark.parametrize("src", [0, 2])
def test_homo_warping_vs_reference_golden(src):
    """ops.homo_warping == the reference's homo_warping on the same inputs (golden from oracle/make_golden.py)."""
    x = ops_inputs()
    D = x["depth_values"].numel()
    ext = torch.inverse(x["poses"]).unsqueeze(0)
    sp, rp = ext[:, src].clone(), ext[:, 1].clone()
    sp[:, :3, :4] = x["K4"] @ ext[:, src, :3, :4]
    rp[:, :3, :4] = x["K4"] @ ext[:, 1, :3, :4]
    got = ops.homo_warping(x["fea"].to(DEV), sp.to(DEV), rp.to(DEV), x["depth_values"].view(1, D, 1, 1).to(DEV))
    want = golden("ops_small.npz")["homo_warp_%d" % src]
    # white-noise features: a 1e-5 px coordinate difference moves a sample by ~1e-5 * |gradient| ~ 5e-5
    assert maxdiff(got, want) < 2e-4
    oracle = orc.homo_warp(x["fea"], sp, rp, x["depth_values"], sampler="explicit")
    assert maxdiff(got, oracle) < 2e-4


def test_warp_cost_fused_equals_pre0_of_cat():
    """K1 == pre0(cat[ref_volume, homo_warping(src)]) (model_hybrid.py:76,90-94), incl. out-of-range planes."""
    g = torch.Generator().manual_seed(3)
    C, D, H, W = 32, 12, 36, 44
    feats = [torch.randn(1, C, H, W, generator=g) for _ in range(3)]
    sd = {"pre0.0.weight": torch.randn(32, 64, 1, 1, 1, generator=g) / 8,
          "pre0.1.weight": torch.rand(32, generator=g) + 0.5, "pre0.1.bias": torch.randn(32, generator=g) / 5,
          "pre0.1.running_mean": torch.randn(32, generator=g) / 5, "pre0.1.running_var": torch.rand(32, generator=g) + 0.5}
    poses = synth.camera_track(3).unsqueeze(0)
    K4 = synth.intrinsics(4 * H, 4 * W).unsqueeze(0).clone()
    K4[:, :2] *= 0.25
    dv = torch.linspace(0.1, 10.0, D)
    w_ref, w_src, bias = packing.split_pre0(sd)
    for s in (0, 2):
        ext = torch.inverse(poses)
        sp, rp = ext[:, s].clone(), ext[:, 1].clone()
        sp[:, :3, :4] = K4 @ ext[:, s, :3, :4]
        rp[:, :3, :4] = K4 @ ext[:, 1, :3, :4]
        warped = orc.homo_warp(feats[s], sp, rp, dv)
        want = orc._cb3(torch.cat([feats[1].unsqueeze(2).repeat(1, 1, D, 1, 1), warped], 1), sd, "pre0")[0]
        ref_mix = ops.premix(feats[1][0].to(DEV), w_ref.to(DEV), bias.to(DEV))
        src_mix = ops.premix(feats[s][0].to(DEV), w_src.to(DEV))
        h12 = ops.homography_setup(poses[0, 1].to(DEV), poses[0, s].to(DEV), K4[0].to(DEV))
        got = from_vol4(ops.warp_cost(ref_mix, src_mix, h12, dv.to(DEV)).cpu())
        assert maxdiff(got, want) < 3e-4
        frac_zero = (warped.abs().sum(1) == 0).float().mean().item()
        assert 0.01 < frac_zero < 0.99, "test must cover both in-range and out-of-range samples"


CONV_CASES = [
    # name, cin segments (chunks), cout real, cout_pad, out segments (chunks), act_split, act_lo, act_hi
    ("32to32", (8,), 32, 32, (8,), 32, "relu", "relu"),
    ("16+16to32", (4, 4), 32, 32, (8,), 16, "none", "none"),
    ("36to40", (8, 1), 33, 40, (9,), 40, "relu", "relu"),
    ("36to16+16", (9,), 32, 32, (4, 4), 16, "tanh", "relu"),
    ("16to16", (4,), 16, 16, (4,), 16, "relu", "relu"),
    ("16+16to16", (4, 4), 16, 16, (4,), 16, "none", "none"),
]


@pytest.mark.parametrize("precision", ["fp32", "3xtf32", "3xf16", "3xf16r", "3xf16r2", "3xf16r2d"])
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
@pytest.mark.parametrize("shape", [(5, 21, 40), (8, 32, 64)], ids=["ragged", "aligned"])
def test_conv3d_vs_torch_cpu(case, shape, precision):
    """K2 (both the exact CUDA-core kernel and the 3xTF32 tcgen05 kernel) against F.conv3d (CPU fp32) + affine +
    activation + residuals, all channel configurations, ragged edges."""
    name, cin_seg, cout, cout_pad, out_seg, act_split, act_lo, act_hi = case
    D, H, W = shape
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 1000)
    cin = 4 * sum(cin_seg)
    x = torch.randn(cin, D, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g) / (cin * 27) ** 0.5
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) / 3
    res0 = torch.randn(4 * sum(out_seg), D, H, W, generator=g)
    res1 = torch.randn(4 * sum(out_seg), D, H, W, generator=g)
    order = list(range(cout)) + [-1] * (cout_pad - cout)
    pw = packing.pack_weight(w, list(range(cin)), order)
    s_pad = torch.zeros(cout_pad)
    b_pad = torch.zeros(cout_pad)
    s_pad[:cout], b_pad[:cout] = scale, shift
    pc = packing.attach_tc(ops.PackedConv(pw.to(DEV), s_pad.to(DEV), b_pad.to(DEV), sum(cin_seg), cout_pad, sum(out_seg),
                                          act_split, act_lo, act_hi))

    y = F.conv3d(x.unsqueeze(0), w, None, 1, 1)[0] * scale.view(-1, 1, 1, 1) + shift.view(-1, 1, 1, 1)
    acts = {"none": lambda t: t, "relu": torch.relu, "tanh": torch.tanh}
    y = torch.cat([acts[act_lo](y[:act_split]), acts[act_hi](y[act_split:])], 0)
    cpad = 4 * sum(out_seg)
    ypad = torch.zeros(cpad, D, H, W)
    ypad[:cout] = y
    want = (ypad + res0 + res1) * 0.5

    xin = to_vol4(x).to(DEV)
    ins = [xin[:cin_seg[0]].contiguous()] + ([xin[cin_seg[0]:].contiguous()] if len(cin_seg) > 1 else [])
    outs = [torch.full((c, D, H, W, 4), float("nan"), device=DEV) for c in out_seg]
    n_ctas = ops.conv3d_num_ctas(pc, D, H, W, precision=precision)
    partials = torch.zeros(n_ctas, 2, 2, device=DEV, dtype=torch.float64)
    ops.conv3d(pc, ins[0], outs[0], in1=ins[1] if len(ins) > 1 else None, out1=outs[1] if len(outs) > 1 else None,
               res0=to_vol4(res0).to(DEV), res1=to_vol4(res1).to(DEV), post_scale=0.5, gn_partials=partials,
               precision=precision)
    got = from_vol4(torch.cat(outs, 0).cpu())
    assert torch.isfinite(got).all()
    err = maxdiff(got, want)
    print("conv3d %s %s %s: max |err| = %.3e" % (name, shape, precision, err))
    # K = 27*cin <= 972 fp32 products of O(1)/sqrt(K) terms: round-off ~ 1e-6; 3e-5 leaves margin for tanh.
    # 3xTF32 drops the x_lo*w_lo products (2^-22 relative each): same bound.
    assert err < 3e-5
    # deterministic GroupNorm partial sums (over the real, written channels of each group)
    tot = partials.sum(0).cpu()
    real = want.clone()
    g0 = real[:min(act_split, cpad)].double()
    assert abs(tot[0, 0].item() - g0.sum().item()) < 1e-3 * max(1.0, g0.abs().sum().item() ** 0.5)
    assert abs(tot[0, 1].item() - (g0 ** 2).sum().item()) < 1e-4 * (g0 ** 2).sum().item()
    if act_split < cpad:
        g1 = real[act_split:].double()
        assert abs(tot[1, 1].item() - (g1 ** 2).sum().item()) < 1e-4 * (g1 ** 2).sum().item()


@pytest.mark.parametrize("precision", ["fp32", "3xtf32", "3xf16", "3xf16r", "3xf16r2", "3xf16r2d"])
def test_conv3d_is_bitwise_deterministic(precision):
    g = torch.Generator().manual_seed(5)
    D, H, W = 6, 24, 64
    x = to_vol4(torch.randn(32, D, H, W, generator=g)).to(DEV)
    w = torch.randn(32, 32, 3, 3, 3, generator=g) / 30
    pc = packing.attach_tc(ops.PackedConv(packing.pack_weight(w, list(range(32)), list(range(32))).to(DEV),
                                          torch.ones(32, device=DEV), torch.zeros(32, device=DEV), 8, 32, 8, 16, "none", "none"))
    outs, parts = [], []
    for _ in range(2):
        y = torch.empty_like(x)
        p = torch.zeros(ops.conv3d_num_ctas(pc, D, H, W, precision=precision), 2, 2, device=DEV, dtype=torch.float64)
        ops.conv3d(pc, x, y, gn_partials=p, precision=precision)
        outs.append(y.cpu())
        parts.append(p.cpu())
    assert torch.equal(outs[0], outs[1]) and torch.equal(parts[0], parts[1])


@pytest.mark.parametrize("ring", ["3xf16r", "3xf16r2"])
@pytest.mark.parametrize("case", ["32to32", "36to32", "16to16", "32to16", "36to33"])
@pytest.mark.parametrize("shape", [(48, 64, 128), (13, 40, 70), (3, 120, 160), (1, 33, 65)],
                         ids=["long_segments", "ragged", "three_planes", "one_plane"])
def test_conv3d_ring_matches_exact_kernel(case, shape, ring):
    """The plane-ring schedule (conv3d_ring.cu) against the exact fp32 CUDA-core kernel on volumes large enough that a
    CTA's range spans several planes and crosses column boundaries (partial first/last planes, ring wrap-around, the
    hand-over of the last plane of a column), with residuals, two input segments and two output tensors."""
    D, H, W = shape
    g = torch.Generator().manual_seed(D * 1000 + H)
    cin_seg = {"32to32": (8,), "36to32": (8, 1), "16to16": (4,), "32to16": (4, 4), "36to33": (8, 1)}[case]
    cout = {"16to16": 16, "32to16": 16, "36to33": 33}.get(case, 32)
    cout_pad = 40 if cout == 33 else cout                 # 33 -> 40 for the exact kernel, 48 on the tensor cores (dres2)
    cin = 4 * sum(cin_seg)
    x = torch.randn(sum(cin_seg), D, H, W, 4, generator=g).to(DEV)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g) / (cin * 27) ** 0.5
    scale, shift = torch.zeros(cout_pad), torch.zeros(cout_pad)
    scale[:cout] = torch.rand(cout, generator=g) + 0.5
    shift[:cout] = torch.randn(cout, generator=g) / 3
    order = list(range(cout)) + [-1] * (cout_pad - cout)
    out_chunks = (cout + 3) // 4
    oc = out_chunks // 2                                  # chunks of the first output tensor; the rest go to the second
    pc = packing.attach_tc(ops.PackedConv(packing.pack_weight(w, list(range(cin)), order).to(DEV), scale.to(DEV), shift.to(DEV),
                                          sum(cin_seg), cout_pad, out_chunks, 8 * (cout // 16), "tanh", "relu"))
    assert pc.weight_ring is not None
    res0 = torch.randn(out_chunks, D, H, W, 4, generator=g).to(DEV)
    if cout % 4:                                          # pad channels of the residual must be zero to compare padded outputs
        res0[-1, ..., cout % 4:] = 0
    ins = [x[:cin_seg[0]].contiguous()] + ([x[cin_seg[0]:].contiguous()] if len(cin_seg) > 1 else [])
    got, want = {}, {}
    if ring == "3xf16r2" and pc.weight_ring2 is None:
        pytest.skip("no CTA-pair specialisation for this shape (runs the single-CTA ring kernel)")
    for precision, store in (("fp32", want), (ring, got)):
        o0 = torch.full((oc, D, H, W, 4), float("nan"), device=DEV)
        o1 = torch.full((out_chunks - oc, D, H, W, 4), float("nan"), device=DEV)
        n = ops.conv3d_num_ctas(pc, D, H, W, precision=precision)
        part = torch.zeros(n, 2, 2, device=DEV, dtype=torch.float64)
        ops.conv3d(pc, ins[0], o0, in1=ins[1] if len(ins) > 1 else None, out1=o1, res0=res0, post_scale=0.5,
                   gn_partials=part, precision=precision)
        store["y"], store["p"] = torch.cat([o0, o1], 0), part.sum(0)
    assert torch.isfinite(got["y"]).all()
    err = (got["y"] - want["y"]).abs().max().item()
    print("conv3d ring %s %s: max |err| vs exact = %.3e" % (case, shape, err))
    # 3 products x 54 (tap, k-step) pairs accumulate in ONE fp32 TMEM accumulator and the tensor core truncates on every
    # accumulate: ~162 x 0.5 ulp of an O(1) sum (the output-stationary kernel keeps the small products apart: 4e-6)
    assert err < 4e-5
    assert torch.allclose(got["p"], want["p"], rtol=1e-5, atol=1e-3)
    ops.check_status(torch.device(DEV))


PLANAR_CASES = [
    # name, cin segments (channels), cout, dilation, maps, H, W
    ("64to64", (64,), 64, 1, 5, 40, 56),
    ("64to64_dil2", (64,), 64, 2, 3, 33, 47),
    ("32to32", (32,), 32, 1, 2, 48, 64),
    ("320to128", (192, 128), 128, 1, 2, 30, 40),
    ("1280to256_small", (256, 1024), 256, 1, 3, 15, 20),
    ("128to33pad", (128,), 48, 1, 1, 16, 16),
    ("1x1_256to64", (256,), 64, 0, 3, 40, 56),
    ("1x1_64to256", (64,), 256, 0, 3, 33, 47),
    ("1x1_512to2048_small", (512,), 2048, 0, 3, 15, 20),
]


@pytest.mark.parametrize("case", PLANAR_CASES, ids=[c[0] for c in PLANAR_CASES])
def test_conv2d_planar_vs_torch_cpu(case):
    """Planar tcgen05 3x3 convolution (conv2d_tc.cu) == F.conv2d (CPU, fp64 accumulate) + affine + ReLU + residual: any
    number of input channels (run-time k-steps), two input segments (the decoder's torch.cat), 64-channel output slices,
    dilation 2, ragged tiles."""
    name, cin_seg, cout, dil, N, H, W = case
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 1000)
    cin = sum(cin_seg)
    k = 3 if dil > 0 else 1                                # dilation 0 marks the 1x1 (pointwise) cases
    x = torch.randn(N, cin, H, W, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / (k * k * cin) ** 0.5
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) / 3
    res = torch.randn(N, cout, H, W, generator=g)
    y = F.conv2d(x.double(), w.double(), None, 1, dil, max(dil, 1)) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    # 3x3 layers: ReLU then residual (PSM blocks, quirk Q12); 1x1 layers: residual then ReLU (ResNet bottleneck)
    want = (torch.relu(y) + res).float() if k == 3 else torch.relu(y + res).float()
    pcs = packing.pack_conv2d(w, scale, shift, "relu" if k == 3 else "add_relu", DEV,
                              cout_slice=64 if cout > 32 else (32 if cout > 16 else 16))
    x4 = ops.nchw_to_vol4(x.to(DEV))
    ins = [x4[:cin_seg[0] // 4].contiguous()] + ([x4[cin_seg[0] // 4:].contiguous()] if len(cin_seg) > 1 else [])
    out4 = torch.full((cout // 4, N, H, W, 4), float("nan"), device=DEV)
    res4 = ops.nchw_to_vol4(res.to(DEV))
    step = pcs[0].cout_pad // 4
    for i, pc in enumerate(pcs):
        lo, hi = step * i, step * i + pc.out_chunks
        ops.conv_planar(pc, ins[0], out4[lo:hi], res0=res4[lo:hi], dilation=max(dil, 1), in1=ins[1] if len(ins) > 1 else None,
                        taps=k * k)
    got = ops.vol4_to_nchw(out4).cpu()
    assert torch.isfinite(got).all()
    err = (got - want).abs().max().item()
    print("conv2d planar %s: max |err| = %.3e" % (name, err))
    # x_hi*w_hi runs through one fp32 TMEM accumulator; the tensor core truncates on each of the 9*cin/16 accumulates
    assert err < 3e-5 * max(1.0, cin / 320.0)
    ops.check_status(torch.device(DEV))


@pytest.mark.parametrize("j", [0, 2])
def test_warp_volume_vs_reference_golden(j):
    """ops.warp_volume (the fused EST gather with N=1) == the reference's warp_volume (utils/homo_utils.py:240)."""
    x = ops_inputs()
    D = x["depth_values"].numel()
    _, C, _, H, W = x["vol"].shape
    rel = torch.matmul(x["poses"][j:j + 1], torch.inverse(x["poses"][1:2]))
    dv = x["depth_values"].view(1, 1, D, 1).repeat(1, 1, 1, H * W)
    got = ops.warp_volume(x["vol"].to(DEV), dv.to(DEV), rel.to(DEV), x["K4"].to(DEV), None, x["depth_min"], x["interval"])
    want = golden("ops_small.npz")["warp_volume_%d" % j]
    assert maxdiff(got, want) < 3e-4        # white-noise volume; see SURVEY.md 8c (4e-5 restatement noise)
    oracle = orc.warp_volume(x["vol"], rel, x["K4"], x["depth_values"], x["depth_min"], x["interval"], sampler="explicit")
    assert maxdiff(got, oracle) < 3e-4
    frac_zero = (torch.as_tensor(want).abs().sum(1) == 0).float().mean().item()
    assert 0.01 < frac_zero < 0.99


@pytest.mark.parametrize("n_src", [1, 2, 3])
def test_est_attend_vs_oracle(n_src):
    """K3 == warp_volume x 2N + attention (epipolar_transformer.py:62-73) on smooth volumes."""
    g = torch.Generator().manual_seed(10 + n_src)
    D, H, W = 8, 24, 40
    dmin, dmax = 0.5, 6.0
    interval = (dmax - dmin) / (D - 1)
    dv = torch.arange(D, dtype=torch.float32) * interval + dmin

    def smooth(c):
        return F.interpolate(torch.randn(1, c, D // 2, H // 4, W // 4, generator=g), size=(D, H, W), mode="trilinear")

    key_t = torch.relu(smooth(16))
    keys = [torch.relu(smooth(16)) for _ in range(n_src)]
    vals = [torch.tanh(smooth(16)) for _ in range(n_src)]
    poses = synth.camera_track(n_src + 1)
    K4 = synth.intrinsics(4 * H, 4 * W).unsqueeze(0).clone()
    K4[:, :2] *= 0.25
    wk, wv, w30 = [], [], []
    for n in range(n_src):
        rel = torch.matmul(poses[n + 1:n + 2], torch.inverse(poses[0:1]))
        wk.append(orc.warp_volume(keys[n], rel, K4, dv, dmin, interval))
        wv.append(orc.warp_volume(vals[n], rel, K4, dv, dmin, interval))
        w30.append(ops.volume_warp_setup(poses[0].to(DEV), poses[n + 1].to(DEV), K4[0].to(DEV)))
    want = orc.est_attention(key_t, wk, wv)[0]
    got = ops.est_attend(to_vol4(key_t[0]).to(DEV), [to_vol4(k[0]).to(DEV) for k in keys],
                         [to_vol4(v[0]).to(DEV) for v in vals], torch.stack(w30), dv.to(DEV), dmin, interval)
    assert maxdiff(from_vol4(got.cpu()), want) < 1e-4


@pytest.mark.parametrize("n_src", [1, 2, 3])
def test_gru_chain_vs_reference_golden(n_src):
    """gate conv -> GroupNorm -> reset -> output conv -> GroupNorm -> blend == reference EpipolarTransformer output."""
    from oracle.ref_loader import reference_available  # noqa: F401  (golden was produced by the reference)
    x = ops_inputs()
    tmpl = {"gate_conv.weight": torch.empty(32, 32, 3, 3, 3), "gate_conv.bias": torch.empty(32),
            "reset_gate_norm.weight": torch.empty(16), "reset_gate_norm.bias": torch.empty(16),
            "update_gate_norm.weight": torch.empty(16), "update_gate_norm.bias": torch.empty(16),
            "output_conv.weight": torch.empty(16, 32, 3, 3, 3), "output_conv.bias": torch.empty(16),
            "output_norm.weight": torch.empty(16), "output_norm.bias": torch.empty(16)}
    sd = {"CostRegNet.epipolar_transformer." + k: v for k, v in synth.synth_state_dict(tmpl, seed=3).items()}
    want = golden("ops_small.npz")["est_n%d" % n_src][0]
    assert maxdiff(orc.est_fuse(sd, x["key_t"], x["wkeys"][:n_src], x["val_t"], x["wvals"][:n_src])[0], want) < 1e-5
    h = orc.est_attention(x["key_t"], x["wkeys"][:n_src], x["wvals"][:n_src])
    _, _, D, H, W = h.shape
    r32 = list(range(32))
    gate = ops.PackedConv(packing.pack_weight(sd["CostRegNet.epipolar_transformer.gate_conv.weight"], r32, r32).to(DEV),
                          torch.ones(32, device=DEV), sd["CostRegNet.epipolar_transformer.gate_conv.bias"].to(DEV),
                          8, 32, 8, 16, "none", "none")
    outc = ops.PackedConv(packing.pack_weight(sd["CostRegNet.epipolar_transformer.output_conv.weight"], r32, list(range(16))).to(DEV),
                          torch.ones(16, device=DEV), sd["CostRegNet.epipolar_transformer.output_conv.bias"].to(DEV),
                          8, 16, 4, 16, "none", "none")
    p = {k.rsplit("transformer.", 1)[1]: v.to(DEV) for k, v in sd.items()}
    v4, h4 = to_vol4(x["val_t"][0]).to(DEV), to_vol4(h[0]).to(DEV)
    f = torch.empty(8, D, H, W, 4, device=DEV)
    pf = torch.zeros(ops.conv3d_num_ctas(gate, D, H, W), 2, 2, device=DEV, dtype=torch.float64)
    ops.conv3d(gate, v4, f, in1=h4, gn_partials=pf)
    sf = ops.gn_finalize(pf, 2, 16.0 * D * H * W)
    rh = ops.gru_reset(f, h4, sf, p["reset_gate_norm.weight"], p["reset_gate_norm.bias"])
    o = torch.empty(4, D, H, W, 4, device=DEV)
    po = torch.zeros(ops.conv3d_num_ctas(outc, D, H, W), 2, 2, device=DEV, dtype=torch.float64)
    ops.conv3d(outc, v4, o, in1=rh, gn_partials=po)
    so = ops.gn_finalize(po, 1, 16.0 * D * H * W)
    fused = ops.gru_blend(f, h4, o, sf, so, p["update_gate_norm.weight"], p["update_gate_norm.bias"],
                          p["output_norm.weight"], p["output_norm.bias"])
    assert maxdiff(from_vol4(fused.cpu()), want) < 2e-5


def test_depthlayer_vs_reference_golden():
    x = ops_inputs()
    D = x["depth_values"].numel()
    dv = x["depth_values"].view(1, D, 1, 1).repeat(1, 1, 9, 11)
    depth, prob = ops.depthlayer(x["logits"].to(DEV), dv.to(DEV))
    gold = golden("ops_small.npz")
    assert maxdiff(depth, gold["depthlayer_depth"]) < 2e-6 * 4.0      # depths up to 4.0, fp32 softmax round-off
    assert maxdiff(prob, gold["depthlayer_prob"]) < 2e-6


def test_head_softargmin_fused_head_upsample_and_argmax():
    """1x1x1 head + x4 replicated soft-argmin; argmax index bit-exact wherever the top-2 logit gap exceeds round-off."""
    g = torch.Generator().manual_seed(7)
    D, H, W = 32, 12, 20
    hid = torch.relu(torch.randn(16, D, H, W, generator=g))
    w = torch.randn(16, generator=g)
    b = torch.randn(1, generator=g)
    dv = torch.linspace(0.1, 10.0, D)
    logits = (hid * w.view(16, 1, 1, 1)).sum(0) + b
    want_d, want_p, want_i = orc.soft_argmin(logits.unsqueeze(0), dv)
    lo = torch.empty(D, H, W, device=DEV)
    d = torch.empty(4 * H, 4 * W, device=DEV)
    p = torch.empty(4 * H, 4 * W, device=DEV)
    i = torch.empty(4 * H, 4 * W, device=DEV, dtype=torch.int32)
    ops.head_softargmin(dv.to(DEV), hidden=to_vol4(hid).to(DEV), head_w=w.to(DEV), head_b=b.to(DEV), logits_out=lo,
                        depth_out=d, prob_out=p, argmax_out=i, up=4)
    assert maxdiff(lo, logits) < 1e-5
    assert maxdiff(d, want_d[0, 0]) < 1e-4
    assert maxdiff(p, want_p[0, 0]) < 1e-5
    top2 = torch.topk(logits, 2, dim=0).values
    safe = F.interpolate(((top2[0] - top2[1]) > 1e-4).float()[None, None], scale_factor=4)[0, 0].bool()
    assert safe.float().mean() > 0.95
    assert torch.equal(i.cpu()[safe].long(), want_i[0, 0][safe])


def test_align_corners_flag_changes_sampling():
    """Quirk Q1: both grid_sample conventions are implemented; True matches F.grid_sample(align_corners=True)."""
    x = ops_inputs()
    D = x["depth_values"].numel()
    ext = torch.inverse(x["poses"]).unsqueeze(0)
    sp, rp = ext[:, 0].clone(), ext[:, 1].clone()
    sp[:, :3, :4] = x["K4"] @ ext[:, 0, :3, :4]
    rp[:, :3, :4] = x["K4"] @ ext[:, 1, :3, :4]
    _, C, H, W = x["fea"].shape
    xn, yn = orc.plane_sweep_grid(sp[0], rp[0], x["depth_values"], H, W)
    grid = torch.stack((xn, yn), 2).view(1, D * H, W, 2)
    want = F.grid_sample(x["fea"], grid, mode="bilinear", padding_mode="zeros", align_corners=True).view(1, C, D, H, W)
    got = ops.homo_warping(x["fea"].to(DEV), sp.to(DEV), rp.to(DEV), x["depth_values"].view(1, D).to(DEV), align_corners=True)
    assert maxdiff(got, want) < 2e-4
    got_false = ops.homo_warping(x["fea"].to(DEV), sp.to(DEV), rp.to(DEV), x["depth_values"].view(1, D).to(DEV))
    assert maxdiff(got_false, want) > 1e-2


# ------------------------------------------------------------------------------------------- pre-split activations (vol4s)
@pytest.mark.parametrize("precision", ["3xf16r", "3xf16r2", "3xf16r2d"])
@pytest.mark.parametrize("cin_chunks,cout,cout_pad", [(8, 32, 32), (9, 33, 48), (4, 16, 16)])
def test_conv3d_presplit_tensors_equal_fp32_tensors(precision, cin_chunks, cout, cout_pad):
    """A producer that writes x_hi | x_lo (out_split) and a consumer that reads it (in_split, res_split) compute exactly what the
    fp32-tensor launch computes: the tensor core sees the same operands, only the stored form differs (22 significant bits)."""
    g = torch.Generator().manual_seed(11)
    D, H, W = 5, 21, 40
    cin = 4 * cin_chunks
    x = torch.randn(cin, D, H, W, generator=g)
    if cin_chunks == 9:
        x[33:] = 0                                                 # the canonical 36-channel tensor: 33 real channels + 3 zero pads
    w = torch.randn(cout, cin, 3, 3, 3, generator=g) / (cin * 27) ** 0.5
    order = list(range(cout)) + [-1] * (cout_pad - cout)
    out_chunks = (cout + 3) // 4
    pc = packing.attach_tc(ops.PackedConv(packing.pack_weight(w, list(range(cin)), order), torch.ones(cout_pad), torch.zeros(cout_pad),
                                          cin_chunks, cout_pad, out_chunks, cout_pad, "relu", "relu")).to(DEV)
    res = torch.randn(4 * out_chunks, D, H, W, generator=g)
    x4, r4 = to_vol4(x).to(DEV), to_vol4(res).to(DEV)
    plain = ops.conv3d(pc, x4, torch.empty(out_chunks, D, H, W, 4, device=DEV), res0=r4, res1=r4, post_scale=0.5, precision=precision)
    xs, rs = ops.to_split(x4), ops.to_split(r4)
    n_out = (out_chunks + 1) // 2 * 2
    split = ops.conv3d(pc, xs, torch.full((n_out, D, H, W, 4), float("nan"), device=DEV), res0=rs, res1=rs, post_scale=0.5,
                       precision=precision, in_split=(True, False), res_split=True, out_split=True)
    got = ops.from_split(split)[:out_chunks]
    assert torch.isfinite(got).all()
    scale = plain.abs().max().item()
    assert (got - plain).abs().max().item() <= 2.0 ** -20 * scale          # 22-bit storage of the output and of the residuals
    # split input + fp32 output: bit-identical to the all-fp32 launch when there is no residual
    a = ops.conv3d(pc, x4, torch.empty(out_chunks, D, H, W, 4, device=DEV), precision=precision)
    b = ops.conv3d(pc, xs, torch.empty(out_chunks, D, H, W, 4, device=DEV), precision=precision, in_split=(True, False))
    assert torch.equal(a, b)
    ops.check_status(torch.device(DEV))


def test_conv3d_presplit_rejected_by_other_kernels():
    pc = packing.attach_tc(ops.PackedConv(packing.pack_weight(torch.randn(32, 32, 3, 3, 3) / 30, list(range(32)), list(range(32))),
                                          torch.ones(32), torch.zeros(32), 8, 32, 8, 32, "none", "none")).to(DEV)
    x = torch.randn(8, 4, 16, 32, 4, device=DEV)
    for precision in ("fp32", "3xf16", "3xtf32"):
        with pytest.raises(RuntimeError, match="plane-ring"):
            ops.conv3d(pc, x, torch.empty_like(x), precision=precision, in_split=(True, False))


def test_conv_planar_presplit_tensors_equal_fp32_tensors():
    g = torch.Generator().manual_seed(4)
    N, H, W, cin, cout = 2, 37, 50, 64, 128
    x = torch.randn(N, cin, H, W, generator=g).relu()
    w = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    res = torch.randn(N, cout, H, W, generator=g)
    pc = packing.pack_conv2d(w, torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) / 3, "add_relu", DEV)[0]
    x4, r4 = ops.nchw_to_vol4(x.to(DEV)), ops.nchw_to_vol4(res.to(DEV))
    plain = ops.conv_planar(pc, x4, torch.empty(cout // 4, N, H, W, 4, device=DEV), res0=r4)
    split = ops.conv_planar(pc, ops.to_split(x4), torch.full((cout // 4, N, H, W, 4), float("nan"), device=DEV), res0=ops.to_split(r4),
                            in_split=(True, False), res_split=True, out_split=True)
    got = ops.from_split(split)
    assert torch.isfinite(got).all()
    assert (got - plain).abs().max().item() <= 2.0 ** -20 * max(1.0, plain.abs().max().item())
    mixed = ops.conv_planar(pc, ops.to_split(x4), torch.empty(cout // 4, N, H, W, 4, device=DEV), res0=r4, in_split=(True, False))
    assert torch.equal(mixed, plain)


@pytest.mark.parametrize("split", [False, True])
def test_conv_planar_upsampled_output_equals_nearest_interpolate(split):
    """`upsample` of hybrid_depth_decoder.py:11-14 (F.interpolate(scale_factor=2, mode="nearest")) written by the producing
    layer's epilogue: bit-identical to the plain output replicated 2 x 2, in fp32 and in pre-split form."""
    g = torch.Generator().manual_seed(8)
    N, H, W, cin, cout = 3, 15, 20, 64, 64
    x = torch.randn(N, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    pc = packing.pack_conv2d(w, torch.ones(cout), torch.randn(cout, generator=g) / 3, "relu", DEV)[0]
    x4 = ops.nchw_to_vol4(x.to(DEV))
    plain = ops.conv_planar(pc, x4, torch.empty(cout // 4, N, H, W, 4, device=DEV), out_split=split)
    up = ops.conv_planar(pc, x4, torch.full((cout // 4, N, 2 * H, 2 * W, 4), float("nan"), device=DEV), out_split=split, out_up2=True)
    want = plain[:, :, :, None, :, None, :].expand(-1, -1, -1, 2, -1, 2, -1).reshape(cout // 4, N, 2 * H, 2 * W, 4)
    assert torch.equal(up, want)
    if not split:
        ref = F.interpolate(ops.vol4_to_nchw(plain), scale_factor=2, mode="nearest")
        assert torch.equal(ops.vol4_to_nchw(up), ref)


@pytest.mark.parametrize("precision", ["3xf16r", "3xf16r2", "3xf16r2d"])
def test_fused_logit_head_equals_head_conv_then_1x1x1(precision):
    """stereo_head: Conv3d(16,16,3)+BN+ReLU then Conv3d(16,1,1,bias) (hybrid_depth_decoder.py:104-112): the logit computed in the
    convolution's epilogue equals the two-kernel path (hidden volume + head_softargmin's own dot product)."""
    g = torch.Generator().manual_seed(6)
    D, H, W = 9, 24, 40
    x = to_vol4(torch.randn(16, D, H, W, generator=g)).to(DEV)
    w = torch.randn(16, 16, 3, 3, 3, generator=g) / (16 * 27) ** 0.5
    pc = packing.attach_tc(ops.PackedConv(packing.pack_weight(w, list(range(16)), list(range(16))), torch.rand(16, generator=g) + 0.5,
                                          torch.randn(16, generator=g) / 3, 4, 16, 4, 16, "relu", "relu")).to(DEV)
    hw, hb = (3.0 * torch.randn(16, generator=g)).to(DEV), torch.randn(1, generator=g).to(DEV)
    dv = (torch.arange(D, dtype=torch.float32) * (9.9 / (D - 1)) + 0.1).to(DEV)
    hidden = ops.conv3d(pc, x, torch.empty_like(x), precision=precision)
    logits_a = torch.empty(D, H, W, device=DEV)
    depth_a, prob_a = torch.empty(4 * H, 4 * W, device=DEV), torch.empty(4 * H, 4 * W, device=DEV)
    ops.head_softargmin(dv, hidden=hidden, head_w=hw, head_b=hb, logits_out=logits_a, depth_out=depth_a, prob_out=prob_a, up=4)
    logits_b = torch.full((D, H, W), float("nan"), device=DEV)
    ops.conv3d(pc, x, None, precision=precision, head=(hw, hb, logits_b))
    depth_b, prob_b = torch.empty_like(depth_a), torch.empty_like(prob_a)
    ops.head_softargmin(dv, logits_in=logits_b, depth_out=depth_b, prob_out=prob_b, up=4)
    want = (from_vol4(hidden.cpu()) * hw.cpu().view(16, 1, 1, 1)).sum(0) + hb.cpu()
    assert (logits_b.cpu() - want).abs().max().item() < 2e-5
    assert (logits_b - logits_a).abs().max().item() < 2e-5
    assert (depth_b - depth_a).abs().max().item() < 1e-4 and (prob_b - prob_a).abs().max().item() < 1e-5


@pytest.mark.parametrize("shape", [(2, 64, 96), (3, 37, 53)], ids=["even", "odd"])
def test_stem_conv_matches_conv2d_bn_relu(shape):
    """psm_submodule.py:42-44: convbn(3, 32, 3, stride 2, pad 1) + ReLU, straight from NCHW images into vol4 (fp32 and pre-split)."""
    N, H, W = shape
    g = torch.Generator().manual_seed(9)
    img = torch.rand(N, 3, H, W, generator=g) * 2 - 1
    w = torch.randn(32, 3, 3, 3, generator=g) / 27 ** 0.5
    b = torch.randn(32, generator=g) / 3
    want = F.relu(F.conv2d(img, w, b, stride=2, padding=1))
    got = ops.stem_conv(img.to(DEV), w.to(DEV), b.to(DEV))
    assert tuple(got.shape) == (8, N, want.shape[2], want.shape[3], 4)
    assert maxdiff(ops.vol4_to_nchw(got).cpu(), want) < 2e-6
    split = ops.stem_conv(img.to(DEV), w.to(DEV), b.to(DEV), out_split=True)
    assert torch.equal(split, ops.to_split(got))
    ops.check_status(torch.device(DEV))


@pytest.mark.parametrize("shape", [(2, 64, 128), (3, 37, 53), (1, 10, 130)], ids=["even", "odd", "wide"])
def test_stem7_conv_matches_conv2d_bn_relu(shape):
    """torchvision ResNet conv1 + bn1 + relu (resnet_encoder.py:40-51): Conv2d(3, 64, 7, stride 2, pad 3) with the BN folded in, straight
    from NCHW images into vol4 (fp32 and pre-split); tiles of 5 x 64 output pixels, so ragged and multi-tile sizes are covered."""
    N, H, W = shape
    g = torch.Generator().manual_seed(10)
    img = torch.rand(N, 3, H, W, generator=g) * 2 - 1
    w = torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5
    b = torch.randn(64, generator=g) / 3
    want = F.relu(F.conv2d(img.double(), w.double(), b.double(), stride=2, padding=3)).float()
    wt = w.permute(1, 2, 3, 0).contiguous().to(DEV)                # tap-major [3, 7, 7, 64]
    got = ops.stem7_conv(img.to(DEV), wt, b.to(DEV))
    assert tuple(got.shape) == (16, N, want.shape[2], want.shape[3], 4)
    assert maxdiff(ops.vol4_to_nchw(got).cpu(), want) < 5e-6
    split = ops.stem7_conv(img.to(DEV), wt, b.to(DEV), out_split=True)
    assert torch.equal(split, ops.to_split(got))
    ops.check_status(torch.device(DEV))


@pytest.mark.parametrize("shape", [(2, 16, 24), (3, 37, 53)], ids=["even", "odd"])
def test_maxpool3x3s2_vol4_matches_max_pool2d(shape):
    """torchvision ResNet maxpool = MaxPool2d(3, stride 2, pad 1) over vol4 maps, fp32 and pre-split on either side (bit-exact: a
    maximum of stored values; the pre-split form holds 22 significant bits of each)."""
    N, H, W = shape
    g = torch.Generator().manual_seed(12)
    x = torch.randn(N, 16, H, W, generator=g)
    want = F.max_pool2d(x, 3, stride=2, padding=1)
    x4 = ops.nchw_to_vol4(x.to(DEV))
    got = ops.maxpool3x3s2_vol4(x4)
    assert torch.equal(ops.vol4_to_nchw(got).cpu(), want)
    xs = ops.to_split(x4)
    want_s = F.max_pool2d(ops.vol4_to_nchw(ops.from_split(xs)).cpu(), 3, stride=2, padding=1)      # the values the split form holds
    got_s = ops.maxpool3x3s2_vol4(xs, in_split=True)
    assert torch.equal(ops.vol4_to_nchw(got_s).cpu(), want_s)
    got_ss = ops.maxpool3x3s2_vol4(xs, in_split=True, out_split=True)
    assert torch.equal(ops.from_split(got_ss), got_s)
    assert torch.equal(ops.maxpool3x3s2_vol4(x4, out_split=True), ops.to_split(got))
    ops.check_status(torch.device(DEV))
