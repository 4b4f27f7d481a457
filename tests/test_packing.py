"""Host-side folding/packing against the oracle's un-fused arithmetic (CPU only)."""
import torch
import torch.nn.functional as F

from estdepth_b200 import packing, synth
from oracle import estdepth_oracle as orc
from tests.helpers import from_vol4, state_template, to_vol4


def _sd():
    _, tmpl = state_template(18, 32)
    return synth.synth_state_dict(tmpl, seed=4)


def test_pre0_split_equals_conv_bn_of_concat():
    sd = _sd()
    g = torch.Generator().manual_seed(0)
    ref = torch.randn(1, 32, 5, 6, 7, generator=g)
    warped = torch.randn(1, 32, 5, 6, 7, generator=g)
    want = orc._cb3(torch.cat([ref, warped], 1), sd, "pre0")
    w_ref, w_src, bias = packing.split_pre0(sd)
    got = torch.einsum("oc,bcdhw->bodhw", w_ref, ref) + torch.einsum("oc,bcdhw->bodhw", w_src, warped) + bias.view(1, -1, 1, 1, 1)
    assert (got - want).abs().max() < 2e-5


def _emulate(pc, x):
    """conv with the packed [27][cin][cout] weights + affine, the way the kernel consumes them."""
    cin = pc.weight.shape[1]
    w = pc.weight.permute(2, 1, 0).reshape(pc.cout_pad, cin, 3, 3, 3)
    y = F.conv3d(x.unsqueeze(0), w, None, 1, 1)[0]
    return y * pc.scale[:pc.cout_pad].view(-1, 1, 1, 1) + pc.shift[:pc.cout_pad].view(-1, 1, 1, 1)


def test_packed_layers_reproduce_the_oracle_chain():
    sd = _sd()
    L = packing.pack_layers(sd, torch.device("cpu"))
    g = torch.Generator().manual_seed(1)
    D, H, W = 4, 6, 8
    m = torch.randn(32, D, H, W, generator=g)
    sem = torch.randn(D, H, W, generator=g)
    # dres2 on cat[sem, m] (reference order) vs packed layer on canonical order [m | sem,0,0,0]
    want_z = orc._cb3(torch.cat([sem.unsqueeze(0), m], 0).unsqueeze(0), sd, "CostRegNet.dres2.0", "relu")[0]      # 33 ch
    canon_in = torch.cat([m, sem.unsqueeze(0), torch.zeros(3, D, H, W)], 0)
    z = torch.relu(_emulate(L["dres2"], canon_in))[:36]
    assert (z[:32] - want_z[1:]).abs().max() < 1e-5 and (z[32] - want_z[0]).abs().max() < 1e-5
    assert z[33:].abs().max() == 0
    # fused value/key layer
    want_v = orc._cb3(want_z.unsqueeze(0), sd, "CostRegNet.value_layer.0", "tanh")[0]
    want_k = orc._cb3(want_z.unsqueeze(0), sd, "CostRegNet.key_layer.0", "relu")[0]
    vk = _emulate(L["value_key"], z)
    assert (torch.tanh(vk[:16]) - want_v).abs().max() < 1e-5 and (torch.relu(vk[16:]) - want_k).abs().max() < 1e-5
    assert L["value_key"].act_split == 16 and L["dres2"].out_chunks == 9 and L["dres2"].cin == 33
    # plain 32->32 and the biased EST convs
    x = torch.randn(32, D, H, W, generator=g)
    assert (torch.relu(_emulate(L["pre1"], x)) - orc._cb3(x.unsqueeze(0), sd, "pre1", "relu")[0]).abs().max() < 1e-5
    p = "CostRegNet.epipolar_transformer"
    want = F.conv3d(x.unsqueeze(0), sd[p + ".gate_conv.weight"], sd[p + ".gate_conv.bias"], 1, 1)[0]
    assert (_emulate(L["gate"], x) - want).abs().max() < 1e-5


def test_tensor_core_packing_is_an_exact_two_term_split():
    g = torch.Generator().manual_seed(2)
    packed = torch.randn(27, 36, 40, generator=g)
    tc = packing.pack_weight_tc(packed)                       # [3][5][9][2][96][4]
    assert tuple(tc.shape) == (3, 5, 9, 2, 96, 4)
    hi, lo = tc[..., :48, :], tc[..., 48:, :]
    # both parts are exactly representable in TF32 (13 low mantissa bits clear) ...
    assert int((hi.contiguous().view(torch.int32) & 0x1FFF).abs().max()) == 0
    assert int((lo.contiguous().view(torch.int32) & 0x1FFF).abs().max()) == 0
    # ... and hi + lo reproduces the fp32 weight to 2^-21 relative, in the kernel's (dd, ks, tap, khalf, n, e) order
    w = torch.zeros(27, 40, 48)
    w[:, :36, :40] = packed
    want = w.reshape(3, 9, 5, 2, 4, 48).permute(0, 2, 1, 3, 5, 4)
    assert ((hi + lo) - want).abs().max() <= 2.0 ** -21 * want.abs().max()
    assert torch.equal(hi[0, 0, 4, 1, 7, 2], (w[4, 6, 7].view(torch.int32) & -8192).view(torch.float32))


def test_vol4_helpers_roundtrip():
    x = torch.arange(16 * 2 * 3 * 5, dtype=torch.float32).reshape(16, 2, 3, 5)
    v = to_vol4(x)
    assert v.shape == (4, 2, 3, 5, 4) and v[1, 0, 0, 0, 2] == x[6, 0, 0, 0]
    assert torch.equal(from_vol4(v), x)


def _simulate_ring(x, ring, k, nks, C, D, grid):
    """Host model of conv3d_ring.cu's schedule for a 1-column volume: x [16*nks, D, H, W]; ring = pack_weight_ring output.
    Walks the flat plane list in `grid` contiguous ranges exactly as the kernel does (segments, partial first/last
    planes, slot masks and runs, hand-over after each plane, zeroing) and returns the [C, D, H, W] result."""
    import torch.nn.functional as F
    _, _, H, W = x.shape
    w16 = ring.view(torch.float16).reshape(3, nks, 9, 2, 2, 3 * C, 8).to(torch.float64)
    wsum = (w16[:, :, :, 0] + w16[:, :, :, 1]) * 2.0 ** -k            # hi + lo: [rot][ks][tap][kg][3C][8]
    out = torch.full((C, D, H, W), float("nan"), dtype=torch.float64)
    xs = x.to(torch.float64)
    total = D
    for b in range(grid):
        f0, f1 = total * b // grid, total * (b + 1) // grid
        if f1 <= f0:
            continue
        z0, z1 = f0, f1                                               # one column: the range is one segment
        acc = torch.zeros(3, C, H, W, dtype=torch.float64)
        for z in range(max(z0 - 1, 0), z1 + 1):
            completes = (z - 1) >= z0
            if z < D:
                o_lo, o_hi = max(z - 1, z0), min(z + 1, z1 - 1)
                mask = 0
                for o in range(o_lo, o_hi + 1):
                    mask |= 1 << (o % 3)
                runs = [(0, 1), (2, 1)] if mask == 5 else [((0 if mask & 1 else 1 if mask & 2 else 2), bin(mask).count("1"))]
                rot = z % 3
                for first, cnt in runs:
                    rows = slice(first * C, (first + cnt) * C)
                    # [ks][tap][kg][rows][8] -> conv2d weight [rows, 16*nks, 3, 3]
                    wt = wsum[rot][:, :, :, rows]                     # [ks][tap][kg][n][8]
                    wt = wt.permute(3, 0, 2, 4, 1).reshape(cnt * C, 16 * nks, 3, 3)
                    y = F.conv2d(xs[:, z].unsqueeze(0), wt, padding=1)[0]
                    acc[first:first + cnt] += y.reshape(cnt, C, H, W)
            if completes:
                slot = (z - 1) % 3
                out[:, z - 1] = acc[slot]
                acc[slot] = 0
        assert float(acc.abs().max()) == 0.0, "a ring slot was written but never handed over"
    return out


def test_ring_packing_and_schedule_reproduce_conv3d():
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(3)
    for cin, D, grids in ((32, 7, (1, 2, 3, 7)), (36, 4, (1, 4)), (32, 1, (1,)), (32, 2, (1, 2))):
        w = torch.randn(32, cin, 3, 3, 3, generator=g) / 10
        packed = packing.pack_weight(w, list(range(cin)), list(range(32)))
        ring, k = packing.pack_weight_ring(packed, 32)
        nks = (cin + 15) // 16
        assert tuple(ring.shape) == (3, nks, 9, 2, 2, 96, 4)
        x = torch.zeros(16 * nks, D, 5, 6)
        x[:cin] = torch.randn(cin, D, 5, 6, generator=g)
        want = F.conv3d(x[:cin].unsqueeze(0).double(), w.double(), padding=1)[0]
        for grid in grids:
            got = _simulate_ring(x, ring, k, nks, 32, D, grid)
            assert torch.isfinite(got).all()
            assert (got - want).abs().max().item() < 1e-5, (cin, D, grid)


def test_ring2_packing_is_the_ring_packing_split_over_the_cta_pair():
    g = torch.Generator().manual_seed(4)
    w = torch.randn(32, 36, 3, 3, 3, generator=g) / 10
    packed = packing.pack_weight(w, list(range(36)), list(range(32)))
    ring, k = packing.pack_weight_ring(packed, 32)
    ring2, k2 = packing.pack_weight_ring2(packed, 32)
    assert k == k2 and tuple(ring2.shape) == (7, 3, 3, 2, 9, 2, 2, 48, 4)
    full = ring2[6].permute(0, 1, 3, 4, 5, 2, 6, 7).reshape(3, 3, 9, 2, 2, 96, 4)         # all taps live: [rot][ks][tap][prod][kg][96][4]
    assert torch.equal(full, ring)
    # mask 1 (only depth tap 0 live): for rotation r the rows of the slots holding taps 1 and 2 are zero
    only0 = ring2[0].permute(0, 1, 3, 4, 5, 2, 6, 7).reshape(3, 3, 9, 2, 2, 96, 4)
    for r in range(3):
        for slot in range(3):
            blk = only0[r][..., 32 * slot:32 * slot + 32, :]
            if (r - slot + 1) % 3 == 0:
                assert torch.equal(blk, ring[r][..., 32 * slot:32 * slot + 32, :])
            else:
                assert float(blk.view(torch.int32).abs().max()) == 0.0


def test_conv_descriptor_templates_are_per_layer_and_per_arithmetic(monkeypatch):
    """Host logic of ops._conv_desc: the layer-constant half of estd_conv3d_desc is cached on the PackedConv as a byte
    template.  A template must never leak between layers that share packed weights (pre2 / pre2_pair differ only in the
    offsets) or between arithmetics of one layer, and the per-launch fields must be filled on a fresh copy every time."""
    import ctypes
    from estdepth_b200 import ops
    from tests.helpers import synth_model_and_state

    def host_ptr(t, dtype=torch.float32):
        if t is None:
            return None
        assert t.dtype == dtype and t.is_contiguous()
        return t.data_ptr()
    flag = torch.zeros(1, dtype=torch.int32)
    monkeypatch.setattr(ops, "_ptr", lambda t, dtype=torch.float32: None if t is None else ctypes.c_void_p(host_ptr(t, dtype)))
    monkeypatch.setattr(ops, "_act_ptr", host_ptr)
    monkeypatch.setattr(ops, "status_flag", lambda device: flag)

    model, _ = synth_model_and_state(18, 32)
    L = model._layers(torch.device("cpu"))
    pre2, pair = L["pre2"], L["pre2_pair"]
    assert pair.weight_ring2 is pre2.weight_ring2 and pair.shift is not pre2.shift
    assert torch.equal(pair.shift, 2 * pre2.shift)
    x, y, r0, r1 = (torch.zeros(8, 2, 8, 8, 4) for _ in range(4))
    d_pre2 = ops._conv_desc(pre2, x, None, y, None, r0, None, 1.0, None, "3xf16r2")
    d_pair = ops._conv_desc(pair, x, None, y, None, r0, r1, 0.5, None, "3xf16r2")
    assert d_pre2.shift == pre2.shift.data_ptr() and d_pair.shift == pair.shift.data_ptr()
    assert d_pre2.weight_tc == d_pair.weight_tc == pre2.weight_ring2.data_ptr()
    assert (d_pre2.res1, d_pre2.post_scale) == (None, 1.0) and (d_pair.res1, d_pair.post_scale) == (r1.data_ptr(), 0.5)
    # a second launch of the same layer starts from the template again, not from the previous launch's descriptor
    d_again = ops._conv_desc(pair, x, None, y, None, None, None, 1.0, None, "3xf16r2")
    assert (d_again.res0, d_again.res1, d_again.post_scale) == (None, None, 1.0)
    # another arithmetic of the same layer: its own template (weights, multiplier array, padded width)
    d_fp32 = ops._conv_desc(pre2, x, None, y, None, None, None, 1.0, None, "fp32")
    assert d_fp32.precision == ops.PRECISION["fp32"] and d_fp32.weight == pre2.weight.data_ptr()
    assert d_fp32.scale == pre2.scale.data_ptr() and d_pre2.scale == pre2.scale_ring.data_ptr()
    assert d_fp32.status is None and d_pre2.status == flag.data_ptr()
    assert set(pre2._desc) == {("3xf16r2", 0, 1, None), ("fp32", 0, 1, None)}
    # shape errors are still caught per launch
    import pytest
    with pytest.raises(RuntimeError):
        ops._conv_desc(pre2, torch.zeros(4, 2, 8, 8, 4), None, y, None, None, None, 1.0, None, "3xf16r2")
