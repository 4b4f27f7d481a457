"""Parameter folding and packing for the CUDA kernels (pure tensor reshuffling; testable on CPU).

* eval-mode BatchNorm3d -> per-channel (scale, shift):  s = gamma / sqrt(var + eps),  b = beta - mean * s
  (networks/layers_op.py:16-39 put a BatchNorm3d after every bias-free Conv3d; eps = 1e-5).
* pre0 (1x1x1, 64->32, hybrid_models/model_hybrid.py:58) splits into the two 32x32 matrices applied to the target
  and source feature maps at 2-D resolution: ``cat([ref_volume, warped])`` puts the target
  features in input channels 0..31 and the warped source in 32..63 (model_hybrid.py:93).
* 3x3x3 weights [Cout,Cin,3,3,3] -> [27][cin_pad][cout_pad] with tap = (kd*3+kh)*3+kw, with the channel
  permutations that the vol4 segment layout of each layer implies (see ``LAYER_PLAN``).

The 33-channel tensors of the reference (context map concatenated as channel 0, hybrid_depth_decoder.py:195) live
in a 36-channel "canonical" order here: [ref channels 1..32 | ref channel 0 | 3 zero pads] so that the 32 matching
channels stay chunk aligned and the context channel is a separate 1-chunk tensor.
"""
import copy

import torch

from .ops import PackedConv

BN_EPS = 1e-5


def fold_bn(sd, prefix, eps=BN_EPS):
    """-> (scale, shift) of the eval-mode BatchNorm at ``prefix``."""
    var = sd[prefix + ".running_var"].to(torch.float64)
    mean = sd[prefix + ".running_mean"].to(torch.float64)
    gamma = sd[prefix + ".weight"].to(torch.float64)
    beta = sd[prefix + ".bias"].to(torch.float64)
    scale = gamma / torch.sqrt(var + eps)
    shift = beta - mean * scale
    return scale.to(torch.float32), shift.to(torch.float32)


def split_pre0(sd):
    """-> (W_ref [32,32], W_src [32,32], bias [32]) with the BN scale folded into the matrices."""
    w = sd["pre0.0.weight"].reshape(sd["pre0.0.weight"].shape[0], -1).to(torch.float64)       # [32, 64]
    scale, shift = fold_bn(sd, "pre0.1")
    w = w * scale.to(torch.float64).unsqueeze(1)
    half = w.shape[1] // 2
    return (w[:, :half].to(torch.float32).contiguous(), w[:, half:].to(torch.float32).contiguous(),
            shift.contiguous())


def pack_weight(weight, cin_order, cout_order):
    """[Cout,Cin,3,3,3] -> [27, len(cin_order), len(cout_order)]; order entries are reference channel indices, -1 = zero pad."""
    cout, cin = weight.shape[0], weight.shape[1]
    w = weight.reshape(cout, cin, 27).permute(2, 1, 0)                     # [27, Cin, Cout]
    out = torch.zeros(27, len(cin_order), len(cout_order), dtype=torch.float32, device=weight.device)
    ci_dst = [i for i, c in enumerate(cin_order) if c >= 0]
    ci_src = [c for c in cin_order if c >= 0]
    co_dst = [i for i, c in enumerate(cout_order) if c >= 0]
    co_src = [c for c in cout_order if c >= 0]
    sub = w[:, ci_src][:, :, co_src]
    out[:, torch.tensor(ci_dst).unsqueeze(1), torch.tensor(co_dst).unsqueeze(0)] = sub.to(torch.float32)
    return out.contiguous()


TF32_MASK = -8192          # 0xFFFFE000 as int32: keeps sign, exponent and the 10 explicit TF32 mantissa bits
AFFINE_PAD = 48            # scale/shift arrays are padded to the widest cout_pad of either kernel


def tc_cout_pad(cout_pad):
    """N of the tensor-core kernel: a multiple of 16 (UMMA M=128 needs N % 16 == 0)."""
    return (cout_pad + 15) // 16 * 16


def pack_weight_tc(packed, cout_pad_tc=None):
    """SIMT packing [27][cin_pad][cout_pad] -> tcgen05 packing [3 dd][nks][9 taps][2 k-halves][2*C rows][4] (fp32).

    Rows 0..C-1 hold w_hi = w truncated to TF32, rows C..2C-1 hold w_lo = (w - w_hi) truncated to TF32; the 4 floats
    of a row are input channels 8*ks + 4*khalf + (0..3); K is zero padded to a multiple of 8 channels.
    """
    taps, cin_pad, cout_pad = packed.shape
    C = tc_cout_pad(cout_pad) if cout_pad_tc is None else cout_pad_tc
    nks = (cin_pad + 7) // 8
    w = torch.zeros(27, 8 * nks, C, dtype=torch.float32, device=packed.device)
    w[:, :cin_pad, :cout_pad] = packed
    hi = (w.view(torch.int32) & TF32_MASK).view(torch.float32)
    lo = ((w - hi).view(torch.int32) & TF32_MASK).view(torch.float32)

    def arrange(x):      # [27 = dd*9+tap9][8*nks = ks*8+k2*4+e][C] -> [dd][ks][tap9][k2][C][e]
        return x.reshape(3, 9, nks, 2, 4, C).permute(0, 2, 1, 3, 5, 4)

    return torch.cat([arrange(hi), arrange(lo)], dim=4).contiguous()


def pack_weight_f16(packed, cout_pad_tc=None):
    """SIMT packing -> fp16-split tcgen05 packing [3 dd][nks][9 taps][2 K-groups][2*C rows][8 x fp16], returned as a
    float32-typed byte buffer ([...,4]) plus the power-of-two exponent k the weights were scaled by.

    w * 2^k = w_hi + w_lo with both parts NORMAL fp16 numbers (k puts max|w| just below 1024, so w_lo ~ 2^-11 w_hi
    stays far above the fp16 subnormal threshold for every weight that matters); the kernel's output must be
    multiplied by 2^-k, which ``attach_tc`` folds into the per-channel scale.  K is zero padded to 16 channels.
    """
    taps, cin_pad, cout_pad = packed.shape             # taps = 27 (3x3x3) or 9 (planar 3x3)
    C = tc_cout_pad(cout_pad) if cout_pad_tc is None else cout_pad_tc
    nks = (cin_pad + 15) // 16
    w = torch.zeros(taps, 16 * nks, C, dtype=torch.float32, device=packed.device)
    w[:, :cin_pad, :cout_pad] = packed
    wmax = float(w.abs().max())
    k = 0 if wmax == 0.0 else max(-14, min(24, int(torch.floor(torch.log2(torch.tensor(1023.0 / wmax))))))
    ws = w * (2.0 ** k)
    hi = ws.to(torch.float16)
    lo = (ws - hi.to(torch.float32)).to(torch.float16)

    tpp = 9 if taps % 9 == 0 else taps                 # taps per plane (a 1x1 convolution has one)

    def arrange(x):      # [taps = dd*9+tap9][16*nks = ks*16+kg*8+e][C] -> [dd][ks][tap9][kg][C][e]
        return x.reshape(taps // tpp, tpp, nks, 2, 8, C).permute(0, 2, 1, 3, 5, 4)

    both = torch.cat([arrange(hi), arrange(lo)], dim=4)                       # [planes][nks][9][2][2C][8] fp16
    return both.contiguous().view(torch.float32), k


# (16-channel k-steps, cout_pad) the plane-ring kernel (conv3d_ring.cu) is specialised for
RING_SHAPES = ((2, 32), (3, 32), (1, 16), (2, 16), (3, 48))


def pack_weight_ring(packed, cout_pad_tc=None, scale=None):
    """SIMT packing [27][cin_pad][cout_pad] -> plane-ring packing of conv3d_ring.cu,
    [3 rotations][nks][9 taps][hi,lo][2 K-groups][3*C rows][8 x fp16] (returned as a float32-typed byte buffer) + the
    power-of-two exponent of ``pack_weight_f16``.

    Row ``slot*C + c`` of rotation ``r`` holds depth tap ``kd = (r - slot + 1) mod 3``: while input plane z (r = z mod 3)
    is stationary, ring slot ``j`` accumulates output plane ``z + 1 - kd`` == j (mod 3).
    """
    taps, cin_pad, cout_pad = packed.shape
    assert taps == 27
    C = tc_cout_pad(cout_pad) if cout_pad_tc is None else cout_pad_tc
    nks = (cin_pad + 15) // 16
    w = torch.zeros(27, 16 * nks, C, dtype=torch.float32, device=packed.device)
    w[:, :cin_pad, :cout_pad] = packed
    if scale is not None:
        # the per-channel multiplier (folded BN scale) goes into the weights: the ring kernels' epilogue applies one
        # uniform factor 2^-k and the per-channel offset only
        w[:, :, :cout_pad] = w[:, :, :cout_pad] * scale[:cout_pad].to(device=w.device, dtype=torch.float32).view(1, 1, -1)
    wmax = float(w.abs().max())
    k = 0 if wmax == 0.0 else max(-14, min(24, int(torch.floor(torch.log2(torch.tensor(1023.0 / wmax))))))
    ws = w * (2.0 ** k)
    hi = ws.to(torch.float16)
    lo = (ws - hi.to(torch.float32)).to(torch.float16)

    def arrange(x):      # [kd*9+tap9][ks*16+kg*8+e][C] -> [kd][ks][tap9][kg][C][e]
        return x.reshape(3, 9, nks, 2, 8, C).permute(0, 2, 1, 3, 5, 4)

    parts = torch.stack([arrange(hi), arrange(lo)], dim=3)                    # [kd][ks][tap9][prod][kg][C][e]
    rots = []
    for r in range(3):
        slots = [parts[(r - j + 1) % 3] for j in range(3)]                    # each [ks][tap9][prod][kg][C][e]
        rots.append(torch.cat(slots, dim=4))                                  # [ks][tap9][prod][kg][3C][e]
    return torch.stack(rots, dim=0).contiguous().view(torch.float32), k


RING2_SHAPES = ((2, 32), (3, 32), (1, 16), (2, 16), (3, 48))


RING2_COMPACT = {(3, 48): 33}    # (k-steps, cout_pad_tc) -> columns per ring slot when the layer's real width allows fewer


def pack_weight_ring2(packed, cout_pad_tc=None, scale=None, cslot=None, dual=False):
    """SIMT packing [27][cin_pad][cout_pad] -> CTA-pair ring packing of conv3d_ring2.cu,
    [7 live-tap masks][3 rotations][nks][2 CTAs][9 taps][hi,lo][2 K-groups][3*C/2 rows][8 x fp16] (float32-typed bytes)
    + the power-of-two exponent of ``pack_weight_f16``.

    Rotation r orders the rows by ring slot exactly like ``pack_weight_ring`` (slot j <- depth tap (r - j + 1) mod 3);
    variant ``mask - 1`` zeroes the taps whose bit is clear in ``mask`` (bit kd: output plane z + 1 - kd belongs to the
    CTA pair's range), so that partial first / last planes issue the same full-N MMA; CTA 0 of the pair holds rows
    [0, 3C/2), CTA 1 rows [3C/2, 3C).

    ``cslot`` (default C): rows per ring slot.  A layer with fewer real output channels than its padded width packs only
    ``cslot`` rows per slot and pads the TOTAL to a multiple of 16 -- the 33-channel layer dres2 (hybrid_depth_decoder.py:90)
    gets N = 112 instead of 3 x 48 = 144: a fifth fewer tensor-core cycles, and 4 M tiles instead of 2 fit TMEM.

    ``dual`` (two accumulators per ring slot, precision 3xf16r2d): per tap, CTA 0 holds a "merged" block with ALL N rows of w_hi
    and CTA 1 one with all N rows of w_lo -- the two halves of the B operand [w_hi | w_lo] of ONE MMA with 2N columns (large
    accumulator | small accumulator) -- followed by a "third" block with that CTA's half of w_hi for the x_lo w_hi product:
    [7 masks][3 rotations][nks][2 CTAs][9 taps]{[2 K-groups][N rows], [2 K-groups][N/2 rows]}[8 x fp16]."""
    taps, cin_pad, cout_pad = packed.shape
    assert taps == 27
    C = tc_cout_pad(cout_pad) if cout_pad_tc is None else cout_pad_tc
    nks = (cin_pad + 15) // 16
    w = torch.zeros(27, 16 * nks, C, dtype=torch.float32, device=packed.device)
    w[:, :cin_pad, :cout_pad] = packed
    if scale is not None:
        # the per-channel multiplier (folded BN scale) goes into the weights: the ring kernels' epilogue applies one
        # uniform factor 2^-k and the per-channel offset only
        w[:, :, :cout_pad] = w[:, :, :cout_pad] * scale[:cout_pad].to(device=w.device, dtype=torch.float32).view(1, 1, -1)
    wmax = float(w.abs().max())
    k = 0 if wmax == 0.0 else max(-14, min(24, int(torch.floor(torch.log2(torch.tensor(1023.0 / wmax))))))
    ws = w * (2.0 ** k)
    hi = ws.to(torch.float16)
    lo = (ws - hi.to(torch.float32)).to(torch.float16)

    def arrange(x):      # [kd*9+tap9][ks*16+kg*8+e][C] -> [kd][ks][tap9][kg][C][e]
        return x.reshape(3, 9, nks, 2, 8, C).permute(0, 2, 1, 3, 5, 4)

    parts = torch.stack([arrange(hi), arrange(lo)], dim=3)                    # [kd][ks][tap9][prod][kg][C][e]
    cslot = C if cslot is None else int(cslot)
    assert cout_pad <= cslot <= C
    parts = parts[:, :, :, :, :, :cslot]
    zero = torch.zeros_like(parts[0])
    n_rows = (3 * cslot + 15) // 16 * 16                                      # N of the MMA
    NH = n_rows // 2
    pad = torch.zeros(nks, 9, 2, 2, n_rows - 3 * cslot, 8, dtype=parts.dtype, device=parts.device)
    variants = []
    for mask in range(1, 8):
        rots = []
        for r in range(3):
            slots = [parts[(r - j + 1) % 3] if (mask >> ((r - j + 1) % 3)) & 1 else zero for j in range(3)]
            full = torch.cat(slots + [pad], dim=4)                            # [ks][tap9][prod][kg][N][e]
            if dual:
                hi_rows, lo_rows = full[:, :, 0], full[:, :, 1]                # [ks][tap9][kg][N][e]
                per_cta = []
                for cta in range(2):
                    merged = (hi_rows if cta == 0 else lo_rows).reshape(nks, 9, 2 * n_rows, 8)
                    third = hi_rows[:, :, :, cta * NH:(cta + 1) * NH].reshape(nks, 9, 2 * NH, 8)
                    per_cta.append(torch.cat([merged, third], dim=2))          # [ks][tap9][2N + 2NH rows][e]
                rots.append(torch.stack(per_cta, dim=1))                       # [ks][cta][tap9][rows][e]
                continue
            halves = full.reshape(nks, 9, 2, 2, 2, NH, 8).permute(0, 4, 1, 2, 3, 5, 6)      # [ks][half][tap9][prod][kg][NH][e]
            rots.append(halves)
        variants.append(torch.stack(rots, dim=0))
    return torch.stack(variants, dim=0).contiguous().view(torch.float32), k


def attach_tc(pc):
    """Adds the tensor-core packings (3xTF32 and fp16-split) to a PackedConv and pads its affine arrays."""
    pc.cout_pad_tc = tc_cout_pad(pc.cout_pad)
    pc.weight_tc = pack_weight_tc(pc.weight, pc.cout_pad_tc)
    pc.weight_f16, k = pack_weight_f16(pc.weight, pc.cout_pad_tc)
    for name in ("scale", "shift"):
        v = getattr(pc, name)
        if v.numel() < AFFINE_PAD:
            setattr(pc, name, torch.cat([v, torch.zeros(AFFINE_PAD - v.numel(), dtype=v.dtype, device=v.device)]).contiguous())
    pc.scale_f16 = (pc.scale * (2.0 ** -k)).contiguous()
    nks = (pc.weight.shape[1] + 15) // 16
    if pc.weight.shape[0] == 27 and (nks, pc.cout_pad_tc) in RING2_SHAPES:
        cslot = RING2_COMPACT.get((nks, pc.cout_pad_tc))
        if cslot is not None and pc.cout <= cslot and pc.weight.shape[2] <= cslot + 7:
            # narrow layer: `cslot` columns per ring slot (conv3d_ring2.cu Shape<3, 33, 4>); the kernel is selected by cout_pad
            pc.cout_pad_ring2 = cslot
            pc.weight_ring2, k_ring2 = pack_weight_ring2(pc.weight[:, :, :cslot].contiguous(), pc.cout_pad_tc, pc.scale, cslot=cslot)
            pc.weight_ring2d, _ = pack_weight_ring2(pc.weight[:, :, :cslot].contiguous(), pc.cout_pad_tc, pc.scale, cslot=cslot, dual=True)
        else:
            pc.weight_ring2, k_ring2 = pack_weight_ring2(pc.weight, pc.cout_pad_tc, pc.scale)
            if pc.cout_pad_tc <= 32:                      # two accumulators per slot fit TMEM (and N = 6 * cout_pad <= 256)
                pc.weight_ring2d, _ = pack_weight_ring2(pc.weight, pc.cout_pad_tc, pc.scale, dual=True)
    if pc.weight.shape[0] == 27 and (nks, pc.cout_pad_tc) in RING_SHAPES:
        pc.weight_ring, k_ring = pack_weight_ring(pc.weight, pc.cout_pad_tc, pc.scale)
        assert pc.weight_ring2 is None or k_ring2 == k_ring
        # uniform multiplier of the ring kernels (read as scale[0]): undoes the power-of-two weight scaling
        pc.scale_ring = torch.full((AFFINE_PAD,), 2.0 ** -k_ring, dtype=torch.float32, device=pc.weight.device)
    return pc


def pack_conv2d(weight, scale, shift, act, device, cout_slice=64):
    """2-D 3x3 conv [Cout,Cin,3,3] with folded per-channel affine -> list with ONE PackedConv for conv2d_tc.cu.

    All arithmetic happens on the HOST (the parameters are copied back once if they live on the GPU) and only the packed
    buffers are uploaded: packing on the device cost ~15 tiny torch kernels per layer, i.e. about a thousand launches before
    a model's first real kernel.

    The planar tensor-core kernel is specialised for 64 / 32 / 16 output channels per accumulator; wider layers are packed
    as Cout/64 slices [slice][nks][9][2][128 rows][16 B] that the kernel runs as independent units of one launch
    (cout_pad = 64 * slices).  (A list is returned for the callers that iterate over per-launch pieces.)"""
    cout, cin = weight.shape[0], weight.shape[1]
    host = torch.device("cpu")
    weight, scale, shift = weight.detach().to(host), scale.detach().to(host), shift.detach().to(host)
    if cout > cout_slice:
        assert cout_slice == 64, "only 64-channel slices can be combined in one launch"
        pieces = _pack_conv2d_slices(weight, scale, shift, act, host, 64)
        total = 64 * len(pieces)
        pc = PackedConv(None, torch.cat([q.scale for q in pieces]), torch.cat([q.shift for q in pieces]), (cin + 3) // 4, total,
                        (cout + 3) // 4, total, act, act, cin=cin, cout=cout, cout_pad_tc=total)
        pc.weight_f16 = torch.cat([q.weight_f16.reshape(-1) for q in pieces]).contiguous()
        pc.scale_f16 = torch.cat([q.scale_f16 for q in pieces]).contiguous()
        return [pc.to(device)]
    return [q.to(device) for q in _pack_conv2d_slices(weight, scale, shift, act, host, cout_slice)]


def _pack_conv2d_slices(weight, scale, shift, act, device, cout_slice):
    cout, cin = weight.shape[0], weight.shape[1]
    out = []
    for c0 in range(0, cout, cout_slice):
        n = min(cout_slice, cout - c0)
        taps = weight.shape[2] * weight.shape[3]           # 9 (3x3) or 1 (1x1)
        w = weight[c0:c0 + n].reshape(n, cin, taps).permute(2, 1, 0).to(device=device, dtype=torch.float32).contiguous()   # [taps][Cin][n]
        # the per-channel multiplier is folded into the weights: the kernel applies one factor per slice (2^-k) + the offsets
        w = w * scale[c0:c0 + n].to(device=device, dtype=torch.float32).view(1, 1, -1)
        wf16, k = pack_weight_f16(w, cout_slice)
        s = torch.full((cout_slice,), 2.0 ** -k, dtype=torch.float32, device=device)
        b = torch.zeros(cout_slice, dtype=torch.float32, device=device)
        b[:n] = shift[c0:c0 + n].to(device)
        pc = PackedConv(None, s, b, (cin + 3) // 4, cout_slice, (n + 3) // 4, cout_slice, act, act, cin=cin, cout=n,
                        cout_pad_tc=cout_slice)
        pc.weight_f16, pc.scale_f16 = wf16, s
        out.append(pc)
    return out


def _affine(scale, shift, cout_order, device):
    s = torch.zeros(len(cout_order), dtype=torch.float32, device=device)
    b = torch.zeros(len(cout_order), dtype=torch.float32, device=device)
    for i, c in enumerate(cout_order):
        if c >= 0:
            s[i] = scale[c]
            b[i] = shift[c]
    return s, b


CANON36 = list(range(1, 33)) + [0, -1, -1, -1]            # canonical order of the 33-channel tensors


def pack_layers(sd, device):
    """All 3-D layers of the hot path -> dict name -> PackedConv (tensors on ``device``).

    Folding / permuting / splitting runs on the host; the packed buffers are uploaded once at the end (see pack_conv2d)."""
    target = torch.device(device)
    device = torch.device("cpu")
    sd = {k: v.detach().to(device) for k, v in sd.items()}
    def conv_bn(prefix, cin_order, cout_order, cout_pad, act_split, act_lo, act_hi, out_chunks):
        w = sd[prefix + ".0.weight"].to(device)
        scale, shift = fold_bn(sd, prefix + ".1")
        order = list(cout_order) + [-1] * (cout_pad - len(cout_order))
        s, b = _affine(scale, shift, order, device)
        return PackedConv(pack_weight(w, cin_order, order), s, b, len(cin_order) // 4, cout_pad, out_chunks,
                          act_split, act_lo, act_hi, cin=w.shape[1], cout=w.shape[0])

    def conv_bias(prefix, cout_pad, act_split, out_chunks):
        w = sd[prefix + ".weight"].to(device)
        order = list(range(w.shape[0]))
        s = torch.ones(cout_pad, dtype=torch.float32, device=device)
        b = sd[prefix + ".bias"].to(device=device, dtype=torch.float32).contiguous()
        return PackedConv(pack_weight(w, list(range(w.shape[1])), order), s, b, w.shape[1] // 4, cout_pad, out_chunks,
                          act_split, "none", "none")

    r32, r16 = list(range(32)), list(range(16))
    layers = {}
    layers["pre1"] = conv_bn("pre1", r32, r32, 32, 32, "relu", "relu", 8)
    layers["pre2"] = conv_bn("pre2", r32, r32, 32, 32, "none", "none", 8)
    for name in ("dres0.0", "dres0.1", "dres1.0", "dres1.1"):
        layers[name] = conv_bn("CostRegNet." + name, r32, r32, 32, 32, "relu", "relu", 8)
    layers["dres2"] = conv_bn("CostRegNet.dres2.0", CANON36, CANON36[:33], 40, 40, "relu", "relu", 9)
    # value (tanh) and key (ReLU) share their input: one 33->32 layer, outputs split into two tensors
    wv, wk = sd["CostRegNet.value_layer.0.0.weight"].to(device), sd["CostRegNet.key_layer.0.0.weight"].to(device)
    sv, bv = fold_bn(sd, "CostRegNet.value_layer.0.1")
    sk, bk = fold_bn(sd, "CostRegNet.key_layer.0.1")
    layers["value_key"] = PackedConv(pack_weight(torch.cat([wv, wk], 0), CANON36, r32),
                                     torch.cat([sv, sk]).to(device).contiguous(), torch.cat([bv, bk]).to(device).contiguous(),
                                     9, 32, 8, 16, "tanh", "relu", cin=33, cout=32)
    for i in (0, 1):
        layers["head%d" % i] = conv_bn("CostRegNet.stereo_head%d.0" % i, r16, r16, 16, 16, "relu", "relu", 4)
        layers["head%d_w" % i] = sd["CostRegNet.stereo_head%d.1.weight" % i].to(device=device, dtype=torch.float32).reshape(-1).contiguous()
        layers["head%d_b" % i] = sd["CostRegNet.stereo_head%d.1.bias" % i].to(device=device, dtype=torch.float32).reshape(-1).contiguous()
    est = "CostRegNet.epipolar_transformer"
    if (est + ".gate_conv.weight") in sd:
        layers["gate"] = conv_bias(est + ".gate_conv", 32, 16, 8)
        layers["output"] = conv_bias(est + ".output_conv", 16, 16, 4)
        for gate, key in (("r", "reset_gate_norm"), ("u", "update_gate_norm"), ("o", "output_norm")):
            layers["gn_%s_w" % gate] = sd["%s.%s.weight" % (est, key)].to(device=device, dtype=torch.float32).contiguous()
            layers["gn_%s_b" % gate] = sd["%s.%s.bias" % (est, key)].to(device=device, dtype=torch.float32).contiguous()
    for v in layers.values():
        if isinstance(v, PackedConv):
            attach_tc(v)
    # pre2 applied ONCE to the sum of both sources' pre1 outputs (conv + eval-BN is affine: pre2(a) + pre2(b) =
    # s * W (a + b) + 2 b): the same packed weights with the offset doubled
    layers["pre2_pair"] = copy.copy(layers["pre2"])
    layers["pre2_pair"].shift = (layers["pre2"].shift * 2.0).contiguous()
    layers["pre2_pair"]._desc = None              # descriptor templates hold the offsets' address: never shared with pre2
    w_ref, w_src, bias = split_pre0(sd)
    layers["pre0_ref"], layers["pre0_src"], layers["pre0_bias"] = w_ref.to(device), w_src.to(device), bias.to(device)
    # both halves stacked: one premix launch per sequence gives every frame's target-side (chunks 0..7) and source-side mix
    layers["pre0_both"] = torch.cat([layers["pre0_ref"], layers["pre0_src"]], 0).contiguous()
    layers["pre0_both_bias"] = torch.cat([layers["pre0_bias"], torch.zeros_like(layers["pre0_bias"])]).contiguous()
    if target.type != "cpu":
        memo = {}
        for name, v in layers.items():
            layers[name] = v.to(target, memo) if isinstance(v, PackedConv) else _upload(v, target, memo)
    return layers


def _upload(t, device, memo=None):
    """Host tensor -> device, once per distinct tensor (layers made with copy.copy share their buffers)."""
    if t is None or not torch.is_tensor(t):
        return t
    if memo is None:
        return t.to(device)
    got = memo.get(id(t))
    if got is None:
        got = memo[id(t)] = (t, t.to(device))          # the source is kept alive so that its id() stays unique
    return got[1]
