"""2-D feeder networks of the ESTDepth hot path.

These stay on cuDNN (they are marked "kept" in the scope table): the B200-native work of this
package is the 3-D plane-sweep / matching / EST path that consumes their outputs.  The modules
below only exist so that

  * ``state_dict`` key names and shapes are identical to the reference's
    (``matchingFeature.*``   <- networks/psm_submodule.py:40-116,
     ``semanticFeature.encoder.*`` <- hybrid_models/resnet_encoder.py:17-51,
     ``CostRegNet.upconv_*`` / ``dispconv_*`` <- hybrid_models/hybrid_depth_decoder.py:56-75),
    so a checkpoint written by the reference's ``train_hybrid.py`` loads with ``strict=True``;
  * the 2-D arithmetic is the same sequence of conv / BN(eval) / ReLU / pool / resize ops.

Parameter containers are registered under the reference's attribute names; the forward code is
written against small functional helpers rather than mirroring the reference's module nesting.
"""
import logging

import numpy as np
import torch
import torch.nn as nn
import weakref

import torch.nn.functional as F
import torchvision.models as tvm


class _FoldedConv(object):
    """Eval-mode conv+BN pairs of the cuDNN-side layers run as ONE convolution with the BN affine folded into weight and
    bias (w' = w * gamma/sqrt(var+eps), b' = beta - mean * gamma/sqrt(var+eps)), optionally with cuDNN's fused
    bias(+residual)+ReLU epilogue.  Saves one full read+write of every feature map per BN and per ReLU.  Folded
    parameters are cached and re-derived when the underlying parameters change (``_version`` counters)."""

    def __init__(self):
        self.cache = {}
        self.channels_last = False      # experiment switch (profiles/bench_feeders.py): NHWC weights for NHWC activations
        self.split_tf32 = False         # experiment switch: three-term TF32 split through cuDNN's tensor-core kernels
        self.cache3 = {}

    @staticmethod
    def _split(t):
        hi = (t.view(torch.int32) & -8192).view(torch.float32)
        lo = ((t - hi).view(torch.int32) & -8192).view(torch.float32)
        return hi, lo

    def _call_split(self, x, conv, bn, relu, residual):
        w, b = self.params(conv, bn)
        ent = self.cache3.get(id(conv))
        if ent is None or ent[0] is not w:
            wh, wl = self._split(w)
            ent = (w, torch.cat([wh, wl, wh], 1).contiguous(memory_format=torch.channels_last))
            self.cache3[id(conv)] = ent
        xh, xl = self._split(x)
        x3 = torch.cat([xh, xh, xl], 1).contiguous(memory_format=torch.channels_last)
        with torch.backends.cudnn.flags(enabled=True, benchmark=True, allow_tf32=True):
            y = F.conv2d(x3, ent[1], b, conv.stride, conv.padding, conv.dilation, 1)
        if residual is not None:
            y = y + residual
        y = F.relu_(y) if relu else y
        return y.contiguous()

    def params(self, conv, bn):
        ver = (conv.weight.data_ptr(), conv.weight._version, bn.weight._version, bn.bias._version,
               bn.running_mean._version, bn.running_var._version)
        ent = self.cache.get(id(conv))
        # the entry remembers WHICH module it was derived from: this cache outlives models, and a new model's layer can get a
        # dead layer's id() -- and, through the caching allocator, its weight address and version counters too
        if ent is None or ent[0] != ver or ent[3]() is not conv:
            with torch.no_grad():
                # folded on the host in fp64, uploaded once (on the device this was ~10 tiny kernels per layer)
                dev = conv.weight.device
                s = bn.weight.detach().cpu().double() / torch.sqrt(bn.running_var.detach().cpu().double() + bn.eps)
                w = (conv.weight.detach().cpu().double() * s.view(-1, 1, 1, 1)).float().contiguous().to(dev)
                b = (bn.bias.detach().cpu().double() - bn.running_mean.detach().cpu().double() * s).float().contiguous().to(dev)
            key = id(conv)
            ent = (ver, w, b, weakref.ref(conv, lambda _, key=key, cache=self.cache: cache.pop(key, None)))
            self.cache[key] = ent
        return ent[1], ent[2]

    def __call__(self, x, conv, bn, relu=False, residual=None):
        if self.split_tf32 and x.is_cuda and conv.groups == 1:
            return self._call_split(x, conv, bn, relu, residual)
        w, b = self.params(conv, bn)
        if self.channels_last:
            w = w.contiguous(memory_format=torch.channels_last)
            if residual is not None:
                residual = residual.contiguous(memory_format=torch.channels_last)
        if x.is_cuda and relu and conv.groups == 1:
            if residual is not None:
                return torch.cudnn_convolution_add_relu(x, w, residual, 1.0, b, conv.stride, conv.padding, conv.dilation, 1)
            return torch.cudnn_convolution_relu(x, w, b, conv.stride, conv.padding, conv.dilation, 1)
        y = F.conv2d(x, w, b, conv.stride, conv.padding, conv.dilation, conv.groups)
        if residual is not None:
            y = y + residual
        return F.relu_(y) if relu else y


_folded = _FoldedConv()
import os as _os  # noqa: E402
_folded.split_tf32 = _os.environ.get("ESTD_FEEDER_SPLIT_TF32", "0") == "1"


def _bn_affine(bn):
    """Eval-mode BatchNorm -> per-channel (scale, shift), computed on the host in fp64."""
    scale = bn.weight.detach().cpu().double() / torch.sqrt(bn.running_var.detach().cpu().double() + bn.eps)
    shift = bn.bias.detach().cpu().double() - bn.running_mean.detach().cpu().double() * scale
    return scale.float(), shift.float()


def _pack3x3(conv, bn, act, device, any_stride=False):
    """conv(3x3 or 1x1, stride 1)+BN (bn=None: conv with bias) -> packed planar tensor-core layer (packing.pack_conv2d), or
    None when the shape is not one the kernel takes (input channels must come in whole 16-channel k-steps).
    any_stride: also pack stride-2 layers -- the caller runs them at stride 1 / on a subsampled input and keeps every other
    row and column (a stride-2 convolution with this padding IS the stride-1 result at the even positions)."""
    from . import packing
    w = conv.weight.detach()
    if conv.kernel_size not in ((3, 3), (1, 1)) or (conv.stride != (1, 1) and not (any_stride and conv.stride == (2, 2))) \
            or conv.padding != ((conv.kernel_size[0] // 2) * conv.dilation[0],) * 2 or conv.groups != 1 or w.shape[1] % 16 \
            or conv.dilation not in ((1, 1), (2, 2)):
        return None
    if bn is not None:
        s, b = _bn_affine(bn)
    else:
        s = torch.ones(w.shape[0])
        b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0])
    cout = w.shape[0]
    return packing.pack_conv2d(w, s, b, act, device, cout_slice=64 if cout > 32 else 32 if cout > 16 else 16)


SPLIT_ACTIVATIONS = _os.environ.get("ESTD_SPLIT_ACT", "1") != "0"


def _is_split(t):
    """True when a vol4 tensor holds PRE-SPLIT activations (vol4s, include/estdepth_b200.h: chunk pairs = x_hi | x_lo of 8
    channels as fp16).  The form is a tag on the tensor object; every helper below that makes a new tensor from a tagged
    one copies the tag."""
    return t is not None and bool(getattr(t, "_estd_split", False))


def _tag(t, split):
    t._estd_split = bool(split)
    return t


def _run3x3(pcs, x4, out4=None, res4=None, dilation=1, in1=None, taps=9, post_scale=1.0, out_split=True, up2=False):
    """One packed planar layer over vol4 maps [C/4, N, H, W, 4] (+ optional second input segment = torch.cat on channels).

    Activations between planar layers travel PRE-SPLIT (``out_split``, the default): the producer's epilogue writes the fp16
    x_hi | x_lo pair the consumer's tensor-core operands need, so the consumer skips its in-place split (a tenth of a stage's
    shared-memory traffic: 29 -> 26 us for 64->64 at 120x160x5, 87 -> 77 us for 128->128).  Inputs may be in either form (their
    tag tells); layers whose output is read by something else than a planar layer pass ``out_split=False``.
    ``up2``: the result is returned nearest-neighbour x2 up-sampled (hybrid_depth_decoder.py:11-14), written that way by the
    layer's own epilogue instead of by a separate copy."""
    from . import ops
    pc = pcs[0]
    out_split = bool(out_split) and SPLIT_ACTIVATIONS and pc.out_chunks % 2 == 0
    if out4 is None:
        f = 2 if up2 else 1
        out4 = torch.empty(pc.out_chunks, x4.shape[1], f * x4.shape[2], f * x4.shape[3], 4, device=x4.device, dtype=torch.float32)
    ops.conv_planar(pc, x4, out4, res0=res4, dilation=dilation, in1=in1, taps=taps, post_scale=post_scale,
                    in_split=(_is_split(x4), _is_split(in1)), res_split=_is_split(res4), out_split=out_split, out_up2=up2)
    return _tag(out4, out_split)


def _sub2(x4):
    """Every other row and column of vol4 maps (what a stride-2 layer keeps of its stride-1 result); form-agnostic."""
    return _tag(x4[:, :, ::2, ::2, :].contiguous(), _is_split(x4))


def _as_vol4(t):
    """Feature map -> vol4: a 5-D tensor already is one (the tensor-core encoder returns its maps that way); an NCHW map is
    converted, reusing the copy its producer left on the tensor when there is one."""
    from . import ops
    if t.dim() == 5:
        return t
    cached = getattr(t, "_estd_vol4", None)
    if cached is not None and cached.device == t.device and cached.shape[0] * 4 == t.shape[1]:
        return cached
    return ops.nchw_to_vol4(t.contiguous())


def _as_nchw(t):
    from . import ops
    if t.dim() != 5:
        return t
    return ops.vol4_to_nchw(ops.from_split(t) if _is_split(t) else t)


def _channels(t):
    return t.shape[0] * 4 if t.dim() == 5 else t.shape[1]


def _conv_bn(cin, cout, k, stride, pad, dilation):
    """conv(bias=False)+BN pair; padding rule of networks/layers_op.py:10-14 (pad = dilation if dilation>1)."""
    return nn.Sequential(
        nn.Conv2d(cin, cout, k, stride=stride, padding=dilation if dilation > 1 else pad,
                  dilation=dilation, bias=False),
        nn.BatchNorm2d(cout))


class _ResidualUnit(nn.Module):
    """Two conv+BN with ReLU after the first only; NO ReLU after the add (psm_submodule.py:14-37, quirk Q12)."""

    def __init__(self, cin, cout, stride, pad, dilation, project):
        super().__init__()
        self.conv1 = nn.Sequential(_conv_bn(cin, cout, 3, stride, pad, dilation), nn.ReLU(inplace=True))
        self.conv2 = _conv_bn(cout, cout, 3, 1, pad, dilation)
        self.downsample = project

    def forward(self, x):
        y = self.conv2(self.conv1(x))
        return y + (x if self.downsample is None else self.downsample(x))


def _stage(cin, cout, n, stride, pad, dilation):
    project = None
    if stride != 1 or cin != cout:
        project = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), nn.BatchNorm2d(cout))
    units = [_ResidualUnit(cin, cout, stride, pad, dilation, project)]
    units += [_ResidualUnit(cout, cout, 1, pad, dilation, None) for _ in range(n - 1)]
    return nn.Sequential(*units)


class MatchingFeatureNet(nn.Module):
    """PSMNet-style 1/4-resolution, 32-channel matching features; output is a raw conv (no BN/ReLU).

    Reference: networks/psm_submodule.py:40-116.  Stage table (channels, units, stride, dilation):
    (32,3,1,1) (64,16,2,1) (128,3,1,1) (128,3,1,2); SPP pools 32/16/8/4; fuse 320->128->32.
    """

    def __init__(self):
        super().__init__()
        stem = []
        for cin, stride in ((3, 2), (32, 1), (32, 1)):
            stem += [_conv_bn(cin, 32, 3, stride, 1, 1), nn.ReLU(inplace=True)]
        self.firstconv = nn.Sequential(*stem)
        self.layer1 = _stage(32, 32, 3, 1, 1, 1)
        self.layer2 = _stage(32, 64, 16, 2, 1, 1)
        self.layer3 = _stage(64, 128, 3, 1, 1, 1)
        self.layer4 = _stage(128, 128, 3, 1, 1, 2)
        for idx, win in ((1, 32), (2, 16), (3, 8), (4, 4)):
            setattr(self, "branch%d" % idx, nn.Sequential(
                nn.AvgPool2d((win, win), stride=(win, win)), _conv_bn(128, 32, 1, 1, 0, 1), nn.ReLU(inplace=True)))
        self.lastconv = nn.Sequential(_conv_bn(320, 128, 3, 1, 1, 1), nn.ReLU(inplace=True),
                                      nn.Conv2d(128, 32, 1, bias=False))
        self.out_channels = [32]

    # ------------------------------------------------------------------ tensor-core path (3x3 convs of layers 2-4, fuse conv)
    @staticmethod
    def _bn_affine(bn):
        return _bn_affine(bn)

    def _packed(self, device):
        from . import packing
        probe = self.layer2[1].conv1[0][0].weight
        key = (str(device), probe.data_ptr(), probe._version, self.layer4[2].conv2[0].weight._version,
               self.lastconv[0][0].weight._version)
        if getattr(self, "_tc_key", None) != key:
            def pack(seq, act):          # seq = Sequential(conv, bn)
                s, b = self._bn_affine(seq[1])
                w = seq[0].weight.detach()
                return packing.pack_conv2d(w, s, b, act, device, cout_slice=64 if w.shape[0] >= 64 else 32)
            P = {}
            P["stem2"], P["stem4"] = pack(self.firstconv[2], "relu"), pack(self.firstconv[4], "relu")
            for i, blk in enumerate(self.layer1):
                P[("layer1", i, 1)], P[("layer1", i, 2)] = pack(blk.conv1[0], "relu"), pack(blk.conv2, "none")
            for name in ("layer2", "layer3", "layer4"):
                for i, blk in enumerate(getattr(self, name)):
                    P[(name, i, 1)] = pack(blk.conv1[0], "relu")        # layer2 block 0 is stride 2: run at stride 1, subsampled
                    P[(name, i, 2)] = pack(blk.conv2, "none")
                    if blk.downsample is not None:
                        P[(name, i, "down")] = pack(blk.downsample, "none")      # 1x1 (+BN) projection shortcut
            P["fuse"] = pack(self.lastconv[0], "relu")
            last = self.lastconv[2].weight.detach()                              # final 1x1, no BN / bias / activation
            P["last"] = packing.pack_conv2d(last, torch.ones(last.shape[0]), torch.zeros(last.shape[0]), "none", device, cout_slice=32)
            self._tc_packed, self._tc_key = P, key
        return self._tc_packed

    @staticmethod
    def _conv_tc(pcs, x4, out4, res4=None, dilation=1, taps=9, out_split=True):
        """Runs a packed planar layer (3x3, or 1x1 with taps=1) from vol4 maps into out4."""
        return _run3x3(pcs, x4, out4, res4=res4, dilation=dilation, taps=taps, out_split=out_split)

    def forward_tc(self, x):
        """Same arithmetic as ``forward`` with every stride-1 3x3 / 1x1 convolution (600 of the net's 679 GFLOP at 480x640) on
        the tcgen05 tensor cores (fp16 two-term split, fp32-class accuracy) with BN / ReLU / residual add fused into their
        epilogues; activations stay in vol4 between them, pre-split (vol4s) wherever the reader is another planar layer."""
        from . import ops
        P = self._packed(x.device)
        # 3->32 stride-2 stem conv + BN + ReLU: the library's direct kernel, NCHW images -> (pre-split) vol4 in one pass
        wf, bf = _folded.params(self.firstconv[0][0], self.firstconv[0][1])
        stem = _tag(ops.stem_conv(x.contiguous(), wf, bf, out_split=SPLIT_ACTIVATIONS), SPLIT_ACTIVATIONS)
        _, N, Hh, Wh, _ = stem.shape
        dev = x.device
        half = lambda: torch.empty(8, N, Hh, Wh, 4, device=dev, dtype=torch.float32)  # noqa: E731
        cur = self._conv_tc(P["stem4"], self._conv_tc(P["stem2"], stem, half()), half())
        tmp = half()
        for i in range(len(self.layer1)):
            self._conv_tc(P[("layer1", i, 1)], cur, tmp)
            cur = self._conv_tc(P[("layer1", i, 2)], tmp, half(), cur)
        H, W = (Hh + 1) // 2, (Wh + 1) // 2

        def vol(chunks):
            return torch.empty(chunks, N, H, W, 4, device=dev, dtype=torch.float32)

        # [raw 64 | skip 128 | branch4..1 32 each]: fp32 throughout (the SPP branches and the pooling read / write it too)
        cat = vol(80)
        # layer2 block 0: the stride-2 3x3 runs at stride 1 and keeps the even rows / columns; the stride-2 1x1 shortcut
        # runs on the subsampled input
        y = _sub2(self._conv_tc(P[("layer2", 0, 1)], cur, torch.empty(16, N, Hh, Wh, 4, device=dev)))
        shortcut = self._conv_tc(P[("layer2", 0, "down")], _sub2(cur), vol(16), taps=1)
        cur = self._conv_tc(P[("layer2", 0, 2)], y, vol(16), shortcut)
        tmp = vol(16)
        n2 = len(self.layer2)
        for i in range(1, n2):
            self._conv_tc(P[("layer2", i, 1)], cur, tmp)
            last = i == n2 - 1
            cur = self._conv_tc(P[("layer2", i, 2)], tmp, cat[0:16] if last else vol(16), cur, out_split=not last)
        raw = cur
        # layer3: block 0 changes the width (64 -> 128) and projects the shortcut with a 1x1 conv
        shortcut = self._conv_tc(P[("layer3", 0, "down")], raw, vol(32), taps=1)
        tmp = self._conv_tc(P[("layer3", 0, 1)], raw, vol(32))
        cur = self._conv_tc(P[("layer3", 0, 2)], tmp, vol(32), shortcut)
        stages = [("layer3", i, 1) for i in range(1, len(self.layer3))] + [("layer4", i, 2) for i in range(len(self.layer4))]
        for k, (name, i, dil) in enumerate(stages):
            self._conv_tc(P[(name, i, 1)], cur, tmp, dilation=dil)
            last = k == len(stages) - 1
            cur = self._conv_tc(P[(name, i, 2)], tmp, cat[16:48] if last else vol(32), cur, dilation=dil, out_split=not last)
        deep = ops.vol4_to_nchw(cur)                                # SPP pooling / 1x1 / bilinear resize: torch
        pooled = None
        for slot, idx in enumerate((4, 3, 2, 1)):
            br = getattr(self, "branch%d" % idx)
            # the pooling windows nest (4, 8, 16, 32): each level is the 2x2 average of the previous one
            pooled = F.avg_pool2d(deep, 4, 4) if pooled is None else F.avg_pool2d(pooled, 2, 2)
            # 1x1 conv + BN + ReLU on the pooled map, bilinear resize, concat: one small GEMM (folded weights) + one kernel that
            # adds the offset, applies the ReLU to the taps, interpolates and writes the branch's chunks of `cat`
            wf, bf = _folded.params(br[1][0], br[1][1])
            n_, c_, ph, pw = pooled.shape
            tf32 = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = False             # strict fp32, whatever the caller's global setting
            y = torch.matmul(wf.view(wf.shape[0], c_), pooled.reshape(n_, c_, ph * pw)).view(n_, wf.shape[0], ph, pw)
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ops.upsample_bilinear_vol4(y, cat[48 + 8 * slot:56 + 8 * slot], bias=bf, relu=True)
        fused = self._conv_tc(P["fuse"], cat, vol(32))
        return ops.vol4_to_nchw(self._conv_tc(P["last"], fused, vol(8), taps=1, out_split=False))

    def forward(self, x):
        if x.is_cuda and getattr(self, "tensor_cores", False):
            return self.forward_tc(x)
        x = self.layer1(self.firstconv(x))
        quarter = self.layer2(x)
        deep = self.layer4(self.layer3(quarter))
        size = deep.shape[-2:]
        # F.upsample(mode='bilinear') of the reference == interpolate(align_corners=False)
        pyramid = [F.interpolate(getattr(self, "branch%d" % i)(deep), size=size, mode="bilinear", align_corners=False)
                   for i in (4, 3, 2, 1)]
        return self.lastconv(torch.cat([quarter, deep] + pyramid, dim=1))


class ContextEncoder(nn.Module):
    """torchvision ResNet trunk returning the 5 feature maps at /2 ... /32 (resnet_encoder.py:17-51).

    The reference asks torchvision for ImageNet weights; there is no network here and eval always loads
    a checkpoint afterwards (eval_hybrid.py:328-333), so the trunk is created un-initialised
    (``weights=None``).  The unused ``fc`` layer is kept because it is part of the checkpoint.
    """

    def __init__(self, num_layers):
        super().__init__()
        ctor = {18: tvm.resnet18, 34: tvm.resnet34, 50: tvm.resnet50, 101: tvm.resnet101, 152: tvm.resnet152}
        if num_layers not in ctor:
            raise ValueError("{} is not a valid number of resnet layers".format(num_layers))
        self.num_ch_enc = np.array([64, 64, 128, 256, 512])
        if num_layers > 34:
            self.num_ch_enc[1:] *= 4
        self.encoder = ctor[num_layers](weights=None)
        self.tensor_cores = False       # set by the owning model (feature_precision="3xf16")
        self._tc_cache = {}

    def _packed_block(self, blk, device):
        """Layers of a torchvision Bottleneck (conv1 1x1, conv2 3x3, conv3 1x1, downsample) or BasicBlock (conv1 3x3, conv2 3x3,
        downsample) packed for the planar tensor-core kernel; cached per block.  -> (kind, p1, p2, p3 | None, pd | None)"""
        last = blk.conv3 if hasattr(blk, "conv3") else blk.conv2
        key = (str(device), blk.conv2.weight.data_ptr(), blk.conv1.weight._version, blk.conv2.weight._version,
               last.weight._version, blk.bn2.running_var._version)
        ent = self._tc_cache.get(id(blk))
        if ent is None or ent[0] != key:
            pd = None if blk.downsample is None else _pack3x3(blk.downsample[0], blk.downsample[1], "none", device, any_stride=True)
            if hasattr(blk, "conv3"):
                packed = ("bottleneck", _pack3x3(blk.conv1, blk.bn1, "relu", device), _pack3x3(blk.conv2, blk.bn2, "relu", device, any_stride=True),
                          _pack3x3(blk.conv3, blk.bn3, "add_relu", device), pd)
            else:
                packed = ("basic", _pack3x3(blk.conv1, blk.bn1, "relu", device, any_stride=True), _pack3x3(blk.conv2, blk.bn2, "add_relu", device),
                          None, pd)
            if any(q is None for q in packed[1:3]) or (packed[0] == "bottleneck" and packed[3] is None) or \
                    (blk.downsample is not None and pd is None):
                # every torchvision ResNet the reference can ask for (resnet_encoder.py:23-31) has channel counts that are
                # multiples of 16; anything else is not a silent cuDNN case but an error (feature_precision="fp32" is the
                # explicit way to run the 2-D nets on cuDNN)
                raise RuntimeError("estdepth_b200: ResNet block with a layer shape the planar tensor-core kernel does not take "
                                   "(channels must be multiples of 16, stride 1 or 2); construct the model with "
                                   "feature_precision='fp32' to run the 2-D networks on cuDNN")
            ent = (key, packed)
            self._tc_cache[id(blk)] = ent
        return ent[1]

    @staticmethod
    def _block(blk, x):
        """torchvision BasicBlock / Bottleneck in eval mode with folded BN and fused bias/residual/ReLU epilogues (cuDNN;
        the feature_precision="fp32" path)."""
        identity = x if blk.downsample is None else _folded(x, blk.downsample[0], blk.downsample[1])
        y = _folded(x, blk.conv1, blk.bn1, relu=True)
        if hasattr(blk, "conv3"):
            y = _folded(y, blk.conv2, blk.bn2, relu=True)
            return _folded(y, blk.conv3, blk.bn3, relu=True, residual=identity)
        return _folded(y, blk.conv2, blk.bn2, relu=True, residual=identity)

    @staticmethod
    def _plain_stem(e):
        """torchvision's stem as the direct kernels implement it: conv 7x7 / 2 / pad 3 without bias, max-pool 3 / 2 / pad 1."""
        c, p = e.conv1, e.maxpool
        pair = lambda v: tuple(v) if isinstance(v, (tuple, list)) else (v, v)  # noqa: E731
        return (tuple(c.weight.shape) == (64, 3, 7, 7) and c.bias is None and c.stride == (2, 2) and c.padding == (3, 3) and
                c.dilation == (1, 1) and c.groups == 1 and pair(p.kernel_size) == (3, 3) and pair(p.stride) == (2, 2) and
                pair(p.padding) == (1, 1) and pair(p.dilation) == (1, 1) and not p.ceil_mode)

    def _stage_tc(self, stage, x4):
        """One ResNet stage (Bottlenecks: ResNet-50/101/152, BasicBlocks: ResNet-18/34) on the planar tcgen05 kernel;
        activations stay in vol4 from the max-pool to the decoder (no NCHW copies between stages).  The stride-2 3x3 of a
        stage's first block runs at stride 1 and keeps the even rows / columns; its stride-2 1x1 shortcut runs on the
        subsampled input.  x4: vol4 [C/4, N, H, W, 4] in and out."""
        for blk in stage:
            kind, p1, p2, p3, pd = self._packed_block(blk, x4.device)
            strided = (blk.conv2 if kind == "bottleneck" else blk.conv1).stride != (1, 1)
            if blk.downsample is None:
                identity4 = x4
            else:
                xs = x4 if blk.downsample[0].stride == (1, 1) else _sub2(x4)
                identity4 = _run3x3(pd, xs, taps=1)
            if kind == "bottleneck":
                y4 = _run3x3(p2, _run3x3(p1, x4, taps=1))
                if strided:
                    y4 = _sub2(y4)
                x4 = _run3x3(p3, y4, res4=identity4, taps=1)
            else:
                y4 = _run3x3(p1, x4)
                if strided:
                    y4 = _sub2(y4)
                x4 = _run3x3(p2, y4, res4=identity4)
        return x4

    def forward(self, x):
        e = self.encoder
        if self.training:
            maps = [e.relu(e.bn1(e.conv1(x)))]
            maps.append(e.layer1(e.maxpool(maps[-1])))
            for stage in (e.layer2, e.layer3, e.layer4):
                maps.append(stage(maps[-1]))
            return maps
        if self.tensor_cores and x.is_cuda:
            # every map is returned in vol4 (5-D); the decoder's tensor-core path consumes them as they are (_as_vol4).
            # conv1 (7x7 stride 2) + BN + ReLU and the 3x3 stride-2 max-pool are the library's direct kernels: NCHW images ->
            # pre-split vol4, no cuDNN convolution and no layout pass left in the context branch
            from . import ops
            if self._plain_stem(e):
                wf, bf = _folded.params(e.conv1, e.bn1)
                cached = getattr(self, "_stem7_w", None)
                if cached is None or cached[0] is not wf:            # tap-major copy of the folded weight, made once
                    cached = self._stem7_w = (wf, wf.permute(1, 2, 3, 0).contiguous())
                stem = _tag(ops.stem7_conv(x.contiguous(), cached[1], bf, out_split=SPLIT_ACTIVATIONS), SPLIT_ACTIVATIONS)
                x4 = _tag(ops.maxpool3x3s2_vol4(stem, in_split=SPLIT_ACTIVATIONS, out_split=SPLIT_ACTIVATIONS), SPLIT_ACTIVATIONS)
                maps = [stem]
            else:
                maps = [_folded(x, e.conv1, e.bn1, relu=True)]
                x4 = ops.nchw_to_vol4(e.maxpool(maps[-1]).contiguous())
            for stage in (e.layer1, e.layer2, e.layer3, e.layer4):
                x4 = self._stage_tc(stage, x4)
                maps.append(x4)
            return maps
        maps = [_folded(x, e.conv1, e.bn1, relu=True)]
        x = e.maxpool(maps[-1])
        for stage in (e.layer1, e.layer2, e.layer3, e.layer4):
            for blk in stage:
                x = self._block(blk, x)
            maps.append(x)
        return maps


class _UpBlock(nn.Module):
    """3x3 conv + BN + ReLU (hybrid_depth_decoder.py:17-30); parameter names ``conv.0`` / ``conv.1``."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = _conv_bn(int(cin), int(cout), 3, 1, 1, 1)

    def forward(self, x):
        if self.training:
            return F.relu(self.conv(x), inplace=True)
        return _folded(x, self.conv[0], self.conv[1], relu=True)


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode="nearest")


class ContextDecoder2D(nn.Module):
    """The 2-D halves of the hybrid decoder (hybrid_depth_decoder.py:56-75, 163-184, 264-290).

    ``context(maps)``      -> ``semantic_vs`` [B*T, D, H/4, W/4]   (row a7)
    ``refine(...)``        -> depth at 1/2 (nearest x2) and full resolution        (row a13)
    Registered under ``CostRegNet`` by the owning module so the keys read ``CostRegNet.upconv_4_0...``.
    """

    def __init__(self, num_ch_enc, ndepths, depth_max):
        super().__init__()
        enc = [int(c) for c in num_ch_enc]
        dec = [16, 32, int(ndepths), 128, 256]
        self.depth_max = float(depth_max)
        self.upconv_4_0 = _UpBlock(enc[4], dec[4])
        self.upconv_4_1 = _UpBlock(dec[4] + enc[3], dec[4])
        self.upconv_3_0 = _UpBlock(dec[4], dec[3])
        self.upconv_3_1 = _UpBlock(dec[3] + enc[2], dec[3])
        self.upconv_2_0 = _UpBlock(dec[3], dec[2])
        self.upconv_2_1 = _UpBlock(dec[2] + enc[1], ndepths)
        self.upconv_1_0 = _UpBlock(dec[2] + ndepths, dec[1])
        self.upconv_1_1 = _UpBlock(dec[1] + enc[0], dec[1])
        self.dispconv_1 = nn.Conv2d(dec[1], 1, 3, 1, 1, 1, bias=True)
        self.upconv_0_0 = _UpBlock(dec[1], dec[0])
        self.upconv_0_1 = _UpBlock(dec[0], dec[0])
        self.dispconv_0 = nn.Conv2d(dec[0], 1, 3, 1, 1, 1, bias=True)

    tensor_cores = False                # set by the owning model (feature_precision="3xf16")

    def _packed(self, device):
        """The decoder's 3x3 conv+BN+ReLU layers packed for the planar tensor-core kernel (None entries: shapes it does not
        take, e.g. a depth-plane count that is not a multiple of 16)."""
        probe = self.upconv_4_0.conv[0].weight
        key = (str(device), probe.data_ptr(), probe._version, self.upconv_2_1.conv[0].weight._version,
               self.upconv_1_1.conv[1].running_var._version)
        if getattr(self, "_tc_key", None) != key:
            names = ("upconv_4_0", "upconv_4_1", "upconv_3_0", "upconv_3_1", "upconv_2_0", "upconv_2_1", "upconv_1_0", "upconv_1_1",
                     "upconv_0_0", "upconv_0_1")
            self._tc_packed = {n: _pack3x3(getattr(self, n).conv[0], getattr(self, n).conv[1], "relu", device) for n in names}
            for n in ("dispconv_1", "dispconv_0"):          # 3x3 conv + bias -> sigmoid, x depth_max in the epilogue
                self._tc_packed[n] = _pack3x3(getattr(self, n), None, "sigmoid", device)
            self._tc_key = key
        return self._tc_packed

    def _use_tc(self, x, names):
        if not (self.tensor_cores and x.is_cuda and not self.training):
            return None
        P = self._packed(x.device)
        return P if all(P[n] is not None for n in names) else None

    def _log_cudnn(self, what, x):
        """The one shape the planar kernel does not take is a depth-plane count that is not a multiple of 16 (the decoder's
        layers then have input channels that are not whole 16-channel k-steps).  That is an explicit decision, logged once
        per module: the stage runs on cuDNN in strict fp32 -- not a silent fallback."""
        if self.tensor_cores and x.is_cuda and not self.training and what not in self.__dict__.setdefault("_cudnn_logged", set()):
            self._cudnn_logged.add(what)
            logging.getLogger("estdepth_b200").warning(
                "%s: ndepths is not a multiple of 16 -> this stage runs on cuDNN (strict fp32) instead of the planar "
                "tcgen05 kernel; choose ndepths %% 16 == 0 for the tensor-core path", what)

    def context(self, maps):
        from . import ops
        P = self._use_tc(maps[4], ("upconv_4_0", "upconv_4_1", "upconv_3_0", "upconv_3_1", "upconv_2_0", "upconv_2_1"))
        if P is not None and all(_channels(m) % 16 == 0 for m in maps[1:]):
            # same layers, planar tcgen05 kernel: torch.cat becomes a second input segment, activations stay in vol4
            # `upsample` (nearest x2) of upconv_N_0's output is written by that layer's epilogue (up2): no separate copy
            x = _run3x3(P["upconv_4_0"], _as_vol4(maps[4]), up2=True)
            x = _run3x3(P["upconv_4_1"], x, in1=_as_vol4(maps[3]))
            x = _run3x3(P["upconv_3_0"], x, up2=True)
            x = _run3x3(P["upconv_3_1"], x, in1=_as_vol4(maps[2]))
            x = _run3x3(P["upconv_2_0"], x, up2=True)
            x = _run3x3(P["upconv_2_1"], x, in1=_as_vol4(maps[1]), out_split=False)    # read by vol4_to_nchw / refine
            out = ops.vol4_to_nchw(x)
            out._estd_vol4 = x                  # refine() takes the vol4 copy (saves a layout pass)
            return out
        self._log_cudnn("context decoder", maps[4])
        maps = [_as_nchw(m) for m in maps]
        x = self.upconv_4_0(maps[4])
        x = self.upconv_4_1(torch.cat([_up2(x), maps[3]], 1))
        x = self.upconv_3_0(x)
        x = self.upconv_3_1(torch.cat([_up2(x), maps[2]], 1))
        x = self.upconv_2_0(x)
        return self.upconv_2_1(torch.cat([_up2(x), maps[1]], 1))

    def refine(self, semantic_vs, fused_logits, skip_half):
        """fused_logits: [B*T, D, H/4, W/4] raw logits of stereo_head1 (ReLU applied here, :268)."""
        from . import ops
        P = self._use_tc(semantic_vs, ("upconv_1_0", "upconv_1_1", "upconv_0_0", "upconv_0_1", "dispconv_1", "dispconv_0"))
        if P is not None and semantic_vs.shape[1] % 16 == 0 and _channels(skip_half) % 16 == 0:
            # whole refinement on the planar tcgen05 kernel: cat -> second input segment, sigmoid * depth_max in the epilogue
            x = _run3x3(P["upconv_1_0"], _as_vol4(semantic_vs), in1=ops.nchw_to_vol4(F.relu(fused_logits)), up2=True)
            x = _run3x3(P["upconv_1_1"], x, in1=_as_vol4(skip_half))
            d1 = _run3x3(P["dispconv_1"], x, post_scale=self.depth_max, out_split=False)  # [1 chunk, N, H/2, W/2, 4]
            depth_half = _up2(d1[0, ..., 0].unsqueeze(1))
            x = _run3x3(P["upconv_0_1"], _run3x3(P["upconv_0_0"], x, up2=True))
            depth_full = _run3x3(P["dispconv_0"], x, post_scale=self.depth_max, out_split=False)[0, ..., 0].unsqueeze(1).contiguous()
            return depth_half, depth_full
        self._log_cudnn("refinement", semantic_vs)
        skip_half = _as_nchw(skip_half)
        x = self.upconv_1_0(torch.cat([semantic_vs, F.relu(fused_logits)], dim=1))
        x = self.upconv_1_1(torch.cat([_up2(x), skip_half], 1))
        depth_half = _up2(self.depth_max * torch.sigmoid(self.dispconv_1(x)))
        x = self.upconv_0_1(_up2(self.upconv_0_0(x)))
        depth_full = self.depth_max * torch.sigmoid(self.dispconv_0(x))
        return depth_half, depth_full
