"""``DepthNetHybrid`` -- the drop-in boundary.

Same constructor, ``forward`` signature, return values and ``state_dict`` keys as the reference's
``hybrid_models.model_hybrid.DepthNetHybrid`` (hybrid_models/model_hybrid.py:15-184) for inference
(``mode='val'``, the mode both eval drivers use: eval_hybrid.py:113-121, eval_hybrid_seq.py:180-183).

What runs where:
  * 2-D feature nets and the 2-D context decoder / refinement (``encoders.py``; the "kept" rows of SURVEY.md 8a): their
    stride-1 3x3 and 1x1 convolutions on the planar tcgen05 kernel of the same library (``feature_precision="3xf16"``), the
    stems / pools / resizes on PyTorch + cuDNN in strict fp32;
  * everything 3-D -- plane-sweep warp + folded pre0 (K1), every 3x3x3 convolution with its BN/activation/residual
    (K2), the EST warp+attention (K3), soft-argmin (K4), GroupNorm/GRU glue (K5), the camera algebra -- runs in the
    hand-written CUDA library through ``ops.py``.  The 3-D ``nn.Conv3d``/``BatchNorm3d``/``GroupNorm`` modules below
    are parameter containers only (they give the reference's state-dict names); their ``forward`` is never called.
    There is no PyTorch fallback for the 3-D path: without the CUDA library ``forward`` raises.

Protocol quirks reproduced on purpose: Q3 first window of a scene has no EST fusion,
Q4 the returned hidden-state pose is the LAST MEMORY pose once a memory exists, Q5 targets are fused in order and
later targets attend to already-fused values, Q6 mean-not-sum attention, Q7 rel_pose = P_j P_i^-1.
"""
import collections
import contextlib
import os

import torch
import torch.nn as nn

from . import ops, packing
from .encoders import ContextDecoder2D, ContextEncoder, MatchingFeatureNet


def _upload(t, dev):
    """Small host tensor (camera matrices, warp tables) -> device WITHOUT blocking the host.  A plain ``.to(dev)`` of pageable
    memory -- and a blocking copy of pinned memory -- first waits for everything already enqueued on the stream: with four such
    copies per step the host could never run ahead of the GPU, and every step started on an idle device."""
    if t.is_cuda:
        return (t if t.device == dev else t.to(dev)).contiguous()
    return t.contiguous().pin_memory().to(dev, non_blocking=True)


@contextlib.contextmanager
def _strict_fp32():
    """The cuDNN-side 2-D layers (stems, pools' neighbours, the fp32 feature path) and the small fp32 GEMMs run in STRICT
    fp32 whatever the process-wide flags say: PyTorch's default is ``cudnn.allow_tf32 = True`` and the reference's drivers
    never change it (eval_hybrid.py:13 only sets ``benchmark``), but single-pass TF32 in these layers breaks the 1e-3 depth
    gate.  The caller's flags are restored on exit."""
    cudnn_tf32, matmul_tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cudnn.allow_tf32 = cudnn_tf32
        torch.backends.cuda.matmul.allow_tf32 = matmul_tf32


def _conv_bn3(cin, cout, k):
    return nn.Sequential(nn.Conv3d(cin, cout, k, padding=(k - 1) // 2, bias=False), nn.BatchNorm3d(cout))


def _conv_bn_act3(cin, cout, act):
    return nn.Sequential(nn.Conv3d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm3d(cout), act)


class _ESTParams(nn.Module):
    """Parameter container with the names of transformer/epipolar_transformer.py:12-29."""

    def __init__(self, channels=16):
        super().__init__()
        self.gate_conv = nn.Conv3d(2 * channels, 2 * channels, 3, padding=1)
        self.reset_gate_norm = nn.GroupNorm(1, channels, 1e-5, True)
        self.update_gate_norm = nn.GroupNorm(1, channels, 1e-5, True)
        self.output_conv = nn.Conv3d(2 * channels, channels, 3, padding=1)
        self.output_norm = nn.GroupNorm(1, channels, 1e-5, True)


class HybridDecoder(ContextDecoder2D):
    """``CostRegNet``: 2-D context decoder (cuDNN) + parameter containers of the 3-D matching net."""

    def __init__(self, num_ch_enc, ndepths, depth_max, est):
        super().__init__(num_ch_enc, ndepths, depth_max)
        if est:
            self.epipolar_transformer = _ESTParams(16)
        relu = lambda: nn.ReLU(inplace=True)  # noqa: E731
        self.dres0 = nn.Sequential(_conv_bn_act3(32, 32, relu()), _conv_bn_act3(32, 32, relu()))
        self.dres1 = nn.Sequential(_conv_bn_act3(32, 32, relu()), _conv_bn_act3(32, 32, relu()))
        self.dres2 = nn.Sequential(_conv_bn_act3(33, 33, relu()))
        self.key_layer = nn.Sequential(_conv_bn_act3(33, 16, relu()))
        self.value_layer = nn.Sequential(_conv_bn_act3(33, 16, nn.Tanh()))
        self.stereo_head0 = nn.Sequential(_conv_bn_act3(16, 16, relu()), nn.Conv3d(16, 1, 1, bias=True))
        self.stereo_head1 = nn.Sequential(_conv_bn_act3(16, 16, relu()), nn.Conv3d(16, 1, 1, bias=True))


class _Workspace(object):
    """Scratch volumes reused across calls (never returned to the caller)."""

    def __init__(self, device, D, H, W, gn_rows):
        def vol(chunks):
            return torch.empty(chunks, D, H, W, 4, device=device, dtype=torch.float32)
        self.key = (str(device), D, H, W)
        self.x0, self.y, self.cost, self.a, self.b = vol(8), vol(8), vol(8), vol(8), vol(8)
        self.sem, self.z = vol(1), vol(10)          # z: 36 channels = 9 chunks; 10 when it is kept pre-split (vol4s)
        self.hid = vol(4)
        self.h, self.rh, self.o = vol(4), vol(4), vol(4)
        self.f = vol(8)
        self.part_f = torch.zeros(gn_rows, 2, 2, device=device, dtype=torch.float64)
        self.part_o = torch.zeros(gn_rows, 2, 2, device=device, dtype=torch.float64)
        self.stats_f = torch.empty(4, device=device, dtype=torch.float32)
        self.stats_o = torch.empty(4, device=device, dtype=torch.float32)
        self.homo = torch.empty(12, device=device, dtype=torch.float32)
        self.warp30 = torch.empty(ops.MAX_SOURCES, 30, device=device, dtype=torch.float32)


class DepthNetHybrid(nn.Module):
    def __init__(self, ndepths=64, depth_min=0.01, depth_max=10.0, resnet=50, IF_EST_transformer=True,
                 align_corners=False, fix_stale_pose=False, precision="3xf16r2d", feature_precision="3xf16", geometry="auto",
                 merged_pre2=True):
        """First five arguments: hybrid_models/model_hybrid.py:15-16.  Extra, keyword-only in practice:

        align_corners   grid_sample semantics of the warps: False = torch >= 1.3 (what the reference computes when run
                        today, and what the oracle pins); True = the torch 1.2 it was written for (quirk Q1).
        fix_stale_pose  opt-in fix of quirk Q4 (return the current target's pose with the hidden state).
        geometry        how the camera matrices of the two warps are derived.  "auto" (default): camera parameters given as
                        HOST tensors -> the reference's own fp32 torch.inverse / matmul sequence on the host (bit-identical
                        to the reference run on the CPU; the small tables are uploaded); given as CUDA tensors (what the
                        eval drivers pass) -> two launches of the library's fp64 geometry kernels per window
                        (estd_homography_table / estd_volume_warp_table), rounded once.  "torch": the reference's op
                        sequence on whatever device the parameters live on (~90 tiny launches per window on the GPU: the
                        reference's own GPU arithmetic).  "fp64": always the kernels.  Matrices from different LUs differ
                        in the last bit; a sampling coordinate within an ulp of the sampling range may then fall on the
                        other side of quirk Q10's cut (about one voxel in ten million) -- the reference shows the same
                        difference between its own CPU and GPU runs.
        precision       arithmetic of the 3-D convolutions: "3xf16" / "3xtf32" = error-compensated two-term splits on the
                        tcgen05 tensor cores (fp32-class accuracy; "3xf16" moves half the operand bytes and needs
                        |activation| <= 65504, which is checked), "3xf16r" = the 3xf16 arithmetic on the plane-ring
                        schedule, "3xf16r2" = the same on CTA pairs (cta_group::2), "3xf16r2d" = the CTA-pair ring with the
                        small products of the split in a second accumulator per slot (the accuracy of the exact kernel),
                        "fp32" = exact fp32 on the CUDA cores.
        """
        super().__init__()
        self.ndepths = int(ndepths)
        self.depth_min = depth_min
        self.depth_max = depth_max
        self.depth_interval = (depth_max - depth_min) / (ndepths - 1)
        # model_hybrid.py:32-33 (fp32 arange * interval + min, evaluated by torch on the CPU)
        self.depth_cands = torch.arange(0, ndepths, requires_grad=False).reshape(1, -1).to(
            torch.float32) * self.depth_interval + self.depth_min
        self.IF_EST_transformer = bool(IF_EST_transformer)
        self.align_corners = bool(align_corners)
        self.fix_stale_pose = bool(fix_stale_pose)
        if geometry not in ("auto", "torch", "fp64"):
            raise ValueError("geometry must be 'auto', 'torch' or 'fp64'")
        self.geometry = geometry
        # feature_precision: "3xf16" = the 3x3 convolutions of the matching-feature net on the tensor cores (fp32-class
        # accuracy), "fp32" = the whole 2-D net on cuDNN's strict-fp32 kernels
        if feature_precision not in ("3xf16", "fp32"):
            raise ValueError("feature_precision must be '3xf16' or 'fp32'")
        self.feature_precision = feature_precision
        if precision not in ops.PRECISION:
            raise ValueError("precision must be one of %s" % sorted(ops.PRECISION))
        self.precision = precision
        # merged_pre2: one pre2 convolution per target on the SUM of both sources' pre1 outputs (see _cost_volume)
        self.merged_pre2 = bool(merged_pre2)
        # overlap_context: context encoder / decoder on a second stream beside the matching-feature net (see prepare)
        self.overlap_context = int(os.environ.get("ESTD_OVERLAP_CONTEXT", "2"))    # 0 off, 1 join after the feature nets, 2 join at dres2
        self.split_activations = os.environ.get("ESTD_SPLIT_ACT", "1") != "0"      # see _ring
        self.fused_head = os.environ.get("ESTD_FUSED_HEAD", "1") != "0"
        self._ctx_done = None

        self.matchingFeature = MatchingFeatureNet()
        self.matchingFeature.tensor_cores = (feature_precision == "3xf16")
        self.semanticFeature = ContextEncoder(resnet)
        self.CostRegNet = HybridDecoder(self.semanticFeature.num_ch_enc, self.ndepths, self.depth_max,
                                        self.IF_EST_transformer)
        # the stride-1 3x3 convolutions of the context encoder's bottlenecks and of the 2-D decoder / refinement follow
        # the same switch as the matching-feature net
        self.semanticFeature.tensor_cores = self.CostRegNet.tensor_cores = (feature_precision == "3xf16")
        self.pre0 = _conv_bn3(64, 32, 1)
        self.pre1 = _conv_bn_act3(32, 32, nn.ReLU(inplace=True))
        self.pre2 = _conv_bn3(32, 32, 3)

        self._homo_table = None
        self._packed = None          # folded / packed 3-D parameters (rebuilt when the state dict changes)
        self._packed_key = None
        self._ws = None
        self._depth_dev = None
        self._feat_cache = collections.OrderedDict()     # (batch slot, frame id) -> matching features [32, H/4, W/4] (frame_ids=)
        self.feature_cache_size = 16

    # ------------------------------------------------------------------ parameter packing
    def _param_fingerprint(self, device):
        p = self.pre0[0].weight
        return (str(device), p.data_ptr(), p._version, self.pre1[0].weight._version,
                self.CostRegNet.dres0[0][0].weight._version)

    def load_state_dict(self, *args, **kwargs):
        self._packed = None
        self._feat_cache.clear()
        return super().load_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self._packed = None
        self._depth_dev = None
        self._ws = None
        if "_feat_cache" in self.__dict__:
            self._feat_cache.clear()
        return super()._apply(fn, *args, **kwargs)

    def repack(self):
        """Re-fold BN and re-pack the 3-D weights (call after mutating parameters in place)."""
        self._packed = None
        self._feat_cache.clear()

    def _matching_features(self, imgs, frame_ids=None):
        """psm_feature_extraction over every view (model_hybrid.py:126-131): imgs [B,V,3,Hi,Wi] (normalised) -> [B,V,32,Hi/4,Wi/4].

        ``frame_ids`` (optional): one hashable id per view -- a sequence of V ids, or B such sequences.
        Consecutive windows of a scene overlap (Joint: 2 of 5 frames, ``eval_hybrid.py:195-196``; ESTM: 2 of 3,
        ``eval_hybrid_seq.py:169-190``) and the reference recomputes the shared frames' features every window; with ids the
        features of a frame are computed once and kept for the next ``feature_cache_size`` frames.  The caller promises that an
        id names the same image; the cache is dropped whenever the parameters change."""
        B, V, _, Hi, Wi = imgs.shape
        if frame_ids is None:
            return self.matchingFeature(imgs.reshape(B * V, 3, Hi, Wi)).reshape(B, V, 32, Hi // 4, Wi // 4)
        ids = list(frame_ids)
        if B == 1 and len(ids) == V and not (V == 1 and isinstance(ids[0], (list, tuple))):
            ids = [ids]
        if len(ids) != B or any(len(row) != V for row in ids):
            raise ValueError("frame_ids must hold one id per view: %d x %d" % (B, V))
        cache = self._feat_cache
        keys = [[(b, ids[b][v], Hi, Wi, str(imgs.device)) for v in range(V)] for b in range(B)]
        pending = collections.OrderedDict()              # key -> (b, v) of the first view that shows this frame
        for b in range(B):
            for v in range(V):
                if keys[b][v] not in cache and keys[b][v] not in pending:
                    pending[keys[b][v]] = (b, v)
        if pending:
            new = self.matchingFeature(torch.stack([imgs[b, v] for b, v in pending.values()]))
            for i, key in enumerate(pending):
                cache[key] = new[i]
        rows = []
        for b in range(B):
            for v in range(V):
                cache.move_to_end(keys[b][v])
            rows.append(torch.stack([cache[keys[b][v]] for v in range(V)]))
        feats = torch.stack(rows)
        while len(cache) > max(self.feature_cache_size, B * V):
            cache.popitem(last=False)
        return feats

    def _layers(self, device):
        key = self._param_fingerprint(device)
        if self._packed is None or self._packed_key != key:
            sd = {k: v.detach() for k, v in self.state_dict().items()
                  if k.startswith(("pre", "CostRegNet.dres", "CostRegNet.key_layer", "CostRegNet.value_layer",
                                   "CostRegNet.stereo_head", "CostRegNet.epipolar_transformer"))}
            self._packed = packing.pack_layers(sd, device)
            self._packed_key = key
        return self._packed

    def _workspace(self, device, D, H, W, L):
        key = (str(device), D, H, W)
        if self._ws is None or self._ws.key != key:
            rows = max(ops.conv3d_num_ctas(L[n], D, H, W, precision=q) for n in ("gate", "output")
                       for q in ops.PRECISION) + 8
            self._ws = _Workspace(device, D, H, W, rows)
        return self._ws

    def _side_stream(self, device, name="geometry"):
        key = (str(device), name)
        streams = self.__dict__.setdefault("_side_streams", {})
        if key not in streams:
            streams[key] = torch.cuda.Stream(device=device)
        return streams[key]

    def _conv(self, pc, *args, **kwargs):
        return ops.conv3d(pc, *args, precision=self.precision, **kwargs)

    def _ring(self, L):
        """True when every 3-D layer runs on a plane-ring kernel: the conv-to-conv activations of the 3-D path are then kept
        PRE-SPLIT (vol4s: the producer's epilogue writes x_hi | x_lo, the consumer skips its in-place split) and the 1x1x1 logit
        heads are fused into the head convolutions' epilogues.  ESTD_SPLIT_ACT=0 / ESTD_FUSED_HEAD=0 turn the two off."""
        return self.precision in ops.RING_PRECISIONS and all(
            (v.precision is None or v.precision in ops.RING_PRECISIONS) and v.weight_ring is not None
            for v in L.values() if isinstance(v, ops.PackedConv))

    def _split3d(self, L):
        """Pre-split (vol4s) conv-to-conv activations on the 3-D path: a gain for the single-accumulator CTA-pair kernels (4 M
        tiles per CTA: 169 -> 164 us per 32->32 layer), a loss for the two-accumulator ones (2 M tiles: 173 -> 181 us,
        profiles/bench_split_r02.txt), which therefore keep fp32 tensors.  ESTD_SPLIT_ACT3D=0|1 forces it."""
        if not (self._ring(L) and self.split_activations):
            return False
        forced = os.environ.get("ESTD_SPLIT_ACT3D")
        return (forced != "0") if forced is not None else self.precision != "3xf16r2d"

    def _head(self, L, ws, which, volume, depth_values, logits_out, depth_out, prob_out):
        """stereo_head{0,1} (3x3x3 conv + BN + ReLU, then Conv3d(16, 1, 1)) + nearest x4 + depthlayer
        (hybrid_depth_decoder.py:104-112, 202-209, 259-260)."""
        name = "head%d" % which
        if self._ring(L) and self.fused_head:
            # the 16-channel hidden volume never reaches memory: the logit is a dot product in the convolution's epilogue
            self._conv(L[name], volume, None, head=(L[name + "_w"], L[name + "_b"], logits_out))
            ops.head_softargmin(depth_values, logits_in=logits_out, depth_out=depth_out, prob_out=prob_out, up=4)
        else:
            self._conv(L[name], volume, ws.hid)
            ops.head_softargmin(depth_values, hidden=ws.hid, head_w=L[name + "_w"], head_b=L[name + "_b"],
                                logits_out=logits_out, depth_out=depth_out, prob_out=prob_out, up=4)

    def scale_cam_intr(self, cam_intr, scale):
        out = cam_intr.clone()
        out[:, :2, :] *= scale
        return out

    # ------------------------------------------------------------------ the 3-D path for one batch element
    def _cost_volume(self, L, ws, ref_mix, src_mix, poses, K4, t, depth_values, out):
        """get_costvolume (model_hybrid.py:62-102) for target view t+1 with sources t and t+2:
        cost = sum_s [x0_s + pre2(pre1(x0_s))] / 2  (:94-97, :100).

        ``merged_pre2`` (default): pre2 = Conv3d + eval-mode BN is affine, so pre2(y_a) + pre2(y_b) = s * W (y_a + y_b) + 2 b
        (SURVEY.md Appendix A.1): the second pre1 adds y_a in its epilogue (after its ReLU) and ONE pre2 with the doubled
        offset runs per target -- 3 instead of 4 convolutions, same value up to fp32 summation order."""
        homo = [self._homo_table[2 * t + n] if self._homo_table is not None else None for n in (0, 1)]
        # conv-to-conv activations are kept pre-split (vol4s, see _ring): pre1's output and the cost volume.  x0 stays fp32 (a
        # K1 that writes it pre-split spills: measured 69 us against 41 us) and so do the tensors that are residuals.
        sp = self._split3d(L)
        s0 = (sp, False)
        if self.merged_pre2:
            assert out is not ws.x0 and out is not ws.a
            h = homo[0] if homo[0] is not None else ops.homography_setup(poses[t + 1], poses[t], K4, ws.homo)
            ops.warp_cost(ref_mix[t + 1], src_mix[t], h, depth_values, ws.x0, self.align_corners)
            self._conv(L["pre1"], ws.x0, ws.y)                                              # y is only ever a residual: fp32
            h = homo[1] if homo[1] is not None else ops.homography_setup(poses[t + 1], poses[t + 2], K4, ws.homo)
            ops.warp_cost(ref_mix[t + 1], src_mix[t + 2], h, depth_values, ws.b, self.align_corners)
            self._conv(L["pre1"], ws.b, ws.a, res0=ws.y, out_split=sp)                      # relu(bn(conv(x0_b))) + y_a
            self._conv(L["pre2_pair"], ws.a, out, res0=ws.x0, res1=ws.b, post_scale=0.5, in_split=s0, out_split=sp)
            return out
        for n, s in enumerate((t, t + 2)):
            h = homo[n] if homo[n] is not None else ops.homography_setup(poses[t + 1], poses[s], K4, ws.homo)
            ops.warp_cost(ref_mix[t + 1], src_mix[s], h, depth_values, ws.x0, self.align_corners)
            self._conv(L["pre1"], ws.x0, ws.y, out_split=sp)
            if n == 0:      # cost = x0 + pre2(pre1(x0))   (fp32: it is a residual of the second application)
                self._conv(L["pre2"], ws.y, ws.cost, res0=ws.x0, in_split=s0)
            else:           # cost = (cost + x0 + pre2(pre1(x0))) / 2
                self._conv(L["pre2"], ws.y, out, res0=ws.x0, res1=ws.cost, post_scale=0.5, in_split=s0, out_split=sp)
        return out

    def _matching(self, L, ws, cost, semantic_vs_t, depth_values, logits_out, depth_out, prob_out):
        """dres0..2, value/key heads, stereo_head0 + soft-argmin (hybrid_depth_decoder.py:187-209)."""
        dev = cost.device
        _, D, H, W, _ = cost.shape
        sp = self._split3d(L)                                   # the cost volume arrives pre-split then (see _cost_volume)
        s0 = (sp, False)
        self._conv(L["dres0.0"], cost, ws.a, in_split=s0, out_split=sp)
        self._conv(L["dres0.1"], ws.a, ws.b, in_split=s0, out_split=sp)
        self._conv(L["dres1.0"], ws.b, ws.a, in_split=s0, out_split=sp)
        self._conv(L["dres1.1"], ws.a, ws.b, in_split=s0, out_split=sp)
        if self._ctx_done is not None:
            torch.cuda.current_stream(dev).wait_event(self._ctx_done)
            self._ctx_done = None
        ops.scalar_to_vol4(semantic_vs_t, ws.sem)
        z = ws.z if sp else ws.z[:9]
        self._conv(L["dres2"], ws.b, z, in1=ws.sem, in_split=s0, out_split=sp)             # the context chunk stays fp32
        value = torch.empty(4, D, H, W, 4, device=dev, dtype=torch.float32)
        key = torch.empty(4, D, H, W, 4, device=dev, dtype=torch.float32)
        self._conv(L["value_key"], z, value, out1=key, in_split=s0)
        self._head(L, ws, 0, value, depth_values, logits_out, depth_out, prob_out)
        return value, key

    def _fuse(self, L, ws, key_i, value_i, src_keys, src_values, pose_i, src_poses, K4, depth_values, warp30=None):
        """EpipolarTransformer.forward (transformer/epipolar_transformer.py:56-83) for one target."""
        _, D, H, W, _ = value_i.shape
        n = len(src_keys)
        if warp30 is None:
            warp30 = ws.warp30
            for k in range(n):
                ops.volume_warp_setup(pose_i, src_poses[k], K4, warp30[k])
        ops.est_attend(key_i, src_keys, src_values, warp30, depth_values, self.depth_min, self.depth_interval,
                       out=ws.h, align_corners=self.align_corners)
        count = 16.0 * D * H * W
        rows_f = ops.conv3d_num_ctas(L["gate"], D, H, W, precision=self.precision)
        self._conv(L["gate"], value_i, ws.f, in1=ws.h, gn_partials=ws.part_f)
        ops.gn_finalize(ws.part_f[:rows_f], 2, count, out=ws.stats_f)
        ops.gru_reset(ws.f, ws.h, ws.stats_f, L["gn_r_w"], L["gn_r_b"], out=ws.rh)
        rows_o = ops.conv3d_num_ctas(L["output"], D, H, W, precision=self.precision)
        self._conv(L["output"], value_i, ws.o, in1=ws.rh, gn_partials=ws.part_o)
        ops.gn_finalize(ws.part_o[:rows_o], 1, count, out=ws.stats_o)
        fused = torch.empty_like(value_i)
        ops.gru_blend(ws.f, ws.h, ws.o, ws.stats_f, ws.stats_o, L["gn_u_w"], L["gn_u_b"], L["gn_o_w"], L["gn_o_b"],
                      out=fused)
        return fused

    @staticmethod
    def _state_to_vol4(t, b):
        """Hidden-state tensor [B,16,D,H,W] handed back by a driver -> this batch element's vol4."""
        cached = getattr(t, "_estd_vol4", None)
        if cached is not None and cached[b] is not None and cached[b].device == t.device:
            return cached[b]
        vol = ops.ncdhw_to_vol4(t[b].detach().to(torch.float32).contiguous())
        # remembered on the tensor object: a state that arrived as plain NCDHW data (cloned by a driver, received from another
        # rank) is converted once, not once per window it stays in the memory.  The caller must not modify a state in place.
        if cached is None:
            cached = [None] * t.shape[0]
            try:
                t._estd_vol4 = cached
            except AttributeError:
                return vol
        cached[b] = vol
        return vol

    # ------------------------------------------------------------------ forward
    def forward(self, imgs, cam_poses, cam_intr, sample=None, pre_costs=None, pre_cam_poses=None, mode='train', frame_ids=None):
        """imgs [B,V,3,H,W] (0..255), cam_poses [B,V,4,4] cam->world, cam_intr [B,3,3]; V-2 target views.

        mode='val' -> (outputs, {"keys": [k], "values": [v]}, [pose]) exactly as the reference
        (hybrid_models/model_hybrid.py:183-184).  'train' / 'test' (loss / metric heads) are outside the inference
        hot path and raise.  ``frame_ids`` (extension, optional): see ``_matching_features``.
        """
        if mode != 'val':
            raise NotImplementedError("estdepth_b200 implements the inference path (mode='val'); got mode=%r" % (mode,))
        if not imgs.is_cuda:
            raise RuntimeError("estdepth_b200.DepthNetHybrid runs on CUDA only (no CPU fallback); imgs is on %s" % imgs.device)
        return self._forward_val(imgs, cam_poses, cam_intr, pre_costs, pre_cam_poses, frame_ids)

    def _forward_val(self, imgs, cam_poses, cam_intr, pre_costs, pre_cam_poses, frame_ids=None):
        memory_poses = pre_cam_poses if (self.IF_EST_transformer and pre_costs is not None) else None
        return self.fuse(self.prepare(imgs, cam_poses, cam_intr, memory_poses=memory_poses, frame_ids=frame_ids),
                         pre_costs, pre_cam_poses)

    def _poll_status(self, device):
        """fp16 range flag of EARLIER launches, examined without draining the GPU (ops.check_status_async): called at the start
        of ``prepare`` and of ``fuse`` so that callers of the split API (the clip pipeline) are covered too.  The check is
        late by design -- a violation in the last call of a run is only caught by ``check()``."""
        if self.precision in ("3xf16",) + ops.RING_PRECISIONS or self.feature_precision == "3xf16":
            ops.check_status_async(device)

    def check(self, device=None):
        """Blocking end-of-sequence check: raises if any fp16-split convolution launched so far saw an activation beyond
        the fp16 range (its outputs, and every hidden state derived from them, are invalid).  One 4-byte read that
        synchronises the device -- call it after the last window of a sequence, before trusting / saving its maps."""
        if device is None:
            device = next(self.parameters()).device
        ops.check_status(device)

    def prepare(self, imgs, cam_poses, cam_intr, memory_poses=None, frame_ids=None):
        """Everything that does not depend on the hidden state (about 89 % of the FLOPs of a step):
        2-D feeders, cost volumes, matching net, key/value volumes, initial depth.  ``forward`` is
        ``fuse(prepare(...), pre_costs, pre_cam_poses)``; the split lets a rank of the ESTM clip pipeline
        (``sharding.py``) run ahead while it waits for its predecessor's memory."""
        self._poll_status(imgs.device)
        with torch.no_grad(), _strict_fp32():
            return self._prepare(imgs, cam_poses, cam_intr, memory_poses, frame_ids)

    def fuse(self, prep, pre_costs=None, pre_cam_poses=None):
        """EST fusion against the memory (or the no-EST path, quirk Q3), stereo_head1 + soft-argmin, 2-D refinement,
        outputs and the hidden state to hand to the next call (hybrid_depth_decoder.py:211-292 / :373-417)."""
        self._poll_status(prep["dev"])
        with torch.no_grad(), _strict_fp32():
            return self._fuse_tail(prep, pre_costs, pre_cam_poses)

    def _prepare(self, imgs, cam_poses, cam_intr, memory_poses=None, frame_ids=None):
        dev = imgs.device
        imgs = 2 * (imgs / 255.) - 1.
        B, V, _, Hi, Wi = imgs.shape
        H, W = Hi // 4, Wi // 4
        assert V > 2  # the views_num should be larger than 2 (model_hybrid.py:123)
        T = V - 2
        D = self.ndepths
        L = self._layers(dev)
        ws = self._workspace(dev, D, H, W, L)
        if self._depth_dev is None or self._depth_dev.device != dev:
            self._depth_dev = self.depth_cands.reshape(-1).to(dev).contiguous()
        depth_values = self._depth_dev

        # Camera parameters may live on the host: the 4x4 / 3x3 algebra is then done with the reference's torch ops ON THE HOST
        # -- bit-identical to the reference run on the CPU (how the parity fixtures were made) -- and only the small tables are
        # uploaded.  With CUDA parameters (what the eval drivers pass after tocuda) the library's fp64 kernels derive every
        # pair's matrices in one launch per warp kind ("auto"), or the reference's own op sequence runs on the GPU ("torch").
        host_geometry = self.geometry in ("auto", "torch") and not cam_poses.is_cuda and not cam_intr.is_cuda
        kernel_geometry = self.geometry == "fp64" or (self.geometry == "auto" and not host_geometry)
        K4_src = self.scale_cam_intr(cam_intr.to(torch.float32), 0.25).contiguous()
        poses_src = cam_poses.to(torch.float32).contiguous()
        K4, poses = _upload(K4_src, dev), _upload(poses_src, dev)
        inputs_ready = None
        if self.geometry == "torch" and not host_geometry:
            inputs_ready = torch.cuda.Event()
            inputs_ready.record(torch.cuda.current_stream(dev))

        # ---- 2-D feeders (cuDNN) ----
        t_prof = ops._bracket_begin("torch_2d_feeders")
        if self.overlap_context:
            # the context branch (ResNet + 2-D decoder on the T target frames) does not meet the matching branch before dres2:
            # it runs on its own stream beside the matching-feature net.  Most of its layers work at 1/8 .. 1/32 resolution
            # and launch fewer CTAs than there are SMs, so the two branches fill each other's idle SMs and launch gaps.
            main = torch.cuda.current_stream(dev)
            ctx = self._side_stream(dev, "context")
            imgs_ready = torch.cuda.Event()
            imgs_ready.record(main)
            ctx.wait_event(imgs_ready)
            with torch.cuda.stream(ctx):
                maps = self.semanticFeature(imgs[:, 1:1 + T].reshape(B * T, 3, Hi, Wi))
                semantic_vs = self.CostRegNet.context(maps).contiguous()             # [B*T, D, H, W]
                ctx_done = torch.cuda.Event()
                ctx_done.record(ctx)
            feats = self._matching_features(imgs, frame_ids)
            for t_ in (semantic_vs, maps[0], getattr(semantic_vs, "_estd_vol4", None)):
                if t_ is not None:
                    t_.record_stream(main)       # allocated on the context stream, consumed (and released) on the main one
            if self.overlap_context >= 2:
                self._ctx_done = ctx_done        # joined where the context map is first needed (dres2 of the first target)
            else:
                main.wait_event(ctx_done)
        else:
            feats = self._matching_features(imgs, frame_ids)
            maps = self.semanticFeature(imgs[:, 1:1 + T].reshape(B * T, 3, Hi, Wi))
            semantic_vs = self.CostRegNet.context(maps).contiguous()                 # [B*T, D, H, W]
        ops._bracket_end("torch_2d_feeders", t_prof)
        # camera algebra of both warps with the reference's own torch ops (~90 tiny launches): issued AFTER the feeders so that
        # the GPU is already busy while the host spends its millisecond on them, and on a side stream so that they run beside
        # the feeders instead of between them and the first warp
        homo_tables, warp_tables = None, None
        pairs = [(t + 1, s) for t in range(T) for s in (t, t + 2)]
        if host_geometry:
            homo_tables = [_upload(ops.homography_table_torch(poses_src[b], K4_src[b], pairs), dev) for b in range(B)]
            if memory_poses is not None and all(not p.is_cuda for p in memory_poses):
                warp_tables = [[_upload(tab, dev) for tab in ops.volume_warp_tables_torch(
                    [poses_src[b, t + 1] for t in range(T)] + [p[b].to(torch.float32).contiguous() for p in memory_poses],
                    T, K4_src[b])] for b in range(B)]
        elif kernel_geometry:
            homo_tables = [ops.homography_table(poses[b], K4[b], pairs) for b in range(B)]
            if memory_poses is not None:
                warp_tables = [ops.volume_warp_tables(
                    [poses[b, t + 1] for t in range(T)] + [_upload(p[b].to(torch.float32), dev) for p in memory_poses],
                    T, K4[b]) for b in range(B)]
        else:
            side = self._side_stream(dev)
            side.wait_event(inputs_ready)
            with torch.cuda.stream(side):
                homo_tables = [ops.homography_table_torch(poses[b], K4[b], pairs) for b in range(B)]
                if memory_poses is not None:
                    # the EST warps' matrices too, when the memory poses are already known (forward(); a clip-pipeline rank
                    # that prepares ahead of its predecessor's state derives them in fuse())
                    warp_tables = [ops.volume_warp_tables_torch(
                        [poses[b, t + 1] for t in range(T)] + [_upload(p[b].to(torch.float32), dev) for p in memory_poses],
                        T, K4[b]) for b in range(B)]
                homo_ready = torch.cuda.Event()
                homo_ready.record(side)
            torch.cuda.current_stream(dev).wait_event(homo_ready)

        init_logits = torch.empty(B * T, D, H, W, device=dev, dtype=torch.float32)
        depth3 = torch.empty(B, T, 1, Hi, Wi, device=dev, dtype=torch.float32)
        init_prob = torch.empty(B, T, 1, Hi, Wi, device=dev, dtype=torch.float32)
        keys, values = [], []
        for b in range(B):
            self._homo_table = None if homo_tables is None else homo_tables[b]
            mix = ops.premix_batch(feats[b], L["pre0_both"], L["pre0_both_bias"])          # [V, 16, H, W, 4], one launch
            ref_mix = [mix[v, :8] for v in range(V)]
            src_mix = [mix[v, 8:] for v in range(V)]
            kb, vb = [], []
            for t in range(T):
                self._cost_volume(L, ws, ref_mix, src_mix, poses[b], K4[b], t, depth_values, ws.cost)
                value, key = self._matching(L, ws, ws.cost, semantic_vs[b * T + t], depth_values,
                                            init_logits[b * T + t], depth3[b, t, 0], init_prob[b, t, 0])
                vb.append(value)
                kb.append(key)
            keys.append(kb)
            values.append(vb)
        return dict(B=B, V=V, T=T, D=D, H=H, W=W, Hi=Hi, Wi=Wi, dev=dev, keys=keys, values=values, poses=poses,
                    cam_poses=cam_poses, K4=K4, semantic_vs=semantic_vs, skip_half=maps[0], depth3=depth3,
                    init_prob=init_prob, depth_values=depth_values, warp_tables=warp_tables,
                    warp_tables_memory=None if (memory_poses is None or warp_tables is None) else len(memory_poses),
                    host_geometry=host_geometry, kernel_geometry=kernel_geometry, poses_host=poses_src if host_geometry else None,
                    K4_host=K4_src if host_geometry else None)

    def _fuse_tail(self, prep, pre_costs=None, pre_cam_poses=None):
        B, T, D, H, W, Hi, Wi, dev = (prep[k] for k in ("B", "T", "D", "H", "W", "Hi", "Wi", "dev"))
        L = self._layers(dev)
        ws = self._workspace(dev, D, H, W, L)
        depth_values, poses, K4, cam_poses = prep["depth_values"], prep["poses"], prep["K4"], prep["cam_poses"]
        fused_logits = torch.empty(B * T, D, H, W, device=dev, dtype=torch.float32)
        depth2 = torch.empty(B, T, 1, Hi, Wi, device=dev, dtype=torch.float32)
        fused_prob = torch.empty(B, T, 1, Hi, Wi, device=dev, dtype=torch.float32)
        use_est = self.IF_EST_transformer and pre_costs is not None              # quirk Q3 (hybrid_depth_decoder.py:423)
        pre_num = len(pre_cam_poses) if use_est else 0
        state_key = torch.empty(B, 16, D, H, W, device=dev, dtype=torch.float32)
        state_value = torch.empty(B, 16, D, H, W, device=dev, dtype=torch.float32)
        state_key._estd_vol4, state_value._estd_vol4 = [None] * B, [None] * B

        for b in range(B):
            values, keys = list(prep["values"][b]), list(prep["keys"][b])
            all_poses = [poses[b, t + 1] for t in range(T)]
            if use_est:
                # memory volumes are appended after the current ones (hybrid_depth_decoder.py:220-224)
                all_poses += [_upload(p[b].to(torch.float32), dev) for p in pre_cam_poses]
                values += [self._state_to_vol4(v, b) for v in pre_costs["values"]]
                keys += [self._state_to_vol4(k, b) for k in pre_costs["keys"]]
            tables = None
            if use_est:
                if prep.get("warp_tables") is not None and prep.get("warp_tables_memory") == pre_num:
                    tables = prep["warp_tables"][b]                  # derived during prepare()
                elif prep.get("host_geometry") and all(not p.is_cuda for p in pre_cam_poses):
                    host_poses = [prep["poses_host"][b, t + 1] for t in range(T)] + [p[b].to(torch.float32).contiguous() for p in pre_cam_poses]
                    tables = [_upload(tab, dev) for tab in ops.volume_warp_tables_torch(host_poses, T, prep["K4_host"][b])]
                elif prep.get("kernel_geometry"):
                    tables = ops.volume_warp_tables(all_poses, T, K4[b])
                else:
                    tables = ops.volume_warp_tables_torch(all_poses, T, K4[b])
            for i in range(T):
                if use_est:
                    others = [j for j in range(T + pre_num) if j != i]
                    fused = self._fuse(L, ws, keys[i], values[i], [keys[j] for j in others], [values[j] for j in others],
                                       all_poses[i], [all_poses[j] for j in others], K4[b], depth_values,
                                       warp30=None if tables is None else tables[i])
                    values[i] = fused                                             # quirk Q5 (:253)
                self._head(L, ws, 1, values[i], depth_values, fused_logits[b * T + i], depth2[b, i, 0], fused_prob[b, i, 0])
            ops.vol4_to_ncdhw(keys[T - 1], state_key[b])
            ops.vol4_to_ncdhw(values[T - 1], state_value[b])
            state_key._estd_vol4[b], state_value._estd_vol4[b] = keys[T - 1], values[T - 1]
        last_pose_src = (T + pre_num - 1) if (use_est and not self.fix_stale_pose) else (T - 1)

        # ---- 2-D refinement (cuDNN) ----
        t_prof = ops._bracket_begin("torch_2d_refine")
        depth_half, depth_full = self.CostRegNet.refine(prep["semantic_vs"], fused_logits, prep["skip_half"])
        ops._bracket_end("torch_2d_refine", t_prof)
        depth_half = depth_half.reshape(B, T, 1, Hi, Wi)
        depth_full = depth_full.reshape(B, T, 1, Hi, Wi)

        outputs = {}
        for t in range(T):
            outputs[("depth", t, 3)] = prep["depth3"][:, t]
            outputs[("init_prob", t)] = prep["init_prob"][:, t]
        for t in range(T):
            outputs[("depth", t, 2)] = depth2[:, t]
            outputs[("fused_prob", t)] = fused_prob[:, t]
        for t in range(T):
            outputs[("depth", t, 1)] = depth_half[:, t]
        for t in range(T):
            outputs[("depth", t, 0)] = depth_full[:, t]

        # hidden state: last target's key and (fused) value; pose = last element of the (extended) pose list -- Q4
        if last_pose_src >= T:
            state_pose = pre_cam_poses[last_pose_src - T]
        else:
            state_pose = cam_poses[:, last_pose_src + 1, :, :]
        return outputs, {"keys": [state_key], "values": [state_value]}, [state_pose]
