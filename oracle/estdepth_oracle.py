"""CPU oracle for the ESTDepth inference hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional (state-dict in, tensors out) fp32 restatement of what the reference computes on the path
``DepthNetHybrid.forward(..., mode='val')``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the product
package ``estdepth_b200`` never does (it fails loudly when its CUDA library is missing).

Parity pinning: the reference ships no tests, fixtures or golden vectors, and its
arithmetic lives in a third-party dependency that is not under /root/reference: PyTorch ATen
(pinned ``pytorch=1.2.0`` in the reference's environment.yml:70; 2.11.0 is what is installed and what
executes here).  The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the
build container by ``oracle/make_golden.py`` (reference imported from /root/reference through the shims
in ``oracle/ref_loader.py``) and committed under ``tests/golden/``; ``tests/test_oracle_golden.py``
re-checks them on every run and ``tests/test_oracle_vs_reference.py`` re-runs the live comparison
whenever /root/reference is present.

Each function cites the reference lines it follows.  Two samplers are provided for the geometric
warps: ``"aten"`` calls ``F.grid_sample`` (the primitive the reference itself calls) and
``"explicit"`` is a closed-form gather restatement of ATen's ``GridSampler`` (unnormalise with
align_corners=False, zeros padding) used to cross-check the CUDA kernels' tap arithmetic.

Everything is B = 1 per call, like the reference.
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


# ----------------------------------------------------------------------------------------------
# small building blocks
# ----------------------------------------------------------------------------------------------

def _bn(x, sd, p):
    """eval-mode BatchNorm{2,3}d (networks/layers_op.py:10-39 wrap every conv in one)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, BN_EPS)


def _cb2(x, sd, p, stride=1, pad=1, dil=1):
    """conv2d(bias=False)+BN, parameters at ``p.0`` / ``p.1``; padding rule of layers_op.py:12."""
    w = sd[p + ".0.weight"]
    return _bn(F.conv2d(x, w, None, stride, dil if dil > 1 else pad, dil), sd, p + ".1")


def _cb3(x, sd, p, act=None):
    """conv3d(bias=False)+BN(+act), parameters at ``p.0`` / ``p.1`` (layers_op.py:16-39)."""
    w = sd[p + ".0.weight"]
    y = _bn(F.conv3d(x, w, None, 1, (w.shape[-1] - 1) // 2), sd, p + ".1")
    if act == "relu":
        y = F.relu(y)
    elif act == "tanh":
        y = torch.tanh(y)
    return y


# ----------------------------------------------------------------------------------------------
# 2-D feeders (kept on cuDNN in the product; restated here so the oracle is self-contained)
# ----------------------------------------------------------------------------------------------

def psm_features(sd, x, prefix="matchingFeature"):
    """networks/psm_submodule.py:93-116 (forward) with blocks of :14-37 and stage table :51-54."""
    p = prefix
    for i, stride in ((0, 2), (2, 1), (4, 1)):
        x = F.relu(_cb2(x, sd, "%s.firstconv.%d" % (p, i), stride, 1, 1))

    def stage(x, name, n, stride, dil):
        for b in range(n):
            q = "%s.%s.%d" % (p, name, b)
            y = F.relu(_cb2(x, sd, q + ".conv1.0", stride if b == 0 else 1, 1, dil))
            y = _cb2(y, sd, q + ".conv2", 1, 1, dil)
            if (q + ".downsample.0.weight") in sd:
                x = _bn(F.conv2d(x, sd[q + ".downsample.0.weight"], None, stride), sd, q + ".downsample.1")
            x = y + x                                   # no ReLU after the add (:35)
        return x

    x = stage(x, "layer1", 3, 1, 1)
    raw = stage(x, "layer2", 16, 2, 1)
    skip = stage(stage(raw, "layer3", 3, 1, 1), "layer4", 3, 1, 2)
    size = skip.shape[-2:]
    br = {}
    for idx, win in ((1, 32), (2, 16), (3, 8), (4, 4)):
        y = F.avg_pool2d(skip, (win, win), (win, win))
        y = F.relu(_cb2(y, sd, "%s.branch%d.1" % (p, idx), 1, 0, 1))
        br[idx] = F.interpolate(y, size=size, mode="bilinear", align_corners=False)
    x = torch.cat((raw, skip, br[4], br[3], br[2], br[1]), 1)
    x = F.relu(_cb2(x, sd, p + ".lastconv.0", 1, 1, 1))
    return F.conv2d(x, sd[p + ".lastconv.2.weight"])


def resnet_maps(sd, x, num_layers, prefix="semanticFeature.encoder"):
    """hybrid_models/resnet_encoder.py:40-51 over a torchvision ResNet-18/34/50 trunk."""
    p = prefix

    def bn(x, q):
        return _bn(x, sd, q)

    x = F.relu(bn(F.conv2d(x, sd[p + ".conv1.weight"], None, 2, 3), p + ".bn1"))
    maps = [x]
    x = F.max_pool2d(x, 3, 2, 1)
    blocks = {18: (2, 2, 2, 2), 34: (3, 4, 6, 3), 50: (3, 4, 6, 3)}[num_layers]
    bottleneck = num_layers >= 50
    for li, n in enumerate(blocks, start=1):
        for b in range(n):
            q = "%s.layer%d.%d" % (p, li, b)
            stride = 2 if (li > 1 and b == 0) else 1
            idt = x
            if bottleneck:          # torchvision Bottleneck (v1.5: stride on the 3x3)
                y = F.relu(bn(F.conv2d(x, sd[q + ".conv1.weight"]), q + ".bn1"))
                y = F.relu(bn(F.conv2d(y, sd[q + ".conv2.weight"], None, stride, 1), q + ".bn2"))
                y = bn(F.conv2d(y, sd[q + ".conv3.weight"]), q + ".bn3")
            else:
                y = F.relu(bn(F.conv2d(x, sd[q + ".conv1.weight"], None, stride, 1), q + ".bn1"))
                y = bn(F.conv2d(y, sd[q + ".conv2.weight"], None, 1, 1), q + ".bn2")
            if (q + ".downsample.0.weight") in sd:
                idt = bn(F.conv2d(x, sd[q + ".downsample.0.weight"], None, stride), q + ".downsample.1")
            x = F.relu(y + idt)
        maps.append(x)
    return maps


def _up2(x):
    return F.interpolate(x, scale_factor=2, mode="nearest")


def _upblock(x, sd, name):
    return F.relu(_cb2(x, sd, "CostRegNet.%s.conv" % name))


def context_decoder(sd, maps):
    """hybrid_depth_decoder.py:163-184 -> semantic_vs [T, D, H/4, W/4]."""
    x = _upblock(maps[4], sd, "upconv_4_0")
    x = _upblock(torch.cat([_up2(x), maps[3]], 1), sd, "upconv_4_1")
    x = _upblock(x, sd, "upconv_3_0")
    x = _upblock(torch.cat([_up2(x), maps[2]], 1), sd, "upconv_3_1")
    x = _upblock(x, sd, "upconv_2_0")
    return _upblock(torch.cat([_up2(x), maps[1]], 1), sd, "upconv_2_1")


def refine_2d(sd, semantic_vs, fused_logits, skip_half, depth_max):
    """hybrid_depth_decoder.py:264-290 -> (depth s=1 [T,1,H,W], depth s=0 [T,1,H,W])."""
    x = _upblock(torch.cat([semantic_vs, F.relu(fused_logits)], 1), sd, "upconv_1_0")
    x = _upblock(torch.cat([_up2(x), skip_half], 1), sd, "upconv_1_1")
    d1 = F.conv2d(x, sd["CostRegNet.dispconv_1.weight"], sd["CostRegNet.dispconv_1.bias"], 1, 1)
    d1 = F.interpolate(depth_max * torch.sigmoid(d1), scale_factor=2)
    x = _upblock(_up2(_upblock(x, sd, "upconv_0_0")), sd, "upconv_0_1")
    d0 = F.conv2d(x, sd["CostRegNet.dispconv_0.weight"], sd["CostRegNet.dispconv_0.bias"], 1, 1)
    return d1, depth_max * torch.sigmoid(d0)


# ----------------------------------------------------------------------------------------------
# geometry: plane-sweep homography warp (row a5) and frustum volume warp (row a10)
# ----------------------------------------------------------------------------------------------

def _unnormalize(coord, size):
    """ATen GridSampler.h grid_sampler_unnormalize, align_corners=False: ((c+1)*size-1)/2."""
    return ((coord + 1.0) * size - 1.0) / 2.0


def bilinear_zeros(src, xn, yn):
    """Explicit restatement of grid_sample(bilinear, zeros, align_corners=False) for one image.

    src [C,H,W]; xn, yn [...] normalised coords.  Returns [C, ...].
    """
    C, H, W = src.shape
    ix = _unnormalize(xn, W)
    iy = _unnormalize(yn, H)
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    out = torch.zeros((C,) + tuple(xn.shape), dtype=src.dtype)
    flat = src.reshape(C, H * W)
    for dy in (0, 1):
        for dx in (0, 1):
            xi = x0 + dx
            yi = y0 + dy
            # ATen: nw = (ix_se - ix)*(iy_se - iy) etc. with ix_se = x0+1, iy_se = y0+1
            wx = (x0 + 1 - ix) if dx == 0 else (ix - x0)
            wy = (y0 + 1 - iy) if dy == 0 else (iy - y0)
            ok = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1)
            idx = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1)).long()
            tap = flat[:, idx.reshape(-1)].reshape((C,) + tuple(xn.shape))
            out = out + tap * (wx * wy * ok.to(src.dtype)).unsqueeze(0)
    return out


def trilinear_zeros(vol, xn, yn, zn):
    """Explicit grid_sample for a [C,D,H,W] volume (5-D 'bilinear' = trilinear, zeros, align_corners=False)."""
    C, D, H, W = vol.shape
    ix, iy, iz = _unnormalize(xn, W), _unnormalize(yn, H), _unnormalize(zn, D)
    x0, y0, z0 = torch.floor(ix), torch.floor(iy), torch.floor(iz)
    out = torch.zeros((C,) + tuple(xn.shape), dtype=vol.dtype)
    flat = vol.reshape(C, D * H * W)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                xi, yi, zi = x0 + dx, y0 + dy, z0 + dz
                wx = (x0 + 1 - ix) if dx == 0 else (ix - x0)
                wy = (y0 + 1 - iy) if dy == 0 else (iy - y0)
                wz = (z0 + 1 - iz) if dz == 0 else (iz - z0)
                ok = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1) & (zi >= 0) & (zi <= D - 1)
                idx = ((zi.clamp(0, D - 1) * H + yi.clamp(0, H - 1)) * W + xi.clamp(0, W - 1)).long()
                tap = flat[:, idx.reshape(-1)].reshape((C,) + tuple(xn.shape))
                out = out + tap * (wx * wy * wz * ok.to(vol.dtype)).unsqueeze(0)
    return out


def _force_outside(c):
    """coords outside [-1,1] are set to 2 so that every tap is out of bounds (homo_utils.py:488-491, :193-198)."""
    return torch.where((c > 1) | (c < -1), torch.full_like(c, 2.0), c)


def plane_sweep_grid(src_proj, ref_proj, depth_values, H, W, homo12=None):
    """Normalised sample coords of utils/homo_utils.py:469-491.  Returns xn, yn [D, H*W] (B=1).

    src_proj/ref_proj [4,4]; depth_values [D].  ``homo12`` (test hook): the 12 numbers [rot (9) | trans (3)] of
    ``src_proj @ inverse(ref_proj)`` computed elsewhere (e.g. by the same torch ops on a GPU, whose LU differs from
    LAPACK's in the last bit) -- the rest of the arithmetic is unchanged.
    """
    # batched (B=1) matmuls on purpose: ATen picks bmm for [1,3,3]x[1,3,HW] and its fp32 summation order
    # differs from the 2-D mm kernel by a few 1e-7 -- enough to move taps by 3e-5 in feature units.
    if homo12 is None:
        proj = torch.matmul(src_proj.unsqueeze(0), torch.inverse(ref_proj.unsqueeze(0)))
        rot, trans = proj[:, :3, :3], proj[:, :3, 3:4]
    else:
        rot, trans = homo12[:9].reshape(1, 3, 3), homo12[9:12].reshape(1, 3, 1)
    y, x = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    xyz = torch.stack((x.reshape(-1), y.reshape(-1), torch.ones(H * W))).unsqueeze(0)   # [1, 3, HW]
    rot_xyz = torch.matmul(rot, xyz)[0]                                                 # [3, HW]
    p = rot_xyz.unsqueeze(1) * depth_values.view(1, -1, 1) + trans.view(3, 1, 1)        # [3, D, HW]
    px = p[0] / (p[2] + 1e-8)
    py = p[1] / (p[2] + 1e-8)
    xn = _force_outside(px / ((W - 1) / 2) - 1)
    yn = _force_outside(py / ((H - 1) / 2) - 1)
    return xn, yn


def homo_warp(src_fea, src_proj, ref_proj, depth_values, sampler="aten", homo12=None, align_corners=False):
    """utils/homo_utils.py:458-504.  src_fea [1,C,H,W] -> [1,C,D,H,W].  ``align_corners``: the grid_sample convention; False =
    what the reference computes under torch >= 1.3 (and what the fixtures pin), True = the torch 1.2 it was written for (Q1)."""
    _, C, H, W = src_fea.shape
    D = depth_values.numel()
    xn, yn = plane_sweep_grid(src_proj[0], ref_proj[0], depth_values.reshape(-1), H, W, homo12)
    if sampler == "aten":
        grid = torch.stack((xn, yn), dim=2).view(1, D * H, W, 2)
        out = F.grid_sample(src_fea, grid, mode="bilinear", padding_mode="zeros", align_corners=align_corners)
        return out.view(1, C, D, H, W)
    assert not align_corners, "the explicit sampler restates align_corners=False only"
    return bilinear_zeros(src_fea[0], xn.view(D, H, W), yn.view(D, H, W)).unsqueeze(0)


def volume_warp_grid(rel_pose, cam_intr, depth_values, D, H, W, depth_min, depth_interval, table30=None):
    """Normalised coords of warp_volume (homo_utils.py:240-271 with helpers :40-62, :26-37, :107-134, :170-205).

    rel_pose [4,4] (= P_j . P_i^-1, hybrid_depth_decoder.py:235), cam_intr [3,3] (1/4-scaled K),
    depth_values [D].  Returns xn, yn, zn [D, H*W].  ``table30`` (test hook): [inverse(K) (9) | first three rows of
    inverse(rel_pose) (12) | K (9)] computed elsewhere, used in place of the two ``torch.inverse`` calls.
    """
    y, x = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    pix = torch.stack((x.reshape(-1), y.reshape(-1), torch.ones(H * W)))             # set_id_grid :7-14
    # the reference broadcasts the pixel grid over D *before* the K^-1 product (pixel grid [B,3,D,HW], :52-54)
    pix = pix.view(1, 3, 1, H * W).repeat(1, 1, D, 1).view(1, 3, -1)
    if table30 is None:
        kinv, rel_inv = torch.inverse(cam_intr.unsqueeze(0)), torch.inverse(rel_pose.unsqueeze(0))
    else:
        kinv = table30[:9].reshape(1, 3, 3)
        rel_inv = torch.cat([table30[9:21].reshape(3, 4), torch.tensor([[0.0, 0.0, 0.0, 1.0]])], 0).unsqueeze(0)
    ray = kinv.bmm(pix).view(3, D, H * W)                                             # pixel2cam :51-54
    cam = ray * depth_values.view(1, D, 1)                                            # [3, D, HW]
    cam4 = torch.cat([cam, torch.ones(1, D, H * W)], 0).reshape(1, 4, -1)
    src = torch.bmm(rel_inv, cam4)                                                    # cam2cam :26-37
    uvw = torch.bmm(cam_intr.unsqueeze(0), src[:, :3])[0]                             # cam2pixel_depth :116
    px = (uvw[0] / (uvw[2] + 1e-10)).view(D, H * W)
    py = (uvw[1] / (uvw[2] + 1e-10)).view(D, H * W)
    pz = uvw[2].view(D, H * W)
    xn = _force_outside(2 * px / (W - 1) - 1)
    yn = _force_outside(2 * py / (H - 1) - 1)
    zn = _force_outside(2 * ((pz - depth_min) / depth_interval) / (D - 1) - 1.0)
    return xn, yn, zn


def warp_volume(vol, rel_pose, cam_intr, depth_values, depth_min, depth_interval, sampler="aten", table30=None, align_corners=False):
    """utils/homo_utils.py:240-279, zeros padding.  vol [1,C,D,H,W] -> same shape."""
    _, C, D, H, W = vol.shape
    xn, yn, zn = volume_warp_grid(rel_pose[0], cam_intr[0], depth_values.reshape(-1), D, H, W, depth_min, depth_interval,
                                  table30)
    if sampler == "aten":
        grid = torch.stack((xn, yn, zn), dim=2).view(1, D, H, W, 3)
        return F.grid_sample(vol, grid, mode="bilinear", padding_mode="zeros", align_corners=align_corners)
    assert not align_corners, "the explicit sampler restates align_corners=False only"
    return trilinear_zeros(vol[0], xn.view(D, H, W), yn.view(D, H, W), zn.view(D, H, W)).unsqueeze(0)


# ----------------------------------------------------------------------------------------------
# cost volume (row a4/a6), matching net (a8), soft-argmin (a9), EST fusion (a11)
# ----------------------------------------------------------------------------------------------

def cost_volume(sd, feats, poses, cam_intr, depth_values, sampler="aten", taps=None, homo=None, align_corners=False):
    """hybrid_models/model_hybrid.py:62-102.  feats: 3 maps [1,32,H,W]; poses [1,3,4,4]; middle = target.
    ``homo`` (test hook): [2, 12] precomputed [rot | trans] of the two sources (see plane_sweep_grid)."""
    ref = feats[1]
    D = depth_values.numel()
    ref_ext = torch.inverse(poses[:, 1])
    ref_volume = ref.unsqueeze(2).repeat(1, 1, D, 1, 1)
    total = torch.zeros_like(ref_volume)
    for v in (0, 2):
        src_ext = torch.inverse(poses[:, v])
        src_proj, ref_proj = src_ext.clone(), ref_ext.clone()
        src_proj[:, :3, :4] = cam_intr @ src_ext[:, :3, :4]
        ref_proj[:, :3, :4] = cam_intr @ ref_ext[:, :3, :4]
        warped = homo_warp(feats[v], src_proj, ref_proj, depth_values, sampler, None if homo is None else homo[v // 2], align_corners)
        x = _cb3(torch.cat([ref_volume, warped], 1), sd, "pre0")
        if taps is not None:
            taps.setdefault("x0", []).append(x)
        x = x + _cb3(_cb3(x, sd, "pre1", "relu"), sd, "pre2")
        total = total + x
    return total / 2


def matching_net(sd, cost_volumes, semantic_vs):
    """hybrid_depth_decoder.py:187-200: dres0, dres1, cat(context as channel 0), dres2, value/key, head0."""
    x = torch.cat(cost_volumes, 0)
    for name in ("dres0.0", "dres0.1", "dres1.0", "dres1.1"):
        x = _cb3(x, sd, "CostRegNet." + name, "relu")
    x = torch.cat([semantic_vs.unsqueeze(1), x], 1)
    x = _cb3(x, sd, "CostRegNet.dres2.0", "relu")
    value = _cb3(x, sd, "CostRegNet.value_layer.0", "tanh")
    key = _cb3(x, sd, "CostRegNet.key_layer.0", "relu")
    return value, key, stereo_head(sd, value, 0)


def stereo_head(sd, vol, which):
    """stereo_head{0,1}: conv3d+BN+ReLU then 1x1x1 conv with bias (hybrid_depth_decoder.py:104-112) -> [T,D,H,W]."""
    p = "CostRegNet.stereo_head%d" % which
    y = _cb3(vol, sd, p + ".0", "relu")
    return F.conv3d(y, sd[p + ".1.weight"], sd[p + ".1.bias"]).squeeze(1)


def soft_argmin(logits_quarter, depth_values, up=4):
    """nearest x4 then depthlayer (hybrid_depth_decoder.py:33-38, 202-204).

    Returns depth [T,1,H,W], prob [T,1,H,W], argmax index [T,1,H,W] (the index the reference discards at :36).
    """
    logits = F.interpolate(logits_quarter, scale_factor=up) if up > 1 else logits_quarter
    p = F.softmax(logits, dim=1)
    depth = torch.sum(p * depth_values.view(1, -1, 1, 1), dim=1, keepdim=True)
    prob, idx = torch.max(p, dim=1, keepdim=True)
    return depth, prob, idx


def est_attention(target_key, warped_keys, warped_values):
    """transformer/epipolar_transformer.py:63-73: per-voxel softmax over N sources, MEAN of weighted values."""
    corr = torch.stack([torch.sum(target_key * k, dim=1, keepdim=True) for k in warped_keys], dim=-1)
    att = F.softmax(corr, dim=-1)
    vals = torch.stack(warped_values, dim=-1)
    return torch.mean(vals * att, dim=-1)


def est_gru(sd, x, h, prefix="CostRegNet.epipolar_transformer"):
    """ConvGRU with GroupNorm(1,16) gates (epipolar_transformer.py:31-54, 80-83)."""
    p = prefix
    f = F.conv3d(torch.cat((x, h), 1), sd[p + ".gate_conv.weight"], sd[p + ".gate_conv.bias"], 1, 1)
    half = f.shape[1] // 2
    r = torch.sigmoid(F.group_norm(f[:, :half], 1, sd[p + ".reset_gate_norm.weight"], sd[p + ".reset_gate_norm.bias"], 1e-5))
    u = torch.sigmoid(F.group_norm(f[:, half:], 1, sd[p + ".update_gate_norm.weight"], sd[p + ".update_gate_norm.bias"], 1e-5))
    o = F.conv3d(torch.cat((x, r * h), 1), sd[p + ".output_conv.weight"], sd[p + ".output_conv.bias"], 1, 1)
    o = F.group_norm(o, 1, sd[p + ".output_norm.weight"], sd[p + ".output_norm.bias"], 1e-5)
    return u * h + (1 - u) * torch.tanh(o)


def est_fuse(sd, target_key, warped_keys, target_value, warped_values):
    """EpipolarTransformer.forward (epipolar_transformer.py:56-83)."""
    if warped_values:
        h = est_attention(target_key, warped_keys, warped_values)
    else:
        h = torch.zeros_like(target_value)
    return est_gru(sd, target_value, h)


# ----------------------------------------------------------------------------------------------
# the whole forward (rows a1, a12)
# ----------------------------------------------------------------------------------------------

def depth_planes(cfg):
    """model_hybrid.py:28-33: d_i = depth_min + i * (depth_max - depth_min)/(D-1), fp32."""
    interval = (cfg["depth_max"] - cfg["depth_min"]) / (cfg["ndepths"] - 1)
    return torch.arange(0, cfg["ndepths"]).to(torch.float32) * interval + cfg["depth_min"], interval


def forward(sd, cfg, imgs, cam_poses, cam_intr, pre_costs=None, pre_cam_poses=None, sampler="aten", taps=None, geometry=None,
            align_corners=False):
    """``DepthNetHybrid.forward(..., mode='val')`` (model_hybrid.py:110-184 + hybrid_depth_decoder.py:138-432).

    cfg = dict(ndepths, depth_min, depth_max, resnet, est=True).  imgs [1,V,3,H,W] in 0..255,
    cam_poses [1,V,4,4] cam->world, cam_intr [1,3,3].  Returns (outputs, state, poses) exactly like the
    reference: state = {"keys": [k], "values": [v]}, poses = [pose] with the stale-pose quirk Q4.
    ``taps`` (optional dict) receives intermediate tensors for seam-level tests.
    ``geometry`` (test hook, optional): {"homo": [2T, 12], "warp": list over targets of [n_sources, 30]} -- the camera
    matrices of the two warps computed elsewhere (layout of estdepth_b200.ops.homography_table_torch /
    volume_warp_tables_torch), used instead of this function's own ``torch.inverse`` products; everything downstream
    of the matrices is unchanged.  Lets a test hand the CPU oracle the matrices a GPU's LU produced.
    ``align_corners``: grid_sample convention of both warps (quirk Q1); False = the reference as it runs today.
    """
    assert imgs.shape[0] == 1, "the reference (and this oracle) run at B=1 (quirk Q16)"
    imgs = 2 * (imgs / 255.) - 1.
    _, V, _, Hi, Wi = imgs.shape
    H, W = Hi // 4, Wi // 4
    assert V > 2
    T = V - 2
    D = cfg["ndepths"]
    feats = psm_features(sd, imgs.view(V, 3, Hi, Wi))
    feats = [feats[v:v + 1] for v in range(V)]
    maps = resnet_maps(sd, imgs[0, 1:1 + T], cfg["resnet"])
    K4 = cam_intr.clone()
    K4[:, :2, :] *= 0.25                                                      # scale_cam_intr :104-108
    depth_values, interval = depth_planes(cfg)
    cvs = [cost_volume(sd, feats[t:t + 3], cam_poses[:, t:t + 3], K4, depth_values, sampler, taps,
                       None if geometry is None else geometry["homo"][2 * t:2 * t + 2], align_corners) for t in range(T)]
    poses = [cam_poses[:, t + 1] for t in range(T)]
    if taps is not None:
        taps["features"] = feats
        taps["cost_volumes"] = cvs

    outputs = {}
    semantic_vs = context_decoder(sd, maps)
    value, key, init_logits = matching_net(sd, cvs, semantic_vs)
    if taps is not None:
        taps.update(semantic_vs=semantic_vs, value=value, key=key, init_logits=init_logits)
    d3, p3, i3 = soft_argmin(init_logits, depth_values)
    for t in range(T):
        outputs[("depth", t, 3)] = d3[t:t + 1]
        outputs[("init_prob", t)] = p3[t:t + 1]
        outputs[("init_argmax", t)] = i3[t:t + 1]
    values = [value[t:t + 1] for t in range(T)]
    keys = [key[t:t + 1] for t in range(T)]
    out_keys, out_values = list(keys), list(values)

    use_est = cfg.get("est", True) and pre_costs is not None                   # quirk Q3 (:423)
    if use_est:
        poses = poses + list(pre_cam_poses)                                      # quirk Q4 (:221)
        values = values + list(pre_costs["values"])
        keys = keys + list(pre_costs["keys"])
        fused_logits = []
        for i in range(T):                                                       # in order: quirk Q5 (:253)
            wk, wv = [], []
            for j in range(len(poses)):
                if j == i:
                    continue
                rel = poses[j] @ torch.inverse(poses[i])                         # quirk Q7 (:235)
                tab = None if geometry is None else geometry["warp"][i][len(wk)]
                wk.append(warp_volume(keys[j], rel, K4, depth_values, cfg["depth_min"], interval, sampler, tab, align_corners))
                wv.append(warp_volume(values[j], rel, K4, depth_values, cfg["depth_min"], interval, sampler, tab, align_corners))
            fused = est_fuse(sd, keys[i], wk, values[i], wv)
            if taps is not None:
                taps.setdefault("h", []).append(est_attention(keys[i], wk, wv))
                taps.setdefault("fused", []).append(fused)
            values[i] = fused
            out_values[i] = fused
            lg = stereo_head(sd, fused, 1)
            fused_logits.append(lg)
            d2, p2, i2 = soft_argmin(lg, depth_values)
            outputs[("depth", i, 2)], outputs[("fused_prob", i)], outputs[("fused_argmax", i)] = d2, p2, i2
        fused_logits = torch.cat(fused_logits, 0)
    else:
        fused_logits = stereo_head(sd, value, 1)                                 # :377
        d2, p2, i2 = soft_argmin(fused_logits, depth_values)
        for t in range(T):
            outputs[("depth", t, 2)] = d2[t:t + 1]
            outputs[("fused_prob", t)] = p2[t:t + 1]
            outputs[("fused_argmax", t)] = i2[t:t + 1]
    if taps is not None:
        taps["fused_logits"] = fused_logits
    d1, d0 = refine_2d(sd, semantic_vs, fused_logits, maps[0], cfg["depth_max"])
    for t in range(T):
        outputs[("depth", t, 1)] = d1[t:t + 1]
        outputs[("depth", t, 0)] = d0[t:t + 1]
    return outputs, {"keys": out_keys[-1:], "values": out_values[-1:]}, poses[-1:]
