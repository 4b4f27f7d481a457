"""Deterministic synthetic weights and inputs (there is no checkpoint or dataset offline).

The published ESTDepth checkpoint is a Google-Drive link (reference README.md:86) and the datasets are
not in the image, so the benchmark, the smoke test and the parity fixtures all use

  * ``synth_state_dict``: a pure function of (key names, shapes, seed) that fills a state dict so that
    activations stay O(1) through the ~25 un-normalised residual adds of the feature nets and the
    depth logits spread over several units (default torch init gives logits with sigma ~ 1e-3, i.e. a
    uniform softmax and depth == 5.05 everywhere whatever the kernels do -- SURVEY.md section 7);
  * ``synth_inputs``: the synthetic 7-Scenes-like window of SURVEY.md section 8(d): low-passed random
    images, intrinsics of data/general_eval.py:168-176 scaled to the image size, a slow camera track
    (translation + yaw / pitch / roll) that keeps every projected depth positive.

Both are bit-reproducible on any machine with the same torch build (CPU generators only).
"""
import math
import zlib

import torch
import torch.nn.functional as F

HEAD_GAIN = 3.0           # scale of the 1x1x1 logit heads


def _gen(seed, key):
    g = torch.Generator(device="cpu")
    g.manual_seed((int(seed) * 1000003 + zlib.crc32(key.encode())) % (2 ** 31 - 1))
    return g


def _is_branch_tail(key):
    """BN whose output is added to a skip connection: keep its gain small so residual sums do not explode."""
    if key.startswith("matchingFeature.layer") and ".conv2.1." in key:
        return True
    if key.startswith("semanticFeature.encoder.layer"):
        # BasicBlock: bn2, Bottleneck: bn3 (bn2 of a Bottleneck is followed by conv3 -- harmless to damp too)
        return ".bn3." in key or (".bn2." in key and ".bn3." not in key)
    return key.startswith("pre2.1.")


def synth_state_dict(reference_state, seed=0, head_gain=HEAD_GAIN):
    """Returns a new state dict with the same keys/shapes/dtypes as ``reference_state``.

    ``reference_state`` only provides names and shapes (any module's ``state_dict()``).
    """
    out = {}
    keys = list(reference_state.keys())
    keyset = set(keys)
    for key in keys:
        ref = reference_state[key]
        shape = tuple(ref.shape)
        g = _gen(seed, key)
        leaf = key.rsplit(".", 1)[-1]
        stem = key.rsplit(".", 1)[0]
        is_norm = (stem + ".running_mean") in keyset or "_norm" in stem
        if leaf == "num_batches_tracked":
            val = torch.zeros(shape, dtype=ref.dtype)
        elif leaf == "running_mean":
            val = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "running_var":
            val = 0.7 + 0.6 * torch.rand(shape, generator=g)
        elif is_norm and leaf == "weight":
            val = 0.7 + 0.6 * torch.rand(shape, generator=g)
            if _is_branch_tail(key):
                val = val * 0.25
        elif is_norm and leaf == "bias":
            val = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "weight" and len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            val = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
            if ".stereo_head" in key and shape[0] == 1:
                val = val * head_gain
        elif leaf == "bias":
            val = 0.1 * torch.randn(shape, generator=g)
        else:
            val = torch.randn(shape, generator=g)
        out[key] = val.to(ref.dtype)
    return out


def camera_track(num_views, start=0, dtype=torch.float32):
    """cam->world poses [V,4,4]: translation (0.05 v, 0.013 v, 0.01 v) m, yaw 0.02 v, pitch 0.011 v and roll 0.007 v rad.

    SURVEY.md 8d proposed translate-in-x/z + yaw only.  That track keeps every image ROW on itself wherever the depth of a
    point is the same in both views, so the warps of the first / last row land exactly on the +-1 cut of the reference's
    sampling range (quirk Q10) and fp32 round-off decides, voxel by voxel, between 'sampled' and 'zero-filled' -- a property
    of the synthetic geometry, not of real sequences.  A little vertical motion, pitch and roll makes the cut generic again.
    """
    poses = torch.zeros(num_views, 4, 4, dtype=torch.float64)
    for i in range(num_views):
        v = start + i
        yaw, pitch, roll = 0.02 * v, 0.011 * v, 0.007 * v
        cy, sy = math.cos(yaw), math.sin(yaw)
        cp, sp = math.cos(pitch), math.sin(pitch)
        cr, sr = math.cos(roll), math.sin(roll)
        Ry = torch.tensor([[cy, 0.0, sy], [0.0, 1.0, 0.0], [-sy, 0.0, cy]], dtype=torch.float64)
        Rx = torch.tensor([[1.0, 0.0, 0.0], [0.0, cp, -sp], [0.0, sp, cp]], dtype=torch.float64)
        Rz = torch.tensor([[cr, -sr, 0.0], [sr, cr, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
        poses[i, :3, :3] = Ry @ Rx @ Rz
        poses[i, :3, 3] = torch.tensor([0.05 * v, 0.013 * v, 0.01 * v], dtype=torch.float64)
        poses[i, 3, 3] = 1.0
    return poses.to(dtype)


def intrinsics(height, width, dtype=torch.float32):
    """K of data/general_eval.py:168-176 (577.87, 319.5, 239.5 at 640x480) scaled to (height, width)."""
    sx, sy = width / 640.0, height / 480.0
    return torch.tensor([[577.87 * sx, 0.0, 319.5 * sx],
                         [0.0, 577.87 * sy, 239.5 * sy],
                         [0.0, 0.0, 1.0]], dtype=dtype)


def synth_images(num_views, height, width, seed=0, start=0):
    """[V,3,H,W] in 0..255: per-frame white noise, 8x8 average-pooled and bilinearly re-expanded.

    Frame ``start+i`` depends only on (seed, start+i) so overlapping windows see identical frames.
    """
    frames = []
    for i in range(num_views):
        g = _gen(seed, "frame%d" % (start + i))
        noise = torch.rand(1, 3, height, width, generator=g)
        low = F.avg_pool2d(noise, 8)
        img = F.interpolate(low, size=(height, width), mode="bilinear", align_corners=False)
        img = (img - img.min()) / (img.max() - img.min() + 1e-12)
        frames.append(255.0 * (0.75 * img + 0.25 * noise))
    return torch.cat(frames, 0)


def synth_inputs(num_views, height, width, seed=0, start=0, batch=1):
    """One synthetic window in the layout the reference's drivers feed ``DepthNetHybrid.forward``.

    Returns (imgs [B,V,3,H,W], cam_poses [B,V,4,4], cam_intr [B,3,3], sample dict).
    """
    imgs = torch.stack([synth_images(num_views, height, width, seed=seed + 7919 * b, start=start)
                        for b in range(batch)], 0)
    poses = camera_track(num_views, start).unsqueeze(0).repeat(batch, 1, 1, 1)
    intr = intrinsics(height, width).unsqueeze(0).repeat(batch, 1, 1)
    sample = {"dmaps": torch.zeros(batch, num_views, 1, height, width),
              "dmasks": torch.ones(batch, num_views, 1, height, width, dtype=torch.bool)}
    return imgs, poses, intr, sample
