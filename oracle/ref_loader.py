"""Import the UNMODIFIED reference (xxlong0/ESTDepth) from /root/reference for oracle validation.

TEST INFRASTRUCTURE ONLY.  This module exists so that, in the build container (where
/root/reference is mounted read-only), the oracle restatement in ``oracle/estdepth_oracle.py`` can be
checked against the reference's own code and golden fixtures can be generated
(``oracle/make_golden.py``).  Nothing here is shipped or imported by the product package, and
nothing here may run on the GPU box (the reference does not exist there).

Shims applied (no edits to the reference tree):
  1. torchvision: ``ResnetEncoder(resnet, "pretrained")`` calls ``models.resnet50("pretrained")``
     (hybrid_models/resnet_encoder.py:35) which fails on torchvision >= 0.13 -> proxy returning
     ``resnetNN(weights=None)``.
  2. ``utils/homo_utils.py:56`` (dead debug line ``tt = depth[:, 0, 62, :]``) raises IndexError for
     ndepths < 63 -> ``pixel2cam`` rebound to an equivalent without that line.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ESTD_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "hybrid_models", "model_hybrid.py"))


_loaded = {}


def load_reference():
    """Returns a namespace with the reference's modules (model_hybrid, homo_utils, ...)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    # The product repo ships a drop-in `hybrid_models` shim package with the same name; make sure
    # the reference's own packages win inside this process.
    for name in list(sys.modules):
        if name.split(".")[0] in ("hybrid_models", "networks", "transformer", "utils"):
            del sys.modules[name]
    # The reference's packages are namespace packages (no __init__.py); a regular package of the same name
    # anywhere on sys.path (the product's `hybrid_models` shim) would win, so hide such entries while importing.
    saved_path = list(sys.path)
    sys.path[:] = [REFERENCE_ROOT] + [q for q in saved_path
                                      if not os.path.isfile(os.path.join(q or ".", "hybrid_models", "__init__.py"))]
    try:
        import torch
        import torchvision.models as tvm
        import hybrid_models.resnet_encoder as ref_resnet_encoder

        class _Models(object):
            def __getattr__(self, name):
                fn = getattr(tvm, name)
                if name.startswith("resnet"):
                    return lambda *a, **k: fn(weights=None)
                return fn

        ref_resnet_encoder.models = _Models()

        import utils.homo_utils as ref_homo

        def _pixel2cam(depth, intrinsics, pixel_coords, is_homogeneous=True):
            # same arithmetic as utils/homo_utils.py:40-62 minus the debug line :56
            b, _, h, w = depth.size()
            kinv = torch.inverse(intrinsics)
            pc = pixel_coords[:, :, :h, :w].expand(b, 3, h, w).contiguous().view(b, 3, -1).to(depth.device)
            cam = kinv.bmm(pc).view(b, 3, h, w) * depth.repeat(1, 3, 1, 1)
            if is_homogeneous:
                cam = torch.cat([cam, torch.ones((b, 1, h, w), dtype=depth.dtype, device=depth.device)], dim=1)
            return cam

        ref_homo.pixel2cam = _pixel2cam

        import hybrid_models.model_hybrid as ref_model
        import hybrid_models.hybrid_depth_decoder as ref_decoder
        import transformer.epipolar_transformer as ref_est
        import networks.psm_submodule as ref_psm
        import networks.layers_op as ref_layers
        # the decoder did `from utils.homo_utils import *` before the rebind -> its global
        # `warp_volume` still resolves pixel2cam through utils.homo_utils' globals (rebound above).
    finally:
        sys.path[:] = saved_path
    _loaded.update(model_hybrid=ref_model, decoder=ref_decoder, est=ref_est, psm=ref_psm,
                   layers=ref_layers, homo=ref_homo, resnet_encoder=ref_resnet_encoder)
    # keep the reference's modules importable only through this namespace
    for name in list(sys.modules):
        if name.split(".")[0] in ("hybrid_models", "networks", "transformer", "utils"):
            _loaded.setdefault("_mods", {})[name] = sys.modules.pop(name)
    return types.SimpleNamespace(**_loaded)
