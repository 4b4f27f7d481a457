"""estdepth_b200 -- B200-native (sm_100a) implementation of ESTDepth's plane-sweep + EST inference hot path.

Public surface:
  DepthNetHybrid   drop-in for hybrid_models.model_hybrid.DepthNetHybrid (inference, mode='val')
  ops              operators over the C ABI (include/estdepth_b200.h), incl. the reference-named
                   homo_warping / warp_volume / depthlayer
  synth            deterministic synthetic weights / inputs used by the bench, the smoke test and the fixtures
"""
from . import ops, packing, synth            # noqa: F401
from .model import DepthNetHybrid            # noqa: F401

__version__ = "0.1.0"
