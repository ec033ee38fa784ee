"""difusion_b200: B200-native implementation of DI-Fusion's per-frame map hot path.

Host side (this package): ``ext`` (operator API = system.ext + torch_scatter names), ``map.DenseIndexedMap``,
``tracker.SDFTracker``, ``motion.Isometry`` -- mirrors of the reference interfaces for this path.
Device side: ``libdifusion_b200.so`` (hand-written sm_100a CUDA behind the C-ABI of include/difusion_b200.h).
The directory name has a hyphen; import it as ``nerf_fusion_b200`` (alias module at the repo root) or with
``importlib.import_module("nerf-fusion_b200")``.
"""
from . import _lib, weights, motion, synth          # noqa: F401  (no torch.cuda needed)
from . import ext, map, tracker, sharded, dataset   # noqa: F401
from .map import DenseIndexedMap                    # noqa: F401
from .tracker import SDFTracker, FrameIntrinsic     # noqa: F401
from .motion import Isometry, Quaternion            # noqa: F401
from ._lib import DfbError                          # noqa: F401

__all__ = ["ext", "map", "tracker", "motion", "weights", "synth", "sharded", "dataset", "DenseIndexedMap", "SDFTracker", "FrameIntrinsic",
           "Isometry", "Quaternion", "DfbError"]
