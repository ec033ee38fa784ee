"""Import the reference's own Python hot path (/root/reference) on CPU.  TEST INFRASTRUCTURE ONLY.

Used in the BUILD CONTAINER by oracle/make_golden.py and by the `-m "not gpu"` pinning tests (skipped when
/root/reference is absent, as it is on the GPU box).  Nothing is copied: the reference modules are imported
from where they lie after installing the shims its missing dependencies need:

  * numpy 2: ``np.product`` alias (map.py:201,408);
  * ``open3d`` / ``matplotlib`` stubs (map.py:6 imports open3d at module scope);
  * ``pyquaternion`` -> oracle/pyquat_shim.py; ``torch_scatter.scatter_mean`` -> index_add based mean;
  * ``system.ext`` pre-seeded with the oracle's CPU restatements of the CUDA ops (oracle/ops.py), so the
    JIT build in system/ext/__init__.py:13-42 is skipped;
  * CUDA-only calls made CPU-safe: ``torch.cuda.Stream``, ``torch.cuda.stream``, ``torch.cuda.device``,
    ``torch.cuda.synchronize``, ``Tensor.cuda`` (map.py:232,626-627; tracker.py:108,202).
"""
import contextlib
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF_ROOT = Path(os.environ.get("DFB_REFERENCE_ROOT", "/root/reference"))


def available():
    return (REF_ROOT / "system" / "map.py").exists()


def _ext_module():
    from . import ops
    m = types.ModuleType("system.ext")

    def groupby_sum(values, indices, C):
        s, c = ops.groupby_sum(values.detach().numpy(), indices.numpy(), int(C))
        return torch.from_numpy(s), torch.from_numpy(c)

    def unproject_depth(depth, fx, fy, cx, cy):
        return torch.from_numpy(ops.unproject_depth(depth.numpy(), fx, fy, cx, cy))

    def remove_radius_outlier(pc, nb, radius):
        return torch.from_numpy(ops.remove_radius_outlier(pc.numpy(), nb, radius))

    def estimate_normals(pc, max_nn, radius, cam):
        return torch.from_numpy(ops.estimate_normals(pc.numpy(), max_nn, radius, cam))

    def gradient_xy(img):
        return torch.from_numpy(ops.gradient_xy(img.numpy()))

    def rgb_odometry(pI, pD, cI, cD, cG, intr, krkinv, kt, mgs, mdd, compute_J):
        out = ops.rgb_odometry(pI.numpy(), pD.numpy(), cI.numpy(), cD.numpy(), cG.numpy(), intr, krkinv, kt, mgs, mdd, compute_J)
        return [torch.from_numpy(t) for t in out]

    def marching_cubes_interp(indexer, valid_blocks, mapping, cube_sdf, cube_std, max_n, n_xyz, max_std):
        t, i, s = ops.marching_cubes_sparse_interp(indexer.numpy(), valid_blocks.numpy(), mapping.numpy(),
                                                   cube_sdf.numpy(), cube_std.numpy(), max_n, n_xyz, max_std)
        return torch.from_numpy(t), torch.from_numpy(i), torch.from_numpy(s)

    for f in (groupby_sum, unproject_depth, remove_radius_outlier, estimate_normals, gradient_xy, rgb_odometry,
              marching_cubes_interp):
        setattr(m, f.__name__, f)
    return m


class _DummyStream:
    def __init__(self, *a, **k):
        pass

    def synchronize(self):
        pass


_installed = False


def install():
    """Idempotent.  Returns a namespace with the reference modules: map, tracker, net_util, motion_util."""
    global _installed
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    if not _installed:
        if not hasattr(np, "product"):
            np.product = np.prod
        o3d = types.ModuleType("open3d")
        o3d.geometry = types.SimpleNamespace(); o3d.utility = types.SimpleNamespace(); o3d.visualization = types.SimpleNamespace()
        sys.modules.setdefault("open3d", o3d)
        if "matplotlib" not in sys.modules:
            try:
                import matplotlib  # noqa: F401
            except Exception:
                mpl = types.ModuleType("matplotlib"); mpl.cm = types.ModuleType("matplotlib.cm")
                sys.modules["matplotlib"] = mpl; sys.modules["matplotlib.cm"] = mpl.cm
        from . import pyquat_shim
        pq = types.ModuleType("pyquaternion"); pq.Quaternion = pyquat_shim.Quaternion
        sys.modules["pyquaternion"] = pq
        ts = types.ModuleType("torch_scatter")

        def scatter_mean(src, index, dim=0):
            from . import ops
            return torch.from_numpy(ops.scatter_mean(src.numpy(), index.numpy(), dim))
        ts.scatter_mean = scatter_mean
        sys.modules["torch_scatter"] = ts
        torch.cuda.Stream = _DummyStream
        torch.cuda.stream = lambda s: contextlib.nullcontext()
        torch.cuda.device = lambda d: contextlib.nullcontext()
        torch.cuda.synchronize = lambda *a, **k: None
        torch.Tensor.cuda = lambda self, *a, **k: self
        sys.path.insert(0, str(REF_ROOT))
        import system  # namespace package of the reference
        sys.modules["system.ext"] = _ext_module()
        system.ext = sys.modules["system.ext"]
        _installed = True
    import system.map as ref_map
    import system.tracker as ref_tracker
    import network.utility as ref_net_util
    import utils.motion_util as ref_motion
    import utils.exp_util as ref_exp
    return types.SimpleNamespace(map=ref_map, tracker=ref_tracker, net_util=ref_net_util, motion=ref_motion, exp=ref_exp)


def load_reference_model():
    """network/utility.py:22-58 without the hard-coded .cuda() (utility.py:46,49)."""
    ref = install()
    import importlib
    args = ref.exp.parse_config_json(REF_ROOT / "ckpt" / "default" / "hyper.json")
    model = ref.net_util.Networks()
    model.decoder = importlib.import_module("network." + args.network_name).Model(args.code_length, **args.network_specs)
    model.encoder = importlib.import_module("network." + args.encoder_name).Model(**args.encoder_specs)
    sd = torch.load(REF_ROOT / "ckpt" / "default" / "model_300.pth.tar", map_location="cpu", weights_only=False)["model_state"]
    model.decoder.load_state_dict(sd)
    se = torch.load(REF_ROOT / "ckpt" / "default" / "encoder_300.pth.tar", map_location="cpu", weights_only=False)["model_state"]
    model.encoder.load_state_dict(se)
    return model, args


def make_reference_map(model, device="cpu"):
    ref = install()
    import yaml
    cfg = yaml.safe_load((REF_ROOT / "configs" / "fusion-lr-kt.yaml").read_text())
    margs = ref.exp.dict_to_args(cfg["mapping"])
    m = ref.map.DenseIndexedMap(model, margs, 29, torch.device(device))
    return m, cfg
