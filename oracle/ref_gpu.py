"""Run the reference's own Python hot path ON A GPU.  TEST INFRASTRUCTURE ONLY (tests/, bench.py's reference legs).

The reference modules (`system/map.py`, `system/tracker.py`, `network/*`, `utils/*`) are imported UNMODIFIED from
`oracle/_ref/pyref/` (staged from /root/reference by oracle/stage_ref_py.py; git-ignored, shipped by gpurun) or from
/root/reference itself when that exists.  Its native layer `system.ext` (system/ext/__init__.py:13-42, a JIT build) is
replaced by a proxy module whose seven ops dispatch to a backend chosen at run time:

  * "reference": the reference's four CUDA extensions, prebuilt from its sources by oracle/build_ref_ext.py
                 (`oracle/_ref/ext_build/*/ref_*.so`) -- the reference CUDA path, as a user of the reference runs it;
  * "dfb":       `nerf-fusion_b200.ext` -- the operator-level drop-in of INTEGRATION.md §2: the reference's map.py and
                 tracker.py executing on this repo's C-ABI ops.

Other shims, for dependencies the image lacks: numpy-2 `np.product`, `open3d` / `matplotlib` stubs (map.py:6 imports
open3d at module scope; only the meshing output and the GUI use it), `pyquaternion` -> oracle/pyquat_shim.py,
`torch_scatter.scatter_mean` -> an index_add based mean on the tensor's device ("reference") or `dfb.ext.scatter_mean`.
"""
import importlib
import importlib.util
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
PYREF = HERE / "_ref" / "pyref"
EXT_BUILD = HERE / "_ref" / "ext_build"
_BACKEND = {"name": None, "ops": None, "scatter_mean": None}
_OPS = ("unproject_depth", "gradient_xy", "rgb_odometry", "groupby_sum", "remove_radius_outlier", "estimate_normals",
        "marching_cubes_interp")
_installed = False


def ref_root():
    if (PYREF / "system" / "map.py").exists():
        return PYREF
    return Path(os.environ.get("DFB_REFERENCE_ROOT", "/root/reference"))


def available(need_ext=True):
    ok = (ref_root() / "system" / "map.py").exists()
    if need_ext:
        ok = ok and all((EXT_BUILD / n / f"ref_{n}.so").exists() for n in ("indexing", "marching_cubes", "imgproc", "pcproc"))
    return ok


def _load_so(name):
    so = EXT_BUILD / name / f"ref_{name}.so"
    spec = importlib.util.spec_from_file_location(f"ref_{name}", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _torch_scatter_mean(src, index, dim=0):
    """torch_scatter.scatter_mean (dim 0): sum by index_add, divided by the clamped count."""
    assert dim == 0
    n = int(index.max().item()) + 1 if index.numel() else 0
    out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).index_add_(0, index, src)
    cnt = torch.zeros((n,), dtype=src.dtype, device=src.device).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
    return out / cnt.clamp_(min=1).view(-1, *([1] * (src.dim() - 1)))


def use_backend(name):
    """Select what the reference's `system.ext` / `torch_scatter` calls run on: "reference" or "dfb"."""
    if name == "reference":
        mc, ip, ix, pc = _load_so("marching_cubes"), _load_so("imgproc"), _load_so("indexing"), _load_so("pcproc")
        ops = dict(marching_cubes_interp=mc.marching_cubes_sparse_interp, unproject_depth=ip.unproject_depth,
                   rgb_odometry=ip.rgb_odometry, gradient_xy=ip.gradient_xy, groupby_sum=ix.groupby_sum,
                   remove_radius_outlier=pc.remove_radius_outlier, estimate_normals=pc.estimate_normals)
        sm = _torch_scatter_mean
    elif name == "dfb":
        ext = importlib.import_module("nerf-fusion_b200").ext
        ops = {k: getattr(ext, k) for k in _OPS}
        sm = ext.scatter_mean
    else:
        raise ValueError(name)
    _BACKEND.update(name=name, ops=ops, scatter_mean=sm)


def _proxy(opname):
    def call(*a, **k):
        return _BACKEND["ops"][opname](*a, **k)
    call.__name__ = opname
    return call


def install(backend="reference"):
    """Idempotent.  Returns a namespace with the reference modules (map, tracker, net_util, motion, exp, FrameIntrinsic)."""
    global _installed
    if not available(need_ext=(backend == "reference")):
        raise RuntimeError("reference python (oracle/_ref/pyref or /root/reference) or its built extensions are missing")
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
        ts.scatter_mean = lambda src, index, dim=0: _BACKEND["scatter_mean"](src, index, dim)
        sys.modules["torch_scatter"] = ts
        sys.path.insert(0, str(ref_root()))
        import system  # the reference's namespace package
        ext = types.ModuleType("system.ext")
        for op in _OPS:
            setattr(ext, op, _proxy(op))
        sys.modules["system.ext"] = ext
        system.ext = ext
        _installed = True
    use_backend(backend)
    import system.map as ref_map
    import system.tracker as ref_tracker
    import network.utility as ref_net_util
    import utils.motion_util as ref_motion
    import utils.exp_util as ref_exp
    from dataset.production import FrameIntrinsic
    return types.SimpleNamespace(map=ref_map, tracker=ref_tracker, net_util=ref_net_util, motion=ref_motion, exp=ref_exp,
                                 FrameIntrinsic=FrameIntrinsic, Quaternion=sys.modules["pyquaternion"].Quaternion)


def load_reference_model(device):
    """network/utility.py:22-58 (load_model) with the checkpoint path resolved against the staged tree."""
    ref = install(_BACKEND["name"] or "reference")
    root = ref_root()
    args = ref.exp.parse_config_json(root / "ckpt" / "default" / "hyper.json")
    model = ref.net_util.Networks()
    model.decoder = importlib.import_module("network." + args.network_name).Model(args.code_length, **args.network_specs).to(device)
    model.encoder = importlib.import_module("network." + args.encoder_name).Model(**args.encoder_specs).to(device)
    sd = torch.load(root / "ckpt" / "default" / "model_300.pth.tar", map_location=device, weights_only=False)["model_state"]
    model.decoder.load_state_dict(sd)
    se = torch.load(root / "ckpt" / "default" / "encoder_300.pth.tar", map_location=device, weights_only=False)["model_state"]
    model.encoder.load_state_dict(se)
    return model, args


def load_config():
    import yaml
    return yaml.safe_load((ref_root() / "configs" / "fusion-lr-kt.yaml").read_text())


def make_reference_system(device, iter_config=None, mapping_over=None):
    """(map, tracker, cfg) exactly as main.py:112-133 builds them (configs/fusion-lr-kt.yaml, ckpt/default, epoch 300)."""
    ref = install(_BACKEND["name"] or "reference")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model, margs = load_reference_model(device)
    cfg = load_config()
    mapping = dict(cfg["mapping"])
    if mapping_over:
        mapping.update(mapping_over)
    m = ref.map.DenseIndexedMap(model, ref.exp.dict_to_args(mapping), margs.code_length, torch.device(device), False, None)
    targs = ref.exp.dict_to_args(cfg["tracking"])
    if iter_config is not None:
        targs.iter_config = iter_config
    trk = ref.tracker.SDFTracker(m, targs)
    return m, trk, cfg
