"""Shared helpers for the parity tests."""
import argparse
import importlib
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"

MAPPING = dict(bound_min=[-3.5, -0.5, -2.5], bound_max=[4.5, 3.5, 5.5], voxel_size=0.1, prune_min_vox_obs=16,
               ignore_count_th=16.0, encoder_count_th=600.0)               # configs/fusion-lr-kt.yaml:27-35
TRACKING = dict(iter_config=[{"n": 10, "type": [["rgb", 2]]}, {"n": 10, "type": [["sdf"], ["rgb", 1]]},
                             {"n": 50, "type": [["sdf"], ["rgb", 0]]}],
                sdf=dict(robust_kernel="huber", robust_k=5.0, subsample=0.5),
                rgb=dict(weight=500.0, robust_kernel=None, robust_k=0.01, min_grad_scale=0.0, max_depth_delta=0.2))


def ns(d):
    a = argparse.Namespace()
    a.__dict__.update(d)
    return a


def pkg():
    return importlib.import_module("nerf-fusion_b200")


def make_map(weights, device="cuda:0", **over):
    d = dict(MAPPING); d.update(over)
    return pkg().DenseIndexedMap(weights, ns(d), 29, torch.device(device))


def make_oracle_map(weights, **over):
    from oracle.map_oracle import OracleMap
    d = dict(MAPPING); d.update(over)
    return OracleMap(weights, d["bound_min"], d["bound_max"], d["voxel_size"], 29, d["prune_min_vox_obs"], d["ignore_count_th"],
                     d["encoder_count_th"])


def synth_cloud(n_frames=1, frame=0, H=480, W=640):
    """Preprocessed (oracle) cloud of a synthetic frame, cached on disk under /tmp to keep the suite fast."""
    from oracle import tracker_oracle
    synth = pkg().synth
    cache = Path("/tmp") / f"dfb_cloud_{frame}_{H}x{W}.npz"
    if cache.exists():
        z = np.load(cache)
        return z["P"], z["N"]
    seq = synth.SyntheticSequence(n_frames=frame + 1, H=H, W=W)
    depth, _ = seq.frame(frame)
    depth = depth.clone(); depth[(depth < 0.5) | (depth > 5.0)] = float("nan")
    calib = tuple(c * H / 480.0 for c in synth.ICL_CALIB)
    P, N = tracker_oracle.preprocess(depth.numpy(), calib)
    np.savez(cache, P=P, N=N)
    return P, N


def to_world(P, N):
    synth = pkg().synth
    R0 = synth.quat_to_R(synth.FIRST_TQ[3:]).astype(np.float32); t0 = np.asarray(synth.FIRST_TQ[:3], np.float32)
    return (P @ R0.T + t0).astype(np.float32), (N @ R0.T).astype(np.float32)


def sort_rows(a):
    a = np.asarray(a)
    idx = np.lexsort(a.reshape(a.shape[0], -1).T[::-1])
    return a[idx], idx


def match_rows(a, b, tol):
    """Order-free comparison of two row sets: every row of `a` has a distinct partner in `b` within `tol` (max-abs).
    Returns (perm, max_err) with a[i] ~ b[perm[i]]."""
    from scipy.spatial import cKDTree
    a = np.asarray(a, dtype=np.float64).reshape(len(a), -1); b = np.asarray(b, dtype=np.float64).reshape(len(b), -1)
    assert a.shape == b.shape, (a.shape, b.shape)
    if len(a) == 0:
        return np.zeros(0, dtype=np.int64), 0.0
    d, j = cKDTree(b).query(a, k=1, p=np.inf)
    assert d.max() < tol, f"unmatched row: max distance {d.max()}"
    assert len(np.unique(j)) == len(j) or np.unique(np.round(a, 5), axis=0).shape[0] < len(a), "matching is not one-to-one"
    return j, float(d.max())
