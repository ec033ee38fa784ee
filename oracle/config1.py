"""BASELINE config 1 as written: fusion-lr-kt.yaml + ckpt/default, N synthetic 640x480 frames, full iter_config
(10 / 10 / 50), integrate every 20 frames -- run through (a) the REFERENCE's own CUDA path (its map.py / tracker.py on its
own system/ext kernels, oracle/ref_gpu.py) and (b) this repo's class-level path, on the same device tensors.
TEST INFRASTRUCTURE ONLY: used by tests/test_gpu_config1.py, tools/config1_parity.py and bench.py's `cuda_reference` /
`parity` legs.  Nothing under nerf-fusion_b200/ imports it.

main.py:42-102 (`refresh`) is restated here without the GUI for both arms: depth cut, track_camera, integrate_keyframe
when frame_id % integrate_interval == 0.
"""
import importlib
import time

import numpy as np
import torch

from . import ref_gpu


def dfb_pkg():
    return importlib.import_module("nerf-fusion_b200")


def make_frames(n, device, seed=0, H=480, W=640):
    """Synthetic frames as a dataset delivers them (uint16 depth at 1/5000 m, uint8 colour) converted to float32 like
    dataset/production/icl_nuim.py:110-114; returns [(depth (H,W), rgb (H,W,3))] on `device`, calib tuple, sequence."""
    dfb = dfb_pkg()
    seq = dfb.synth.SyntheticSequence(n_frames=n, H=H, W=W, device=device, seed=seed)
    calib = tuple(c * H / 480.0 for c in dfb.synth.ICL_CALIB)
    frames = []
    for i in range(n):
        depth, rgb = seq.frame(i)
        d16 = torch.round(depth * 5000.0).to(torch.int32)
        c8 = torch.round(rgb.clamp(0.0, 1.0) * 255.0).to(torch.uint8)
        frames.append(((d16.float() / 5000.0).contiguous(), (c8.float() / 255.0).contiguous()))
    return frames, calib, seq


def _cut(depth, lo=0.5, hi=5.0):
    d = depth.clone()
    d[torch.logical_or(d < lo, d > hi)] = float("nan")       # main.py:56-57
    return d


def _pose_arrays(p):
    return np.asarray(p.q.rotation_matrix, dtype=np.float64).copy(), np.asarray(p.t, dtype=np.float64).copy()


def _snapshot_map(m):
    n = int(m.n_occupied)
    return dict(n_occupied=n, pos=m.latent_vecs_pos[:n].cpu().numpy().copy(), latent=m.latent_vecs[:n].cpu().numpy().copy(),
                count=m.voxel_obs_count[:n].cpu().numpy().copy())


def run_reference(frames, calib, device, iter_config=None, integrate_interval=20, backend="reference", keep_clouds=True):
    """The reference's SDFTracker + DenseIndexedMap on `device` (backend "reference": its own CUDA ops; "dfb": this repo's
    ops under the reference's unmodified Python = the operator-level drop-in)."""
    dfb = dfb_pkg()
    ref = ref_gpu.install(backend)
    m, trk, cfg = ref_gpu.make_reference_system(device, iter_config)
    counts = {"sdf": 0, "rgb": 0}
    o_sdf, o_rgb = trk.compute_sdf_Hg, trk.compute_rgb_Hg

    def c_sdf(*a, **k):
        counts["sdf"] += 1
        return o_sdf(*a, **k)

    def c_rgb(*a, **k):
        counts["rgb"] += 1
        return o_rgb(*a, **k)
    trk.compute_sdf_Hg, trk.compute_rgb_Hg = c_sdf, c_rgb
    first = ref.motion.Isometry(q=ref.Quaternion(array=dfb.synth.FIRST_TQ[3:]), t=np.array(dfb.synth.FIRST_TQ[:3]))
    fi = ref.FrameIntrinsic(*calib, 5000.0)
    out = dict(poses=[], clouds=[], frame_ms=[], n_points=[])
    for i, (depth, rgb) in enumerate(frames):
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        d = _cut(depth)
        pose = trk.track_camera(rgb, d, fi, first if len(trk.all_pd_pose) == 0 else None)
        pc, nrm = trk.last_processed_pc
        if i % integrate_interval == 0:
            m.integrate_keyframe(pose @ pc, pose.rotation @ nrm, async_optimize=False, do_optimize=False)
        torch.cuda.synchronize(device)
        out["frame_ms"].append((time.perf_counter() - t0) * 1e3)
        out["poses"].append(_pose_arrays(pose))
        out["n_points"].append(int(pc.size(0)))
        if keep_clouds:
            out["clouds"].append((pc.clone(), nrm.clone()))
    out["map"] = _snapshot_map(m)
    out["n_sdf"], out["n_rgb"] = counts["sdf"], counts["rgb"]
    out["map_obj"], out["tracker_obj"] = m, trk
    return out


def make_ours(device, iter_config=None, div_mode=1, mapping_over=None):
    """div_mode 1 (DFB_DIV_RECIP): torch CUDA divides by a Python scalar as a reciprocal multiply, which is what the
    reference executes on the GPU; 0 (DFB_DIV_IEEE) reproduces torch CPU."""
    import argparse
    import yaml
    dfb = dfb_pkg()
    from pathlib import Path
    gold = Path(__file__).resolve().parent.parent / "tests" / "golden"
    W = dfb.weights.load_npz(gold / "weights.npz")
    if ref_gpu.available(need_ext=False):
        cfg = ref_gpu.load_config()
        mapping, tracking = cfg["mapping"], cfg["tracking"]
    else:                                             # the same numbers, as tests/util.py holds them
        import sys
        sys.path.insert(0, str(gold.parent))
        from util import MAPPING, TRACKING
        mapping, tracking = dict(MAPPING), dict(TRACKING)
    if iter_config is not None:
        tracking = dict(tracking); tracking["iter_config"] = iter_config
    if mapping_over:
        mapping = dict(mapping); mapping.update(mapping_over)

    def ns(d):
        a = argparse.Namespace(); a.__dict__.update(d); return a
    m = dfb.DenseIndexedMap(W, ns(mapping), 29, torch.device(device), div_mode=div_mode)
    trk = dfb.SDFTracker(m, ns(tracking))
    return m, trk


def run_ours(frames, calib, device, iter_config=None, integrate_interval=20, keep_clouds=False):
    """This repo's class-level path (what bench.py times), engines as currently selected."""
    dfb = dfb_pkg()
    m, trk = make_ours(device, iter_config)
    first = dfb.Isometry(q=dfb.Quaternion(array=dfb.synth.FIRST_TQ[3:]), t=np.array(dfb.synth.FIRST_TQ[:3]))
    fi = dfb.FrameIntrinsic(*calib)
    out = dict(poses=[], clouds=[], frame_ms=[], n_points=[])
    for i, (depth, rgb) in enumerate(frames):
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        pose = trk.track_camera(rgb, depth, fi, first if len(trk.all_pd_pose) == 0 else None, depth_cut=(0.5, 5.0))
        pc, nrm = trk.last_processed_pc
        if i % integrate_interval == 0:
            m.integrate_keyframe(pose @ pc, pose.rotation @ nrm, do_optimize=False)
        torch.cuda.synchronize(device)
        out["frame_ms"].append((time.perf_counter() - t0) * 1e3)
        out["poses"].append(_pose_arrays(pose))
        out["n_points"].append(int(pc.size(0)))
        if keep_clouds:
            out["clouds"].append((pc.clone(), nrm.clone()))
    out["map"] = _snapshot_map(m)
    out["n_sdf"], out["n_rgb"] = trk.n_sdf_evals, trk.n_rgb_evals
    out["map_obj"], out["tracker_obj"] = m, trk
    return out


def run_ours_on_reference_points(frames, calib, device, ref_run, iter_config=None, integrate_interval=20):
    """Solver parity where it is well-posed: every frame's Gauss-Newton is fed the REFERENCE's preprocessed cloud, its
    previous pose and a map integrated from the reference's keyframe clouds at the reference's poses; what differs is
    what this repo computes in the solve (pyramids, photometric term, SDF term, device-resident Gauss-Newton)."""
    dfb = dfb_pkg()
    m, trk = make_ours(device, iter_config)
    fi = dfb.FrameIntrinsic(*calib)
    poses = []
    for i, (depth, rgb) in enumerate(frames):
        Rr, tr = ref_run["poses"][i]
        pose_ref = dfb.Isometry.from_matrix(Rr, tr)
        pc, nrm = ref_run["clouds"][i]
        d = _cut(depth)
        Ic, Dc, Gc = trk._make_image_pyramid(rgb.mean(-1), d)
        if i == 0:
            pose = pose_ref
        else:
            last = dfb.Isometry.from_matrix(*ref_run["poses"][i - 1])
            trk.all_pd_pose = [last]
            pose = trk.gauss_newton(last.dot(dfb.Isometry()), Ic, Dc, Gc, pc.contiguous(), fi)
        trk.last_intensity, trk.last_depth = Ic, Dc
        poses.append(_pose_arrays(pose))
        if i % integrate_interval == 0:
            m.integrate_keyframe(pose_ref @ pc, pose_ref.rotation @ nrm, do_optimize=False)
    return dict(poses=poses, map=_snapshot_map(m), n_sdf=trk.n_sdf_evals, n_rgb=trk.n_rgb_evals)


def rot_angle(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1.0) / 2.0
    return float(np.arccos(np.clip(c, -1.0, 1.0)))


def compare(a, b):
    """Deltas between two runs: poses (max |dt| in metres, max |dR| entry, max rotation angle), map ids / counts / latents."""
    dt = [float(np.abs(pa[1] - pb[1]).max()) for pa, pb in zip(a["poses"], b["poses"])]
    dR = [float(np.abs(pa[0] - pb[0]).max()) for pa, pb in zip(a["poses"], b["poses"])]
    ang = [rot_angle(pa[0], pb[0]) for pa, pb in zip(a["poses"], b["poses"])]
    out = dict(pose_t_max=max(dt), pose_R_max=max(dR), pose_angle_max=max(ang), pose_t_per_frame=dt)
    ma, mb = a["map"], b["map"]
    ids_equal = ma["n_occupied"] == mb["n_occupied"] and np.array_equal(ma["pos"], mb["pos"])
    out["map_ids_equal"] = bool(ids_equal)
    out["n_occupied"] = (ma["n_occupied"], mb["n_occupied"])
    common, ia, ib = np.intersect1d(ma["pos"], mb["pos"], return_indices=True)
    out["map_common_voxels"] = int(common.size)
    if common.size:
        la, lb = ma["latent"][ia], mb["latent"][ib]
        out["latent_max_abs"] = float(np.abs(la - lb).max())
        out["latent_rel"] = float(np.abs(la - lb).max() / max(np.abs(la).max(), 1e-12))
        out["count_equal_frac"] = float((ma["count"][ia] == mb["count"][ib]).mean())
    if "n_points" in a and "n_points" in b:
        out["n_points_max_diff"] = int(max(abs(x - y) for x, y in zip(a["n_points"], b["n_points"])))
    if a.get("clouds") and b.get("clouds"):                       # preprocessed clouds, frame by frame (same order in both paths)
        pd, nd, rows = 0.0, 0.0, 0.0
        for (pa, na), (pb, nb) in zip(a["clouds"], b["clouds"]):
            if pa.shape != pb.shape:
                continue
            dp = (pa - pb).abs().max(1).values; dn = (na - nb).abs().max(1).values
            pd = max(pd, float(dp.max())); nd = max(nd, float(dn.max()))
            rows = max(rows, float(((dp > 1e-6) | (dn > 1e-4)).float().mean()))
        out["cloud_point_max_abs"] = pd; out["cloud_normal_max_abs"] = nd; out["cloud_rows_differing_frac_max"] = rows
    return out


def decoder_deltas_on_reference_map(ref_run, dev, engines=(1, 0)):
    """Reference map -> our map (cold_vars file), then get_sdf / compute_sdf_Hg of both engines vs the reference's."""
    import tempfile
    from pathlib import Path
    dfb = dfb_pkg()
    lib = dfb._lib.load()
    rmap, rtrk = ref_run["map_obj"], ref_run["tracker_obj"]
    ref = ref_gpu.install("reference")
    m, trk = make_ours(dev)
    with tempfile.TemporaryDirectory() as td:
        p = Path(td) / "map.pt"
        rmap.save(p)
        m.load(p)
    pc = ref_run["clouds"][1][0]
    last_R, last_t = ref_run["poses"][0]
    xi = np.array([0.004, -0.003, 0.005, 0.002, -0.0015, 0.001])
    r_last = ref.motion.Isometry(q=ref.Quaternion(matrix=last_R), t=last_t)
    r_delta = ref.motion.Isometry.from_twist(xi)
    o_last = dfb.Isometry.from_matrix(last_R, last_t)
    o_delta = dfb.Isometry.from_twist(xi)
    Hr, gr, er = rtrk.compute_sdf_Hg(0, r_last, r_delta, pc)
    world = (r_last.dot(r_delta)) @ pc
    with torch.no_grad():
        sr, dr, vr = rmap.get_sdf(world)
    out = {}
    for eng in engines:
        lib.dfb_set_decoder_engine(eng)
        Ho, go, eo = trk.compute_sdf_Hg(0, o_last, o_delta, pc.contiguous())
        so, do, vo = m.get_sdf(world.contiguous())
        vr_ = vr.cpu().numpy().astype(bool); vo_ = vo.cpu().numpy().astype(bool)
        both = vr_ & vo_
        sdf_ref = np.zeros(len(vr_), np.float32); sdf_ref[vr_] = sr.detach().cpu().numpy().reshape(-1)
        std_ref = np.zeros(len(vr_), np.float32); std_ref[vr_] = dr.detach().cpu().numpy().reshape(-1)
        so_ = so.detach().cpu().numpy().reshape(-1); do_ = do.detach().cpu().numpy().reshape(-1)
        if so_.shape[0] != len(vr_):                       # compact outputs
            t = np.zeros(len(vr_), np.float32); t[vo_] = so_; so_ = t
            t = np.zeros(len(vr_), np.float32); t[vo_] = do_; do_ = t
        out[f"engine{eng}"] = dict(
            H_rel=float(np.abs(Ho - Hr).max() / np.abs(Hr).max()), g_rel=float(np.abs(go - gr).max() / np.abs(gr).max()),
            E_rel=float(abs(eo - er) / abs(er)), valid_equal=bool(np.array_equal(vr_, vo_)),
            sdf_max_abs_network_units=float(np.abs(so_[both] - sdf_ref[both]).max()),
            sdf_max_abs_m=float(np.abs(so_[both] - sdf_ref[both]).max() * 0.1),
            std_max_abs=float(np.abs(do_[both] - std_ref[both]).max()), n_queries=int(len(vr_)), n_valid=int(both.sum()))
    lib.dfb_set_decoder_engine(1)
    return out
