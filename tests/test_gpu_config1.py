"""-m gpu: BASELINE config 1 as written -- fusion-lr-kt.yaml + ckpt/default, 20 synthetic 640x480 frames, full iter_config
(10 / 10 / 50), integrate every 20 frames -- through the REFERENCE'S OWN CUDA PATH (its unmodified map.py / tracker.py on its
own system/ext kernels, staged under oracle/_ref by oracle/stage_ref_py.py + oracle/build_ref_ext.py) and through this repo
with the DEFAULT engines (the ones bench.py measures), on the same device tensors.  Skipped when oracle/_ref is absent.

What "parity" can mean here is bounded by the reference itself: its kd-tree / scatter kernels use atomics and its
Gauss-Newton loop stops on `energy > last_energy` (tracker.py:269), so two runs of the unmodified reference differ from
each other by up to ~1e-4..1e-3 m on single frames (profiles/r02_config1_parity*.json, `reference_run_to_run`).  The
assertions therefore are: bit-exact where the domain is integer (point counts, voxel ids), the north-star tolerances
where the comparison is deterministic (latents, SDF, H, g on the reference's map; poses on the low-resolution golden
sequence in tests/test_gpu_tracker.py), and median / worst-frame bounds at the reference's own noise floor for the
full-resolution end-to-end poses.
"""
import sys
import tempfile
from pathlib import Path

import numpy as np
import pytest
import torch

from util import ROOT, pkg

sys.path.insert(0, str(ROOT))
from oracle import config1 as C1, ref_gpu  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
N_FRAMES = 20


@pytest.fixture(scope="module")
def runs():
    if not ref_gpu.available():
        pytest.skip("oracle/_ref (staged reference python + built reference extensions) is absent")
    lib = pkg()._lib.load()
    lib.dfb_set_decoder_engine(1); lib.dfb_set_encoder_engine(1)                 # defaults: what bench.py measures
    frames, calib, seq = C1.make_frames(N_FRAMES, DEV)
    C1.run_reference(frames[:2], calib, DEV, keep_clouds=False)                   # warm-up
    ref = C1.run_reference(frames, calib, DEV)
    C1.run_ours(frames[:4], calib, DEV)
    ours = C1.run_ours(frames, calib, DEV)
    onref = C1.run_ours_on_reference_points(frames, calib, DEV, ref)
    return dict(frames=frames, calib=calib, seq=seq, ref=ref, ours=ours, onref=onref)


def _summary(tag, cmp_):
    dt = np.array(cmp_["pose_t_per_frame"][1:])
    print(f"{tag}: pose |dt| median {np.median(dt):.2e} max {dt.max():.2e} m, angle max {cmp_['pose_angle_max']:.2e} rad, "
          f"ids equal {cmp_['map_ids_equal']}, latent rel {cmp_.get('latent_rel', float('nan')):.2e}, "
          f"counts equal {cmp_.get('count_equal_frac', float('nan')):.4f}")
    return dt


def test_config1_preprocessing_and_map(runs):
    ref, ours = runs["ref"], runs["ours"]
    assert ours["n_points"] == ref["n_points"]                                    # preprocessing: same row count on every frame
    cmp_ = C1.compare(ours, ref)
    _summary("ours (default engines) vs reference CUDA", cmp_)
    a, b = ours["map"], ref["map"]
    # (one bench run on a slow box saw the reference leave 0.5 % of its counts different and a few voxels on the other side of
    # the pruning threshold -- its scatter kernels race: bound the id difference at 0.2 % of the map instead of 2 voxels)
    assert len(np.setxor1d(a["pos"], b["pos"])) <= max(2, len(b["pos"]) // 500)   # voxel ids (bit-exact unless the reference's own
    assert cmp_["count_equal_frac"] >= 0.99                                       #  atomics moved a point across a threshold)
    common, ia, ib = np.intersect1d(a["pos"], b["pos"], return_indices=True)
    same = a["count"][ia] == b["count"][ib]
    la, lb = a["latent"][ia][same], b["latent"][ib][same]
    assert np.abs(la - lb).max() <= 1e-3 * np.abs(lb).max()                       # north star: latents 1e-3 relative


def test_config1_poses(runs):
    ref, ours, onref, seq = runs["ref"], runs["ours"], runs["onref"], runs["seq"]
    # Bounds = the reference's own run-to-run floor (two runs of the UNMODIFIED reference on identical inputs: median 1.3e-5,
    # worst frame 1.5e-4 m, profiles/r02_config1_parity.json), with headroom for its tail: over the GPU runs of round 2 the medians
    # of these comparisons ranged from 8e-6 to 5e-5 (the operator-level drop-in below, whose arithmetic did not change between
    # those runs, included), because WHICH iteration ends a group (tracker.py:269) flips on the last bits of the energy.
    dt = _summary("end to end", C1.compare(ours, ref))
    assert np.median(dt) < 1e-4 and dt.max() < 2e-3
    dt2 = _summary("fed the reference's points", C1.compare(onref, ref))
    assert np.median(dt2) < 5e-5 and dt2.max() < 2e-3
    gt = max(float(np.abs(p[1] - seq.poses[i][1]).max()) for i, p in enumerate(ours["poses"]))
    gt_ref = max(float(np.abs(p[1] - seq.poses[i][1]).max()) for i, p in enumerate(ref["poses"]))
    print(f"max |t - ground truth|: ours {gt:.2e} m, reference {gt_ref:.2e} m")
    assert gt < gt_ref + 1e-3


@pytest.mark.parametrize("dec_engine", [1, 0])
def test_config1_decoder_on_the_reference_map(runs, dec_engine):
    """The reference's map (its cold_vars file) loaded into this repo's map: get_sdf and compute_sdf_Hg of the reference
    (torch autograd on CUDA) against the fused kernel.  North star: SDF 1e-4 m, H / g 1e-4 relative."""
    d = pkg()
    lib = d._lib.load()
    ref_run = runs["ref"]
    ref = ref_gpu.install("reference")
    rmap, rtrk = ref_run["map_obj"], ref_run["tracker_obj"]
    m, trk = C1.make_ours(DEV)
    with tempfile.TemporaryDirectory() as td:
        rmap.save(Path(td) / "map.pt")
        m.load(Path(td) / "map.pt")
    pc = ref_run["clouds"][1][0].contiguous()
    last_R, last_t = ref_run["poses"][0]
    xi = np.array([0.004, -0.003, 0.005, 0.002, -0.0015, 0.001])
    r_last, r_delta = ref.motion.Isometry(q=ref.Quaternion(matrix=last_R), t=last_t), ref.motion.Isometry.from_twist(xi)
    o_last, o_delta = d.Isometry.from_matrix(last_R, last_t), d.Isometry.from_twist(xi)
    Hr, gr, er = rtrk.compute_sdf_Hg(0, r_last, r_delta, pc)
    world = ((r_last.dot(r_delta)) @ pc).contiguous()
    with torch.no_grad():
        sr, dr, vr = rmap.get_sdf(world)
    lib.dfb_set_decoder_engine(dec_engine)
    try:
        Ho, go, eo = trk.compute_sdf_Hg(0, o_last, o_delta, pc)
        so, do, vo = m.get_sdf(world)
    finally:
        lib.dfb_set_decoder_engine(1)
    assert torch.equal(vr.bool().cpu(), vo.bool().cpu())
    e_sdf = float((so - sr.reshape(-1)).abs().max()); e_std = float((do - dr.reshape(-1)).abs().max())
    print(f"engine {dec_engine}: H rel {np.abs(Ho - Hr).max() / np.abs(Hr).max():.2e} g rel {np.abs(go - gr).max() / np.abs(gr).max():.2e} "
          f"E rel {abs(eo - er) / abs(er):.2e} sdf {e_sdf:.2e} std {e_std:.2e} (network units)")
    assert np.abs(Ho - Hr).max() <= 1e-4 * np.abs(Hr).max()
    assert np.abs(go - gr).max() <= 1e-4 * np.abs(gr).max()
    assert abs(eo - er) <= 1e-4 * abs(er)
    assert e_sdf < 1e-4 and e_std < 1e-4                                          # 1e-5 m


def test_operator_level_dropin(runs):
    """INTEGRATION.md §2 as evidence: the reference's UNMODIFIED DenseIndexedMap / SDFTracker (integrate_keyframe, get_sdf,
    track_camera, compute_sdf_Hg, gauss_newton) executing on this repo's C-ABI ops installed as `system.ext` and
    `torch_scatter` -- against the same Python on the reference's own kernels."""
    ref = runs["ref"]
    drop = C1.run_reference(runs["frames"], runs["calib"], DEV, backend="dfb", keep_clouds=False)
    ref_gpu.use_backend("reference")
    assert drop["n_points"] == ref["n_points"]
    cmp_ = C1.compare(drop, ref)
    dt = _summary("reference python on dfb ops vs on its own ops", cmp_)
    assert len(np.setxor1d(drop["map"]["pos"], ref["map"]["pos"])) <= max(2, len(ref["map"]["pos"]) // 500) and cmp_["count_equal_frac"] >= 0.99
    assert np.median(dt) < 1e-4 and dt.max() < 2e-3                  # (bounds: see test_config1_poses)
    # the class-level path and the operator-level path agree with each other as well
    dt2 = _summary("class-level path vs operator-level drop-in", C1.compare(runs["ours"], drop))
    assert np.median(dt2) < 1e-4 and dt2.max() < 2e-3


def test_extract_mesh_twice_matches_reference(runs):
    """X3: DenseIndexedMap.extract_mesh() itself, called after the first and after a second keyframe so that the mesh-cache
    splice (map.py:703-715) runs, against the reference's extract_mesh on the same keyframes (its own CUDA marching cubes,
    torch decoder).  Triangles compared as sets."""
    d = pkg()
    ref = ref_gpu.install("reference")
    ref.map.DenseIndexedMap._make_mesh_from_cache = lambda self: None             # open3d is a stub here; the cache holds the mesh
    rmap, _, _ = ref_gpu.make_reference_system(DEV)
    m, _ = C1.make_ours(DEV)
    pose = d.Isometry.from_matrix(*runs["ref"]["poses"][0])
    pc, nrm = runs["ref"]["clouds"][0]
    Pw, Nw = (pose @ pc).contiguous(), (pose.rotation @ nrm).contiguous()
    shift = torch.tensor([0.013, -0.007, 0.021], device=DEV)
    for k, P in enumerate((Pw, Pw + shift)):
        rmap.integrate_keyframe(P, Nw, async_optimize=False, do_optimize=False)
        m.integrate_keyframe(P, Nw, do_optimize=False)
        rmap.extract_mesh(4, int(4e6), max_std=0.15, extract_async=False, interpolate=True)
        mesh = m.extract_mesh(4, int(4e6), max_std=0.15, extract_async=False, interpolate=True)
        rv, ov = rmap.mesh_cache.vertices, m.mesh_cache.vertices
        assert ov.shape[0] == mesh.triangles.shape[0] and ov.shape[0] > 1000
        # band membership / sign flips at fp noise level move a handful of triangles: counts agree to 1 %, and the triangles of
        # the voxels both meshes triangulated identically agree to 1e-4 m
        assert abs(ov.shape[0] - rv.shape[0]) <= max(5, 0.01 * rv.shape[0]), (k, ov.shape, rv.shape)
        rid, oid = rmap.mesh_cache.vertices_flatten_id, m.mesh_cache.vertices_flatten_id
        assert np.array_equal(np.unique(rid), np.unique(oid)) or len(np.setxor1d(np.unique(rid), np.unique(oid))) <= 3
        cr, co = np.unique(rid, return_counts=True), np.unique(oid, return_counts=True)
        cnt_r = dict(zip(*cr)); cnt_o = dict(zip(*co))
        good = np.array([v for v in cnt_r if cnt_o.get(v) == cnt_r[v]])
        assert len(good) >= 0.97 * len(cnt_r)
        sel_r, sel_o = np.isin(rid, good), np.isin(oid, good)
        # order-free comparison: every triangle centroid has a partner in the other mesh.  A sample whose interpolated |sdf|
        # sits within fp noise of the 0.05 refine band (map.py:662) is re-decoded exactly in one run and stays interpolated in
        # the other, which moves a few vertices by a fraction of a 12.5 mm cell: bound the bulk tightly and the tail loosely.
        from scipy.spatial import cKDTree
        ca, cb = rv[sel_r].mean(1), ov[sel_o].mean(1)
        dist = np.maximum(cKDTree(cb).query(ca, k=1)[0], cKDTree(ca).query(cb, k=1)[0])
        print(f"keyframe {k}: {ov.shape[0]} triangles (reference {rv.shape[0]}), {len(good)}/{len(cnt_r)} voxels with equal "
              f"triangle counts, centroid distance median {np.median(dist):.2e} p99.9 {np.quantile(dist, 0.999):.2e} max {dist.max():.2e} m")
        assert np.median(dist) < 1e-5 and np.quantile(dist, 0.999) < 5e-4 and dist.max() < 5e-3


def test_latent_optimiser_matches_reference(runs):
    """integrate_keyframe(do_optimize=True) -- voxels that crossed encoder_count_th are refined by Adam on perturbed
    surface samples (map.py:456-517, OptimizeProcess.do_optimize :81-113) -- against the reference's own do_optimize (torch
    autograd + torch.optim.Adam on CUDA) with the same torch seed, so both draw the same random offsets along the normals."""
    d = pkg()
    over = dict(encoder_count_th=120.0, optim_n_iters=5, code_regularization=True, code_reg_lambda=1e-4)
    ref_gpu.install("reference")
    rmap, _, _ = ref_gpu.make_reference_system(DEV, mapping_over=over)
    m, _ = C1.make_ours(DEV, mapping_over=over)
    pose = d.Isometry.from_matrix(*runs["ref"]["poses"][0])
    pc, nrm = runs["ref"]["clouds"][0]
    Pw, Nw = (pose @ pc).contiguous(), (pose.rotation @ nrm).contiguous()
    for k in range(3):                                              # counts rise past the threshold
        rmap.integrate_keyframe(Pw, Nw, async_optimize=False, do_optimize=False)
        m.integrate_keyframe(Pw, Nw, do_optimize=False)
    assert int((m.voxel_obs_count[:m.n_occupied] >= 120.0).sum()) > 100
    before = m.latent_vecs[:m.n_occupied].clone()
    torch.manual_seed(123)
    rmap.integrate_keyframe(Pw, Nw, async_optimize=False, do_optimize=True)
    torch.manual_seed(123)
    m.integrate_keyframe(Pw, Nw, do_optimize=True)
    n = m.n_occupied
    assert n == rmap.n_occupied and torch.equal(m.latent_vecs_pos[:n], rmap.latent_vecs_pos[:n])
    opt_r, opt_o = rmap.voxel_optimized[:n], m.voxel_optimized[:n]
    assert torch.equal(opt_r, opt_o) and int(opt_o.sum()) > 100
    moved = (m.latent_vecs[:n][opt_o] - before[opt_o]).abs().max().item()
    la, lb = m.latent_vecs[:n][opt_o], rmap.latent_vecs[:n][opt_r]
    diff = (la - lb).abs().flatten()
    q50, q99, mx = (float(torch.quantile(diff, q)) for q in (0.5, 0.99, 1.0))
    print(f"latent optimiser: {int(opt_o.sum())} voxels refined, moved by up to {moved:.3e}; |ours - reference| median {q50:.2e} "
          f"p99 {q99:.2e} max {mx:.2e}")
    assert moved > 1e-2                                             # Adam did move them (5 steps of lr 1e-2)
    # The objective's gradient is DISCONTINUOUS (clamp(sdf, +-0.2), ReLU masks) and Adam's first steps are +-lr * sign-like,
    # so a single sample on the other side of a clamp flips the direction of a small gradient component: the reference's
    # own do_optimize and a torch restatement of it (same ops, oracle networks) already differ by 1.6e-4 after one step and
    # 4e-3 .. 1.3e-2 after five on these inputs (measured).  The kernels' gradient itself is pinned to 1e-5 and one Adam step to
    # 1e-6 against torch autograd in tests/test_gpu_ops.py::test_latent_adam_step_vs_torch.
    assert q50 < 2e-4 and q99 < 1e-2 and mx < 6e-2
    assert torch.equal(m.voxel_obs_count[:n], rmap.voxel_obs_count[:n])
