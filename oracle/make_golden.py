"""Generate tests/golden/*.npz by running the REFERENCE's own Python (imported from /root/reference through
oracle/ref_shims.py) on seeded synthetic inputs.  TEST INFRASTRUCTURE ONLY; run in the build container:

    python -m oracle.make_golden

Fixtures (all small):
  weights.npz        ckpt/default/{model_300,encoder_300}.pth.tar tensors, verbatim, keys 'dec.*' / 'enc.*'
  map_golden.npz     two integrate_keyframe calls, get_sdf, compute_sdf_Hg, meshing cubes of the reference map
  track_golden.npz   3 low-resolution synthetic frames through SDFTracker.track_camera + integrate (poses, clouds)
The CUDA-extension ops inside those runs are the oracle restatements (oracle/ops.py) because the reference's
kernels cannot execute without a GPU; map/network/tracker Python is the reference's, unmodified.
"""
import importlib
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_shims, tracker_oracle  # noqa: E402

GOLD = ROOT / "tests" / "golden"
synth = importlib.import_module("nerf-fusion_b200.synth")


def main():
    torch.manual_seed(0); np.random.seed(0)
    torch.set_num_threads(1)          # deterministic reductions in the reference run
    ref = ref_shims.install()
    model, _ = ref_shims.load_reference_model()
    GOLD.mkdir(parents=True, exist_ok=True)

    W = {}
    for k, v in model.decoder.state_dict().items():
        W["dec." + k] = v.numpy().copy()
    for k, v in model.encoder.state_dict().items():
        W["enc." + k] = v.numpy().copy()
    np.savez_compressed(GOLD / "weights.npz", **W)

    # ------------------------------------------------------------------ map golden
    rmap, cfg = ref_shims.make_reference_map(model)
    seq = synth.SyntheticSequence(n_frames=2)
    depth, _ = seq.frame(0)
    depth = depth.clone(); depth[(depth < 0.5) | (depth > 5.0)] = float("nan")
    P, N = tracker_oracle.preprocess(depth.numpy(), synth.ICL_CALIB)
    sel = (P[:, 0] > -0.9) & (P[:, 0] < 0.35)                # crop: keeps the fixture small but dense
    P, N = torch.from_numpy(P[sel]), torch.from_numpy(N[sel])
    R0 = torch.from_numpy(synth.quat_to_R(synth.FIRST_TQ[3:])).float(); t0 = torch.tensor(synth.FIRST_TQ[:3]).float()
    Pw = (P @ R0.T + t0).contiguous(); Nw = (N @ R0.T).contiguous()
    out = {"Pc": P.numpy(), "Pw": Pw.numpy(), "Nw": Nw.numpy()}

    def snap(tag):
        n = rmap.n_occupied
        out[f"{tag}_n_occupied"] = np.int64(n)
        out[f"{tag}_pos"] = rmap.latent_vecs_pos[:n].numpy().copy()
        out[f"{tag}_count"] = rmap.voxel_obs_count[:n].numpy().copy()
        out[f"{tag}_latent"] = rmap.latent_vecs[:n].numpy().copy()
        out[f"{tag}_capacity"] = np.int64(rmap.latent_vecs.size(0))
        out[f"{tag}_updated"] = rmap.mesh_cache.updated_vec_id.numpy().copy()

    out["k1_mask"] = rmap.integrate_keyframe(Pw, Nw).numpy().copy(); snap("k1")
    shift = torch.tensor([0.013, -0.007, 0.021])
    out["k2_shift"] = shift.numpy()
    out["k2_mask"] = rmap.integrate_keyframe(Pw + shift, Nw).numpy().copy(); snap("k2")

    # get_sdf on displaced camera-space points
    Iso = ref.motion.Isometry; Q = sys.modules["pyquaternion"].Quaternion
    last = Iso(q=Q(array=synth.FIRST_TQ[3:]), t=np.array(synth.FIRST_TQ[:3]))
    xi = np.array([0.004, -0.003, 0.005, 0.002, -0.0015, 0.001])
    delta = Iso.from_twist(xi)
    out["q_xi"] = xi
    world = (last.dot(delta)) @ P
    sdf, std, valid = rmap.get_sdf(world)
    out["q_world"] = world.numpy().copy(); out["q_sdf"] = sdf.detach().numpy().copy()
    out["q_std"] = std.detach().numpy().copy(); out["q_valid"] = valid.numpy().copy()
    trk = ref.tracker.SDFTracker(rmap, ref.exp.dict_to_args(cfg["tracking"]))
    H, g, e = trk.compute_sdf_Hg(0, last, delta, P)
    _, _, e_ng = trk.compute_sdf_Hg(-1, last, delta, P, True)
    out["hg_H"] = H; out["hg_g"] = g; out["hg_e"] = np.float64(e); out["hg_e_nograd"] = np.float64(e_ng)
    out["hg_last_R"] = last.q.rotation_matrix; out["hg_last_t"] = last.t
    out["hg_delta_R"] = delta.q.rotation_matrix; out["hg_delta_t"] = delta.t

    # meshing: capture what the reference hands to marching_cubes_interp
    cap = {}
    ext = sys.modules["system.ext"]
    orig_mc = ext.marching_cubes_interp

    def spy(indexer, valid_blocks, mapping, cube_sdf, cube_std, max_n, n_xyz, max_std):
        cap.update(valid_blocks=valid_blocks.numpy().copy(), mapping=mapping.numpy().copy(),
                   cube_sdf=cube_sdf.numpy().copy(), cube_std=cube_std.numpy().copy())
        sub = valid_blocks[:48]
        return orig_mc(indexer, sub, mapping, cube_sdf, cube_std, max_n, n_xyz, max_std)

    ext.marching_cubes_interp = spy
    ref.map.DenseIndexedMap._make_mesh_from_cache = lambda self: None
    rmap.extract_mesh(4, int(4e6), max_std=0.15, extract_async=False, interpolate=True)
    ext.marching_cubes_interp = orig_mc
    keep = 64                                                   # first rows only: keeps the file small
    out["mesh_valid_blocks"] = cap["valid_blocks"]; out["mesh_mapping"] = cap["mapping"]
    out["mesh_cube_sdf_head"] = cap["cube_sdf"][:keep]; out["mesh_cube_std_head"] = cap["cube_std"][:keep]
    out["mesh_B"] = np.int64(cap["cube_sdf"].shape[0])
    out["mesh_cube_sdf_sum"] = cap["cube_sdf"].astype(np.float64).sum(axis=(1, 2, 3))
    out["mesh_cube_std_sum"] = cap["cube_std"].astype(np.float64).sum(axis=(1, 2, 3))
    out["mesh_tri_first48"] = rmap.mesh_cache.vertices.copy()           # metres, blocks valid_blocks[:48]
    out["mesh_tri_id_first48"] = rmap.mesh_cache.vertices_flatten_id.copy()
    out["mesh_tri_std_first48"] = rmap.mesh_cache.vertices_std.copy()
    np.savez_compressed(GOLD / "map_golden.npz", **out)
    print("map golden:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})

    # ------------------------------------------------------------------ tracker golden (low resolution)
    model2, _ = ref_shims.load_reference_model()
    rmap2, cfg2 = ref_shims.make_reference_map(model2)
    targs = ref.exp.dict_to_args(cfg2["tracking"])
    targs.iter_config = [{"n": 3, "type": [["rgb", 2]]}, {"n": 3, "type": [["sdf"], ["rgb", 1]]}, {"n": 8, "type": [["sdf"], ["rgb", 0]]}]
    trk2 = ref.tracker.SDFTracker(rmap2, targs)
    Hh, Ww = 240, 320
    calib = tuple(c * 0.5 for c in synth.ICL_CALIB)
    seq2 = synth.SyntheticSequence(n_frames=3, H=Hh, W=Ww, step_trans=0.006, step_rot_deg=0.15)
    seq2.calib = calib
    from dataset.production import FrameIntrinsic
    first = Iso(q=Q(array=synth.FIRST_TQ[3:]), t=np.array(synth.FIRST_TQ[:3]))
    tout = {"calib": np.array(calib), "iter_config_n": np.array([3, 3, 8])}
    for i in range(3):
        depth, rgb = seq2.frame(i)
        d16 = torch.round(depth * 5000.0).to(torch.int32).numpy().astype(np.uint16)
        r8 = torch.round(rgb * 255.0).to(torch.uint8).numpy()
        tout[f"f{i}_depth_u16"] = d16; tout[f"f{i}_rgb_u8"] = r8
        depth = torch.from_numpy(d16.astype(np.float32)) / 5000.0
        rgb = torch.from_numpy(r8).float() / 255.
        depth[torch.logical_or(depth < 0.5, depth > 5.0)] = np.nan
        pose = trk2.track_camera(rgb, depth, FrameIntrinsic(*calib, 5000.0), first if i == 0 else None)
        pc, nrm = trk2.last_processed_pc
        tout[f"f{i}_pc"] = pc.numpy().copy(); tout[f"f{i}_normal"] = nrm.numpy().copy()
        tout[f"f{i}_pose_R"] = pose.q.rotation_matrix.copy(); tout[f"f{i}_pose_t"] = pose.t.copy()
        gt = seq2.poses[i]
        tout[f"f{i}_gt_R"] = gt[0]; tout[f"f{i}_gt_t"] = gt[1]
        if i == 0:
            rmap2.integrate_keyframe(pose @ pc, pose.rotation @ nrm)
            tout["n_occupied_after_f0"] = np.int64(rmap2.n_occupied)
        print("frame", i, "pose t", pose.t, "gt t", gt[1], "N", pc.shape[0])
    np.savez_compressed(GOLD / "track_golden.npz", **tout)


if __name__ == "__main__":
    main()
