"""The CPU oracle (oracle/map_oracle.py, oracle/tracker_oracle.py) against the golden vectors the REFERENCE's own code produced
(oracle/make_golden.py ran /root/reference/system/map.py and tracker.py on these inputs): this is what pins the checker the
GPU parity tests lean on."""
import numpy as np
import pytest
import torch

from oracle import tracker_oracle as TO
from util import GOLD, make_oracle_map, pkg


@pytest.fixture(scope="module")
def G():
    return dict(np.load(GOLD / "map_golden.npz"))


@pytest.fixture(scope="module")
def W():
    return pkg().weights.load_npz(GOLD / "weights.npz")


@pytest.fixture(scope="module")
def omap(W, G):
    om = make_oracle_map(W)
    Pw, Nw = torch.from_numpy(G["Pw"]), torch.from_numpy(G["Nw"])
    m1 = om.integrate_keyframe(Pw, Nw)
    s1 = dict(mask=m1.numpy().copy(), n=om.n_occupied, pos=om.latent_vecs_pos[:om.n_occupied].numpy().copy(),
              cnt=om.voxel_obs_count[:om.n_occupied].numpy().copy(), lat=om.latent_vecs[:om.n_occupied].numpy().copy(),
              cap=om.latent_vecs.size(0))
    m2 = om.integrate_keyframe(Pw + torch.from_numpy(G["k2_shift"]), Nw)
    return om, s1, m2


def test_oracle_integrate_matches_reference_golden(omap, G):
    om, s1, m2 = omap
    assert np.array_equal(s1["mask"], G["k1_mask"]) and s1["n"] == int(G["k1_n_occupied"]) and s1["cap"] == int(G["k1_capacity"])
    assert np.array_equal(s1["pos"], G["k1_pos"]) and np.array_equal(s1["cnt"], G["k1_count"])            # ids, slot order, counts: bit-exact
    assert np.abs(s1["lat"] - G["k1_latent"]).max() <= 1e-5 * np.abs(G["k1_latent"]).max()
    n = om.n_occupied
    assert np.array_equal(m2.numpy(), G["k2_mask"]) and n == int(G["k2_n_occupied"])
    assert np.array_equal(om.latent_vecs_pos[:n].numpy(), G["k2_pos"]) and np.array_equal(om.voxel_obs_count[:n].numpy(), G["k2_count"])
    assert np.abs(om.latent_vecs[:n].numpy() - G["k2_latent"]).max() <= 1e-5 * np.abs(G["k2_latent"]).max()


def test_oracle_get_sdf_and_hg_match_reference_golden(omap, G):
    om = omap[0]
    sdf, std, valid = om.get_sdf(torch.from_numpy(G["q_world"]))
    assert np.array_equal(valid.numpy(), G["q_valid"])
    assert np.abs(sdf.detach().numpy() - G["q_sdf"]).max() < 1e-5 and np.abs(std.detach().numpy() - G["q_std"]).max() < 1e-5
    last = TO.Pose(TO.Quaternion(matrix=G["hg_last_R"]), G["hg_last_t"]); delta = TO.Pose(TO.Quaternion(matrix=G["hg_delta_R"]), G["hg_delta_t"])
    H, g, e, _ = TO.compute_sdf_Hg(om, last, delta, torch.from_numpy(G["Pc"]))
    # fp32 sums over 12 148 rows in a different order than the reference's einsum: 3e-5 measured
    assert np.abs(H - G["hg_H"]).max() <= 1e-4 * np.abs(G["hg_H"]).max() and np.abs(g - G["hg_g"]).max() <= 1e-4 * np.abs(G["hg_g"]).max()
    assert abs(e - float(G["hg_e"])) <= 1e-5 * abs(float(G["hg_e"]))
    _, _, e_ng, _ = TO.compute_sdf_Hg(om, last, delta, torch.from_numpy(G["Pc"]), no_grad=True)
    assert abs(e_ng - float(G["hg_e_nograd"])) <= 1e-5 * abs(float(G["hg_e_nograd"]))


def _golden_frame(T, i):
    depth = torch.from_numpy(T[f"f{i}_depth_u16"].astype(np.float32)) / 5000.0
    rgb = torch.from_numpy(T[f"f{i}_rgb_u8"]).float() / 255.
    depth[torch.logical_or(depth < 0.5, depth > 5.0)] = float("nan")
    return rgb, depth


def test_oracle_tracking_matches_reference_golden(W):
    """Preprocessing (unproject, radius filter, PCA normals, box filter) bit-identical to what the reference's tracker produced,
    and the full Gauss-Newton solve (SDF + photometric terms, three groups) lands on the reference's pose."""
    T = dict(np.load(GOLD / "track_golden.npz"))
    K4 = T["calib"][:4].tolist()
    for i in (0, 1):
        P, N = TO.preprocess(_golden_frame(T, i)[1].numpy(), K4)
        assert np.array_equal(P, T[f"f{i}_pc"]) and np.array_equal(N, T[f"f{i}_normal"])
    om = make_oracle_map(W)
    pose0 = TO.Pose(TO.Quaternion(matrix=T["f0_pose_R"]), T["f0_pose_t"])
    om.integrate_keyframe(pose0.apply(torch.from_numpy(T["f0_pc"])), torch.from_numpy(T["f0_normal"]) @ torch.from_numpy(pose0.R).float().T)
    assert om.n_occupied == int(T["n_occupied_after_f0"])
    (rgb0, d0), (rgb1, d1) = _golden_frame(T, 0), _golden_frame(T, 1)
    I0, D0, _ = TO.image_pyramid(rgb0.mean(-1).numpy(), d0.numpy())
    I1, D1, G1 = TO.image_pyramid(rgb1.mean(-1).numpy(), d1.numpy())
    n = T["iter_config_n"]
    cfg = [{"n": int(n[0]), "type": [["rgb", 2]]}, {"n": int(n[1]), "type": [["sdf"], ["rgb", 1]]}, {"n": int(n[2]), "type": [["sdf"], ["rgb", 0]]}]
    pose, n_eval, _ = TO.gauss_newton(om, pose0, pose0, torch.from_numpy(T["f1_pc"]), cfg, rgb=dict(state=(I0, D0), cur=(I1, D1, G1), K4=K4))
    assert n_eval > 0
    assert np.abs(pose.t - T["f1_pose_t"]).max() < 1e-5 and np.abs(pose.R - T["f1_pose_R"]).max() < 1e-5      # measured 1e-7


def test_oracle_meshing_matches_reference_golden(omap, G):
    """Voxel selection, cube decoding (low pass, trilinear upsampling, refined band) and sparse marching cubes with std-weighted
    blending against what the reference's extract_mesh produced (its marching cubes ran through the oracle's restatement of
    mc_interp_kernel.cu inside the reference's own do_meshing)."""
    from oracle import ops
    from util import match_rows
    om = omap[0]
    focused, occ, mapping = om.meshing_batch(torch.from_numpy(G["k2_updated"]))
    assert np.array_equal(focused.numpy(), G["mesh_valid_blocks"]) and occ.numel() == int(G["mesh_B"])
    assert np.array_equal(mapping.numpy(), G["mesh_mapping"])
    cs, cd = om.decode_cubes(occ, 4)
    head = G["mesh_cube_sdf_head"].shape[0]
    assert np.abs(cs[:head].numpy() - G["mesh_cube_sdf_head"]).max() < 1e-5 and np.abs(cd[:head].numpy() - G["mesh_cube_std_head"]).max() < 1e-5
    assert np.abs(cs.numpy().astype(np.float64).sum(axis=(1, 2, 3)) - G["mesh_cube_sdf_sum"]).max() < 2e-3
    tri, fid, tstd = ops.marching_cubes_sparse_interp(om.indexer.view(*om.n_xyz).numpy(), focused[:48].numpy(), mapping.numpy(), cs.numpy(),
                                                      cd.numpy(), int(4e6), om.n_xyz, 0.15)
    tri_m = tri * np.float32(0.1) + om.bound_min.numpy()
    assert tri_m.shape == G["mesh_tri_first48"].shape
    match_rows(tri_m, G["mesh_tri_first48"], 1e-5)
    assert set(np.unique(fid).tolist()) <= set(G["mesh_valid_blocks"][:48].tolist())
