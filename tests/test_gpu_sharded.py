"""-m gpu: the sharded map's kernels and protocol (csrc/sharded.cu, nerf-fusion_b200/sharded.py) with ALL ranks of a world
emulated in one process on one GPU (`LocalFabric`: peer pointers are plain pointers, phases run rank after rank -- no
kernel waits for another, so this is exactly what each rank executes between two barriers).  Parity with the single-GPU
DenseIndexedMap on {linear voxel id -> (count, latent)}: ids, counts AND latents bit-exact (fixed-point encoder sums).
The multi-process path (CUDA IPC peer mappings + NCCL barrier) is exercised by tools/sharded_bench.py on 2..8 GPUs."""
import numpy as np
import pytest
import torch

from util import GOLD, make_map, ns, pkg, MAPPING

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _single_state(m):
    n = m.n_occupied
    o = torch.argsort(m.latent_vecs_pos[:n])
    return m.latent_vecs_pos[:n][o], m.voxel_obs_count[:n][o], m.latent_vecs[:n][o]


def _check(fab, one, tag):
    ids, cnt, lat = fab.gather_state()
    rid, rcnt, rlat = _single_state(one)
    assert torch.equal(ids, rid), f"{tag}: voxel ids differ ({ids.numel()} vs {rid.numel()})"
    assert torch.equal(cnt, rcnt), f"{tag}: counts differ"
    err = float((lat - rlat).abs().max() / rlat.abs().max())
    # the encoder sums are 64-bit fixed-point (integer adds commute) and a sample's encoding does not depend on the rank or tile
    # that computes it: the sharded map is BIT-IDENTICAL to the single-GPU map, for any split of the points
    assert torch.equal(lat, rlat), (tag, err)
    for m in fab.maps:                                               # zero invariants restored, no overflow
        st = m.read_stats()
        assert int(m.mem.grid_count.abs().sum()) == 0 and int(m.mem.acc_n.sum()) == 0 and int(m.mem.acc.abs().sum()) == 0
    return err


@pytest.mark.parametrize("world", [1, 2, 4])
def test_emulated_ranks_match_single_map_golden_keyframes(weights, world):
    d = pkg()
    G = dict(np.load(GOLD / "map_golden.npz"))
    Pw, Nw = torch.from_numpy(G["Pw"]).to(DEV), torch.from_numpy(G["Nw"]).to(DEV)
    one = make_map(weights)
    fab = d.sharded.LocalFabric.create(weights, ns(dict(MAPPING)), DEV, world, max_points_per_rank=Pw.shape[0], capacity=1 << 15)
    for k, shift in enumerate((torch.zeros(3), torch.from_numpy(G["k2_shift"]))):
        P = (Pw + shift.to(DEV)).contiguous()
        one.integrate_keyframe(P, Nw)
        fab.integrate_keyframe([(P[r::world], Nw[r::world]) for r in range(world)])       # interleaved split
        err = _check(fab, one, f"world {world} keyframe {k}")
    # against the reference's golden as well (ids / counts bit-exact)
    ids, cnt, lat = fab.gather_state()
    r = np.argsort(G["k2_pos"])
    assert np.array_equal(ids.cpu().numpy(), G["k2_pos"][r]) and np.array_equal(cnt.cpu().numpy(), G["k2_count"][r])
    assert sum(m.n_occupied for m in fab.maps) == int(G["k2_n_occupied"])
    if world > 1:
        assert sum(m.n_occupied > 100 for m in fab.maps) >= 2         # the scene spreads over several ranks' bricks
        assert sum(m.last_stats["samples_sent_remote"] for m in fab.maps) > 0
    print(f"world {world}: latent rel err {err:.2e}")


@pytest.mark.parametrize("seed,voxel,bmin,bmax,prune", [
    (1, 0.07, [-0.5, 0.0, -2.0], [1.3, 0.9, -0.6], 8),          # non-cubic grid, n_xyz not multiples of 8: partial bricks
    (2, 0.25, [0.0, 0.0, 0.0], [3.0, 2.0, 1.0], 0),            # pruning disabled
    (3, 0.05, [-0.4, -0.4, -0.4], [0.4, 0.4, 0.4], 30),
])
def test_emulated_ranks_random_scenes_uneven_splits_and_threshold_crossing(weights, seed, voxel, bmin, bmax, prune):
    """Three keyframes of a random surface + a clump at the grid border; encoder_count_th = 120 so voxels LEAVE the candidate
    set between keyframes (the removal deltas), uneven point splits including an empty share."""
    d = pkg()
    rng = np.random.RandomState(seed)
    over = dict(bound_min=bmin, bound_max=bmax, voxel_size=voxel, prune_min_vox_obs=prune, encoder_count_th=120.0)
    args = dict(MAPPING); args.update(over)
    one = make_map(weights, **over)
    world = 3
    fab = d.sharded.LocalFabric.create(weights, ns(args), DEV, world, max_points_per_rank=40000, capacity=1 << 15)
    lo, hi = np.asarray(bmin, np.float32), np.asarray(bmax, np.float32)
    for k in range(3):
        n = 30000
        uv = rng.rand(n, 2).astype(np.float32)
        p = np.stack([uv[:, 0], uv[:, 1], 0.35 + 0.25 * uv[:, 0] + 0.1 * np.sin(5 * uv[:, 1] + k)], 1).astype(np.float32)
        p = lo + p * (hi - lo) * np.float32(0.98) + np.float32(0.01) * (hi - lo)
        clump = (hi - np.float32(1e-4)) - rng.rand(2000, 3).astype(np.float32) * np.float32(1.5 * voxel)
        outside = hi + np.float32(0.5)                               # one point outside the grid: dropped by every path
        p = np.concatenate([p, clump, outside[None]]).astype(np.float32)
        nr = rng.randn(p.shape[0], 3).astype(np.float32); nr /= np.linalg.norm(nr, axis=1, keepdims=True)
        P, N = torch.from_numpy(p).to(DEV), torch.from_numpy(nr).to(DEV)
        one.integrate_keyframe(P[:-1].contiguous(), N[:-1].contiguous())
        cut = [0, 25000, 25000, p.shape[0]] if k == 1 else [0, 9000, 20000, p.shape[0]]     # rank 1 gets nothing in keyframe 1
        fab.integrate_keyframe([(P[cut[r]:cut[r + 1]], N[cut[r]:cut[r + 1]]) for r in range(world)])
        _check(fab, one, f"seed {seed} keyframe {k}")
    assert (one.voxel_obs_count[:one.n_occupied] >= 120.0).any()      # the candidate threshold was crossed
