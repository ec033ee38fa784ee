"""World-size-2 gloo test (CPU) of the sharded map's routing logic: two ranks, each fed half of the points, must
reproduce the single-map oracle as {voxel id -> (count, latent)}.  The encoder is the oracle's (CPU test only)."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
    import importlib
    import torch.distributed as dist
    from oracle import nets
    from util import GOLD, MAPPING, ns
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    sharded = importlib.import_module("nerf-fusion_b200.sharded")
    W = nets.load_weights(GOLD / "weights.npz")
    G = dict(np.load(GOLD / "map_golden.npz"))
    m = sharded.ShardedMap(W, ns(dict(MAPPING)), "cpu", encoder_fn=lambda x: nets.encoder_forward(W, x))
    Pw, Nw = torch.from_numpy(G["Pw"]), torch.from_numpy(G["Nw"])
    for shift in (torch.zeros(3), torch.from_numpy(G["k2_shift"])):
        st = m.integrate_keyframe((Pw + shift)[rank::world], Nw[rank::world])      # interleaved split of the keyframe
    ids, cnt, lat = m.gather_state()
    torch.save({"ids": ids, "cnt": cnt, "lat": lat, "stats": st}, os.path.join(out_dir, f"shard{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_sharded_map_matches_single_map_golden(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    G = dict(np.load(ROOT / "tests" / "golden" / "map_golden.npz"))
    shards = [torch.load(tmp_path / f"shard{r}.pt", weights_only=False) for r in range(world)]
    ids = torch.cat([s["ids"] for s in shards]); cnt = torch.cat([s["cnt"] for s in shards]); lat = torch.cat([s["lat"] for s in shards])
    assert len(torch.unique(ids)) == len(ids)                     # every voxel owned by exactly one rank
    assert all(s["ids"].numel() > 100 for s in shards)            # both ranks own part of the scene
    assert all(s["stats"]["a2a_bytes"] > 0 for s in shards)
    o = torch.argsort(ids)
    ids, cnt, lat = ids[o].numpy(), cnt[o].numpy(), lat[o].numpy()
    # reference golden (single map): slot order differs, compare by voxel id
    r = np.argsort(G["k2_pos"])
    assert np.array_equal(ids, G["k2_pos"][r])
    assert np.array_equal(cnt, G["k2_count"][r])
    ref = G["k2_latent"][r]
    assert np.abs(lat - ref).max() <= 1e-3 * np.abs(ref).max()
