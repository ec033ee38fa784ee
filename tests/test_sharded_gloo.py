"""World-size-2 gloo tests (CPU) of the sharded map: (1) the routing semantics (oracle/sharded_model.py: two ranks, each fed
half of the points, must reproduce the single-map golden as {voxel id -> (count, latent)}); (2) the host logic of the product
path (nerf-fusion_b200/sharded.py): handle exchange and the segment tables every rank derives from it."""
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
    from oracle import sharded_model as sharded
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


def _fabric_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT))
    import importlib
    import types
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module("nerf-fusion_b200.sharded")
    # a stand-in for one rank's memory: only what the handle exchange and the table construction look at
    mem = types.SimpleNamespace(inbox_ptr=0x10000000 * (rank + 1), handle=bytes([rank]) * 64,
                                off={"pts": 0, "smp": 4096, "ids": 8192, "dlt": 12288, "cnt": 16384, "flg": 16512}, pts_cap=16, smp_cap=32, ids_cap=8, dlt_cap=8)
    mem.cap = lambda ch: getattr(mem, ch + "_cap")
    mem.rec_bytes = lambda ch: 32 if ch in ("pts", "smp") else 4
    handles = [None] * world
    dist.all_gather_object(handles, (rank, mem.handle))
    m = types.SimpleNamespace(rank=rank, world=world, mem=mem, device="cpu", lib=None)
    bases, opened = sh.IpcFabric.open_peers(m, dict(handles), opener=lambda h: 0x10000000 * (h[0] + 1) + 0x1000000)   # "mapped" address of a peer
    tables = {ch: sh.segment_tables(rank, world, [b + mem.off[ch] for b in bases], mem.cap(ch), mem.rec_bytes(ch)) for ch in ("pts", "ids", "smp", "dlt")}
    torch.save({"bases": bases, "opened": opened, "tables": tables}, os.path.join(out_dir, f"fabric{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_fabric_handle_exchange_and_segment_tables(tmp_path):
    world = 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_fabric_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    out = [torch.load(tmp_path / f"fabric{r}.pt", weights_only=False) for r in range(world)]
    for r in range(world):
        own = 0x10000000 * (r + 1)
        assert out[r]["bases"][r] == own and len(out[r]["opened"]) == world - 1
        for d in range(world):
            if d != r:                                             # the peer's handle was the one opened
                assert out[r]["bases"][d] == 0x10000000 * (d + 1) + 0x1000000
        # rank r writes segment r of every destination: segments of different sources never overlap inside a destination
        for ch, rec, cap, off in (("pts", 32, 16, 0), ("smp", 32, 32, 4096), ("ids", 4, 8, 8192), ("dlt", 4, 8, 12288)):
            for d in range(world):
                assert out[r]["tables"][ch][d] == out[r]["bases"][d] + off + r * cap * rec
    a, b = out[0]["tables"]["pts"], out[1]["tables"]["pts"]
    assert abs((a[0] - out[0]["bases"][0]) - (b[0] - out[1]["bases"][0])) == 16 * 32
