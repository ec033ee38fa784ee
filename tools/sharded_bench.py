"""BASELINE.json config 5: large synthetic scene, spatially sharded map, point all-to-all over NCCL/NVLink.
Run under torchrun:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_bench.py
Each keyframe = `--points` points (split evenly over the ranks) on a wavy 40 x 40 m floor + walls, 0.05 m voxels over
a 40 x 10 x 40 m volume (G = 800*200*800 = 128 M cells).  Reports points/s integrated (device time, max over ranks),
all-to-all bytes, and checks the sharded state against a single-GPU DenseIndexedMap on a small case (rank 0)."""
import argparse, importlib, json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch, torch.distributed as dist
from util import GOLD, MAPPING, ns

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=1_000_000)
ap.add_argument("--keyframes", type=int, default=8)
ap.add_argument("--voxel", type=float, default=0.05)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
dfb = importlib.import_module("nerf-fusion_b200")
W = dfb.weights.load_npz(GOLD / "weights.npz")


def scene_points(n, seed, lo, hi):
    """points + normals on a wavy floor patch inside [lo, hi] (x, z), dense enough for the > 16 obs/voxel prune"""
    g = torch.Generator(device=dev); g.manual_seed(seed)
    x = torch.rand(n, generator=g, device=dev) * (hi[0] - lo[0]) + lo[0]
    z = torch.rand(n, generator=g, device=dev) * (hi[1] - lo[1]) + lo[1]
    y = 1.0 + 0.3 * torch.sin(0.9 * x) * torch.cos(0.7 * z)
    nx, nz = -0.27 * torch.cos(0.9 * x) * torch.cos(0.7 * z), 0.21 * torch.sin(0.9 * x) * torch.sin(0.7 * z)
    nrm = torch.stack([nx, torch.ones_like(nx), nz], 1); nrm = nrm / nrm.norm(dim=1, keepdim=True)
    return torch.stack([x, y, z], 1).contiguous(), nrm.contiguous()


# ---- parity: sharded (all ranks) vs single map (rank 0), default 0.1 m grid
args_small = ns(dict(MAPPING))
sm = dfb.sharded.ShardedMap(W, args_small, dev)
P, N = scene_points(400_000, 1, (-3.0, -2.0), (4.0, 5.0))
sm.integrate_keyframe(P[rank::world], N[rank::world])
ids, cnt, lat = sm.gather_state()
sizes = [torch.zeros(1, dtype=torch.long, device=dev) for _ in range(world)]
dist.all_gather(sizes, torch.tensor([ids.numel()], device=dev))
mx = int(max(s.item() for s in sizes))
pad = lambda t, fill: torch.cat([t, torch.full((mx - t.shape[0],) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=dev)])
gi = [torch.zeros(mx, dtype=torch.long, device=dev) for _ in range(world)]; dist.all_gather(gi, pad(ids, -1))
gc = [torch.zeros(mx, device=dev) for _ in range(world)]; dist.all_gather(gc, pad(cnt, 0.0))
gl = [torch.zeros((mx, 29), device=dev) for _ in range(world)]; dist.all_gather(gl, pad(lat, 0.0))
parity = None
if rank == 0:
    one = dfb.DenseIndexedMap(W, args_small, 29, torch.device(dev))
    one.integrate_keyframe(P, N)
    n1 = one.n_occupied
    o = torch.argsort(one.latent_vecs_pos[:n1])
    ai = torch.cat(gi); keep = ai >= 0; ai = ai[keep]; ac = torch.cat(gc)[keep]; al = torch.cat(gl)[keep]
    so = torch.argsort(ai)
    same_ids = torch.equal(ai[so], one.latent_vecs_pos[:n1][o]); same_cnt = torch.equal(ac[so], one.voxel_obs_count[:n1][o])
    err = (al[so] - one.latent_vecs[:n1][o]).abs().max().item() / one.latent_vecs[:n1].abs().max().item()
    parity = {"voxels": int(n1), "ids_equal": bool(same_ids), "counts_equal": bool(same_cnt), "latent_rel_err": err}
    assert same_ids and same_cnt and err < 1e-3, parity

# ---- throughput: config 5
cfg = dict(MAPPING); cfg.update(bound_min=[-20.0, -2.0, -20.0], bound_max=[20.0, 8.0, 20.0], voxel_size=a.voxel)
big = dfb.sharded.ShardedMap(W, ns(cfg), dev)
per_rank = a.points // world
tiles = [(-19.0 + 9.5 * (k % 4), -19.0 + 9.5 * (k // 4 % 4)) for k in range(a.keyframes + 2)]
total_ms, total_bytes, total_samples = 0.0, 0, 0
for k, (x0, z0) in enumerate(tiles):
    P, N = scene_points(per_rank, 100 + 17 * k + rank, (x0, z0), (x0 + 9.0, z0 + 9.0))
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st = big.integrate_keyframe(P, N); e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    b = torch.tensor([st["a2a_bytes"], st["samples_in"]], device=dev, dtype=torch.long); dist.all_reduce(b)
    if k >= 2:                                   # two warm-up keyframes
        total_ms += ms.item(); total_bytes += int(b[0]); total_samples += int(b[1])
nv = torch.tensor([big.n_occupied], device=dev); dist.all_reduce(nv)
if rank == 0:
    out = {"config": "sharded map, %d keyframes x %d points, voxel %.3f m, grid cells %d" % (a.keyframes, a.points, a.voxel, big.G),
           "n_gpus": world, "points_per_s": a.keyframes * a.points / (total_ms * 1e-3), "ms_per_keyframe": total_ms / a.keyframes,
           "samples_per_keyframe": total_samples / a.keyframes, "a2a_bytes_per_keyframe": total_bytes / a.keyframes,
           "voxels_total": int(nv.item()), "parity_vs_single_gpu": parity}
    print(json.dumps(out))
dist.destroy_process_group()
