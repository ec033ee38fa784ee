"""BASELINE.json config 5: large synthetic scene, spatially sharded map, records pushed into the owners' receive buffers
over NVLink (csrc/sharded.cu).  Run under torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_bench.py
or with N = 1 directly.  Each keyframe = `--points` points (split evenly over the ranks) on a wavy floor patch of a
40 x 10 x 40 m volume at 0.05 m voxels (128 M cells; three storeys of 4 x 4 patches: > 4 M voxels allocated over 48 keyframes).  Reports points/s integrated (device
time, max over ranks), peer-store bytes, and checks the sharded state against a single-GPU DenseIndexedMap (rank 0).
bench.py imports `run()` for its N > 1 line."""
import argparse, importlib, json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np, torch, torch.distributed as dist


def scene_points(n, seed, lo, hi, dev, y0=1.0):
    """points + normals on a wavy floor patch inside [lo, hi] (x, z), dense enough for the > 16 obs/voxel prune"""
    g = torch.Generator(device=dev); g.manual_seed(seed)
    x = torch.rand(n, generator=g, device=dev) * (hi[0] - lo[0]) + lo[0]
    z = torch.rand(n, generator=g, device=dev) * (hi[1] - lo[1]) + lo[1]
    y = y0 + 0.3 * torch.sin(0.9 * x) * torch.cos(0.7 * z)
    nx, nz = -0.27 * torch.cos(0.9 * x) * torch.cos(0.7 * z), 0.21 * torch.sin(0.9 * x) * torch.sin(0.7 * z)
    nrm = torch.stack([nx, torch.ones_like(nx), nz], 1); nrm = nrm / nrm.norm(dim=1, keepdim=True)
    return torch.stack([x, y, z], 1).contiguous(), nrm.contiguous()


def run(points=1_000_000, keyframes=48, voxel=0.05, parity=True, force_world1=False, barrier="peer"):
    """Collective over the default process group (initialised by the caller; world size 1 works without one).
    force_world1: run the one-rank configuration on THIS GPU only (the N = 1 baseline inside a multi-rank launch)."""
    from util import GOLD, MAPPING, ns
    dfb = importlib.import_module("nerf-fusion_b200")
    world = dist.get_world_size() if dist.is_initialized() and not force_world1 else 1
    rank = dist.get_rank() if dist.is_initialized() and not force_world1 else 0
    dev = f"cuda:{torch.cuda.current_device()}"
    W = dfb.weights.load_npz(GOLD / "weights.npz")
    sh = dfb.sharded

    def make(args, max_pts, capacity):
        if world == 1:
            return sh.LocalFabric.create(W, args, dev, 1, max_pts, capacity)
        return sh.IpcFabric.create(W, args, dev, max_pts, capacity, barrier=barrier)

    def integrate(fab, P, N):
        if world == 1:
            fab.integrate_keyframe([(P, N)])
        else:
            fab.integrate_keyframe(P, N)

    def the_map(fab):
        return fab.maps[0] if world == 1 else fab.map

    par = None
    if parity:                         # sharded (all ranks) vs single map (rank 0), default 0.1 m grid, two keyframes
        args_small = ns(dict(MAPPING))
        fab = make(args_small, 400_000, 1 << 16)
        one = dfb.DenseIndexedMap(W, args_small, 29, torch.device(dev)) if rank == 0 else None
        for k in range(2):
            P, N = scene_points(400_000, 1 + k, (-3.0 + k, -2.0), (4.0, 5.0 - k), dev)
            integrate(fab, P[rank::world], N[rank::world])
            if rank == 0:
                one.integrate_keyframe(P, N)
        the_map(fab).read_stats()
        ids, cnt, lat = the_map(fab).gather_state()
        if world > 1:
            sizes = [torch.zeros(1, dtype=torch.long, device=dev) for _ in range(world)]
            dist.all_gather(sizes, torch.tensor([ids.numel()], device=dev))
            mx = int(max(s.item() for s in sizes))
            pad = lambda t, fill: torch.cat([t, torch.full((mx - t.shape[0],) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=dev)])  # noqa: E731
            gi = [torch.zeros(mx, dtype=torch.long, device=dev) for _ in range(world)]; dist.all_gather(gi, pad(ids, -1))
            gc = [torch.zeros(mx, device=dev) for _ in range(world)]; dist.all_gather(gc, pad(cnt, 0.0))
            gl = [torch.zeros((mx, 29), device=dev) for _ in range(world)]; dist.all_gather(gl, pad(lat, 0.0))
            ids, cnt, lat = torch.cat(gi), torch.cat(gc), torch.cat(gl)
            keep = ids >= 0; ids, cnt, lat = ids[keep], cnt[keep], lat[keep]
        if rank == 0:
            n1 = one.n_occupied
            o = torch.argsort(one.latent_vecs_pos[:n1]); so = torch.argsort(ids)
            same_ids = torch.equal(ids[so], one.latent_vecs_pos[:n1][o]); same_cnt = same_ids and torch.equal(cnt[so], one.voxel_obs_count[:n1][o])
            err = (lat[so] - one.latent_vecs[:n1][o]).abs().max().item() / one.latent_vecs[:n1].abs().max().item() if same_ids else float("nan")
            par = {"voxels": int(n1), "ids_equal": bool(same_ids), "counts_equal": bool(same_cnt), "latent_rel_err": err}
            assert same_ids and same_cnt and err < 1e-3, par
        if world > 1:
            fab.close()
        del fab, one
        torch.cuda.empty_cache()

    # ---- throughput: config 5
    cfg = dict(MAPPING); cfg.update(bound_min=[-20.0, -2.0, -20.0], bound_max=[20.0, 8.0, 20.0], voxel_size=voxel)
    per_rank = points // world
    big = make(ns(cfg), per_rank, int(1.4 * 100_000 * (keyframes + 2) / world) + (1 << 16))
    m = the_map(big)
    # 4 x 4 patches of 9 x 9 m on three storeys (y0 = -0.5 / 2.75 / 6.0): 48 keyframes allocate > 4 M voxels
    tiles = [(-19.0 + 9.5 * (k % 4), -19.0 + 9.5 * (k // 4 % 4), -0.5 + 3.25 * (k // 16 % 3)) for k in range(keyframes + 2)]
    # the clouds are generated up front; the timed region is `keyframes` integrate calls enqueued BACK TO BACK (no host
    # synchronisation inside: the data path has none), device time between two events, max over ranks
    clouds = [scene_points(per_rank, 100 + 17 * k + rank, (x0, z0), (x0 + 9.0, z0 + 9.0), dev, y0) for k, (x0, z0, y0) in enumerate(tiles)]
    for k in range(2):                               # two warm-up keyframes
        integrate(big, *clouds[k])
    m.read_stats()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    snaps = []
    e0.record()
    for k in range(2, keyframes + 2):
        integrate(big, *clouds[k])
        snaps.append(m.mem.stats.clone())            # device-side copy of the keyframe's statistics, read after the timed region
    e1.record()
    torch.cuda.synchronize()
    m.read_stats()                                   # raises if a segment overflowed / capacity was exceeded
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    W8 = dfb._lib.SHARD_MAX_WORLD
    st = torch.stack(snaps).long()
    remote = (st[:, 8:8 + world].sum() - st[:, 8 + rank].sum()) + (st[:, 8 + W8:8 + W8 + world].sum() - st[:, 8 + W8 + rank].sum())
    b = torch.stack([remote * 32, st[:, 1].sum()])
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(b)
    total_ms, total_bytes, total_samples = ms.item(), int(b[0]), int(b[1])
    nv = torch.tensor([m.n_occupied], device=dev)
    if world > 1:
        dist.all_reduce(nv)
        big.close()
    out = {"workload": "config 5: sharded map, %d keyframes x %d points, voxel %.3f m, %d grid cells" % (keyframes, points, voxel, m.G),
           "n_gpus": world, "points_per_s": keyframes * points / (total_ms * 1e-3), "ms_per_keyframe": total_ms / keyframes,
           "samples_per_keyframe": total_samples / keyframes, "peer_store_bytes_per_keyframe": total_bytes / keyframes,
           "voxels_total": int(nv.item()), "transport": "peer stores from the producing kernels (CUDA IPC over NVLink), 4 barriers per keyframe (%s), no host sync" % ("epoch flags exchanged by peer stores" if barrier == "peer" else "4-byte NCCL all-reduce"),
           "parity_vs_single_gpu": par}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1_000_000)
    ap.add_argument("--keyframes", type=int, default=48)
    ap.add_argument("--voxel", type=float, default=0.05)
    ap.add_argument("--barrier", default="peer", choices=["peer", "nccl"])
    a = ap.parse_args()
    world, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    res = run(a.points, a.keyframes, a.voxel, barrier=a.barrier)
    if int(os.environ.get("RANK", 0)) == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()
