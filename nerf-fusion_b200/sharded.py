"""Spatially sharded latent voxel map for large scenes (BASELINE config 5; new functionality, the reference
has no multi-GPU map).  One process per GPU; the voxel id space is partitioned in 8^3-voxel bricks dealt round-robin to
the ranks and every rank keeps indexer / latents / counts for ITS bricks only.

This module is the host side of csrc/sharded.cu: it owns the device memory of one rank, maps the peers' receive buffers
(CUDA IPC over NVLink / NVSwitch, `IpcFabric`) and drives the five phases of a keyframe with a stream-ordered barrier between
them.  The data path has no host decisions: records are written by the producing kernels straight into the owner's
receive buffer (peer stores), per-pair counts travel with the barrier (a one-block kernel exchanging epoch flags by peer
stores; optionally a 4-byte NCCL all-reduce) -- no count read-back, no torch index glue, no host sync inside a keyframe.

`LocalFabric` runs all ranks of a world inside ONE process on ONE device (peer pointers are plain pointers): the same
kernels and protocol, phase by phase over all ranks; it is how the multi-rank logic is tested on a single GPU and what
world size 1 uses.

Semantics: integrate_keyframe of system/map.py:341-453 on the union of all ranks' points.  Parity with the single-GPU
map is defined on {linear voxel id -> (latent, count)} (slot numbers are per shard).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, weights as W_
from ._lib import Shard, check

L = 29
REC = 32                     # bytes of a point / sample record
_CHANNELS = ("pts", "ids", "smp", "dlt")


def segment_tables(rank, world, bases, caps, rec_bytes):
    """Where `rank` writes: the address of ITS segment inside every rank's receive buffer.
    bases[d] = base address of rank d's buffer (as mapped in this process), caps = records per segment."""
    return [int(bases[d]) + rank * caps * rec_bytes for d in range(world)]


class _RankMemory:
    """Device memory of one rank.  Receive buffers and their count slots live in ONE allocation (`inbox`) so that a single
    IPC handle per rank maps everything a peer writes."""

    def __init__(self, lib, device, n_xyz, world, max_points_per_rank, capacity, delta_cap, ipc):
        self.device = torch.device(device)
        G = int(np.prod(n_xyz))
        self.n_local = int(lib.dfb_shard_local_cells(n_xyz[0], n_xyz[1], n_xyz[2], world))
        self.capacity, self.delta_cap, self.world = int(capacity), int(delta_cap), world
        # a source can send all its points to one owner, an owner can emit 8 samples per point it received
        self.pts_cap = int(max_points_per_rank)
        self.smp_cap = 8 * self.pts_cap * world
        self.ids_cap = 6 * self.pts_cap if world > 1 else 16
        self.dlt_cap = self.delta_cap
        z = lambda *shape, dtype=torch.int32: torch.zeros(shape, dtype=dtype, device=self.device)   # noqa: E731
        self.indexer_local = torch.full((self.n_local,), -1, dtype=torch.int32, device=self.device)
        self.latent_vecs = z(self.capacity, L, dtype=torch.float32)
        self.latent_vecs_pos = torch.full((self.capacity,), -1, dtype=torch.int32, device=self.device)
        self.voxel_obs_count = z(self.capacity, dtype=torch.float32)
        self.cand_bits = z((G + 31) // 32 + 1)
        self.grid_count = z(self.n_local)
        self.acc = z(self.capacity, L, dtype=torch.int64)           # fixed-point encoder sums (2^-34 units): order-independent
        self.acc_n = z(self.capacity)
        self.touched = z(self.capacity)
        self.counters = z(int(lib.dfb_shard_counter_ints()))
        self.delta_list, self.next_delta, self.n_next_delta = z(self.delta_cap), z(self.delta_cap), z(1)
        self.stats = z(8 + 2 * _lib.SHARD_MAX_WORLD)
        # inbox layout (bytes): [pts | smp | ids | dlt | counts(4 channels x world ints)]
        self.off = {}
        o = 0
        for name, nbytes in (("pts", world * self.pts_cap * REC), ("smp", world * self.smp_cap * REC), ("ids", world * self.ids_cap * 4),
                             ("dlt", world * self.dlt_cap * 4), ("cnt", 4 * 4 * world), ("flg", 4 * world)):
            self.off[name] = o
            o += (nbytes + 255) // 256 * 256
        self.inbox_bytes = o
        self.handle = None
        if ipc:
            ptr = C.c_void_p()
            handle = (C.c_ubyte * 64)()
            with torch.cuda.device(self.device):
                check(lib.dfb_peer_alloc(self.inbox_bytes, C.byref(ptr), handle))
            self.inbox_ptr, self.handle, self._owned = int(ptr.value), bytes(handle), True
        else:
            self._inbox = torch.zeros((self.inbox_bytes,), dtype=torch.uint8, device=self.device)
            self.inbox_ptr, self._owned = self._inbox.data_ptr(), False

    def rec_bytes(self, ch):
        return REC if ch in ("pts", "smp") else 4

    def cap(self, ch):
        return getattr(self, ch + "_cap")


class ShardedMap:
    """One rank of the sharded map.  Construct through `IpcFabric.create` (one process per GPU) or `LocalFabric.create`."""

    def __init__(self, weights, args, device, rank, world, max_points_per_rank, capacity, div_mode, ipc):
        if torch.device(device).type != "cuda":
            raise RuntimeError("difusion_b200 runs on CUDA devices only (no CPU fallback)")
        if not 1 <= world <= _lib.SHARD_MAX_WORLD:
            raise ValueError(f"world size must be 1..{_lib.SHARD_MAX_WORLD}")
        self.lib = _lib.load()
        self.device, self.rank, self.world, self.args = torch.device(device), rank, world, args
        self.voxel_size = float(args.voxel_size)
        self.n_xyz = np.ceil((np.asarray(args.bound_max) - np.asarray(args.bound_min)) / args.voxel_size).astype(int).tolist()
        self.G = int(np.prod(self.n_xyz))
        if self.G >= 1 << 30:
            raise ValueError("grid too large (candidate deltas carry 31-bit voxel ids)")
        self.encoder_blob = torch.from_numpy(W_.pack_encoder(weights)).to(self.device)
        self.mem = _RankMemory(self.lib, device, self.n_xyz, world, max_points_per_rank, capacity, max(capacity, 1 << 16), ipc)
        self.S = Shard()
        s, m = self.S, self.mem
        s.nx, s.ny, s.nz = self.n_xyz
        s.bound_min = (C.c_float * 3)(*[float(v) for v in args.bound_min])
        s.voxel_size, s.div_mode = self.voxel_size, int(div_mode)
        s.prune_min_vox_obs, s.encoder_count_th = int(args.prune_min_vox_obs), float(args.encoder_count_th)
        s.world, s.rank = world, rank
        for name in ("indexer_local", "latent_vecs", "latent_vecs_pos", "voxel_obs_count", "cand_bits", "grid_count", "acc", "acc_n",
                     "touched", "counters", "delta_list", "next_delta", "n_next_delta"):
            setattr(s, name, getattr(m, name).data_ptr())
        s.capacity, s.delta_cap = m.capacity, m.delta_cap
        for i, ch in enumerate(_CHANNELS):
            setattr(s, ch + "_inbox", m.inbox_ptr + m.off[ch])
            setattr(s, ch + "_cap", m.cap(ch))
            setattr(s, ch + "_count", m.inbox_ptr + m.off["cnt"] + 4 * world * i)
        s.flags = m.inbox_ptr + m.off["flg"]
        self.epoch = 0
        self.last_stats = None

    def connect(self, inbox_bases):
        """inbox_bases[d]: address of rank d's inbox as mapped in THIS process (own inbox for d == rank)."""
        s, m = self.S, self.mem
        for i, ch in enumerate(_CHANNELS):
            seg = segment_tables(self.rank, self.world, [b + m.off[ch] for b in inbox_bases], m.cap(ch), m.rec_bytes(ch))
            cnt = [inbox_bases[d] + m.off["cnt"] + 4 * self.world * i + 4 * self.rank for d in range(self.world)]
            for d in range(self.world):
                getattr(s, "peer_" + ch)[d] = seg[d]
                getattr(s, "peer_" + ch + "_count")[d] = cnt[d]
        for d in range(self.world):
            s.peer_flags[d] = inbox_bases[d] + m.off["flg"] + 4 * self.rank

    # ---- phases (each asynchronous on the current stream) ----------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def phase(self, k, xyz=None, normal=None):
        with torch.cuda.device(self.device):
            if k == 1:
                xyz = xyz.contiguous().float(); normal = normal.contiguous().float()
                self._keep = (xyz, normal)                                    # alive until the kernel has run
                check(self.lib.dfb_shard_phase1(C.byref(self.S), C.c_void_p(xyz.data_ptr()), C.c_void_p(normal.data_ptr()), xyz.size(0), self._stream()))
            elif k == 2:
                check(self.lib.dfb_shard_phase2(C.byref(self.S), self._stream()))
            elif k == 3:
                check(self.lib.dfb_shard_phase3(C.byref(self.S), self._stream()))
            elif k == 4:
                check(self.lib.dfb_shard_phase4(C.byref(self.S), self._stream()))
            else:
                check(self.lib.dfb_shard_phase5(C.byref(self.S), C.c_void_p(self.encoder_blob.data_ptr()), C.c_void_p(self.mem.stats.data_ptr()),
                                                self._stream()))

    def read_stats(self):
        """One host read, AFTER the keyframe (statistics only; nothing on the data path depends on it)."""
        st = self.mem.stats.cpu().numpy()
        W = _lib.SHARD_MAX_WORLD
        if st[3] & 1:
            raise _lib.DfbError("sharded map: a receive-buffer segment overflowed (raise max_points_per_rank)")
        if st[3] & 2:
            raise _lib.DfbError("sharded map: slot capacity exceeded (raise capacity)")
        remote_pts = int(st[8:8 + self.world].sum() - st[8 + self.rank]); remote_smp = int(st[8 + W:8 + W + self.world].sum() - st[8 + W + self.rank])
        self.last_stats = {"points_in": int(st[0]), "samples_in": int(st[1]), "allocated": int(st[2]), "n_occupied": int(st[4]),
                           "voxels_updated": int(st[5]), "points_sent_remote": remote_pts, "samples_sent_remote": remote_smp,
                           "peer_store_bytes": REC * (remote_pts + remote_smp)}
        return self.last_stats

    @property
    def n_occupied(self):
        return int(self.mem.counters[3 * _lib.SHARD_MAX_WORLD + 3].item())

    def gather_state(self):
        """{voxel id -> (count, latent)} of this shard as tensors sorted by id: (ids int64, counts, latents)."""
        n = self.n_occupied
        ids = self.mem.latent_vecs_pos[:n].long()
        o = torch.argsort(ids)
        return ids[o], self.mem.voxel_obs_count[:n][o], self.mem.latent_vecs[:n][o]


class LocalFabric:
    """All ranks of a world in one process on one device.  Same kernels, same protocol; the barrier between phases is the
    stream order itself."""

    def __init__(self, maps):
        self.maps = maps
        bases = [m.mem.inbox_ptr for m in maps]
        for m in maps:
            m.connect(bases)

    @classmethod
    def create(cls, weights, args, device, world, max_points_per_rank, capacity, div_mode=0):
        return cls([ShardedMap(weights, args, device, r, world, max_points_per_rank, capacity, div_mode, ipc=False) for r in range(world)])

    def integrate_keyframe(self, clouds):
        """clouds[r] = (xyz, normal): rank r's share of the keyframe."""
        for k in (1, 2, 3, 4, 5):
            for m, (p, n) in zip(self.maps, clouds):
                m.phase(k, p, n)
        return self

    def gather_state(self):
        parts = [m.gather_state() for m in self.maps]
        ids = torch.cat([p[0] for p in parts]); o = torch.argsort(ids)
        return ids[o], torch.cat([p[1] for p in parts])[o], torch.cat([p[2] for p in parts])[o]


class IpcFabric:
    """One rank per process (torch.distributed, one GPU each, one node).  Receive buffers are cudaMalloc'ed, exported with
    CUDA IPC and mapped by every peer once; the barrier is a 4-byte all-reduce on the map's stream (NCCL), so a keyframe is
    enqueued without a single host synchronisation."""

    def __init__(self, m, group=None, barrier="peer"):
        import torch.distributed as dist
        self.dist, self.group, self.map, self.barrier_kind = dist, group, m, barrier
        m.S.fuse_publish = 1 if barrier == "peer" else 0       # the barrier kernel also carries the per-pair record counts
        handles = [None] * m.world
        dist.all_gather_object(handles, (m.rank, m.mem.handle), group=group)
        self.bases, self._opened = self.open_peers(m, dict(handles))
        m.connect(self.bases)
        self._token = torch.zeros((1,), dtype=torch.int32, device=m.device)
        dist.all_reduce(self._token, group=group)                 # everybody has mapped everybody before the first peer store
        torch.cuda.synchronize(m.device)

    @staticmethod
    def open_peers(m, handle_of, opener=None):
        """Map every peer's inbox; returns (bases, opened pointers).  `opener(handle) -> address` defaults to dfb_peer_open."""
        def _open(h):
            ptr = C.c_void_p()
            buf = (C.c_ubyte * 64).from_buffer_copy(h)
            with torch.cuda.device(m.device):
                check(m.lib.dfb_peer_open(buf, C.byref(ptr)))
            return int(ptr.value)
        opener = opener or _open
        bases, opened = [], []
        for d in range(m.world):
            if d == m.rank:
                bases.append(m.mem.inbox_ptr)
            else:
                a = opener(handle_of[d]); bases.append(a); opened.append(a)
        return bases, opened

    @classmethod
    def create(cls, weights, args, device, max_points_per_rank, capacity, div_mode=0, group=None, barrier="peer"):
        import torch.distributed as dist
        m = ShardedMap(weights, args, device, dist.get_rank(group), dist.get_world_size(group), max_points_per_rank, capacity, div_mode, ipc=True)
        return cls(m, group, barrier)

    def barrier(self, after_phase=0):
        """Stream-ordered barrier between two phases.  "peer": a one-block kernel exchanging epoch flags by peer stores (a few
        microseconds); "nccl": a 4-byte all-reduce (~20 us per barrier, measured 62 % -> see DESIGN.md scaling table)."""
        m = self.map
        if self.barrier_kind == "nccl":
            self.dist.all_reduce(self._token, group=self.group)
            return
        m.epoch += 1
        with torch.cuda.device(m.device):
            check(m.lib.dfb_shard_barrier(C.byref(m.S), m.epoch, after_phase, m._stream()))

    def integrate_keyframe(self, xyz, normal):
        """This rank's share of the keyframe (any split).  Collective: every rank calls it.  Asynchronous."""
        m = self.map
        for k in (1, 2, 3, 4):
            m.phase(k, xyz, normal)
            self.barrier(k)
        m.phase(5)
        return self

    def close(self):
        torch.cuda.synchronize(self.map.device)
        self.dist.barrier(group=self.group)
        for a in self._opened:
            self.map.lib.dfb_peer_close(C.c_void_p(a))
        self._opened = []
        self.dist.barrier(group=self.group)
        if self.map.mem._owned:
            self.map.lib.dfb_peer_free(C.c_void_p(self.map.mem.inbox_ptr)); self.map.mem._owned = False
