"""Executable model of the sharded map's routing semantics in plain torch ops (CPU / gloo).  TEST INFRASTRUCTURE ONLY.

The product path is nerf-fusion_b200/sharded.py + csrc/sharded.cu (kernels writing into peer receive buffers); this model
states WHAT that path must compute -- which rank owns which voxel, where the prune is decided, which samples are accepted
(map.py:390-436 on the union of all ranks' points) -- with torch.distributed collectives, so that the semantics can be
checked against the single-map golden on CPU with world size 2 (tests/test_sharded_gloo.py).  Nothing under
nerf-fusion_b200/ imports it.

Per keyframe (integrate_keyframe, semantics of system/map.py:341-453 on the union of all ranks' points):
  1. every rank normalises its share of the points and routes them to the owner of their HOME voxel, so per-voxel
     observation counts are complete where the prune is decided;
  2. owners prune (`> prune_min_vox_obs`), and request allocation of unseen home voxels and their 6 clamped face
     neighbours; ids owned elsewhere travel to their owners; slots are numbered per shard;
  3. the candidate set (allocated, obs_count < encoder_count_th) is made global;
  4. owners of the points build the (point, offset) samples exactly like map.py:390-436 and route each ACCEPTED sample
     to the owner of its voxel;
  5. owners run the encoder on what they received and apply the running mean (map.py:446-452).
Parity with the single-GPU map is defined on {linear voxel id -> (latent, count)} (slot numbers are per shard).
"""
import numpy as np
import torch
import torch.distributed as dist

BRICK = 8
OFFSETS8 = [(-0.5, -0.5, -0.5), (-0.5, -0.5, 0.5), (-0.5, 0.5, -0.5), (-0.5, 0.5, 0.5),
            (0.5, -0.5, -0.5), (0.5, -0.5, 0.5), (0.5, 0.5, -0.5), (0.5, 0.5, 0.5)]
FACE6 = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]


def _all_to_all_rows(rows, dest, world, group=None):
    """Route `rows[i]` (2-D tensor or list of tensors with equal first dim) to rank dest[i].  Returns received rows."""
    single = not isinstance(rows, (list, tuple))
    tensors = [rows] if single else list(rows)
    order = torch.argsort(dest, stable=True)
    send_counts = torch.bincount(dest, minlength=world)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    ss, rs = send_counts.tolist(), recv_counts.tolist()
    out = []
    for t in tensors:
        t = t[order].contiguous()
        r = torch.empty((int(sum(rs)),) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_to_all_single(r, t, output_split_sizes=rs, input_split_sizes=ss, group=group)
        out.append(r)
    nbytes = sum(int(t.element_size() * t[0].numel()) for t in tensors if t.shape[0] > 0) * int(sum(ss))
    return (out[0] if single else out), nbytes


class ShardedMap:
    def __init__(self, weights, args, device, encoder_fn=None, group=None):
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.group = group
        self.device = torch.device(device)
        self.args = args
        self.voxel_size = args.voxel_size
        self.n_xyz = np.ceil((np.asarray(args.bound_max) - np.asarray(args.bound_min)) / args.voxel_size).astype(int).tolist()
        self.bound_min = torch.tensor(args.bound_min, device=self.device).float()
        self.G = int(np.prod(self.n_xyz))
        self.L = 29
        self.nb = [(n + BRICK - 1) // BRICK for n in self.n_xyz]
        if encoder_fn is None:
            from . import nets
            encoder_fn = lambda x: nets.encoder_forward(weights, x)               # noqa: E731
        self.encoder_fn = encoder_fn
        self.indexer = torch.full((self.G,), -1, dtype=torch.long, device=self.device)
        self.latent_vecs = torch.zeros((0, self.L), device=self.device)
        self.latent_vecs_pos = torch.zeros((0,), dtype=torch.long, device=self.device)
        self.voxel_obs_count = torch.zeros((0,), device=self.device)
        self.n_occupied = 0
        self.stats = {}

    # ------------------------------------------------------------------------------------------ helpers
    def _lin(self, ijk):
        return ijk[:, 2] + self.n_xyz[2] * ijk[:, 1] + (self.n_xyz[2] * self.n_xyz[1]) * ijk[:, 0]

    def _unlin(self, idx):
        return torch.stack([idx // (self.n_xyz[1] * self.n_xyz[2]), (idx // self.n_xyz[2]) % self.n_xyz[1], idx % self.n_xyz[2]], -1)

    def owner_of(self, idx):
        """8^3 bricks, hashed round-robin over the ranks."""
        p = self._unlin(idx) // BRICK
        brick = p[:, 2] + self.nb[2] * (p[:, 1] + self.nb[1] * p[:, 0])
        return brick % self.world

    def _dilate6(self, ids):
        pos = self._unlin(ids)
        out = [ids]
        for off in FACE6:
            q = pos + torch.tensor([off], device=ids.device)
            for d in range(3):
                q[:, d].clamp_(0, self.n_xyz[d] - 1)
            out.append(self._lin(q))
        return torch.unique(torch.cat(out))

    def _allocate(self, ids):
        ids = ids[self.indexer[ids] == -1]
        k = ids.numel()
        if k == 0:
            return 0
        slots = torch.arange(self.n_occupied, self.n_occupied + k, device=self.device)
        self.indexer[ids] = slots
        self.latent_vecs = torch.cat([self.latent_vecs, torch.zeros((k, self.L), device=self.device)])
        self.latent_vecs_pos = torch.cat([self.latent_vecs_pos, ids])
        self.voxel_obs_count = torch.cat([self.voxel_obs_count, torch.zeros((k,), device=self.device)])
        self.n_occupied += k
        return k

    # ------------------------------------------------------------------------------------------ integrate
    def integrate_keyframe(self, surface_xyz, surface_normal):
        """This rank's share of the keyframe's points (any split).  Collective: every rank must call it."""
        dev, W = self.device, self.world
        vs = torch.tensor(self.voxel_size, dtype=torch.float32, device=dev)
        xn = (surface_xyz.float() - self.bound_min.unsqueeze(0)) / vs               # tensor/tensor: IEEE divide on CPU and CUDA
        cell = torch.ceil(xn).long() - 1
        inb = ((cell >= 0) & (cell < torch.tensor(self.n_xyz, device=dev))).all(1)     # out-of-grid points are dropped
        xn, nrm, gid = xn[inb], surface_normal.float()[inb], self._lin(cell[inb])
        # 1. route points to the owner of their home voxel
        (xn, nrm), b1 = _all_to_all_rows([xn, nrm], self.owner_of(gid), W, self.group)
        gid = self._lin(torch.ceil(xn).long() - 1)
        # prune (map.py:373-379)
        if self.args.prune_min_vox_obs > 0 and gid.numel() > 0:
            _, inv, cnt = torch.unique(gid, return_inverse=True, return_counts=True)
            keep = (cnt > self.args.prune_min_vox_obs)[inv]
            xn, nrm, gid = xn[keep], nrm[keep], gid[keep]
        # 2. allocation requests: unseen home voxels (mine by construction) + their 6 neighbours (maybe remote)
        fresh = torch.unique(gid[self.indexer[gid] == -1]) if gid.numel() else gid
        req = self._dilate6(fresh) if fresh.numel() else fresh
        req, b2 = _all_to_all_rows(req.unsqueeze(1), self.owner_of(req), W, self.group)
        n_new = self._allocate(torch.unique(req.squeeze(1)))
        # 3. global candidate map (allocated and obs_count < encoder_count_th, map.py:410-412)
        cand = torch.zeros((self.G,), dtype=torch.uint8, device=dev)
        cand[self.latent_vecs_pos[self.voxel_obs_count < self.args.encoder_count_th]] = 1
        dist.all_reduce(cand, op=dist.ReduceOp.MAX, group=self.group)
        # 4. focus prune + samples (map.py:390-436), accepted samples go to the owner of their voxel
        if gid.numel():
            focus = self._focus(gid, cand)
            pxn, pn = xn[focus], nrm[focus]
        else:
            pxn, pn = xn, nrm
        ids, recs = [], []
        for off in OFFSETS8:
            g = torch.ceil(pxn + torch.tensor(off, device=dev)) - 1
            for d in range(3):
                g[:, d].clamp_(0, self.n_xyz[d] - 1)
            rel = pxn - g - 0.5
            lg = self._lin(g.long())
            ok = cand[lg] == 1
            ids.append(lg[ok]); recs.append(torch.cat([rel[ok], pn[ok]], -1))
        ids = torch.cat(ids) if ids else torch.zeros((0,), dtype=torch.long, device=dev)
        recs = torch.cat(recs) if recs else torch.zeros((0, 6), device=dev)
        (ids, recs), b3 = _all_to_all_rows([ids.unsqueeze(1), recs], self.owner_of(ids), W, self.group)
        ids = ids.squeeze(1)
        # 5. encoder + running mean on the owner (map.py:446-452)
        if ids.numel():
            slots = self.indexer[ids]
            mapping, pinds, pcounts = torch.unique(slots, return_inverse=True, return_counts=True)
            enc = self.encoder_fn(recs)
            s = torch.zeros((mapping.numel(), self.L), device=dev).index_add_(0, pinds, enc)
            s += self.latent_vecs[mapping] * self.voxel_obs_count[mapping].unsqueeze(-1)
            self.voxel_obs_count[mapping] += pcounts.float()
            self.latent_vecs[mapping] = s / self.voxel_obs_count[mapping].unsqueeze(-1)
        self.stats = {"points_in": int(xn.shape[0]), "samples_in": int(ids.numel()), "allocated": n_new,
                      "a2a_bytes": int(b1 + b2 + b3)}
        return self.stats

    def _focus(self, gid, cand):
        """home voxel in dilate6(candidates)  <=>  home or an in-range face neighbour is a candidate."""
        pos = self._unlin(gid)
        f = cand[gid] == 1
        for off in FACE6:
            q = pos + torch.tensor([off], device=gid.device)
            inr = ((q >= 0) & (q < torch.tensor(self.n_xyz, device=gid.device))).all(1)
            lq = self._lin(q.clamp(min=0)).clamp(max=self.G - 1)
            f |= inr & (cand[lq] == 1)
        return f

    def gather_state(self):
        """{voxel id -> (count, latent)} of this shard as sorted tensors (ids, counts, latents)."""
        o = torch.argsort(self.latent_vecs_pos)
        return self.latent_vecs_pos[o], self.voxel_obs_count[o], self.latent_vecs[o]
