"""Oracle restatement of the dense-indexed latent voxel map.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/system/map.py (DenseIndexedMap): state (map.py:199-211), growth (:263-285),
integrate_keyframe (:341-453, do_optimize=False path), get_sdf (:560-580), do_meshing sampling (:625-688).
torch-CPU fp32/int64 so integer results are bit-comparable and float results follow the same op order.
"""

import numpy as np
import torch

from . import nets

OFFSETS8 = [(-0.5, -0.5, -0.5), (-0.5, -0.5, 0.5), (-0.5, 0.5, -0.5), (-0.5, 0.5, 0.5),
            (0.5, -0.5, -0.5), (0.5, -0.5, 0.5), (0.5, 0.5, -0.5), (0.5, 0.5, 0.5)]        # map.py:186-189
FACE6 = [(-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]               # map.py:193-196


class OracleMap:
    def __init__(self, weights, bound_min, bound_max, voxel_size, latent_dim=29, prune_min_vox_obs=16,
                 ignore_count_th=16.0, encoder_count_th=600.0, divide="ieee"):
        self.W = weights
        self.vs = voxel_size
        self.n_xyz = np.ceil((np.asarray(bound_max) - np.asarray(bound_min)) / voxel_size).astype(int).tolist()  # map.py:178
        self.bound_min = torch.tensor(bound_min).float()
        self.L = latent_dim
        self.prune_min = prune_min_vox_obs
        self.ignore_th = ignore_count_th
        self.enc_th = encoder_count_th
        self.divide = divide
        G = int(np.prod(self.n_xyz))
        self.n_occupied = 0
        self.indexer = torch.full((G,), -1, dtype=torch.long)
        self.latent_vecs = torch.zeros((1, latent_dim))
        self.latent_vecs_pos = torch.full((1,), -1, dtype=torch.long)
        self.voxel_obs_count = torch.zeros((1,))
        self.updated = torch.empty((0,), dtype=torch.long)

    # -- helpers -------------------------------------------------------------------------------
    def normalize(self, xyz):
        z = xyz - self.bound_min.unsqueeze(0)
        if self.divide == "ieee":                      # torch CPU: IEEE fp32 divide by float(0.1)
            return z / self.vs
        return z * torch.tensor(1.0 / np.float32(self.vs), dtype=torch.float32)   # torch CUDA: x * (1/vs)

    def lin(self, ijk):                                  # map.py:287-292
        return ijk[:, 2] + self.n_xyz[2] * ijk[:, 1] + (self.n_xyz[2] * self.n_xyz[1]) * ijk[:, 0]

    def unlin(self, idx):                                # map.py:294-301
        return torch.stack([idx // (self.n_xyz[1] * self.n_xyz[2]), (idx // self.n_xyz[2]) % self.n_xyz[1],
                            idx % self.n_xyz[2]], dim=-1)

    def dilate6(self, ids, ensure_valid=False):          # map.py:546-558
        pos = self.unlin(ids)
        out = [ids]
        for off in FACE6:
            q = pos + torch.tensor([off])
            for d in range(3):
                q[:, d].clamp_(0, self.n_xyz[d] - 1)
            q = self.lin(q)
            if ensure_valid:
                q = q[self.indexer[q] != -1]
            out.append(q)
        return torch.unique(torch.cat(out))

    def _grow(self, count):                              # map.py:263-285
        target = self.n_occupied + count
        cap = self.latent_vecs.size(0)
        if cap < target:
            new = cap
            while new < target:
                new *= 2
            lv = torch.zeros((new, self.L)); lv[:cap] = self.latent_vecs
            lp = torch.full((new,), -1, dtype=torch.long); lp[:cap] = self.latent_vecs_pos
            oc = torch.zeros((new,)); oc[:cap] = self.voxel_obs_count
            self.latent_vecs, self.latent_vecs_pos, self.voxel_obs_count = lv, lp, oc
        ids = torch.arange(self.n_occupied, target, dtype=torch.long)
        self.n_occupied = target
        return ids

    # -- integrate -----------------------------------------------------------------------------
    def integrate_keyframe(self, xyz, normal, return_debug=False):
        xn = self.normalize(xyz)                                         # map.py:367-368
        gid = self.lin(torch.ceil(xn).long() - 1)                        # :369-370
        unq_mask = None
        if self.prune_min > 0:                                           # :374-379
            _, inv, cnt = torch.unique(gid, return_counts=True, return_inverse=True)
            unq_mask = (cnt > self.prune_min)[inv]
            xn, gid, normal = xn[unq_mask], gid[unq_mask], normal[unq_mask]
        fresh = self.indexer[gid] == -1                                  # :382-388
        if fresh.sum() > 0:
            ids = self.dilate6(torch.unique(gid[fresh]))
            ids = ids[self.indexer[ids] == -1]
            slots = self._grow(ids.size(0))
            self.latent_vecs_pos[slots] = ids
            self.indexer[ids] = slots
        G = self.indexer.numel()
        status = torch.zeros(G, dtype=torch.short)                       # :408-412
        cand = self.latent_vecs_pos[torch.logical_and(self.voxel_obs_count < self.enc_th, self.latent_vecs_pos >= 0)]
        status[cand] = 1
        dbg = {}
        if cand.size(0) > 0:
            focus = torch.zeros(G, dtype=torch.long)                     # :390-398
            focus[self.dilate6(cand)] = 1
            fm = focus[gid] == 1
            pxn, pn = xn[fm], normal[fm]
            s_slots, s_in = [], []
            for off in OFFSETS8:                                         # :422-436
                g = torch.ceil(pxn + torch.tensor(off)) - 1
                for d in range(3):
                    g[:, d].clamp_(0, self.n_xyz[d] - 1)
                rel = pxn - g - torch.tensor([[0.5, 0.5, 0.5]])
                lg = self.lin(g.long())
                keep = status[lg] >= 1
                s_slots.append(self.indexer[lg][keep])
                s_in.append(torch.cat([rel[keep], pn[keep]], dim=-1))
            s_in = torch.cat(s_in); s_slots = torch.cat(s_slots)
            mapping, pinds, pcounts = torch.unique(s_slots, return_inverse=True, return_counts=True)   # :438-440
            pcounts = pcounts.float()
            enc = nets.encoder_forward(self.W, s_in)                     # :446-447
            ssum = torch.zeros((mapping.size(0), self.L)).index_add_(0, pinds, enc)              # :449 (groupby sum)
            ssum += self.latent_vecs[mapping] * self.voxel_obs_count[mapping].unsqueeze(-1)      # :450
            self.voxel_obs_count[mapping] += pcounts                                              # :451
            self.latent_vecs[mapping] = ssum / self.voxel_obs_count[mapping].unsqueeze(-1)        # :452
            self.updated = torch.unique(torch.cat([self.updated, mapping]))                       # :453,303-308
            dbg = {"sample_in": s_in, "sample_slot": s_slots, "mapping": mapping, "pcounts": pcounts}
        return (unq_mask, dbg) if return_debug else unq_mask

    # -- query ---------------------------------------------------------------------------------
    def get_sdf(self, xyz):
        """map.py:560-580.  xyz may require grad.  Returns sdf (M',), std (M',), valid (N,) bool."""
        xn = self.normalize(xyz)
        with torch.no_grad():
            gid3 = torch.ceil(xn.detach()).long() - 1
            slot = self.indexer[self.lin(gid3)]
            valid = slot != -1
            vv = self.voxel_obs_count[slot[valid]] > self.ignore_th
            valid[valid.clone()] = vv
            lat = self.latent_vecs[slot[valid]]
        rel = xn[valid] - gid3[valid] - torch.tensor([[0.5, 0.5, 0.5]])
        sdf, std = nets.decoder_forward(self.W, torch.cat([lat, rel], dim=1))
        return sdf, std, valid

    # -- meshing -------------------------------------------------------------------------------
    @staticmethod
    def sample_lattice(r, a, b):                         # utility.py:129-149
        idx = torch.arange(0, r ** 3, dtype=torch.long)
        vsize = (b - a) / (r - 1)
        s = torch.zeros(r ** 3, 3)
        s[:, 0] = (idx // (r * r)) * vsize + a
        s[:, 1] = ((idx // r) % r) * vsize + a
        s[:, 2] = (idx % r) * vsize + a
        return s

    def meshing_batch(self, updated_vec_id):
        """map.py:628-636: voxels to decode, and slot -> batch row mapping (int32, -1 = absent)."""
        focused = self.latent_vecs_pos[updated_vec_id]
        occ = self.indexer[self.dilate6(focused, ensure_valid=True)]
        occ = occ[self.voxel_obs_count[occ] > self.ignore_th]
        mapping = torch.full((int(occ.max().item()) + 1,), -1, dtype=torch.int)
        mapping[occ] = torch.arange(0, occ.size(0), dtype=torch.int)
        return focused, occ, mapping

    def decode_cubes(self, occ, voxel_resolution, refine_band=0.05, chunk=1 << 18):
        """map.py:637-688 (fast=True): low-res decode, trilinear x2 (align_corners=True), re-decode the
        |sdf|<band samples, negate.  Returns cube_sdf, cube_std (B, 2r, 2r, 2r)."""
        lat = self.latent_vecs[occ]
        B = lat.size(0)
        r = voxel_resolution
        a = -(r // 2) * (1. / r)
        b = 1. + (r - 1) // 2 * (1. / r)
        R = 2 * r
        low = self.sample_lattice(r, a, b) - torch.tensor([[0.5, 0.5, 0.5]])

        def run(latents, pts):
            so, st = [], []
            for i in range(0, latents.size(0), chunk):
                s, d = nets.decoder_forward(self.W, torch.cat([latents[i:i + chunk], pts[i:i + chunk]], dim=1))
                so.append(s); st.append(d)
            return torch.cat(so), torch.cat(st)

        lsdf, lstd = run(lat.unsqueeze(1).repeat(1, r ** 3, 1).view(-1, self.L), low.unsqueeze(0).repeat(B, 1, 1).view(-1, 3))
        F = torch.nn.functional
        hs = F.interpolate(lsdf.reshape(B, 1, r, r, r), mode="trilinear", size=(R, R, R), align_corners=True).reshape(B, R ** 3)
        hd = F.interpolate(lstd.reshape(B, 1, r, r, r), mode="trilinear", size=(R, R, R), align_corners=True).reshape(B, R ** 3)
        vb, vs_ = torch.where(hs.abs() < refine_band)
        if vb.size(0) > 0:
            high = self.sample_lattice(R, a, b) - torch.tensor([[0.5, 0.5, 0.5]])
            s, d = run(lat[vb], high[vs_])
            hs[vb, vs_] = s; hd[vb, vs_] = d
        return -hs.reshape(B, R, R, R), hd.reshape(B, R, R, R)

    def state(self):
        return {"n_occupied": self.n_occupied, "indexer": self.indexer, "latent_vecs": self.latent_vecs,
                "latent_vecs_pos": self.latent_vecs_pos, "voxel_obs_count": self.voxel_obs_count}
