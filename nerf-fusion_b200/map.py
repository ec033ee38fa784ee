"""DenseIndexedMap: host-side mirror of the reference's map class (system/map.py:158-833) for the hot path.

Same constructor, method names, argument meaning and state layout (``cold_vars``: indexer int64 (G,), latent_vecs
f32 (cap,29), latent_vecs_pos int64 (cap,), voxel_obs_count f32 (cap,), voxel_optimized bool (cap,), n_occupied;
capacity doubles from 1 like map.py:263-285, so ``save()`` files are interchangeable with the reference's).
All arithmetic runs in libdifusion_b200.so through the C-ABI; torch holds the buffers and the stream.

Not provided (out of scope): the Adam latent optimiser (``do_optimize=True``), the async
optimiser process and the async meshing thread (``run_async: false`` in fusion-lr-kt.yaml); Open3D geometry is
optional (a plain SimpleMesh is returned when open3d is not installed).
"""
import argparse
import ctypes as C
import logging
from pathlib import Path

import numpy as np
import torch

from . import _lib, ext, weights as W_
from ._lib import MapParams, check
from .ext import _p, _stream

DIV_IEEE, DIV_RECIP = 0, 1


class MeshExtractCache:
    """map.py:116-133 (host-side mesh cache; updated voxel ids are kept as a device flag array here)."""

    def __init__(self, owner):
        self._owner = owner
        self.vertices = None
        self.vertices_flatten_id = None
        self.vertices_std = None

    @property
    def updated_vec_id(self):
        """Sorted unique slot ids touched since the last extraction (map.py:303-308)."""
        f = self._owner._updated_flag
        return torch.nonzero(f[:self._owner.n_occupied]).squeeze(-1)

    def clear_updated_vec(self):
        self._owner._updated_flag.zero_()

    def clear_all(self):
        self.vertices = None
        self.vertices_flatten_id = None
        self.vertices_std = None
        self.clear_updated_vec()


class SimpleMesh:
    """Stand-in for o3d.geometry.TriangleMesh when open3d is absent: float64 vertices, int32 triangles, rgb colours."""

    def __init__(self, vertices, triangles, vertex_colors=None, vertex_std=None):
        self.vertices = vertices
        self.triangles = triangles
        self.vertex_colors = vertex_colors
        self.vertex_std = vertex_std


class SimpleLineSet:
    """Stand-in for open3d.geometry.LineSet when open3d is not installed."""

    def __init__(self, points, lines, colors=None):
        self.points, self.lines, self.colors = points, lines, colors


class _GetSdfFn(torch.autograd.Function):
    """sdf/std of every query row with the reverse pass done by the same kernel (replaces the autograd tape the
    reference builds through forward_model(no_detach=True), map.py:578-579)."""

    @staticmethod
    def forward(ctx, xyz, owner):
        sdf, std, valid = owner._get_sdf_raw(xyz.detach(), None, None)
        ctx.owner = owner
        ctx.save_for_backward(xyz.detach())
        ctx.mark_non_differentiable(valid)
        return sdf, std, valid

    @staticmethod
    def backward(ctx, g_sdf, g_std, _g_valid):
        (xyz,) = ctx.saved_tensors
        g_sdf = torch.zeros(xyz.size(0), device=xyz.device) if g_sdf is None else g_sdf.contiguous().float()
        g_std = torch.zeros(xyz.size(0), device=xyz.device) if g_std is None else g_std.contiguous().float()
        grad = ctx.owner._get_sdf_raw(xyz, g_sdf, g_std)
        return grad, None


class DenseIndexedMap:
    def __init__(self, model, args: argparse.Namespace, latent_dim: int, device: torch.device, enable_async: bool = False,
                 optimization_device: torch.device = None, div_mode: int = DIV_IEEE, reserve_voxels: int = 1 << 17):
        """model: an object with .decoder/.encoder torch modules (the reference's Networks), or a dict of checkpoint
        tensors with 'dec.'/'enc.' prefixed keys (weights.load_checkpoint / load_npz).
        div_mode: how x/voxel_size is rounded -- DIV_IEEE reproduces the reference on CPU (the parity oracle),
        DIV_RECIP reproduces torch-CUDA's multiply-by-reciprocal.
        reserve_voxels: rows of backing storage reserved up front (33 MB at the default 2^17).  The per-voxel tensors
        in cold_vars are views of the first `capacity` rows, and capacity still doubles from 1 like map.py:263-285, so
        growing the map is a re-slice -- no allocation or copy in the frame loop until the reservation is exceeded."""
        if enable_async:
            raise NotImplementedError("the asynchronous optimisation process / meshing thread are not provided (run_async: false); "
                                      "the optimiser itself runs synchronously (integrate_keyframe(do_optimize=True))")
        if latent_dim != 29:
            raise ValueError("kernels are specialised for the shipped checkpoint (latent dim 29, hyper.json)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("difusion_b200 runs on CUDA devices only (no CPU fallback)")
        self.lib = _lib.load()
        self.model = model
        self.args = args
        self.latent_dim = latent_dim
        self.voxel_size = args.voxel_size
        self.n_xyz = np.ceil((np.asarray(args.bound_max) - np.asarray(args.bound_min)) / args.voxel_size).astype(int).tolist()
        logging.info(f"Map size Nx = {self.n_xyz[0]}, Ny = {self.n_xyz[1]}, Nz = {self.n_xyz[2]}")
        self.bound_min = torch.tensor(args.bound_min, device=self.device).float()
        self.bound_max = self.bound_min + self.voxel_size * torch.tensor(self.n_xyz, device=self.device)
        self.extract_mesh_std_range = None
        self.div_mode = div_mode

        tensors = self._weights_of(model)
        self.decoder_blob = torch.from_numpy(W_.pack_decoder(tensors)).to(self.device)
        self.encoder_blob = torch.from_numpy(W_.pack_encoder(tensors)).to(self.device)
        assert self.decoder_blob.numel() == self.lib.dfb_decoder_blob_floats()
        assert self.encoder_blob.numel() == self.lib.dfb_encoder_blob_floats()

        G = int(np.prod(self.n_xyz))
        self.n_cells = G
        self.cold_vars = {
            "n_occupied": 0,
            "indexer": torch.full((G,), -1, device=self.device, dtype=torch.long),
        }
        self._backing = None
        self._reserve(max(1, int(reserve_voxels)), 1)
        # persistent zero-invariant scratch (include/difusion_b200.h, "Persistent per-map scratch")
        self._grid_count = torch.zeros((G,), dtype=torch.int32, device=self.device)
        self._grid_bits = torch.zeros(((G + 31) // 32 + 1,), dtype=torch.int32, device=self.device)
        self._n_new = torch.zeros((1,), dtype=torch.int32, device=self.device)
        self._stats = torch.zeros((4,), dtype=torch.int32, device=self.device)
        self._ws = None
        self.mesh_cache = MeshExtractCache(self)
        self.last_integrate_stats = None

        p = MapParams()
        p.nx, p.ny, p.nz = self.n_xyz
        bm = [float(np.float32(v)) for v in args.bound_min]
        p.bound_min[0], p.bound_min[1], p.bound_min[2] = bm
        p.voxel_size = float(args.voxel_size)
        p.div_mode = int(div_mode)
        p.prune_min_vox_obs = int(args.prune_min_vox_obs)
        p.ignore_count_th = float(args.ignore_count_th)
        p.encoder_count_th = float(args.encoder_count_th)
        self._params = p

    # ------------------------------------------------------------------------------------------ plumbing
    @staticmethod
    def _weights_of(model):
        if isinstance(model, dict):
            return model
        out = {}
        for k, v in model.decoder.state_dict().items():
            out["dec." + k] = v.detach().float().cpu()
        for k, v in model.encoder.state_dict().items():
            out["enc." + k] = v.detach().cpu()
        return out

    _PER_VOXEL = ("latent_vecs", "latent_vecs_pos", "voxel_obs_count", "voxel_optimized")

    def _reserve(self, rows, cap):
        """(Re)allocate the backing storage with `rows` rows, keep the first min(old capacity, cap) rows of the current
        per-voxel tensors, and point cold_vars at views of the first `cap` rows.  New rows are zero / -1 (map.py:263-285)."""
        rows = max(rows, cap)
        new = {
            "latent_vecs": torch.zeros((rows, self.latent_dim), dtype=torch.float32, device=self.device),
            "latent_vecs_pos": torch.full((rows,), -1, dtype=torch.long, device=self.device),
            "voxel_obs_count": torch.zeros((rows,), dtype=torch.float32, device=self.device),
            "voxel_optimized": torch.zeros((rows,), dtype=torch.bool, device=self.device),
        }
        flag = torch.zeros((rows,), dtype=torch.uint8, device=self.device)
        for k in self._PER_VOXEL:
            old = self.cold_vars.get(k)
            if old is not None:
                keep = min(old.size(0), cap)
                new[k][:keep] = old[:keep].to(self.device)
        old_flag = getattr(self, "_updated_flag", None)
        if old_flag is not None:
            keep = min(old_flag.numel(), rows)
            flag[:keep] = old_flag[:keep]
        self._backing = new
        self._rows = rows
        # per-slot scratch of the integrate kernels (zero-invariant between calls) and the meshing dirty flags
        self._acc = torch.zeros((rows, self.latent_dim), dtype=torch.int64, device=self.device)      # fixed-point sums (2^-34 units)
        self._acc_n = torch.zeros((rows,), dtype=torch.int32, device=self.device)
        self._touched = torch.zeros((rows,), dtype=torch.int32, device=self.device)
        self._updated_flag = flag
        self._view(cap)

    def _view(self, cap):
        for k in self._PER_VOXEL:
            self.cold_vars[k] = self._backing[k][:cap]

    def _is_backed(self):
        return all(self.cold_vars[k].data_ptr() == self._backing[k].data_ptr() for k in self._PER_VOXEL)

    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes * 1.5) + 1024, dtype=torch.uint8, device=self.device)
        return self._ws

    # cold vars as attributes, like map.py:216-220
    n_occupied = property(lambda s: s.cold_vars["n_occupied"], lambda s, v: s.cold_vars.__setitem__("n_occupied", v))
    indexer = property(lambda s: s.cold_vars["indexer"], lambda s, v: s.cold_vars.__setitem__("indexer", v))
    latent_vecs = property(lambda s: s.cold_vars["latent_vecs"], lambda s, v: s.cold_vars.__setitem__("latent_vecs", v))
    latent_vecs_pos = property(lambda s: s.cold_vars["latent_vecs_pos"], lambda s, v: s.cold_vars.__setitem__("latent_vecs_pos", v))
    voxel_obs_count = property(lambda s: s.cold_vars["voxel_obs_count"], lambda s, v: s.cold_vars.__setitem__("voxel_obs_count", v))
    voxel_optimized = property(lambda s: s.cold_vars["voxel_optimized"], lambda s, v: s.cold_vars.__setitem__("voxel_optimized", v))

    def save(self, path):
        """map.py:222-224.  The per-voxel views are cloned so the file holds `capacity` rows, not the reservation."""
        cv = {k: (v.clone() if k in self._PER_VOXEL else v) for k, v in self.cold_vars.items()}
        with Path(path).open("wb") as f:
            torch.save(cv, f)

    def load(self, path):
        with Path(path).open("rb") as f:
            cv = torch.load(f, map_location=self.device, weights_only=False)
        self.cold_vars = cv
        cap = self.latent_vecs.size(0)
        self._updated_flag = None
        self._reserve(max(self._rows, cap), cap)          # adopt the loaded tensors into the backing storage
        self._updated_flag[:self.n_occupied] = 1          # every loaded voxel is dirty for the meshing cache

    def _inflate_latent_buffer(self, count: int):
        """map.py:263-285: capacity doubles until it holds n_occupied + count; new rows are zero / -1."""
        target = self.n_occupied + count
        cap = self.latent_vecs.size(0)
        new = cap
        while new < target:
            new *= 2
        if not self._is_backed():                       # a caller replaced a cold_vars tensor: adopt it
            self._reserve(max(self._rows, new), cap)
        if new > self._rows:
            self._reserve(max(new, 2 * self._rows), cap)
        if new != cap:
            self._view(new)

    def _linearize_id(self, xyz):
        return xyz[:, 2] + self.n_xyz[-1] * xyz[:, 1] + (self.n_xyz[-1] * self.n_xyz[-2]) * xyz[:, 0]

    def _unlinearize_id(self, idx):
        return torch.stack([idx // (self.n_xyz[1] * self.n_xyz[2]), (idx // self.n_xyz[2]) % self.n_xyz[1], idx % self.n_xyz[2]], dim=-1)

    # ------------------------------------------------------------------------------------------ integrate
    def integrate_keyframe(self, surface_xyz: torch.Tensor, surface_normal: torch.Tensor, do_optimize: bool = False,
                           async_optimize: bool = False):
        """map.py:341-520 with do_optimize=False.  Returns unq_mask (N,) bool (None when prune_min_vox_obs <= 0)."""
        assert surface_xyz.device == surface_normal.device == self.device, \
            f"Device of map {self.device} and input observation {surface_xyz.device, surface_normal.device} must be the same."
        if async_optimize and do_optimize:
            raise NotImplementedError("asynchronous latent optimisation (a second process, map.py:29-79) is not provided; "
                                      "do_optimize=True runs synchronously like the reference's async_optimize=False")
        xyz = surface_xyz.detach().contiguous().float()
        nrm = surface_normal.detach().contiguous().float()
        n = xyz.size(0)
        mask = torch.empty((n,), dtype=torch.bool, device=self.device)
        if n == 0:
            return mask if self.args.prune_min_vox_obs > 0 else None
        lib = self.lib
        ws = self._workspace(lib.dfb_integrate_ws_bytes(n, self.n_cells))
        with torch.cuda.device(self.device):
            st = _stream()
            check(lib.dfb_integrate_plan(C.byref(self._params), _p(xyz), _p(nrm), n, _p(self.indexer), _p(self._grid_count),
                                         _p(self._grid_bits), _p(mask), _p(self._n_new), _p(ws), ws.numel(), st))
            n_new = int(self._n_new.item())                      # the one host read of the call
            self._inflate_latent_buffer(n_new)
            check(lib.dfb_integrate_commit(C.byref(self._params), _p(nrm), _p(mask), n, _p(self.indexer), _p(self.latent_vecs),
                                           _p(self.latent_vecs_pos), _p(self.voxel_obs_count), _p(self._updated_flag),
                                           self.n_occupied, self.latent_vecs.size(0), n_new, _p(self._grid_bits), _p(self._acc),
                                           _p(self._acc_n), _p(self._touched), _p(self.encoder_blob), _p(self._stats), _p(ws),
                                           ws.numel(), st))
        self.n_occupied = self.n_occupied + n_new
        self.last_integrate_stats = self._stats
        if do_optimize and getattr(self.args, "optim_n_iters", 0) > 0:
            self._optimize_latents(xyz, nrm, mask if self.args.prune_min_vox_obs > 0 else None)
        return mask if self.args.prune_min_vox_obs > 0 else None

    def _optimize_latents(self, xyz, nrm, unq_mask):
        """map.py:456-517 with async_optimize=False: voxels whose confidence reached encoder_count_th and that were not optimised
        yet get their latents refined by `optim_n_iters` Adam steps on perturbed surface samples (do_optimize, map.py:81-113).
        The sample gathering is the reference's torch code line for line (so that, with the same torch seed, the random
        offsets along the normals are the same numbers); decoder forward / reverse pass and Adam are dfb_latent_adam_step."""
        args = self.args
        n_occ = self.n_occupied
        sel = torch.logical_and(self.voxel_obs_count[:n_occ] >= args.encoder_count_th, ~self.voxel_optimized[:n_occ])
        optim_voxel_pos = self.latent_vecs_pos[:n_occ][sel]
        optim_voxel_pos = optim_voxel_pos[optim_voxel_pos > 0]                  # (sic: map.py:467 drops voxel id 0)
        if optim_voxel_pos.size(0) == 0:
            return
        vs = self.voxel_size
        zeroed = xyz - self.bound_min.unsqueeze(0)
        # tensor / tensor is an IEEE divide on CUDA, tensor / python-float multiplies by the reciprocal (what the reference executes)
        xn = zeroed / torch.tensor(vs, dtype=torch.float32, device=self.device) if self.div_mode == DIV_IEEE else zeroed / vs
        gid = self._linearize_id(torch.ceil(xn).long() - 1)
        if unq_mask is not None:
            xn, gid, nrm = xn[unq_mask], gid[unq_mask], nrm[unq_mask]
        map_status = torch.zeros((self.n_cells,), device=self.device, dtype=torch.short)
        map_status[optim_voxel_pos] |= 1
        exp_indexer = torch.zeros((self.n_cells,), device=self.device, dtype=torch.long)
        exp_indexer[self._expand_flatten_id(optim_voxel_pos, False)] = 1
        focus = exp_indexer[gid] == 1
        pxn, pn = xn[focus], nrm[focus]
        inds, rels, sdfs = [], [], []
        n_xyz = self.n_xyz
        for off in ([-0.5, -0.5, -0.5], [-0.5, -0.5, 0.5], [-0.5, 0.5, -0.5], [-0.5, 0.5, 0.5],
                    [0.5, -0.5, -0.5], [0.5, -0.5, 0.5], [0.5, 0.5, -0.5], [0.5, 0.5, 0.5]):
            g = torch.ceil(pxn + torch.tensor(off, device=self.device, dtype=torch.float32)) - 1
            for dim in range(3):
                g[:, dim].clamp_(0, n_xyz[dim] - 1)
            rel = pxn - g - 0.5
            lin = self._linearize_id(g.long())
            ok = map_status[lin] >= 1
            cur_rel, cur_n = rel[ok], pn[ok]
            cur_sdf = torch.randn(cur_rel.size(0), device=self.device, dtype=torch.float32) * 0.05
            inds.append(self.indexer[lin][ok]); rels.append(cur_rel + cur_sdf.unsqueeze(-1) * cur_n); sdfs.append(cur_sdf)
        inds, rels, sdfs = torch.cat(inds), torch.cat(rels).contiguous(), torch.cat(sdfs).contiguous()
        if inds.numel() == 0:
            return
        uniq, inv = torch.unique(inds, return_inverse=True)
        lat = self.latent_vecs[uniq].contiguous()
        self.latent_vecs[uniq] = self.optimize_latents(lat, inv.contiguous(), sdfs, rels)
        self._updated_flag[uniq] = 1                                           # _mark_updated_vec_id (map.py:336)
        self.voxel_optimized[uniq] = True

    def optimize_latents(self, latent_vecs_unique, latent_id_inv_mapping, gathered_sdf, gathered_relative_xyz):
        """OptimizeProcess.do_optimize (map.py:81-113): Adam(lr 1e-2) for args.optim_n_iters iterations on the Gaussian
        log-likelihood of the clamped SDF targets (+ the optional code regulariser); returns the optimised latents."""
        args = self.args
        lat = latent_vecs_unique.detach().clone().contiguous().float()
        n_rows, n = lat.size(0), int(latent_id_inv_mapping.size(0))
        grad, m1, m2 = torch.zeros_like(lat), torch.zeros_like(lat), torch.zeros_like(lat)
        chunks = max(1, -(-n // int(1.5e6)))                                   # forward_model's max_sample: the regulariser is added per chunk
        reg = float(args.code_reg_lambda) * chunks / n if getattr(args, "code_regularization", False) else 0.0
        inv = latent_id_inv_mapping.contiguous().long(); xyz = gathered_relative_xyz.contiguous().float(); gt = gathered_sdf.contiguous().float()
        with torch.cuda.device(self.device):
            for it in range(int(args.optim_n_iters)):
                check(self.lib.dfb_latent_adam_step(_p(lat), n_rows, _p(inv), _p(xyz), _p(gt), n, _p(self.decoder_blob), _p(grad), _p(m1),
                                                    _p(m2), it + 1, 1.0e-2, reg, _stream()))
        return lat

    # ------------------------------------------------------------------------------------------ query
    def _get_sdf_raw(self, xyz, g_sdf, g_std):
        xyz = xyz.contiguous().float()
        n = xyz.size(0)
        with torch.cuda.device(self.device):
            if g_sdf is None:
                sdf = torch.empty((n,), dtype=torch.float32, device=self.device)
                std = torch.empty((n,), dtype=torch.float32, device=self.device)
                valid = torch.empty((n,), dtype=torch.bool, device=self.device)
                check(self.lib.dfb_get_sdf(C.byref(self._params), _p(xyz), n, _p(self.indexer), _p(self.latent_vecs),
                                           _p(self.voxel_obs_count), _p(self.decoder_blob), _p(sdf), _p(std), _p(valid),
                                           None, None, None, _stream()))
                return sdf, std, valid
            grad = torch.empty((n, 3), dtype=torch.float32, device=self.device)
            valid = torch.empty((n,), dtype=torch.bool, device=self.device)
            check(self.lib.dfb_get_sdf(C.byref(self._params), _p(xyz), n, _p(self.indexer), _p(self.latent_vecs),
                                       _p(self.voxel_obs_count), _p(self.decoder_blob), None, None, _p(valid),
                                       _p(g_sdf), _p(g_std), _p(grad), _stream()))
            return grad

    def get_sdf(self, xyz: torch.Tensor):
        """map.py:560-580.  Returns sdf (M,), std (M,), valid_mask (N,) with M valid rows; differentiable w.r.t. xyz."""
        if xyz.requires_grad:
            sdf, std, valid = _GetSdfFn.apply(xyz, self)
        else:
            sdf, std, valid = self._get_sdf_raw(xyz, None, None)
        return sdf[valid], std[valid], valid

    # ------------------------------------------------------------------------------------------ meshing
    def _expand_flatten_id(self, base_flatten_id, ensure_valid=True):
        """map.py:546-558 (host-side glue: a few thousand ids per call)."""
        out = [base_flatten_id]
        pos = self._unlinearize_id(base_flatten_id)
        for off in ([-1, 0, 0], [1, 0, 0], [0, -1, 0], [0, 1, 0], [0, 0, -1], [0, 0, 1]):
            q = pos + torch.tensor([off], device=self.device)
            for d in range(3):
                q[:, d].clamp_(0, self.n_xyz[d] - 1)
            q = self._linearize_id(q)
            if ensure_valid:
                q = q[self.indexer[q] != -1]
            out.append(q)
        return torch.unique(torch.cat(out))

    def decode_cubes(self, occupied_vec_id, voxel_resolution, refine_band=0.05):
        """map.py:637-688 (fast=True).  occupied_vec_id (B,) int64 slots -> cube_sdf, cube_std (B,2r,2r,2r), sdf negated."""
        B = occupied_vec_id.size(0)
        r = int(voxel_resolution)
        R = 2 * r
        cube_sdf = torch.empty((B, R, R, R), dtype=torch.float32, device=self.device)
        cube_std = torch.empty((B, R, R, R), dtype=torch.float32, device=self.device)
        occ = occupied_vec_id.contiguous()
        ws = ext._WS.get(self.device, self.lib.dfb_decode_cubes_ws_bytes(B, r))
        with torch.cuda.device(self.device):
            check(self.lib.dfb_decode_cubes(_p(self.latent_vecs), _p(occ), B, r, float(refine_band), _p(self.decoder_blob),
                                            _p(cube_sdf), _p(cube_std), _p(ws), ws.numel(), _stream()))
        return cube_sdf, cube_std

    def extract_mesh(self, voxel_resolution: int, max_n_triangles: int, fast: bool = True, max_std: float = 2000.0,
                     extract_async: bool = False, no_cache: bool = False, interpolate: bool = True):
        """map.py:582-724.  Synchronous only; `fast=False` / `interpolate=False` are not on the hot path (the reference's
        interpolate=False branch calls an unexported function, map.py:694)."""
        if not fast or not interpolate:
            raise NotImplementedError("only fast=True, interpolate=True is supported (the path main.py uses)")
        updated = self.mesh_cache.updated_vec_id
        if updated.size(0) == 0 and not no_cache:
            return self._make_mesh_from_cache()
        if no_cache:
            updated_vec_id = torch.arange(self.n_occupied, device=self.device)
            self.mesh_cache.clear_all()
        else:
            updated_vec_id = updated
            self.mesh_cache.clear_updated_vec()
        if updated_vec_id.size(0) == 0:
            return self._make_mesh_from_cache()

        focused_flatten_id = self.latent_vecs_pos[updated_vec_id]
        occupied_flatten_id = self._expand_flatten_id(focused_flatten_id)
        occupied_vec_id = self.indexer[occupied_flatten_id]
        occupied_vec_id = occupied_vec_id[self.voxel_obs_count[occupied_vec_id] > self.args.ignore_count_th]
        if occupied_vec_id.size(0) == 0:
            return self._make_mesh_from_cache()
        mapping = torch.full((int(occupied_vec_id.max().item()) + 1,), -1, device=self.device, dtype=torch.int)
        mapping[occupied_vec_id] = torch.arange(0, occupied_vec_id.size(0), device=self.device, dtype=torch.int)
        cube_sdf, cube_std = self.decode_cubes(occupied_vec_id, voxel_resolution)
        vertices, vertices_flatten_id, vertices_std = ext.marching_cubes_interp(
            self.indexer.view(self.n_xyz), focused_flatten_id.contiguous(), mapping, cube_sdf, cube_std, max_n_triangles,
            self.n_xyz, max_std)
        vertices = vertices * self.voxel_size + self.bound_min
        vertices = vertices.cpu().numpy()
        vertices_std = vertices_std.cpu().numpy()
        vertices_flatten_id = vertices_flatten_id.cpu().numpy()
        mc = self.mesh_cache
        if mc.vertices is None:
            mc.vertices, mc.vertices_flatten_id, mc.vertices_std = vertices, vertices_flatten_id, vertices_std
        else:
            keep = ~np.isin(mc.vertices_flatten_id, np.unique(vertices_flatten_id))       # map.py:709-715
            mc.vertices = np.concatenate([mc.vertices[keep], vertices], axis=0)
            mc.vertices_flatten_id = np.concatenate([mc.vertices_flatten_id[keep], vertices_flatten_id], axis=0)
            mc.vertices_std = np.concatenate([mc.vertices_std[keep], vertices_std], axis=0)
        return self._make_mesh_from_cache()

    def get_fast_preview_visuals(self):
        """map.py:726-750 (the GUI's block wireframe, main.py:89 under --vis): one cube outline per allocated voxel + the map's
        bounding box.  Returns [blocks, bbox]: open3d LineSets when open3d is installed, else SimpleLineSet(points, lines, colors)."""
        occupied = torch.where(self.indexer != -1)[0]
        base = self._unlinearize_id(occupied).float() * self.voxel_size + self.bound_min
        vs = self.voxel_size
        offs = [[0.0, 0.0, 0.0], [0.0, 0.0, vs], [0.0, vs, 0.0], [0.0, vs, vs], [vs, 0.0, 0.0], [vs, 0.0, vs], [vs, vs, 0.0], [vs, vs, vs]]
        verts = torch.cat([base + torch.tensor(o, dtype=torch.float32, device=base.device).unsqueeze(0) for o in offs], 0)
        n = base.size(0)
        ar = np.arange(n, dtype=np.int32)
        edges = np.concatenate([np.stack([ar + a * n, ar + b * n], 1) for a, b in
                                ([0, 1], [0, 2], [0, 4], [1, 3], [1, 5], [2, 3], [2, 6], [3, 7], [4, 5], [4, 6], [5, 7], [6, 7])], 0)
        verts = verts.cpu().numpy().astype(float)
        lo, hi = self.bound_min.cpu().numpy().astype(float), self.bound_max.cpu().numpy().astype(float)
        bb_pts = np.asarray([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
        bb_lines = np.asarray([[0, 1], [2, 3], [4, 5], [6, 7], [0, 4], [1, 5], [2, 6], [3, 7], [0, 2], [4, 6], [1, 3], [5, 7]], np.int32)
        bb_col = np.repeat(np.asarray([[0.5804, 0.4039, 0.7412]]), 12, 0)          # matplotlib tab10[4] (vis_util.py:120, color_id=4)
        try:
            import open3d as o3d
            if not hasattr(o3d.geometry, "LineSet"):
                raise ImportError("open3d.geometry.LineSet")
            blk = o3d.geometry.LineSet(points=o3d.utility.Vector3dVector(verts), lines=o3d.utility.Vector2iVector(edges))
            bb = o3d.geometry.LineSet(points=o3d.utility.Vector3dVector(bb_pts), lines=o3d.utility.Vector2iVector(bb_lines))
            bb.colors = o3d.utility.Vector3dVector(bb_col)
            return [blk, bb]
        except ImportError:
            return [SimpleLineSet(verts, edges, None), SimpleLineSet(bb_pts, bb_lines, bb_col)]

    def _make_mesh_from_cache(self):
        """map.py:522-544."""
        mc = self.mesh_cache
        if mc.vertices is None:
            return SimpleMesh(np.zeros((0, 3)), np.zeros((0, 3), np.int32))
        vertices = mc.vertices.reshape((-1, 3))
        triangles = np.arange(vertices.shape[0]).reshape((-1, 3))
        std = mc.vertices_std.reshape((-1,)).astype(float)
        colors = None
        if vertices.shape[0] > 0:
            if self.extract_mesh_std_range is not None:
                lo, hi = self.extract_mesh_std_range
                c = np.clip(std, lo, hi)
            else:
                lo, hi = std.min(), std.max()
                c = std
            c = (c - lo) / max(hi - lo, 1e-12)
            colors = np.stack([np.clip(1.5 - np.abs(4 * c - 3), 0, 1), np.clip(1.5 - np.abs(4 * c - 2), 0, 1),
                               np.clip(1.5 - np.abs(4 * c - 1), 0, 1)], axis=1)                    # jet colour map
        try:
            import open3d as o3d
            if not hasattr(o3d.geometry, "TriangleMesh"):          # a stub / partial install: fall back to SimpleMesh
                raise ImportError("open3d.geometry.TriangleMesh")
            mesh = o3d.geometry.TriangleMesh()
            mesh.vertices = o3d.utility.Vector3dVector(vertices.astype(float))
            mesh.triangles = o3d.utility.Vector3iVector(triangles.astype(np.int32))
            if colors is not None:
                mesh.vertex_colors = o3d.utility.Vector3dVector(colors)
            return mesh
        except ImportError:
            return SimpleMesh(vertices.astype(float), triangles.astype(np.int32), colors, std)
