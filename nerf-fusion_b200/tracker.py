"""SDFTracker: host-side mirror of system/tracker.py (preprocessing + Gauss-Newton over SDF and photometric terms).

Same class/method names, arguments and return conventions as the reference.  What changed underneath:
  * compute_sdf_Hg (tracker.py:179-223): ONE kernel (transform, map lookup, decoder forward + reverse pass, Jacobian,
    robust weights, J^T J / J^T r / energy reduction) and ONE 352-byte device->host read, instead of ~40 torch ops,
    an autograd tape, an (M,6,6) einsum tensor and 3 syncs;
  * compute_rgb_Hg (tracker.py:136-177): the per-pixel residual/Jacobian kernel reduces in-kernel as well;
  * preprocessing (tracker.py:89-120): unproject / radius filter / normals / box filter are the C-ABI ops.
The 6x6 solve, pose composition and the accept/rollback logic stay on the host in float64, as in the reference.
"""
import copy

import numpy as np
import torch

from . import ext
from ._lib import check, fptr
from .ext import _p, _stream
from .motion import Isometry

import ctypes as C


class FrameIntrinsic:
    """dataset/production/__init__.py:4-18."""

    def __init__(self, fx, fy, cx, cy, dscale=5000.0):
        self.cx, self.cy, self.fx, self.fy, self.dscale = cx, cy, fx, fy, dscale

    def to_K(self):
        return np.asarray([[self.fx, 0.0, self.cx], [0.0, self.fy, self.cy], [0.0, 0.0, 1.0]])


def _args_of(d):
    import argparse
    if isinstance(d, dict):
        ns = argparse.Namespace()
        ns.__dict__.update(d)
        return ns
    return d


def point_box_filter(points: torch.Tensor, normals: torch.Tensor, voxel_size: float, div_mode: int = 0):
    """tracker.py:14-24 (unique + 2x scatter_mean fused into one deterministic segmented mean)."""
    return ext.point_box_filter(points.contiguous(), normals.contiguous(), voxel_size, div_mode)


_ROBUST = {None: 0, "huber": 1, "tukey": 2}


class SDFTracker:
    def __init__(self, map, args):
        self.map = map
        self.args = _args_of(args)
        self.sdf_args = _args_of(self.args.sdf)
        self.rgb_args = _args_of(self.args.rgb)
        self.last_intensity = None
        self.last_depth = None
        self.all_pd_pose = []
        self.last_processed_pc = None
        self.cur_gt_pose = None
        self.last_colored_pcd = None
        self.n_unstable = 0
        dev = map.device
        self._hg_dev = torch.zeros((80,), dtype=torch.float64, device=dev)
        self._hg_host = torch.zeros((80,), dtype=torch.float64).pin_memory()
        self.n_sdf_evals = 0
        self.n_rgb_evals = 0
        # True: the whole Gauss-Newton solve of a frame is one C call (dfb_gauss_newton); False: the Python loop below
        # (same control flow, kept for A/B tests and as executable documentation of tracker.py:225-288)
        self.native_gn = True
        self.fused_preprocess = True       # one C call for tracker.py:89-120 (False: the op-by-op path, same results)
        self.fused_images = True           # intensity / depth cut / pyramids / gradients in two launches (False: the torch ops, same bits)
        # True: the per-frame front end (intensity, image pyramid, gradients, fused preprocessing: ~55 launches, no host
        # decisions) is captured once per (image size, intrinsics) into a CUDA graph and replayed from static buffers
        self.graph_frontend = True
        self._fe_graphs = {}               # key -> dict(graph, static inputs/outputs)
        self._fe_ws = {}                   # depth shape -> workspace owned by the front end (its address is baked into the graphs)
        self._fe_seen = {}                 # key -> eager calls so far (the first call warms kernels and workspaces up)
        self._fe_busy = {}                 # key -> graph set whose static outputs last_intensity / last_depth still alias
        self._fe_choice = None             # (key, set) the last graphed front-end call replayed
        self._fe_cur = {}                  # key -> graph set the frame being tracked right now uses
        # Frame pipelining: track_camera(..., next_frame=(rgb, depth)) / prefetch_frame() enqueue the NEXT frame's front end on
        # a side stream (when: see prefetch_mode below), so that the GPU does not idle while the host does a frame's bookkeeping.
        # Three graph sets rotate: last committed frame (read as last_*), current frame, prefetched frame.
        self._pf = None                    # pending prefetch: dict(rgb, depth, base, idx, ent, event)
        self._pf_stream = None
        # The pose solve is the critical path of a frame; with a prefetched front end running beside it, it is launched on a
        # HIGH-priority stream so that a waiting solve CTA (a whole SM each) is placed before the front end's small blocks,
        # which then fill only what the solve leaves idle.  DFB_SOLVE_PRIO=0 keeps the solve on the caller's stream.
        import os
        self._solve_prio = int(os.environ.get("DFB_SOLVE_PRIO", "-1"))
        # Optionally the evaluation kernel (one whole-SM CTA per SM: nothing else fits beside it) leaves some SMs to the front
        # end.  Measured (tools/pipeline_timeline.py): without reserved SMs the two time-slice (front end 320 -> 900 us, solve
        # 700 -> 850 us, frame 1.19 -> 1.13 ms); reserving 16..48 SMs shortens the front end to ~650 us but lengthens the solve
        # more (the long normals blocks still land on the solve's SMs between evaluations): default 0.
        self._reserve_sms = int(os.environ.get("DFB_GN_RESERVE_SMS", "0"))
        # "after" (default): the next frame's front end is queued the moment this frame's solve has returned, BEFORE the host does
        # its bookkeeping for the frame (pose algebra, copies of the cloud, a keyframe's integration): the GPU never idles between
        # frames and nothing competes with the solve for SMs.  "before": queued ahead of the solve, runs beside it (measured: the
        # two time-slice rather than overlap, see above).
        self.prefetch_mode = os.environ.get("DFB_PREFETCH", "after")
        self._pf_deferred = None
        self._solve_stream = None
        self.time_kernels = False          # bench.py: CUDA-event timing of the SDF-term launches inside the C driver
        self.sdf_kernel_us = 0; self.sdf_queries_J = 0; self.sdf_queries_noJ = 0
        self._gn_pinned = torch.zeros((64,), dtype=torch.float64).pin_memory()       # DFB_GN_PINNED_DOUBLES
        self._gn_dev = torch.zeros((160,), dtype=torch.float64, device=dev)          # DFB_GN_SCRATCH_DOUBLES

    # ------------------------------------------------------------------------------------------ preprocessing
    def _make_image_pyramid(self, intensity_img, depth_img):
        """tracker.py:42-57 (resampling is torch plumbing; gradients are the C-ABI op)."""
        F = torch.nn.functional
        d0_w, d0_h = intensity_img.size(1), intensity_img.size(0)
        d1_w, d1_h = d0_w // 2, d0_h // 2
        d2_w, d2_h = d1_w // 2, d1_h // 2
        d0_i = intensity_img.view(1, 1, d0_h, d0_w)
        d0_d = depth_img.view(1, 1, d0_h, d0_w)
        d1_i = F.interpolate(d0_i, (d1_h, d1_w), mode="bilinear", align_corners=True)
        d1_d = F.interpolate(d0_d, (d1_h, d1_w), mode="nearest")
        d2_i = F.interpolate(d1_i, (d2_h, d2_w), mode="bilinear", align_corners=True)
        d2_d = F.interpolate(d1_d, (d2_h, d2_w), mode="nearest")
        Is = [t.squeeze(0).squeeze(0).contiguous() for t in (d0_i, d1_i, d2_i)]
        Ds = [t.squeeze(0).squeeze(0).contiguous() for t in (d0_d, d1_d, d2_d)]
        return Is, Ds, [ext.gradient_xy(t) for t in Is]

    def preprocess_depth(self, depth_data, calib):
        """tracker.py:89-120 (geometry half): returns (points (N,3), normals (N,3)) in camera space."""
        pc_scale = self.sdf_args.subsample
        if self.fused_preprocess and pc_scale == 0.5:
            return ext.preprocess_frame(depth_data.contiguous(), calib.fx, calib.fy, calib.cx, calib.cy, 16, 0.05, 16, 0.1,
                                        (0.0, 0.0, 0.0), 0.02, self.map.div_mode)
        pc_data = torch.nn.functional.interpolate(depth_data.unsqueeze(0).unsqueeze(0), scale_factor=pc_scale, mode="nearest",
                                                  recompute_scale_factor=False).squeeze(0).squeeze(0).contiguous()
        pc_data = ext.unproject_depth(pc_data, calib.fx * pc_scale, calib.fy * pc_scale, calib.cx * pc_scale, calib.cy * pc_scale)
        pc_data = torch.cat([pc_data, torch.zeros((pc_data.size(0), pc_data.size(1), 1), device=pc_data.device)], dim=-1)
        pc_data = pc_data.reshape(-1, 4)
        pc_data = pc_data[~torch.isnan(pc_data[..., 0])].contiguous()
        with torch.cuda.device(self.map.device):
            pc_data = pc_data[ext.remove_radius_outlier(pc_data, 16, 0.05)].contiguous()
            normal_data = ext.estimate_normals(pc_data, 16, 0.1, [0.0, 0.0, 0.0])
            ok = ~torch.isnan(normal_data[..., 0])
            normal_data = normal_data[ok].contiguous()
            pc_data = pc_data[ok, :3].contiguous()
        return point_box_filter(pc_data, normal_data, 0.02, self.map.div_mode)

    def _frontend(self, rgb_data, depth_data, calib, depth_cut=None):
        """Everything of tracker.py:75-120 that needs no host decision: intensity, pyramids, gradients, preprocessing
        (and, when depth_cut = (near, far) is given, the caller's depth clipping of main.py:56-57).
        Returns (Is, Ds, Gs, points (n_max,3), normals (n_max,3), count i32[1]); no host sync."""
        if self.fused_images and depth_data.size(0) >= 8 and depth_data.size(1) >= 8:
            Is, Ds, Gs = ext.frame_images(rgb_data, depth_data, depth_cut)        # 2 launches instead of ~13 torch kernels, same bits
        else:
            if depth_cut is not None:
                depth_data = torch.where((depth_data < depth_cut[0]) | (depth_data > depth_cut[1]),
                                         torch.full_like(depth_data, float("nan")), depth_data)
            cur_intensity = torch.mean(rgb_data, dim=-1)
            Is, Ds, Gs = self._make_image_pyramid(cur_intensity, depth_data)
        # the front end owns its workspace: the captured graphs keep its address (ext.preprocess_frame, `ws`)
        shape = tuple(Ds[0].shape)
        ws = self._fe_ws.get(shape)
        if ws is None:
            ws = torch.empty(int(self.map.lib.dfb_preprocess_ws_bytes(shape[0], shape[1])) + 1024, dtype=torch.uint8, device=self.map.device)
            self._fe_ws[shape] = ws
        out_p, out_n, cnt = ext.preprocess_frame(Ds[0].contiguous(), calib.fx, calib.fy, calib.cx, calib.cy, 16, 0.05, 16, 0.1,
                                                 (0.0, 0.0, 0.0), 0.02, self.map.div_mode, sync=False, ws=ws)
        return Is, Ds, Gs, out_p, out_n, cnt

    N_FE_SETS = 3

    def _fe_base(self, rgb_data, depth_data, calib, depth_cut):
        return (tuple(rgb_data.shape), tuple(depth_data.shape), calib.fx, calib.fy, calib.cx, calib.cy, self.map.div_mode,
                None if depth_cut is None else tuple(depth_cut))

    def _fe_entry(self, base, idx, rgb_data, depth_data, calib, depth_cut):
        """Graph set `idx` of key `base`, captured on first use (static inputs / outputs + the graph)."""
        key = base + (idx,)
        ent = self._fe_graphs.get(key)
        if ent is None:
            ent = dict(rgb=torch.empty_like(rgb_data), depth=torch.empty_like(depth_data))
            ent["rgb"].copy_(rgb_data); ent["depth"].copy_(depth_data)
            cur = torch.cuda.current_stream(self.map.device)
            side = torch.cuda.Stream(self.map.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                self._frontend(ent["rgb"], ent["depth"], calib, depth_cut)
            cur.wait_stream(side)
            from . import _lib
            before = dict(_lib.CALLS)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                ent["out"] = self._frontend(ent["rgb"], ent["depth"], calib, depth_cut)
            ent["graph"] = graph
            ent["calls"] = {k: v - before.get(k, 0) for k, v in _lib.CALLS.items() if v != before.get(k, 0)}   # C calls inside one replay
            self._fe_graphs[key] = ent
        return ent

    def _fe_replay(self, ent, rgb_data, depth_data):
        """Copies the frame into the set's static inputs and replays its graph on the current stream."""
        ent["rgb"].copy_(rgb_data); ent["depth"].copy_(depth_data)
        ent["graph"].replay()
        from . import _lib
        for k, v in ent["calls"].items():                         # the replay launches the kernels of these calls again
            _lib.CALLS[k] = _lib.CALLS.get(k, 0) + v
        return ent["out"]

    def _frontend_graphed(self, rgb_data, depth_data, calib, depth_cut=None):
        """Replays the captured front end on copies of the inputs.  The first call per key runs eagerly (loads kernels,
        sizes workspaces); later calls capture / replay.  Several graph sets per key exist because their outputs are static
        buffers and the pyramids of the last COMMITTED frame are still read as `last_*` while the next frame is processed:
        a call replays a set `last_*` does not alias (track_camera marks a set busy when it commits its pyramids), so
        calls that commit nothing (for_pc=True, a failed solve) can never make a frame's photometric term read itself.
        A frame whose front end was prefetched (prefetch_frame) only waits for the side stream."""
        base = self._fe_base(rgb_data, depth_data, calib, depth_cut)
        self._fe_choice = None
        pf, self._pf = self._pf, None
        if pf is not None:
            torch.cuda.current_stream(self.map.device).wait_event(pf["event"])     # also orders the shared workspace
            if pf["rgb"] is rgb_data and pf["depth"] is depth_data and pf["base"] == base:
                self._fe_seen[base] = self._fe_seen.get(base, 0) + 1
                self._fe_choice = (base, pf["idx"])
                self._fe_cur[base] = pf["idx"]
                return pf["ent"]["out"]
        seen = self._fe_seen.get(base, 0)
        self._fe_seen[base] = seen + 1
        if seen == 0:
            self._fe_cur.pop(base, None)
            return self._frontend(rgb_data, depth_data, calib, depth_cut)
        idx = 1 if self._fe_busy.get(base) == 0 else 0
        self._fe_choice = (base, idx)
        self._fe_cur[base] = idx
        return self._fe_replay(self._fe_entry(base, idx, rgb_data, depth_data, calib, depth_cut), rgb_data, depth_data)

    def _launch_deferred_prefetch(self):
        d, self._pf_deferred = self._pf_deferred, None
        if d is not None:
            self.prefetch_frame(*d)

    @property
    def prefetch_stream(self):
        """The side stream prefetched front ends run on (callers that produce the next frame on the device, e.g. an ingest
        kernel, can enqueue that work here so that it is ordered before the prefetch without touching the main stream)."""
        if self._pf_stream is None:
            self._pf_stream = torch.cuda.Stream(self.map.device)
        return self._pf_stream

    def prefetch_frame(self, rgb_data, depth_data, calib, depth_cut=None):
        """Enqueues the front end (tracker.py:84-120: intensity, pyramids, gradients, preprocessing) of the frame that will be
        tracked NEXT on the side stream and returns at once; the following track_camera call with the same tensor objects
        only waits for it.  Results are bit-identical to the unprefetched path (same kernels, same inputs).  Returns False when
        nothing was queued (graph front end off, or this image size has not been seen yet)."""
        if not (self.graph_frontend and self.fused_preprocess and self.sdf_args.subsample == 0.5):
            return False
        if not (rgb_data.is_contiguous() and depth_data.is_contiguous()):
            return False
        base = self._fe_base(rgb_data, depth_data, calib, depth_cut)
        if self._fe_seen.get(base, 0) == 0:
            return False
        dev = self.map.device
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream(dev)
            side = self.prefetch_stream
            if self._pf is not None:                              # an unconsumed prefetch: its outputs are dropped
                self._pf = None
            busy, cur = self._fe_busy.get(base), self._fe_cur.get(base)
            idx = next(i for i in range(self.N_FE_SETS) if i != busy and i != cur)
            ent = self._fe_entry(base, idx, rgb_data, depth_data, calib, depth_cut)    # may capture (first use of the set)
            side.wait_stream(main)                                # inputs produced on the main stream; the set's last readers
            with torch.cuda.stream(side):
                self._fe_replay(ent, rgb_data, depth_data)
                ev = torch.cuda.Event()
                ev.record(side)
        self._pf = dict(rgb=rgb_data, depth=depth_data, base=base, idx=idx, ent=ent, event=ev)
        return True

    def track_camera(self, rgb_data, depth_data, calib, set_pose: Isometry = None, for_pc=False, depth_cut=None, next_frame=None):
        """tracker.py:75-134.  rgb (H,W,3) f32, depth (H,W) f32 with NaN = invalid.
        depth_cut = (near, far), optional: depths outside the range become NaN here (what main.py:56-57 does before the
        call), so the clipping is part of the captured front end instead of five eager launches.
        next_frame = (rgb, depth), optional: the frame that will be tracked next (tensors that stay untouched until the call that
        tracks them); its front end is queued on the side stream (prefetch_frame) the moment this frame's pose solve has returned
        (prefetch_mode "after", default) or ahead of the solve ("before")."""
        if self.fused_preprocess and self.sdf_args.subsample == 0.5:
            graphed = self.graph_frontend
            fe = self._frontend_graphed if graphed else self._frontend
            if not graphed and self._pf is not None:              # a prefetch queued before the graph front end was switched off
                torch.cuda.current_stream(self.map.device).wait_event(self._pf["event"])
                self._pf = None
            with torch.cuda.device(self.map.device):
                cur_intensity, cur_depth, cur_dIdxy, out_p, out_n, cnt = fe(rgb_data.contiguous(), depth_data.contiguous(), calib,
                                                                            depth_cut)
                self._pf_deferred = None
                if next_frame is not None and graphed:
                    if self.prefetch_mode == "before":
                        self.prefetch_frame(next_frame[0], next_frame[1], calib, depth_cut)
                    else:                                         # queued the moment the pose solve has returned (_gauss_newton_native)
                        self._pf_deferred = (next_frame[0], next_frame[1], calib, depth_cut)
            # The row count stays on the device while the pose solve is queued behind the front end (the kernels read it
            # there); it is read back once the solve has returned.
            defer = self.native_gn and set_pose is None and not for_pc and len(self.all_pd_pose) > 0
            final_pose = None
            if defer:
                try:
                    final_pose = self.gauss_newton(self.all_pd_pose[-1].dot(Isometry()), cur_intensity, cur_depth, cur_dIdxy, out_p, calib,
                                                   obs_count=cnt)
                finally:
                    m = int(cnt.item())
                    if m < 0:
                        raise RuntimeError("preprocess_frame: box-filter key range exceeds the bitmap capacity")
            else:
                m = int(cnt.item())                               # the one host read of the front end
                if m < 0:
                    raise RuntimeError("preprocess_frame: box-filter key range exceeds the bitmap capacity")
            pc_data, normal_data = out_p[:m], out_n[:m]
        else:
            graphed = False
            final_pose = None
            if depth_cut is not None:
                depth_data = torch.where((depth_data < depth_cut[0]) | (depth_data > depth_cut[1]),
                                         torch.full_like(depth_data, float("nan")), depth_data)
            cur_intensity = torch.mean(rgb_data, dim=-1)
            cur_intensity, cur_depth, cur_dIdxy = self._make_image_pyramid(cur_intensity, depth_data)
            pc_data, normal_data = self.preprocess_depth(cur_depth[0], calib)
        # (graph outputs are static buffers reused a few frames later: hand out copies)
        self.last_processed_pc = [pc_data.clone(), normal_data.clone()] if graphed else [pc_data, normal_data]
        if for_pc:
            self._launch_deferred_prefetch()
            return self.last_processed_pc
        if set_pose is not None:
            final_pose = set_pose
        elif final_pose is None:
            assert len(self.all_pd_pose) > 0
            final_pose = self.gauss_newton(self.all_pd_pose[-1].dot(Isometry()), cur_intensity, cur_depth, cur_dIdxy, pc_data, calib)
        self.last_intensity = cur_intensity
        self.last_depth = cur_depth
        if graphed:
            if self._fe_choice is not None:            # last_* now alias this set's static outputs
                self._fe_busy[self._fe_choice[0]] = self._fe_choice[1]
            else:                                      # eager outputs are fresh tensors: no set is aliased any more
                self._fe_busy.clear()
        self.all_pd_pose.append(final_pose)
        self._launch_deferred_prefetch()                   # (paths that did not go through the native solve)
        return final_pose

    # ------------------------------------------------------------------------------------------ GN terms
    def _read_hg(self, scale_of_count):
        """Copies the 44 result doubles to pinned host memory and unpacks (H, g, sum_wr2, count)."""
        with torch.cuda.device(self.map.device):
            self._hg_host.copy_(self._hg_dev, non_blocking=True)
            torch.cuda.current_stream(self.map.device).synchronize()
        v = self._hg_host.numpy()
        return v[:36].reshape(6, 6).copy(), v[36:42].copy(), float(v[42]), float(v[43])

    def compute_rgb_Hg(self, pyramid_level, cur_delta_pose, cur_intensity_pyramid, cur_depth_pyramid, cur_dIdxy_pyramid, calib,
                       no_grad=False):
        """tracker.py:136-177."""
        cur_R = cur_delta_pose.q.rotation_matrix
        cur_t = cur_delta_pose.t
        K = calib.to_K()
        KRKinv = K @ cur_R @ np.linalg.inv(K)
        Kt = K @ cur_t
        ext.rgb_hg(self.last_intensity[pyramid_level], self.last_depth[pyramid_level], cur_intensity_pyramid[pyramid_level],
                   cur_depth_pyramid[pyramid_level], cur_dIdxy_pyramid[pyramid_level], [calib.fx, calib.fy, calib.cx, calib.cy],
                   KRKinv.flatten().tolist(), Kt.flatten().tolist(), self.rgb_args.min_grad_scale, self.rgb_args.max_depth_delta,
                   _ROBUST[self.rgb_args.robust_kernel], self.rgb_args.robust_k, not no_grad, out=self._hg_dev)
        self.n_rgb_evals += 1
        H, g, e, cnt = self._read_hg(None)
        error_scale = 1. / cnt * self.rgb_args.weight if cnt > 0 else float("nan")
        if no_grad:
            return None, None, float(e * error_scale)
        return H * error_scale, g * error_scale, float(e * error_scale)

    def compute_sdf_Hg(self, n_iter, last_pose, cur_delta_pose, obs_xyz, no_grad=False):
        """tracker.py:179-223.  Returns (H (6,6) f64 | None, g (6,) f64 | None, energy float)."""
        m = self.map
        total = last_pose.dot(cur_delta_pose)
        pose = np.concatenate([total.q.rotation_matrix.reshape(-1), total.t, cur_delta_pose.q.rotation_matrix.reshape(-1),
                               cur_delta_pose.t, last_pose.q.rotation_matrix.reshape(-1)]).astype(np.float32)
        obs = obs_xyz.contiguous()
        with torch.cuda.device(m.device):
            check(m.lib.dfb_sdf_hg(C.byref(m._params), _p(obs), obs.size(0), fptr(pose.tolist()), _p(m.indexer), _p(m.latent_vecs),
                                   _p(m.voxel_obs_count), _p(m.decoder_blob), _ROBUST[self.sdf_args.robust_kernel],
                                   float(self.sdf_args.robust_k), int(not no_grad), _p(self._hg_dev), _stream()))
        self.n_sdf_evals += 1
        H, g, e, cnt = self._read_hg(None)
        error_scale = 1.0 / cnt if cnt > 0 else float("nan")
        if no_grad:
            return None, None, float(e * error_scale)
        return H * error_scale, g * error_scale, float(e * error_scale)

    # ------------------------------------------------------------------------------------------ Gauss-Newton
    def _gauss_newton_native(self, last_pose, delta_pose, Is, Ds, Gs, obs_xyz, calib, obs_count=None):
        from . import _lib
        from ._lib import GnConfig, RgbLevel
        m = self.map
        cfg = GnConfig()
        groups = self.args.iter_config
        if len(groups) > 8:
            raise ValueError("at most 8 iteration groups")
        cfg.n_groups = len(groups)
        need_rgb = False
        for i, gcfg in enumerate(groups):
            cfg.n_iter[i] = int(gcfg["n"])
            cfg.use_sdf[i] = 0
            cfg.rgb_level[i] = -1
            for term in gcfg["type"]:
                if term[0] == "sdf":
                    cfg.use_sdf[i] = 1
                elif term[0] == "rgb":
                    cfg.rgb_level[i] = int(term[1]); need_rgb = True
                else:
                    raise NotImplementedError(term[0])
        cfg.sdf_robust = _ROBUST[self.sdf_args.robust_kernel]; cfg.sdf_robust_k = float(self.sdf_args.robust_k)
        cfg.rgb_robust = _ROBUST[self.rgb_args.robust_kernel]; cfg.rgb_robust_k = float(self.rgb_args.robust_k)
        cfg.rgb_weight = float(self.rgb_args.weight); cfg.rgb_min_grad_scale = float(self.rgb_args.min_grad_scale)
        cfg.rgb_max_depth_delta = float(self.rgb_args.max_depth_delta)
        levels = (RgbLevel * 3)()
        if need_rgb:
            for l in range(3):
                levels[l].prev_I = self.last_intensity[l].data_ptr(); levels[l].prev_D = self.last_depth[l].data_ptr()
                levels[l].cur_I = Is[l].data_ptr(); levels[l].cur_D = Ds[l].data_ptr(); levels[l].cur_G = Gs[l].data_ptr()
                levels[l].H, levels[l].W = int(Is[l].size(0)), int(Is[l].size(1))
        intr = (C.c_double * 4)(calib.fx, calib.fy, calib.cx, calib.cy) if calib is not None else None
        lastp = (C.c_double * 12)(*last_pose.q._rot9(), *last_pose.t.tolist())
        deltap = (C.c_double * 12)(*delta_pose.q._rot9(), *delta_pose.t.tolist())
        stats = (C.c_int32 * 8)()
        if self.time_kernels:
            stats[4] = 0x54494d45
        obs = obs_xyz.contiguous()
        with torch.cuda.device(m.device):
            cur = torch.cuda.current_stream(m.device)
            solve = cur
            if self._solve_prio != 0 and self._pf is not None:    # a front end runs beside this solve: give the solve priority
                if self._solve_stream is None:
                    self._solve_stream = torch.cuda.Stream(m.device, priority=self._solve_prio)
                solve = self._solve_stream
                solve.wait_stream(cur)
            reserve = self._reserve_sms if self._pf is not None else 0
            try:
                m.lib.dfb_set_gn_reserved_sms(reserve)
                with torch.cuda.stream(solve):
                    check(m.lib.dfb_gauss_newton(C.byref(m._params), C.byref(cfg), _p(obs), obs.size(0),
                                                 _p(obs_count) if obs_count is not None else None, _p(m.indexer), _p(m.latent_vecs),
                                                 _p(m.voxel_obs_count), _p(m.decoder_blob), levels, intr, lastp, deltap, _p(self._gn_dev),
                                                 C.c_void_p(self._gn_pinned.data_ptr()), stats, _stream()))
            finally:
                m.lib.dfb_set_gn_reserved_sms(0)
                if solve is not cur:
                    cur.wait_stream(solve)
            self._launch_deferred_prefetch()
        self.n_sdf_evals += stats[1]; self.n_rgb_evals += stats[2]
        if self.time_kernels:
            self.sdf_kernel_us += stats[4]; self.sdf_queries_J += stats[5]; self.sdf_queries_noJ += stats[6]
        _lib.CALLS["gn_kernel"] = _lib.CALLS.get("gn_kernel", 0) + stats[7]            # launches made inside the C driver
        d = np.array(list(deltap), dtype=np.float64)
        new_delta = Isometry.from_matrix(d[:9].reshape(3, 3), d[9:12])
        if stats[0] >= 10:
            self.n_unstable += 1
            if self.n_unstable >= 3:
                self.rgb_args.weight = max(self.rgb_args.weight, 500.)
        return last_pose.dot(new_delta)

    def gauss_newton(self, init_pose, cur_intensity_pyramid, cur_depth_pyramid, cur_dIdxy_pyramid, obs_xyz, calib, obs_count=None):
        """tracker.py:225-288 (control flow unchanged: rollback+break when the energy rises, one evaluation-only pass per
        group, instability counter)."""
        last_pose = self.all_pd_pose[-1]
        cur_delta_pose = last_pose.inv().dot(init_pose)
        if self.native_gn:
            return self._gauss_newton_native(last_pose, cur_delta_pose, cur_intensity_pyramid, cur_depth_pyramid,
                                             cur_dIdxy_pyramid, obs_xyz, calib, obs_count)
        if obs_count is not None:
            obs_xyz = obs_xyz[:int(obs_count.item())]
        last_delta_pose = copy.deepcopy(cur_delta_pose)
        i_iter = 0
        for group in self.args.iter_config:
            last_energy = np.inf
            for i_iter in list(range(group["n"])) + [-1]:
                H = np.zeros((6, 6), dtype=float)
                g = np.zeros((6,), dtype=float)
                cur_energy = 0.0
                for loss_config in group["type"]:
                    if loss_config[0] == "sdf":
                        tH, tg, te = self.compute_sdf_Hg(i_iter, last_pose, cur_delta_pose, obs_xyz, i_iter == -1)
                    elif loss_config[0] == "rgb":
                        tH, tg, te = self.compute_rgb_Hg(loss_config[1], cur_delta_pose, cur_intensity_pyramid, cur_depth_pyramid,
                                                         cur_dIdxy_pyramid, calib, i_iter == -1)
                    else:
                        raise NotImplementedError(loss_config[0])     # 'motion' is referenced but undefined in the reference
                    cur_energy += te
                    if i_iter != -1:
                        H += tH
                        g += tg
                if cur_energy > last_energy:
                    cur_delta_pose = last_delta_pose
                    break
                last_delta_pose = copy.deepcopy(cur_delta_pose)
                last_energy = cur_energy
                if i_iter != -1:
                    xi = np.linalg.solve(H, -g)
                    cur_delta_pose = Isometry.from_twist(xi) @ cur_delta_pose
        if i_iter >= 10:
            self.n_unstable += 1
            if self.n_unstable >= 3:
                self.rgb_args.weight = max(self.rgb_args.weight, 500.)
        return last_pose.dot(cur_delta_pose)
