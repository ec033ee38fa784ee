"""Operator API: the names the reference binds in system/ext/__init__.py:17-42 plus torch_scatter.scatter_mean,
with the same argument meaning and error behaviour (RuntimeError for non-CUDA / non-contiguous inputs, like the
reference's TORCH_CHECKs in imgproc/common.cuh:7-9, pcproc.cu:10-12, indexing.cu:4-6, mc_data.cuh:7-9).

Every function takes torch CUDA tensors, allocates its outputs with torch (callee-allocates, like the reference),
and calls the C-ABI in include/difusion_b200.h on torch's CURRENT stream.  PyTorch is plumbing here (memory and
streams); all arithmetic happens in libdifusion_b200.so.  There is no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, fptr

# dead in the reference (never called by any Python there) and therefore not provided:
#   filter_depth, compute_normal_weight, compute_normal_weight_robust, pack_batch


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t, name, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"{name} must have dtype {dtype}, got {t.dtype}")
    return t


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Workspace:
    """Growable per-device scratch buffer handed to the C-ABI (`ws`, `ws_bytes`)."""

    def __init__(self):
        self.buf = {}

    def get(self, device, nbytes):
        key = (device.type, device.index)
        b = self.buf.get(key)
        if b is None or b.numel() < nbytes:
            b = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
            self.buf[key] = b
        return b


_WS = Workspace()


# ----------------------------------------------------------------------------------------------- frame ingest
def ingest_frame(depth_raw=None, color_raw=None, depth_scale=5000.0, depth_cut=None, bgr=False, div_mode=1, out=None):
    """dataset/production/icl_nuim.py:110-114 (+ main.py:56-57 when depth_cut is given) on the device: depth_raw uint16[H,W]
    and / or color_raw uint8[H,W,3] (CUDA copies of the decoded images) -> (depth f32[H,W] | None, rgb f32[H,W,3] | None).
    div_mode 1 = multiply by the fp32 reciprocal (torch CUDA's tensor / scalar, what the reference runs), 0 = IEEE divide.
    out = (depth f32[H,W], rgb f32[H,W,3]): optional preallocated outputs (a streaming reader reuses two sets)."""
    ref = depth_raw if depth_raw is not None else color_raw
    if ref is None:
        raise ValueError("ingest_frame needs depth_raw and / or color_raw")
    if depth_raw is not None:
        if depth_raw.dtype not in (torch.uint16, torch.int16):
            raise RuntimeError("depth_raw must be a 16-bit integer tensor (the PNG's uint16 payload)")
        _chk(depth_raw, "depth_raw", depth_raw.dtype)
    if color_raw is not None:
        _chk(color_raw, "color_raw", torch.uint8)
        if color_raw.dim() != 3 or color_raw.size(2) != 3:
            raise RuntimeError("color_raw must be (H, W, 3)")
    H, W = int(ref.shape[0]), int(ref.shape[1])
    depth = rgb = None
    if depth_raw is not None:
        depth = out[0] if out is not None else torch.empty((H, W), dtype=torch.float32, device=ref.device)
        _chk(depth, "out depth", torch.float32)
    if color_raw is not None:
        rgb = out[1] if out is not None else torch.empty((H, W, 3), dtype=torch.float32, device=ref.device)
        _chk(rgb, "out rgb", torch.float32)
    lo, hi = (float(depth_cut[0]), float(depth_cut[1])) if depth_cut is not None else (0.0, 0.0)
    with torch.cuda.device(ref.device):
        check(_lib.load().dfb_ingest_frame(_p(depth_raw) if depth_raw is not None else None,
                                           _p(color_raw) if color_raw is not None else None, H, W, float(depth_scale), int(div_mode),
                                           lo, hi, int(bool(bgr)), _p(depth) if depth is not None else None,
                                           _p(rgb) if rgb is not None else None, _stream()))
    return depth, rgb


def transform_points(xyz, R, t=None):
    """Isometry @ xyz (motion_util.py:323-328): xyz f32[N,3] CUDA, R 3x3 / t (3,) host arrays -> R xyz + t, f32[N,3]."""
    _chk(xyz, "xyz", torch.float32)
    out = torch.empty_like(xyz)
    Rf = fptr(np.asarray(R, dtype=np.float32).reshape(-1).tolist())
    tf = fptr(np.asarray(t, dtype=np.float32).reshape(-1).tolist()) if t is not None else None
    with torch.cuda.device(xyz.device):
        check(_lib.load().dfb_transform_points(_p(xyz), xyz.size(0), Rf, tf, _p(out), _stream()))
    return out


# ----------------------------------------------------------------------------------------------- imgproc
def unproject_depth(depth, fx, fy, cx, cy):
    """system.ext.unproject_depth (imgproc.cpp:3): depth f32[H,W] -> f32[H,W,3]."""
    _chk(depth, "depth", torch.float32)
    H, W = depth.shape
    pc = torch.empty((H, W, 3), dtype=torch.float32, device=depth.device)
    with torch.cuda.device(depth.device):
        check(_lib.load().dfb_unproject_depth(_p(depth), H, W, fx, fy, cx, cy, _p(pc), _stream()))
    return pc


def gradient_xy(intensity):
    """system.ext.gradient_xy (imgproc.cpp:21): f32[H,W] -> f32[H,W,2]."""
    _chk(intensity, "cur_intensity", torch.float32)
    H, W = intensity.shape
    g = torch.empty((H, W, 2), dtype=torch.float32, device=intensity.device)
    with torch.cuda.device(intensity.device):
        check(_lib.load().dfb_gradient_xy(_p(intensity), H, W, _p(g), _stream()))
    return g


def frame_images(rgb, depth, depth_cut=None):
    """Image half of the tracker front end (tracker.py:42-57, 84; main.py:56-57) in two launches: depth clipping, intensity =
    mean over the colour axis, the 3-level pyramid (bilinear align_corners=True / nearest) and gradient_xy of every level.
    rgb f32[H,W,3], depth f32[H,W] -> (Is, Ds, Gs), three levels each; bit-identical to the torch ops of the reference."""
    _chk(rgb, "rgb", torch.float32); _chk(depth, "depth", torch.float32)
    H, W = depth.shape
    if tuple(rgb.shape) != (H, W, 3):
        raise RuntimeError("frame_images: rgb must be (H, W, 3)")
    dev = depth.device
    dims = [(H, W), (H // 2, W // 2), (H // 2 // 2, W // 2 // 2)]
    Is = [torch.empty(d, dtype=torch.float32, device=dev) for d in dims]
    Ds = [torch.empty(d, dtype=torch.float32, device=dev) for d in dims]
    Gs = [torch.empty(d + (2,), dtype=torch.float32, device=dev) for d in dims]
    near, far = (float(depth_cut[0]), float(depth_cut[1])) if depth_cut is not None else (0.0, 0.0)
    with torch.cuda.device(dev):
        check(_lib.load().dfb_frame_images(_p(rgb), _p(depth), H, W, near, far, int(depth_cut is not None), _p(Is[0]), _p(Ds[0]),
                                           _p(Is[1]), _p(Ds[1]), _p(Is[2]), _p(Ds[2]), _p(Gs[0]), _p(Gs[1]), _p(Gs[2]), _stream()))
    return Is, Ds, Gs


def rgb_odometry(prev_intensity, prev_depth, cur_intensity, cur_depth, cur_dIdxy, intr, krkinv_data, kt_data,
                 min_grad_scale, max_depth_delta, compute_J):
    """system.ext.rgb_odometry (imgproc.cpp:14-20): returns [f] or [f, J]."""
    for t, n in ((prev_intensity, "prev_intensity"), (prev_depth, "prev_depth"), (cur_intensity, "cur_intensity"),
                 (cur_depth, "cur_depth"), (cur_dIdxy, "cur_dIdxy")):
        _chk(t, n, torch.float32)
    H, W = cur_intensity.shape
    dev = cur_intensity.device
    f = torch.empty((H, W), dtype=torch.float32, device=dev)
    J = torch.empty((H, W, 6), dtype=torch.float32, device=dev) if compute_J else None
    with torch.cuda.device(dev):
        check(_lib.load().dfb_rgb_odometry(_p(prev_intensity), _p(prev_depth), _p(cur_intensity), _p(cur_depth), _p(cur_dIdxy),
                                           H, W, fptr(intr), fptr(krkinv_data), fptr(kt_data), min_grad_scale, max_depth_delta,
                                           _p(f), _p(J), _stream()))
    return [f, J] if compute_J else [f]


def rgb_hg(prev_intensity, prev_depth, cur_intensity, cur_depth, cur_dIdxy, intr, krkinv_data, kt_data,
           min_grad_scale, max_depth_delta, robust, robust_k, compute_J, out=None):
    """Fused compute_rgb_Hg reduction (tracker.py:136-177).  Returns an 80-double device tensor; [0:44] is the result."""
    H, W = cur_intensity.shape
    dev = cur_intensity.device
    if out is None:
        out = torch.empty(80, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(_lib.load().dfb_rgb_hg(_p(prev_intensity), _p(prev_depth), _p(cur_intensity), _p(cur_depth), _p(cur_dIdxy), H, W,
                                     fptr(intr), fptr(krkinv_data), fptr(kt_data), min_grad_scale, max_depth_delta,
                                     int(robust), float(robust_k), int(bool(compute_J)), _p(out), _stream()))
    return out


# ----------------------------------------------------------------------------------------------- pcproc
def remove_radius_outlier(input_pc, nb_points, radius):
    """system.ext.remove_radius_outlier (pcproc.cpp:3-7): f32[N,4] -> bool[N]."""
    _chk(input_pc, "input_pc", torch.float32)
    if input_pc.dim() != 2 or input_pc.size(1) != 4:
        raise RuntimeError("input_pc must be (N, 4) (the kd-tree of the reference reads float4, cuda_kdtree.cu:1311)")
    n = input_pc.size(0)
    mask = torch.empty((n,), dtype=torch.bool, device=input_pc.device)
    lib = _lib.load()
    nbytes = lib.dfb_pcproc_ws_bytes(n)
    ws = _WS.get(input_pc.device, nbytes)
    with torch.cuda.device(input_pc.device):
        check(lib.dfb_remove_radius_outlier(_p(input_pc), n, int(nb_points), float(radius), _p(mask), _p(ws), ws.numel(), _stream()))
    return mask


def estimate_normals(input_pc, max_nn, radius, cam_xyz):
    """system.ext.estimate_normals (pcproc.cpp:9-14): f32[N,4] -> f32[N,3] (NaN rows where < 5 neighbours)."""
    _chk(input_pc, "input_pc", torch.float32)
    if input_pc.dim() != 2 or input_pc.size(1) != 4:
        raise RuntimeError("input_pc must be (N, 4)")
    n = input_pc.size(0)
    out = torch.empty((n, 3), dtype=torch.float32, device=input_pc.device)
    lib = _lib.load()
    nbytes = lib.dfb_pcproc_ws_bytes(n)
    ws = _WS.get(input_pc.device, nbytes)
    with torch.cuda.device(input_pc.device):
        check(lib.dfb_estimate_normals(_p(input_pc), n, int(max_nn), float(radius), fptr(cam_xyz), _p(out), _p(ws), ws.numel(), _stream()))
    return out


# ----------------------------------------------------------------------------------------------- indexing / scatter
def groupby_sum(values, indices, C_):
    """system.ext.groupby_sum (indexing.cpp:3-4): (f32[N,L], i64[N], C) -> [sum f32[C,L], count i32[C]]."""
    _chk(values, "values", torch.float32)
    _chk(indices, "indices", torch.int64)
    n, L = values.shape
    Cn = int(C_)
    s = torch.empty((Cn, L), dtype=torch.float32, device=values.device)
    c = torch.empty((Cn,), dtype=torch.int32, device=values.device)
    with torch.cuda.device(values.device):
        check(_lib.load().dfb_groupby_sum(_p(values), _p(indices), n, L, Cn, _p(s), _p(c), _stream()))
    return [s, c]


def scatter_mean(src, index, dim=0):
    """torch_scatter.scatter_mean(src, index, dim=0) (tracker.py:22-23).  Deterministic row-order sums."""
    if dim != 0:
        raise NotImplementedError("only dim=0 is on the hot path (tracker.py:22-23)")
    _chk(src, "src", torch.float32)
    _chk(index, "index", torch.int64)
    n = src.size(0)
    d = src.numel() // max(n, 1) if n > 0 else (src.size(1) if src.dim() > 1 else 1)
    n_out = int(index.max().item()) + 1 if n > 0 else 0          # torch_scatter sizes its output the same way
    out = torch.zeros((n_out,) + tuple(src.shape[1:]), dtype=torch.float32, device=src.device)
    if n_out == 0:
        return out
    lib = _lib.load()
    ws = _WS.get(src.device, lib.dfb_scatter_mean_ws_bytes(n, n_out))
    with torch.cuda.device(src.device):
        check(lib.dfb_scatter_mean(_p(src), _p(index), n, d, n_out, _p(out), _p(ws), ws.numel(), _stream()))
    return out


def point_box_filter(points, normals, voxel_size, div_mode=0):
    """tracker.point_box_filter (tracker.py:14-24) in one call.  Returns (filtered_pc, filtered_normal)."""
    _chk(points, "points", torch.float32)
    _chk(normals, "normals", torch.float32)
    n = points.size(0)
    dev = points.device
    if n == 0:
        return points.new_zeros((0, 3)), points.new_zeros((0, 3))
    out_p = torch.empty((n, 3), dtype=torch.float32, device=dev)
    out_n = torch.empty((n, 3), dtype=torch.float32, device=dev)
    cnt = torch.zeros((1,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    ws = _WS.get(dev, lib.dfb_box_filter_ws_bytes(n))
    with torch.cuda.device(dev):
        check(lib.dfb_point_box_filter(_p(points), _p(normals), n, float(voxel_size), int(div_mode), _p(out_p), _p(out_n),
                                       _p(cnt), _p(ws), ws.numel(), _stream()))
    m = int(cnt.item())
    if m < 0:
        raise RuntimeError("point_box_filter: cell key range exceeds the bitmap capacity (scene extent too large for voxel_size)")
    return out_p[:m], out_n[:m]


def preprocess_frame(depth, fx, fy, cx, cy, nb_points=16, outlier_radius=0.05, max_nn=16, normal_radius=0.1,
                     cam_xyz=(0.0, 0.0, 0.0), box_voxel=0.02, div_mode=0, sync=True, ws=None):
    """tracker.py:89-120 (geometry half) fused: full-resolution depth (H,W) with NaN = invalid and its intrinsics ->
    (points (N,3), normals (N,3)) in camera space.  One device->host read (the row count).
    sync=False: no host read (CUDA-graph capturable); returns (points (n_max,3), normals (n_max,3), count i32[1]).
    ws: caller-owned workspace (uint8, >= dfb_preprocess_ws_bytes(H, W)).  A captured graph bakes the workspace address in,
    so it must not be the shared growable buffer, which is released whenever a later call needs a bigger one."""
    _chk(depth, "depth", torch.float32)
    H, W = depth.shape
    dev = depth.device
    n_max = (H // 2) * (W // 2)
    out_p = torch.empty((n_max, 3), dtype=torch.float32, device=dev)
    out_n = torch.empty((n_max, 3), dtype=torch.float32, device=dev)
    cnt = torch.empty((1,), dtype=torch.int32, device=dev)      # always written (the box filter's scan total; -1 on overflow)
    lib = _lib.load()
    if ws is None:
        ws = _WS.get(dev, lib.dfb_preprocess_ws_bytes(H, W))
    elif ws.numel() < lib.dfb_preprocess_ws_bytes(H, W) or ws.device != dev:
        raise RuntimeError("preprocess_frame: workspace too small or on another device")
    with torch.cuda.device(dev):
        check(lib.dfb_preprocess_frame(_p(depth), H, W, fx, fy, cx, cy, int(nb_points), float(outlier_radius), int(max_nn),
                                       float(normal_radius), fptr(cam_xyz), float(box_voxel), int(div_mode), _p(out_p), _p(out_n),
                                       _p(cnt), _p(ws), ws.numel(), _stream()))
    if not sync:
        return out_p, out_n, cnt
    m = int(cnt.item())
    if m < 0:
        raise RuntimeError("preprocess_frame: box-filter key range exceeds the bitmap capacity")
    return out_p[:m], out_n[:m]


# ----------------------------------------------------------------------------------------------- networks
def encoder_forward(x, encoder_blob):
    _chk(x, "x", torch.float32)
    m = x.size(0)
    out = torch.empty((m, 29), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().dfb_encoder_forward(_p(x), m, _p(encoder_blob), _p(out), _stream()))
    return out


def decoder_forward(x, decoder_blob):
    """forward_model(decoder, network_input=x) (utility.py:61): x f32[N,32] -> (sdf f32[N], std f32[N])."""
    _chk(x, "network_input", torch.float32)
    n = x.size(0)
    sdf = torch.empty((n,), dtype=torch.float32, device=x.device)
    std = torch.empty((n,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.load().dfb_decoder_forward(_p(x), n, _p(decoder_blob), _p(sdf), _p(std), _stream()))
    return sdf, std


# ----------------------------------------------------------------------------------------------- marching cubes
def marching_cubes_interp(indexer, valid_blocks, vec_batch_mapping, cube_sdf, cube_std, max_n_triangles, n_xyz, max_std):
    """system.ext.marching_cubes_interp (mc.cpp:3-12).  Returns [triangles f32[T,3,3], flatten_id i64[T], std f32[T,3]]."""
    _chk(indexer, "indexer", torch.int64); _chk(valid_blocks, "valid_blocks", torch.int64)
    _chk(vec_batch_mapping, "vec_batch_mapping", torch.int32)
    _chk(cube_sdf, "cube_sdf", torch.float32); _chk(cube_std, "cube_std", torch.float32)
    assert max_n_triangles > 0
    dev = cube_sdf.device
    U = valid_blocks.size(0)
    B = cube_sdf.size(0)
    r = cube_sdf.size(1) // 2
    lib = _lib.load()
    cnt = torch.zeros((1,), dtype=torch.int32, device=dev)
    # two launches when the first guess is too small would change nothing observable; we size the output from a
    # bound instead: at most 5 triangles per sub-cell, capped by max_n_triangles.
    cap = int(min(int(max_n_triangles), 5 * U * r * r * r))
    tri = torch.empty((max(cap, 1), 3, 3), dtype=torch.float32, device=dev)
    fid = torch.empty((max(cap, 1),), dtype=torch.int64, device=dev)
    tstd = torch.empty((max(cap, 1), 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.dfb_marching_cubes(_p(indexer), int(n_xyz[0]), int(n_xyz[1]), int(n_xyz[2]), _p(valid_blocks), U,
                                     _p(vec_batch_mapping), vec_batch_mapping.size(0), _p(cube_sdf), _p(cube_std), B, r,
                                     float(max_std), cap, _p(tri), _p(fid), _p(tstd), _p(cnt), _stream()))
    T = int(cnt.item())
    if T > cap:
        import sys
        print(f"Warning from marching cube: the max triangle number is too small {T} vs {max_n_triangles}", file=sys.stderr)
        T = cap
    return [tri[:T], fid[:T], tstd[:T]]
