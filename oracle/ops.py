"""Oracle restatement of the reference's native ops (system/ext/*, torch_scatter).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  numpy fp32 unless stated; every function cites
the reference lines it follows.  Where the reference's nvcc build contracts ``a*b + c`` into an FMA
the oracle states which contraction it assumes (``_fma``) so the comparison can be bit-exact.
"""
import numpy as np

f32 = np.float32


def _fma(a, b, c):
    """fp32 fused multiply-add emulated through float64 (a*b is exact in f64)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


# ----------------------------------------------------------------------------------------------
# imgproc
# ----------------------------------------------------------------------------------------------
def ingest_frame(depth_raw, color_raw, depth_scale=5000.0, depth_cut=None, bgr=False, recip=True):
    """dataset/production/icl_nuim.py:110-114: depth = float32(raw) / calib[4] and rgb = float32(raw) / 255. -- both executed
    by torch on CUDA tensors in the reference, i.e. as a multiply by the fp32 reciprocal (recip=True); numpy's true divide
    with recip=False.  main.py:56-57: depths outside [cut_min, cut_max] -> NaN.  bgr: cv2.cvtColor(BGR2RGB), :112."""
    out_d = out_c = None
    if depth_raw is not None:
        d = np.asarray(depth_raw).astype(f32)
        out_d = (d * (f32(1.0) / f32(depth_scale))).astype(f32) if recip else (d / f32(depth_scale)).astype(f32)
        if depth_cut is not None:
            out_d = out_d.copy()
            out_d[(out_d < f32(depth_cut[0])) | (out_d > f32(depth_cut[1]))] = np.nan
    if color_raw is not None:
        c = np.asarray(color_raw).astype(f32)
        if bgr:
            c = c[..., ::-1]
        out_c = (c * (f32(1.0) / f32(255.0))).astype(f32) if recip else (c / f32(255.0)).astype(f32)
        out_c = np.ascontiguousarray(out_c)
    return out_d, out_c


def unproject_depth(depth, fx, fy, cx, cy):
    """imgproc.cu:5-23.  depth (H,W) f32, NaN = invalid.  Returns (H,W,3) f32.
    The reference writes only channel 0 (=NaN) for invalid pixels and leaves channels 1,2
    uninitialised (torch::empty, imgproc.cu:33); the oracle writes NaN to all three."""
    depth = np.asarray(depth, dtype=f32)
    H, W = depth.shape
    u = np.arange(W, dtype=f32)[None, :]
    v = np.arange(H, dtype=f32)[:, None]
    x = ((u - f32(cx)) / f32(fx) * depth).astype(f32)
    y = ((v - f32(cy)) / f32(fy) * depth).astype(f32)
    out = np.stack([np.broadcast_to(x, depth.shape), np.broadcast_to(y, depth.shape), depth], axis=-1).astype(f32)
    out[np.isnan(depth)] = np.nan
    return out


def gradient_xy(img):
    """photometric.cu:3-22.  Sobel/8 with NaN border.  (H,W) -> (H,W,2)."""
    I = np.asarray(img, dtype=f32)
    H, W = I.shape
    out = np.full((H, W, 2), np.nan, dtype=f32)
    c = I[1:-1, 1:-1]
    ud1 = I[:-2, 2:] - I[:-2, :-2]
    ud2 = I[1:-1, 2:] - I[1:-1, :-2]
    ud3 = I[2:, 2:] - I[2:, :-2]
    out[1:-1, 1:-1, 0] = ((ud1 + f32(2) * ud2) + ud3) / f32(8)
    vd1 = I[2:, :-2] - I[:-2, :-2]
    vd2 = I[2:, 1:-1] - I[:-2, 1:-1]
    vd3 = I[2:, 2:] - I[:-2, 2:]
    out[1:-1, 1:-1, 1] = ((vd1 + f32(2) * vd2) + vd3) / f32(8)
    del c
    return out


def rgb_odometry(prev_I, prev_D, cur_I, cur_D, cur_dIdxy, intr, krkinv, kt, min_grad_scale, max_depth_delta,
                 compute_J=True):
    """photometric.cu:24-77.  Returns f (H,W) [, J (H,W,6)]; NaN where invalid.  No FMA contraction is
    assumed (plain fp32 ops, left-to-right); pixels whose warped coordinate lands within 1e-3 of a
    rounding boundary may differ from a contracted build and are excluded by the parity tests."""
    prev_I = np.asarray(prev_I, f32); prev_D = np.asarray(prev_D, f32)
    cur_I = np.asarray(cur_I, f32); cur_D = np.asarray(cur_D, f32); G = np.asarray(cur_dIdxy, f32)
    H, W = cur_I.shape
    K = np.asarray(krkinv, f32).reshape(3, 3); kt = np.asarray(kt, f32)
    fx, fy, cx, cy = [f32(t) for t in intr]
    u = np.broadcast_to(np.arange(W, dtype=f32)[None, :], (H, W))
    v = np.broadcast_to(np.arange(H, dtype=f32)[:, None], (H, W))
    f = np.full((H, W), np.nan, dtype=f32)
    J = np.full((H, W, 6), np.nan, dtype=f32) if compute_J else None
    dIx, dIy = G[..., 0], G[..., 1]
    with np.errstate(all="ignore"):
        mTwo = dIx * dIx + dIy * dIy
        ok = ~((mTwo < f32(min_grad_scale)) | np.isnan(mTwo))
        d1 = cur_D
        ok &= ~np.isnan(d1)
        wd = d1 * ((K[2, 0] * u + K[2, 1] * v) + K[2, 2]) + kt[2]
        uf = (d1 * ((K[0, 0] * u + K[0, 1] * v) + K[0, 2]) + kt[0]) / wd
        vf = (d1 * ((K[1, 0] * u + K[1, 1] * v) + K[1, 2]) + kt[1]) / wd
        ufc = np.where(np.isfinite(uf), uf, f32(-1e9)); vfc = np.where(np.isfinite(vf), vf, f32(-1e9))
        u0 = np.rint(np.clip(ufc, -2e9, 2e9)).astype(np.int64)
        v0 = np.rint(np.clip(vfc, -2e9, 2e9)).astype(np.int64)
        ok &= (u0 >= 0) & (u0 < W) & (v0 >= 0) & (v0 < H)
        u0c = np.clip(u0, 0, W - 1); v0c = np.clip(v0, 0, H - 1)
        d0 = prev_D[v0c, u0c]
        ok &= ~np.isnan(d0) & (np.abs(wd - d0) <= f32(max_depth_delta)) & (d0 > 0)
        fv = cur_I - prev_I[v0c, u0c]
        f[ok] = fv[ok]
        if compute_J:
            Gx = d0 * (u0c.astype(f32) - cx) / fx
            Gy = d0 * (v0c.astype(f32) - cy) / fy
            Gz = d0
            p0 = dIx * fx / Gz
            p1 = dIy * fy / Gz
            p2 = -(p0 * Gx + p1 * Gy) / Gz
            Jv = np.stack([p0, p1, p2, -Gz * p1 + Gy * p2, Gz * p0 - Gx * p2, -Gy * p0 + Gx * p1], axis=-1).astype(f32)
            J[ok] = Jv[ok]
    return (f, J) if compute_J else (f,)


# ----------------------------------------------------------------------------------------------
# pcproc: exact kNN (cuda_kdtree.cu:994-1065,130-214 semantics) + radius filter + PCA normals
# ----------------------------------------------------------------------------------------------
def _dist2(a, b):
    """cuda_kdtree.cu:1152-1155 + cutil_math.h:1127-1130: diff = a-b (float4, w=0), dot(diff,diff) =
    x*x + y*y + z*z + w*w.  Assumed nvcc contraction: fma(z,z, fma(y,y, x*x)) (w term adds exact 0)."""
    d = (a - b).astype(f32)
    s = (d[..., 0] * d[..., 0]).astype(f32)
    s = _fma(d[..., 1], d[..., 1], s)
    s = _fma(d[..., 2], d[..., 2], s)
    return s


def knn(points, k, max_radius=None):
    """Exact k nearest neighbours (self included, ascending distance; ties in arbitrary order like the
    reference's heap sort).  points (N,3|4) f32.  Returns dist2 (N,k) f32 [inf padded], idx (N,k) int64 [-1 padded].
    max_radius (optional) bounds the candidate search; entries beyond it come back as inf/-1, which is
    indistinguishable for both callers (they only test entries against radius^2)."""
    from scipy.spatial import cKDTree
    P = np.asarray(points, dtype=f32)[:, :3]
    N = P.shape[0]
    dist = np.full((N, k), np.inf, dtype=f32)
    idx = np.full((N, k), -1, dtype=np.int64)
    if N == 0:
        return dist, idx
    tree = cKDTree(P.astype(np.float64))
    if max_radius is None:
        dd, ii = tree.query(P.astype(np.float64), k=min(k, N))
        ii = ii.reshape(N, -1)
        cand_i = np.repeat(np.arange(N), ii.shape[1]); cand_j = ii.reshape(-1)
    else:
        pairs = tree.query_pairs(float(max_radius) * 1.0001 + 1e-6, output_type="ndarray")
        cand_i = np.concatenate([pairs[:, 0], pairs[:, 1], np.arange(N)])
        cand_j = np.concatenate([pairs[:, 1], pairs[:, 0], np.arange(N)])
    d2 = _dist2(P[cand_j], P[cand_i])
    # order candidates per query by (d2, prefer self first on ties, then index)
    is_self = (cand_i != cand_j).astype(np.int8)
    order = np.lexsort((cand_j, is_self, d2, cand_i))
    cand_i = cand_i[order]; cand_j = cand_j[order]; d2 = d2[order]
    start = np.searchsorted(cand_i, np.arange(N), side="left")
    rank = np.arange(cand_i.shape[0]) - start[cand_i]
    keep = rank < k
    dist[cand_i[keep], rank[keep]] = d2[keep]
    idx[cand_i[keep], rank[keep]] = cand_j[keep]
    return dist, idx


def remove_radius_outlier(points, nb_points, radius):
    """pcproc.cu:98-105,160-187: mask = dist2[nb_points-1] < radius*radius (fp32 product)."""
    dist, _ = knn(points, nb_points, max_radius=radius)
    r2 = f32(radius) * f32(radius)
    return dist[:, nb_points - 1] < r2


def _sym3eig_smallest(c):
    """pcproc.cu:21-96 for a batch: c (N,3,3) f32 symmetric.  Returns (N,3) eigenvector of the smallest
    eigenvalue.  Mixed precision exactly as written there: M_PI terms promote to double
    (pcproc.cu:41-52), everything else fp32."""
    x1 = c[:, 0].copy(); x2 = c[:, 1].copy(); x3 = c[:, 2].copy()
    with np.errstate(all="ignore"):
        p1 = x1[:, 1] * x1[:, 1] + x1[:, 2] * x1[:, 2] + x2[:, 2] * x2[:, 2]
        q = (x1[:, 0] + x2[:, 1] + x3[:, 2]) / f32(3)
        p2 = (x1[:, 0] - q) * (x1[:, 0] - q) + (x2[:, 1] - q) * (x2[:, 1] - q) + (x3[:, 2] - q) * (x3[:, 2] - q) + f32(2) * p1
        p = np.sqrt(p2 / f32(6)).astype(f32)
        ip = (f32(1) / p).astype(f32)
        b11 = ip * (x1[:, 0] - q); b12 = ip * x1[:, 1]; b13 = ip * x1[:, 2]
        b21 = ip * x2[:, 0]; b22 = ip * (x2[:, 1] - q); b23 = ip * x2[:, 2]
        b31 = ip * x3[:, 0]; b32 = ip * x3[:, 1]; b33 = ip * (x3[:, 2] - q)
        r = b11 * b22 * b33 + b12 * b23 * b31 + b13 * b21 * b32 - b13 * b22 * b31 - b12 * b21 * b33 - b11 * b23 * b32
        r = (r / f32(2)).astype(f32)
        phi = np.where(r <= -1, f32(np.pi / 3.0), np.where(r >= 1, f32(0), (np.arccos(np.clip(r, -1, 1)).astype(f32) / f32(3)))).astype(f32)
        lam = (q.astype(np.float64) + (f32(2) * p).astype(np.float64) * np.cos(phi.astype(np.float64) + 2 * np.pi / 3)).astype(f32)
        x1[:, 0] -= lam; x2[:, 1] -= lam; x3[:, 2] -= lam

        def cr(a, b):
            return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
                             a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                             a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1).astype(f32)
        r12, r13, r23 = cr(x1, x2), cr(x1, x3), cr(x2, x3)
        d1 = (r12 * r12).sum(1, dtype=f32); d2 = (r13 * r13).sum(1, dtype=f32); d3 = (r23 * r23).sum(1, dtype=f32)
        # pcproc.cu:68-79: note i_max=2 is chosen when d3 > max(d1,d2) (d_max not updated after that)
        dmax = d1.copy(); imax = np.zeros(len(d1), dtype=np.int64)
        m = d2 > dmax; dmax[m] = d2[m]; imax[m] = 1
        m = d3 > dmax; imax[m] = 2
        vec = np.where((imax == 0)[:, None], r12 / np.sqrt(d1)[:, None],
                       np.where((imax == 1)[:, None], r13 / np.sqrt(d2)[:, None], r23 / np.sqrt(d3)[:, None]))
    return vec.astype(f32)


def estimate_normals(points, max_nn, radius, cam_xyz):
    """pcproc.cu:107-158,189-210.  Mean and (unnormalised) covariance over the sorted neighbours 1..max_nn-1
    with dist2 < radius^2 (stop at the first miss), >=5 needed, smallest-eigenvalue eigenvector, flipped so
    dot(n, p - cam) <= 0.  Sums are sequential in ascending-distance order, fp32, no contraction assumed."""
    P = np.asarray(points, dtype=f32)[:, :3]
    N = P.shape[0]
    out = np.full((N, 3), np.nan, dtype=f32)
    if N == 0:
        return out
    dist, idx = knn(points, max_nn, max_radius=radius)
    r2 = f32(radius) * f32(radius)
    ok = dist[:, 1:] < r2
    ok = np.logical_and.accumulate(ok, axis=1)            # break at first miss
    cnt = ok.sum(1).astype(f32)
    nb = P[np.clip(idx[:, 1:], 0, N - 1)]                 # (N, k-1, 3)
    mean = np.zeros((N, 3), dtype=f32)
    for j in range(max_nn - 1):
        mean = np.where(ok[:, j:j + 1], (mean + nb[:, j]).astype(f32), mean)
    valid = cnt >= 5
    with np.errstate(all="ignore"):
        mean = (mean / cnt[:, None]).astype(f32)
    cov = np.zeros((N, 3, 3), dtype=f32)
    for j in range(max_nn - 1):
        d = (nb[:, j] - mean).astype(f32)
        outer = (d[:, :, None] * d[:, None, :]).astype(f32)
        cov = np.where(ok[:, j, None, None], (cov + outer).astype(f32), cov)
    n = _sym3eig_smallest(cov)
    cam = np.asarray(cam_xyz, dtype=f32)
    dp = (P - cam).astype(f32)
    flip = (n[:, 0] * dp[:, 0] + n[:, 1] * dp[:, 1] + n[:, 2] * dp[:, 2]) > 0
    n = np.where(flip[:, None], -n, n)
    out[valid] = n[valid]
    return out


# ----------------------------------------------------------------------------------------------
# indexing / torch_scatter
# ----------------------------------------------------------------------------------------------
def groupby_sum(values, indices, C):
    """indexing.cu:59-71,89-109: per-group sum (fp32) and int32 "count".  Sequential order here; the
    reference's atomics are order-nondeterministic, so callers compare with a tolerance.  Note the reference
    increments the count from every one of the L threads of a row's block (indexing.cu:69-70), so its count is
    L x (rows in the group) -- confirmed against the reference's own kernel on a B200 (tests/test_gpu_refext.py)."""
    values = np.asarray(values, dtype=f32); indices = np.asarray(indices, dtype=np.int64)
    s = np.zeros((int(C), values.shape[1]), dtype=f32)
    np.add.at(s, indices, values)
    c = (np.bincount(indices, minlength=int(C)) * values.shape[1]).astype(np.int32)
    return s, c


def scatter_mean(src, index, dim=0):
    """torch_scatter.scatter_mean(src, index, dim=0) (call sites tracker.py:22-23): out[i] = mean of rows with
    index == i, output has max(index)+1 rows, empty groups 0.  Sums are sequential in row order
    (torch_scatter's CPU kernel), fp32; count is clamped to >=1 before the divide."""
    assert dim == 0
    src = np.asarray(src, dtype=f32); index = np.asarray(index, dtype=np.int64)
    n = int(index.max()) + 1 if index.size else 0
    s = np.zeros((n,) + src.shape[1:], dtype=f32)
    np.add.at(s, index, src)
    c = np.bincount(index, minlength=n).astype(f32)
    c = np.maximum(c, f32(1))
    return (s / c.reshape((-1,) + (1,) * (src.ndim - 1))).astype(f32)


def point_box_filter(points, normals, voxel_size, divide="ieee"):
    """tracker.py:14-24.  Returns (filtered_pc, filtered_normal, cell_key_sorted).  ``divide`` selects how
    ``x / voxel_size`` is evaluated: 'ieee' (torch CPU) or 'recip' (torch CUDA: x * (1/vs))."""
    P = np.asarray(points, dtype=f32); Nn = np.asarray(normals, dtype=f32)
    vs = f32(voxel_size)
    half = f32(np.float64(voxel_size) * 0.5)
    mn = P.min(0, keepdims=True) - half
    mx = P.max(0, keepdims=True) + half
    if divide == "ieee":
        q = (P - mn) / vs
        ext = (mx - mn) / vs
    else:
        inv = f32(1.0) / vs
        q = (P - mn) * inv
        ext = (mx - mn) * inv
    coord = np.floor(q).astype(np.int64)
    nx, ny, nz = (np.floor(ext).astype(np.int64) + 16)[0].tolist()
    key = coord[:, 0] + coord[:, 1] * nx + coord[:, 2] * nx * ny
    uq, inv_ind = np.unique(key, return_inverse=True)
    return scatter_mean(P, inv_ind), scatter_mean(Nn, inv_ind), uq


# ----------------------------------------------------------------------------------------------
# marching cubes (mc_interp_kernel.cu)
# ----------------------------------------------------------------------------------------------
def _mc_tables():
    from .mc_tables import TRI_TABLE
    tri = np.full((256, 16), -1, dtype=np.int64)
    edge = np.zeros(256, dtype=np.int64)
    for c, s in enumerate(TRI_TABLE):
        for i, ch in enumerate(s):
            e = int(ch, 16)
            tri[c, i] = e
            edge[c] |= 1 << e
    return edge, tri


def _blend_sdf(indexer3, mapping, cube_sdf, cube_std, bpos, rpos, r):
    """mc_interp_kernel.cu:34-185 (get_sdf, STD_W_SDF variant) for arrays of (bpos (n,3) int, rpos (n,3) int).
    Returns sdf (n,), std (n,) fp32, NaN where the reference returns NaN."""
    nx, ny, nz = indexer3.shape
    bsize = np.array([nx, ny, nz])
    bpos = bpos.copy(); rpos = rpos.copy()
    for a in range(3):
        over = bpos[:, a] >= bsize[a]
        bpos[over, a] = bsize[a] - 1
        rpos[over, a] = r - 1
    rbound = (r - 1) // 2
    rstart = r // 2
    rmid = f32(r / 2.0)
    n = bpos.shape[0]
    lo = rpos <= rbound                                    # (n,3) "zero_*" flags
    rf = rpos.astype(f32)
    w_p = np.where(lo, rf + rmid, rf - rmid).astype(f32)
    w_m = np.where(lo, rmid - rf, rmid + f32(r) - rf).astype(f32)
    w_m = (w_m / f32(r)).astype(f32); w_p = (w_p / f32(r)).astype(f32)
    b_m = np.where(lo, -1, 0); r_m = np.where(lo, r, 0)
    b_p = np.where(lo, 0, 1); r_p = np.where(lo, 0, -r)
    rp = rpos + rstart
    zero_det = lo[:, 0] * 4 + lo[:, 1] * 2 + lo[:, 2] * 1
    tot_sdf = np.zeros(n, dtype=f32); tot_w_sdf = np.zeros(n, dtype=f32)
    tot_std = np.zeros(n, dtype=f32); tot_w = np.zeros(n, dtype=f32)
    dead = np.zeros(n, dtype=bool)
    max_vec = mapping.shape[0]
    corner = 0
    for sx in (0, 1):
        for sy in (0, 1):
            for sz in (0, 1):
                bb = np.stack([bpos[:, 0] + (b_p if sx else b_m)[:, 0],
                               bpos[:, 1] + (b_p if sy else b_m)[:, 1],
                               bpos[:, 2] + (b_p if sz else b_m)[:, 2]], 1)
                rr = np.stack([rp[:, 0] + (r_p if sx else r_m)[:, 0],
                               rp[:, 1] + (r_p if sy else r_m)[:, 1],
                               rp[:, 2] + (r_p if sz else r_m)[:, 2]], 1)
                w = ((w_p if sx else w_m)[:, 0] * (w_p if sy else w_m)[:, 1]).astype(f32)
                w = (w * (w_p if sz else w_m)[:, 2]).astype(f32)
                inb = np.all((bb >= 0) & (bb < bsize), axis=1)      # uint wrap of -1 fails the >= size test
                bbc = np.clip(bb, 0, bsize - 1)
                vec = indexer3[bbc[:, 0], bbc[:, 1], bbc[:, 2]]
                ok = inb & (vec != -1) & (vec < max_vec)
                batch = np.where(ok, mapping[np.clip(vec, 0, max_vec - 1)], -1)
                ok &= batch != -1
                bc = np.clip(batch, 0, cube_sdf.shape[0] - 1)
                rrc = np.clip(rr, 0, cube_sdf.shape[1] - 1)
                s = cube_sdf[bc, rrc[:, 0], rrc[:, 1], rrc[:, 2]]
                sd = cube_std[bc, rrc[:, 0], rrc[:, 1], rrc[:, 2]]
                ok &= ~np.isnan(s)
                with np.errstate(all="ignore"):
                    tot_sdf = np.where(ok, (tot_sdf + ((s * w).astype(f32) * sd).astype(f32)).astype(f32), tot_sdf)
                    tot_w_sdf = np.where(ok, (tot_w_sdf + (w * sd).astype(f32)).astype(f32), tot_w_sdf)
                    tot_std = np.where(ok, (tot_std + (w * sd).astype(f32)).astype(f32), tot_std)
                    tot_w = np.where(ok, (tot_w + w).astype(f32), tot_w)
                dead |= (~ok) & (zero_det == corner)
                corner += 1
    with np.errstate(all="ignore"):
        sdf = (tot_sdf / tot_w_sdf).astype(f32)
        std = (tot_std / tot_w).astype(f32)
    sdf[dead] = np.nan; std[dead] = np.nan
    return sdf, std


_CORNER = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]])
_EDGE_ENDS = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]


def marching_cubes_sparse_interp(indexer3, valid_blocks, mapping, cube_sdf, cube_std, max_n_triangles, n_xyz, max_std):
    """mc_interp_kernel.cu:202-382.  Returns triangles (T,3,3) f32 [voxel units], flatten_id (T,) i64,
    std (T,3) f32 in a canonical order (block, sub-cell, table order); the reference's order is set by an
    atomic counter, so parity is on the sorted triangle set."""
    edge_tab, tri_tab = _mc_tables()
    indexer3 = np.asarray(indexer3, dtype=np.int64)
    valid_blocks = np.asarray(valid_blocks, dtype=np.int64)
    mapping = np.asarray(mapping, dtype=np.int64)
    cube_sdf = np.asarray(cube_sdf, dtype=f32); cube_std = np.asarray(cube_std, dtype=f32)
    nx, ny, nz = [int(t) for t in n_xyz]
    r = cube_sdf.shape[1] // 2
    r3 = r ** 3
    U = valid_blocks.shape[0]
    sbs = f32(1.0) / f32(r)
    lif = np.repeat(np.arange(U), r3)
    sub = np.tile(np.arange(r3), U)
    vb = valid_blocks[lif]
    bpos = np.stack([(vb // (ny * nz)) % nx, (vb // nz) % ny, vb % nz], 1)
    rxyz = np.stack([sub // (r * r), (sub // r) % r, sub % r], 1)
    n = lif.shape[0]
    sdf = np.zeros((n, 8), dtype=f32); std = np.zeros((n, 8), dtype=f32)
    pts = np.zeros((n, 8, 3), dtype=f32)
    alive = np.ones(n, dtype=bool)
    for c in range(8):
        rp = rxyz + _CORNER[c]
        s, sd = _blend_sdf(indexer3, mapping, cube_sdf, cube_std, bpos, rp, r)
        alive &= ~np.isnan(s)
        sdf[:, c] = s; std[:, c] = sd
        pts[:, c] = (bpos.astype(f32) + (rp.astype(f32) * sbs).astype(f32)).astype(f32)
    ctype = np.zeros(n, dtype=np.int64)
    for c in range(8):
        ctype |= ((sdf[:, c] < 0) & alive).astype(np.int64) << c
    ctype[~alive] = 0
    tris, ids, stds = [], [], []
    cells = np.nonzero(edge_tab[ctype] != 0)[0]
    for ci in cells:
        ct = ctype[ci]
        verts = {}
        for e in range(12):
            if edge_tab[ct] >> e & 1:
                a, b = _EDGE_ENDS[e]
                verts[e] = _sdf_interp(pts[ci, a], pts[ci, b], std[ci, a], std[ci, b], sdf[ci, a], sdf[ci, b])
        row = tri_tab[ct]
        i = 0
        while i < 16 and row[i] != -1:
            vp = [verts[int(row[i + k])] for k in range(3)]
            i += 3
            if any(v[3] > f32(max_std) for v in vp):
                continue
            tris.append([v[:3] for v in vp]); stds.append([v[3] for v in vp]); ids.append(valid_blocks[lif[ci]])
    T = min(len(tris), int(max_n_triangles))
    tri = np.asarray(tris[:T], dtype=f32).reshape(T, 3, 3)
    return tri, np.asarray(ids[:T], dtype=np.int64), np.asarray(stds[:T], dtype=f32).reshape(T, 3)


def _sdf_interp(p1, p2, s1, s2, v1, v2):
    """mc_interp_kernel.cu:187-200."""
    if abs(f32(0) - v1) < f32(1e-5):
        return np.array([p1[0], p1[1], p1[2], s1], dtype=f32)
    if abs(f32(0) - v2) < f32(1e-5):
        return np.array([p2[0], p2[1], p2[2], s2], dtype=f32)
    if abs(v1 - v2) < f32(1e-5):
        return np.array([p1[0], p1[1], p1[2], s1], dtype=f32)
    w2 = f32((f32(0) - v1) / (v2 - v1))
    w1 = f32(f32(1) - w2)
    return np.array([f32(p1[0] * w1) + f32(p2[0] * w2), f32(p1[1] * w1) + f32(p2[1] * w2),
                     f32(p1[2] * w1) + f32(p2[2] * w2), f32(s1 * w1) + f32(s2 * w2)], dtype=f32)
