"""Oracle restatement of the tracker side of the hot path.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/system/tracker.py: preprocessing (:75-120), compute_sdf_Hg (:179-223),
compute_rgb_Hg (:136-177), gauss_newton (:225-288) and the SE(3) helpers of utils/motion_util.py
(:205-228 from_twist, :43-57 left Jacobian, :275-279 inv/dot, :323-328 '@').
Poses are (q, t) with q a pyquat_shim.Quaternion, float64, exactly like the reference's Isometry.
"""
import copy

import numpy as np
import torch

from . import ops
from .pyquat_shim import Quaternion


class Pose:
    """Restatement of motion_util.Isometry (q: unit quaternion, t: (3,) float64)."""

    def __init__(self, q=None, t=None):
        self.q = q if q is not None else Quaternion()
        self.t = np.zeros(3) if t is None else np.asarray(t, dtype=np.float64)

    @property
    def R(self):
        return self.q.rotation_matrix

    def inv(self):                                        # motion_util.py:275-277
        qi = self.q.inverse
        return Pose(qi, -(qi.rotate(self.t)))

    def dot(self, right):                                 # :278-279
        return Pose(self.q * right.q, self.q.rotate(right.t) + self.t)

    def apply(self, pts):                                 # :323-328 (torch branch): fp32 R, t
        R = torch.from_numpy(self.R).float(); t = torch.from_numpy(self.t).float()
        return pts @ R.t() + t.unsqueeze(0)

    @staticmethod
    def from_twist(xi):                                   # :205-228
        rho, phi = xi[:3], xi[3:6]
        angle = np.linalg.norm(phi)
        if np.isclose(angle, 0.):
            q = Quaternion(matrix=np.identity(3) + _wedge(phi))
            J = np.identity(3) + 0.5 * _wedge(phi)
        else:
            axis = phi / angle
            s, c = np.sin(angle), np.cos(angle)
            q = Quaternion(matrix=c * np.identity(3) + (1 - c) * np.outer(axis, axis) + s * _wedge(axis))
            J = (s / angle) * np.identity(3) + (1 - s / angle) * np.outer(axis, axis) + ((1 - c) / angle) * _wedge(axis)
        return Pose(q, J @ rho)


def _wedge(p):                                            # motion_util.py:20-33
    return np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]], dtype=np.float64)


def robust_weight(x, kind, k):                            # tracker.py:60-73
    if kind == "huber":
        w = torch.ones_like(x)
        xa = torch.abs(x)
        m = xa > k
        w[m] = k / xa[m]
        return w
    if kind == "tukey":
        w = torch.zeros_like(x)
        m = torch.abs(x) <= k
        w[m] = (1 - (x[m] / k) ** 2) ** 2
        return w
    raise NotImplementedError


def compute_sdf_Hg(omap, last_pose, delta_pose, obs_xyz, no_grad=False, robust_kernel="huber", robust_k=5.0):
    """tracker.py:179-223.  Returns (H (6,6) f64 | None, g (6,) | None, energy float, n_valid)."""
    cur = last_pose.dot(delta_pose).apply(obs_xyz)
    cur.requires_grad_(not no_grad)
    sdf, std, valid = omap.get_sdf(cur)
    r = sdf / std.detach()
    J = None
    if not no_grad:
        grad = torch.autograd.grad(r, [cur], grad_outputs=torch.ones_like(r))[0][valid]
        r = r.detach()
        dxyz = delta_pose.apply(obs_xyz)[valid]
        Lt = torch.from_numpy(last_pose.R.astype(np.float32).T)
        Lai = torch.mm(grad, Lt)
        Lbi = torch.cross(dxyz, Lai, dim=-1)
        J = torch.cat([Lai, Lbi], dim=-1)
    r = r.detach()
    Wf = r
    JW = J
    if robust_kernel is not None:
        w = robust_weight(r, robust_kernel, robust_k)
        Wf = Wf * w
        JW = JW * w.unsqueeze(1) if JW is not None else None
    M = Wf.size(0)
    scale = 1.0 / M
    energy = (r * Wf).sum().item() * scale
    if no_grad:
        return None, None, float(energy), M
    H = torch.einsum('na,nb->nab', JW, J).sum(0) * scale
    g = (J * Wf.unsqueeze(1)).sum(0) * scale
    return H.numpy().astype(float), g.numpy().astype(float), float(energy), M


def compute_rgb_Hg(state, level, delta_pose, cur_I, cur_D, cur_G, K4, weight=500.0, min_grad_scale=0.0,
                   max_depth_delta=0.2, no_grad=False, robust_kernel=None, robust_k=0.01):
    """tracker.py:136-177.  state = (last_intensity pyramid, last_depth pyramid); K4 = (fx,fy,cx,cy) of the
    full-resolution camera (the reference passes the level-0 intrinsics at every pyramid level)."""
    fx, fy, cx, cy = K4
    K = np.asarray([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    KRKinv = K @ delta_pose.R @ np.linalg.inv(K)
    Kt = K @ delta_pose.t
    out = ops.rgb_odometry(state[0][level], state[1][level], cur_I[level], cur_D[level], cur_G[level],
                           [fx, fy, cx, cy], KRKinv.flatten().tolist(), Kt.flatten().tolist(),
                           min_grad_scale, max_depth_delta, not no_grad)
    f = torch.from_numpy(out[0])
    m = ~torch.isnan(f)
    f = f[m]
    Wf = f
    J = JW = None
    if not no_grad:
        J = -torch.from_numpy(out[1])[m]
        JW = J
    if robust_kernel is not None:
        w = robust_weight(f, robust_kernel, robust_k)
        Wf = Wf * w
        JW = JW * w.unsqueeze(1) if JW is not None else None
    scale = 1. / Wf.size(0) * weight
    energy = (f * Wf).sum().item() * scale
    if no_grad:
        return None, None, float(energy)
    H = torch.einsum('na,nb->nab', JW, J).sum(0) * scale
    g = (J * Wf.unsqueeze(1)).sum(0) * scale
    return H.numpy().astype(float), g.numpy().astype(float), float(energy)


def preprocess(depth, K4, subsample=0.5, divide="ieee"):
    """tracker.py:89-120 geometry half: nearest x0.5 depth, unproject, drop NaN, radius-outlier filter, PCA
    normals, drop NaN normals, 2 cm box filter.  depth (H,W) torch/np fp32 with NaN = invalid.
    Returns (points (N',3), normals (N',3)) as numpy fp32."""
    fx, fy, cx, cy = K4
    d = np.asarray(depth, dtype=np.float32)
    step = int(round(1.0 / subsample))
    d = d[::step, ::step]                                     # F.interpolate(nearest, scale 0.5) == depth[2i, 2j]
    pc = ops.unproject_depth(d, fx * subsample, fy * subsample, cx * subsample, cy * subsample).reshape(-1, 3)
    pc = pc[~np.isnan(pc[:, 0])]
    pc4 = np.concatenate([pc, np.zeros((pc.shape[0], 1), np.float32)], 1)
    pc4 = pc4[ops.remove_radius_outlier(pc4, 16, 0.05)]
    nrm = ops.estimate_normals(pc4, 16, 0.1, [0.0, 0.0, 0.0])
    ok = ~np.isnan(nrm[:, 0])
    P, Nn, _ = ops.point_box_filter(pc4[ok, :3], nrm[ok], 0.02, divide=divide)
    return P, Nn


def image_pyramid(intensity, depth):
    """tracker.py:42-57: 3-level pyramid (bilinear align_corners for intensity, nearest for depth) + Sobel/8."""
    F = torch.nn.functional
    I0 = torch.as_tensor(intensity).float()[None, None]; D0 = torch.as_tensor(depth).float()[None, None]
    h, w = I0.shape[-2:]
    I1 = F.interpolate(I0, (h // 2, w // 2), mode="bilinear", align_corners=True)
    D1 = F.interpolate(D0, (h // 2, w // 2), mode="nearest")
    I2 = F.interpolate(I1, (h // 4, w // 4), mode="bilinear", align_corners=True)
    D2 = F.interpolate(D1, (h // 4, w // 4), mode="nearest")
    Is = [t[0, 0].numpy() for t in (I0, I1, I2)]
    Ds = [t[0, 0].numpy() for t in (D0, D1, D2)]
    return Is, Ds, [ops.gradient_xy(t) for t in Is]


def gauss_newton(omap, last_pose, init_pose, obs_xyz, iter_config, rgb=None, sdf_args=None, rgb_args=None):
    """tracker.py:225-288.  rgb = dict(state=(last_I, last_D), cur=(I, D, G), K4=...) or None (rgb terms skipped
    only if the config has none).  Returns (pose, n_sdf_evals, trace) where trace lists the per-iteration energies."""
    sdf_args = sdf_args or {"robust_kernel": "huber", "robust_k": 5.0}
    rgb_args = rgb_args or {"weight": 500.0, "robust_kernel": None, "robust_k": 0.01, "min_grad_scale": 0.0, "max_depth_delta": 0.2}
    delta = last_pose.inv().dot(init_pose)
    last_delta = copy.deepcopy(delta)
    trace = []
    n_eval = 0
    for group in iter_config:
        last_energy = np.inf
        for it in list(range(group["n"])) + [-1]:
            H = np.zeros((6, 6)); g = np.zeros(6); energy = 0.0
            for term in group["type"]:
                if term[0] == "sdf":
                    h_, g_, e_, _ = compute_sdf_Hg(omap, last_pose, delta, obs_xyz, it == -1, sdf_args["robust_kernel"], sdf_args["robust_k"])
                    n_eval += 1
                elif term[0] == "rgb":
                    h_, g_, e_ = compute_rgb_Hg(rgb["state"], term[1], delta, *rgb["cur"], rgb["K4"], rgb_args["weight"],
                                                rgb_args["min_grad_scale"], rgb_args["max_depth_delta"], it == -1,
                                                rgb_args["robust_kernel"], rgb_args["robust_k"])
                else:
                    raise NotImplementedError(term[0])
                energy += e_
                if it != -1:
                    H += h_; g += g_
            trace.append(energy)
            if energy > last_energy:
                delta = last_delta
                break
            last_delta = copy.deepcopy(delta)
            last_energy = energy
            if it != -1:
                xi = np.linalg.solve(H, -g)
                delta = Pose.from_twist(xi).dot(delta)
    return last_pose.dot(delta), n_eval, trace
