"""Synthetic ICL-NUIM-shaped RGB-D stream (replaces dataset/production/icl_nuim.py:54-123 for benchmarks and
parity runs; there is no dataset in the image).  Analytic ray casting of a box room with a few boxes and
spheres, procedural texture, pinhole fx 481.2 fy 480 cx 319.5 cy 239.5 (icl_nuim.py:60), depth quantised to
1/5000 m like the PNGs (icl_nuim.py:111), z^2 noise and dropout.  Pure torch, runs on CPU or CUDA; this is
input synthesis (plumbing), never part of a timed region.
"""
import math

import numpy as np
import torch

ICL_CALIB = (481.2, 480.0, 319.5, 239.5)
FIRST_TQ = [-1.4, 1.5, 1.5, 0.0, -1.0, 0.0, 0.0]                 # configs/fusion-lr-kt.yaml:7 (t, then q = w,x,y,z)
BOUND_MIN = [-3.5, -0.5, -2.5]
BOUND_MAX = [4.5, 3.5, 5.5]


def quat_to_R(q):
    w, x, y, z = [float(t) for t in q]
    n = math.sqrt(w * w + x * x + y * y + z * z)
    w, x, y, z = w / n, x / n, y / n, z / n
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=np.float64)


def _rot(axis, ang):
    axis = np.asarray(axis, dtype=np.float64); axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)


class SyntheticRoom:
    """Room = inside of an axis-aligned box (bounds minus margin) + solid boxes + spheres."""

    def __init__(self, bound_min=BOUND_MIN, bound_max=BOUND_MAX, margin=0.3, seed=0):
        rng = np.random.RandomState(seed)
        self.lo = np.asarray(bound_min, dtype=np.float64) + margin
        self.hi = np.asarray(bound_max, dtype=np.float64) - margin
        # furniture in front of the start camera (which looks towards -z from (-1.4, 1.5, 1.5))
        self.boxes = [
            (np.array([-2.6, self.lo[1], -1.6]), np.array([-1.6, self.lo[1] + 0.9, -0.7])),
            (np.array([-0.9, self.lo[1], -2.0]), np.array([0.3, self.lo[1] + 1.6, -1.5])),
            (np.array([0.8, self.lo[1], -1.2]), np.array([1.6, self.lo[1] + 0.6, -0.2])),
        ]
        self.spheres = [(np.array([-1.5, 1.2, -1.3]), 0.45), (np.array([-0.2, 2.2, -1.0]), 0.35)]
        self.tex_phase = rng.uniform(0, 2 * np.pi, size=(6, 3))
        self.tex_freq = rng.uniform(2.0, 9.0, size=(6, 3))

    def _texture(self, p):
        """Sum of sinusoids in world position -> (N,3) rgb in [0,1]; smooth so that Sobel gradients are well-posed."""
        out = []
        for c in range(3):
            v = 0.0
            for k in range(2):
                f = torch.tensor(self.tex_freq[2 * c + k], dtype=p.dtype, device=p.device)
                ph = torch.tensor(self.tex_phase[2 * c + k], dtype=p.dtype, device=p.device)
                v = v + torch.sin(p * f + ph).sum(-1)
            out.append(0.5 + v / 12.0)
        return torch.stack(out, -1).clamp(0, 1)

    def render(self, R, t, H=480, W=640, calib=ICL_CALIB, device="cpu"):
        """Returns ideal depth (H,W) [z-depth in camera frame] and rgb (H,W,3), float32."""
        fx, fy, cx, cy = calib
        dt = torch.float64
        u = torch.arange(W, device=device, dtype=dt)[None, :].expand(H, W)
        v = torch.arange(H, device=device, dtype=dt)[:, None].expand(H, W)
        dc = torch.stack([(u - cx) / fx, (v - cy) / fy, torch.ones_like(u)], -1).reshape(-1, 3)   # z = 1 rays
        Rt = torch.tensor(R, dtype=dt, device=device); o = torch.tensor(t, dtype=dt, device=device)
        d = dc @ Rt.T
        inf = torch.full((d.shape[0],), float("inf"), dtype=dt, device=device)
        best = inf.clone()
        # room walls (inside-out box): far slab hit
        lo = torch.tensor(self.lo, dtype=dt, device=device); hi = torch.tensor(self.hi, dtype=dt, device=device)
        inv = 1.0 / torch.where(d.abs() < 1e-12, torch.full_like(d, 1e-12), d)
        t1 = (lo - o) * inv; t2 = (hi - o) * inv
        tfar = torch.maximum(t1, t2).min(-1).values
        best = torch.minimum(best, torch.where(tfar > 0, tfar, inf))
        for blo, bhi in self.boxes:
            bl = torch.tensor(blo, dtype=dt, device=device); bh = torch.tensor(bhi, dtype=dt, device=device)
            a = (bl - o) * inv; b = (bh - o) * inv
            tn = torch.minimum(a, b).max(-1).values; tf = torch.maximum(a, b).min(-1).values
            hit = (tn <= tf) & (tn > 0)
            best = torch.minimum(best, torch.where(hit, tn, inf))
        for c, rad in self.spheres:
            cc = torch.tensor(c, dtype=dt, device=device)
            oc = o - cc
            bq = (d * oc).sum(-1); aq = (d * d).sum(-1); cq = (oc * oc).sum() - rad * rad
            disc = bq * bq - aq * cq
            ts = (-bq - torch.sqrt(disc.clamp(min=0))) / aq
            hit = (disc > 0) & (ts > 0)
            best = torch.minimum(best, torch.where(hit, ts, inf))
        p = o + d * best.unsqueeze(-1)
        rgb = self._texture(p.float()).reshape(H, W, 3)
        depth = best.reshape(H, W).float()                     # rays have unit camera-z, so t == z-depth
        return depth, rgb


class SyntheticSequence:
    """Iterator with the reference's RGBDSequence surface (frame_id, first_iso-like start pose, __next__ ->
    depth (H,W) f32 metres, rgb (H,W,3) f32 in [0,1], calib).  Poses are (R (3,3), t (3,)) float64."""

    def __init__(self, n_frames=20, H=480, W=640, seed=0, device="cpu", noise=True, first_tq=FIRST_TQ,
                 step_trans=0.012, step_rot_deg=0.3):
        self.n_frames, self.H, self.W = n_frames, H, W
        self.device = device
        self.room = SyntheticRoom(seed=seed)
        self.calib = ICL_CALIB
        self.noise = noise
        self.gen = torch.Generator(device="cpu"); self.gen.manual_seed(seed)
        R0 = quat_to_R(first_tq[3:]); t0 = np.asarray(first_tq[:3], dtype=np.float64)
        self.poses = []
        for i in range(n_frames):
            s = i / 60.0
            # smooth Lissajous-like drift, <= step_trans m and <= step_rot_deg deg per frame
            t = t0 + step_trans * 60.0 * np.array([0.35 * math.sin(2.1 * s), 0.12 * math.sin(1.3 * s), 0.30 * (1 - math.cos(1.7 * s))]) / 2.1
            ang = math.radians(step_rot_deg) * 60.0 * np.array([0.25 * math.sin(1.1 * s), 0.45 * math.sin(1.9 * s), 0.1 * math.sin(0.7 * s)]) / 1.9
            Rw = _rot([0, 1, 0], ang[1]) @ _rot([1, 0, 0], ang[0]) @ _rot([0, 0, 1], ang[2])
            self.poses.append((Rw @ R0, t))
        self.frame_id = 0

    def __len__(self):
        return self.n_frames

    def __iter__(self):
        return self

    def frame(self, i):
        R, t = self.poses[i]
        depth, rgb = self.room.render(R, t, self.H, self.W, self.calib, self.device)
        if self.noise:
            g = torch.Generator(device="cpu"); g.manual_seed(1000 + i)
            n = torch.randn(depth.shape, generator=g).to(depth.device)
            drop = (torch.rand(depth.shape, generator=g) < 0.02).to(depth.device)
            depth = depth + 0.001 * depth * depth * n
            depth = torch.where(drop, torch.zeros_like(depth), depth)
        depth = torch.where(torch.isfinite(depth), depth, torch.zeros_like(depth))
        depth = torch.round(depth.clamp(0, 13.0) * 5000.0) / 5000.0      # uint16 PNG quantisation
        return depth.contiguous(), rgb.contiguous()

    def __next__(self):
        if self.frame_id >= self.n_frames:
            raise StopIteration
        out = self.frame(self.frame_id)
        self.frame_id += 1
        return out
