"""RGB-D sequence readers for the per-frame path (dataset/production/__init__.py:21-40, icl_nuim.py:54-124).

The reference decodes each PNG with OpenCV, converts to float32 on the host and uploads 16 bytes per pixel.  Here the
decoded 16-bit depth and 8-bit colour images are copied to the GPU as they are (5 bytes per pixel, from pinned buffers)
and ``ext.ingest_frame`` does scale, BGR -> RGB and (optionally) the depth clipping of main.py:56-57 in one kernel.
"""
import os
from pathlib import Path

import numpy as np
import torch

from . import ext
from .motion import Isometry, Quaternion
from .tracker import FrameIntrinsic


class FrameData:
    """dataset/production/__init__.py:21-26."""

    def __init__(self):
        self.rgb = None
        self.depth = None
        self.gt_pose = None
        self.calib = None


class RGBDSequence:
    """dataset/production/__init__.py:29-40."""

    def __init__(self):
        self.frame_id = 0

    def __iter__(self):
        return self

    def __len__(self):
        raise NotImplementedError

    def __next__(self):
        raise NotImplementedError


def read_raw(depth_path, color_path):
    """Host half of icl_nuim.py:106-112: decoded images as stored (uint16 depth, uint8 BGR colour)."""
    import cv2
    depth = cv2.imread(str(depth_path), cv2.IMREAD_UNCHANGED)
    color = cv2.imread(str(color_path))
    if depth is None or color is None:
        raise FileNotFoundError(f"cannot decode {depth_path} / {color_path}")
    if depth.dtype != np.uint16:
        raise ValueError(f"{depth_path}: expected a 16-bit depth PNG, got {depth.dtype}")
    return depth, color


def parse_traj_file(traj_path):
    """icl_nuim.py:86-99: TUM-format ground truth (index tx ty tz qx qy qz qw) -> list of Isometry in the reference's frame
    convention (y flipped, 180 degrees about z)."""
    camera_ext = {}
    traj_data = np.genfromtxt(traj_path)
    cano_quat = Isometry(q=Quaternion(axis=[0.0, 0.0, 1.0], degrees=180.0))
    for cur_p in traj_data:
        qx, qy, qz, qw = cur_p[4], cur_p[5], cur_p[6], cur_p[7]
        cur_q = Quaternion(array=[qw, qx, qy, qz]).rotation_matrix.copy()
        cur_t = cur_p[1:4].copy()
        cur_q[1] = -cur_q[1]
        cur_q[:, 1] = -cur_q[:, 1]
        cur_t[1] = -cur_t[1]
        camera_ext[int(cur_p[0])] = cano_quat.dot(Isometry(q=Quaternion(matrix=cur_q), t=cur_t))
    camera_ext[0] = camera_ext[1]
    return [camera_ext[t] for t in range(len(camera_ext))]


class ICLNUIMSequence(RGBDSequence):
    """icl_nuim.py:54-124 (same constructor, same FrameData).  `device` (new) selects the GPU; `depth_cut` (new, optional)
    applies main.py:56-57 inside the ingest kernel."""

    def __init__(self, path: str, start_frame: int = 0, end_frame: int = -1, first_tq: list = None, load_gt: bool = False,
                 device="cuda:0", depth_cut=None):
        super().__init__()
        self.path = Path(path)
        self.device = torch.device(device)
        self.depth_cut = depth_cut
        self.color_names = sorted([f"rgb/{t}" for t in os.listdir(self.path / "rgb")], key=lambda t: int(t[4:].split(".")[0]))
        self.depth_names = [f"depth/{t}.png" for t in range(len(self.color_names))]
        self.calib = [481.2, 480.0, 319.50, 239.50, 5000.0]
        if first_tq is not None:
            self.first_iso = Isometry(q=Quaternion(array=first_tq[3:]), t=np.array(first_tq[:3]))
        else:
            self.first_iso = Isometry(q=Quaternion(array=[0.0, -1.0, 0.0, 0.0]))
        if end_frame == -1:
            end_frame = len(self.color_names)
        self.color_names = self.color_names[start_frame:end_frame]
        self.depth_names = self.depth_names[start_frame:end_frame]
        if load_gt:
            gt_traj_path = (list(self.path.glob("*.freiburg")) + list(self.path.glob("groundtruth.txt")))[0]
            traj = parse_traj_file(gt_traj_path)[start_frame:end_frame]
            change_iso = self.first_iso.dot(traj[0].inv())
            self.gt_trajectory = [change_iso.dot(t) for t in traj]
            assert len(self.gt_trajectory) == len(self.color_names)
        else:
            self.gt_trajectory = None
        self._pin_d = self._pin_c = None

    def __len__(self):
        return len(self.color_names)

    def _pinned(self, depth, color):
        if self._pin_d is None or self._pin_d.shape != depth.shape:
            self._pin_d = torch.empty(depth.shape, dtype=torch.uint16).pin_memory()
            self._pin_c = torch.empty(color.shape, dtype=torch.uint8).pin_memory()
        self._pin_d.numpy()[...] = depth
        self._pin_c.numpy()[...] = color
        return self._pin_d, self._pin_c

    def __next__(self):
        if self.frame_id >= len(self):
            raise StopIteration
        depth, color = read_raw(self.path / self.depth_names[self.frame_id], self.path / self.color_names[self.frame_id])
        pd, pc = self._pinned(depth, color)
        d_dev = pd.to(self.device, non_blocking=True)
        c_dev = pc.to(self.device, non_blocking=True)
        frame_data = FrameData()
        frame_data.depth, frame_data.rgb = ext.ingest_frame(d_dev, c_dev, self.calib[4], self.depth_cut, bgr=True)
        torch.cuda.current_stream(self.device).synchronize()        # the pinned staging buffers are reused by the next frame
        frame_data.gt_pose = self.gt_trajectory[self.frame_id] if self.gt_trajectory is not None else None
        frame_data.calib = FrameIntrinsic(self.calib[0], self.calib[1], self.calib[2], self.calib[3], self.calib[4])
        self.frame_id += 1
        return frame_data
