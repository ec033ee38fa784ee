"""SE(3) poses for the tracker: the subset of utils/motion_util.py (Isometry, :163-340) and of the un-vendored
``pyquaternion`` package that the hot path touches, float64 on the host exactly like the reference (the 6x6
Gauss-Newton solve and pose composition stay in f64 so poses match to ~1e-12 given the same H, g).

API kept: Isometry(q=, t=), .q.rotation_matrix, .t, .rotation, .matrix, .inv(), .dot(), from_twist(),
from_matrix(), ``iso @ x`` for torch (N,3) tensors / numpy arrays / Isometry.
"""
import math

import numpy as np


class Quaternion:
    """Unit quaternion (w, x, y, z).  Constructors used by the reference: (), (array=), (matrix=), (axis=, degrees=)."""

    __slots__ = ("q",)

    def __init__(self, *args, array=None, matrix=None, axis=None, degrees=None, angle=None):
        if matrix is not None:
            self.q = _quat_from_matrix(np.asarray(matrix, dtype=np.float64))
        elif array is not None:
            self.q = np.asarray(array, dtype=np.float64).copy()
        elif axis is not None:
            ax = np.asarray(axis, dtype=np.float64)
            ax = ax / np.linalg.norm(ax)
            ang = math.radians(degrees) if degrees is not None else float(angle or 0.0)
            self.q = np.concatenate([[math.cos(ang / 2)], ax * math.sin(ang / 2)])
        elif len(args) == 4:
            self.q = np.asarray(args, dtype=np.float64)
        else:
            self.q = np.array([1.0, 0.0, 0.0, 0.0])

    def _unit(self):
        n = np.linalg.norm(self.q)
        return self.q / n if n > 0 else self.q

    @property
    def rotation_matrix(self):
        w, x, y, z = self._unit()
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], dtype=np.float64)

    @property
    def transformation_matrix(self):
        m = np.eye(4)
        m[:3, :3] = self.rotation_matrix
        return m

    @property
    def inverse(self):
        w, x, y, z = self.q
        return Quaternion(array=np.array([w, -x, -y, -z]) / float(np.dot(self.q, self.q)))

    def __mul__(self, o):
        a, b = self.q, o.q
        return Quaternion(array=np.array([
            a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
            a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
            a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]]))

    def rotate(self, v):
        return self.rotation_matrix @ np.asarray(v, dtype=np.float64)

    def __repr__(self):
        return f"Quaternion({self.q[0]!r}, {self.q[1]!r}, {self.q[2]!r}, {self.q[3]!r})"


def _quat_from_matrix(M):
    R = M[:3, :3]
    if not np.allclose(R @ R.T, np.eye(3), rtol=1e-5, atol=1e-8) or not np.isclose(np.linalg.det(R), 1.0, rtol=1e-5, atol=1e-8):
        raise ValueError("Matrix must be special orthogonal")      # same acceptance test as pyquaternion
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    return np.asarray(q, dtype=np.float64)


def so3_wedge(p):
    return np.array([[0.0, -p[2], p[1]], [p[2], 0.0, -p[0]], [-p[1], p[0], 0.0]])


class Isometry:
    def __init__(self, q=None, t=None):
        self.q = q if q is not None else Quaternion()
        t = np.zeros(3) if t is None else np.asarray(t, dtype=np.float64)
        assert t.shape == (3,)
        self.t = t

    def __repr__(self):
        return f"Isometry: t = {self.t}, q = {self.q}"

    @property
    def rotation(self):
        return Isometry(q=self.q)

    @property
    def matrix(self):
        m = self.q.transformation_matrix
        m[:3, 3] = self.t
        return m

    @staticmethod
    def from_matrix(mat, t_component=None):
        mat = np.asarray(mat, dtype=np.float64)
        if t_component is None:
            return Isometry(q=Quaternion(matrix=mat[:3, :3]), t=mat[:3, 3])
        return Isometry(q=Quaternion(matrix=mat), t=t_component)

    @staticmethod
    def from_twist(xi):
        """motion_util.py:205-228: Rodrigues rotation + left Jacobian (first-order branch when the angle ~ 0)."""
        xi = np.asarray(xi, dtype=np.float64)
        rho, phi = xi[:3], xi[3:6]
        angle = np.linalg.norm(phi)
        if np.isclose(angle, 0.0):
            R = np.identity(3) + so3_wedge(phi)
            J = np.identity(3) + 0.5 * so3_wedge(phi)
        else:
            ax = phi / angle
            s, c = math.sin(angle), math.cos(angle)
            R = c * np.identity(3) + (1 - c) * np.outer(ax, ax) + s * so3_wedge(ax)
            J = (s / angle) * np.identity(3) + (1 - s / angle) * np.outer(ax, ax) + ((1 - c) / angle) * so3_wedge(ax)
        return Isometry(q=Quaternion(matrix=R), t=J @ rho)

    def inv(self):
        qi = self.q.inverse
        return Isometry(q=qi, t=-(qi.rotate(self.t)))

    def dot(self, right):
        return Isometry(q=self.q * right.q, t=self.q.rotate(right.t) + self.t)

    def torch_matrices(self, device):
        import torch
        return (torch.from_numpy(self.q.rotation_matrix).to(device).float(), torch.from_numpy(self.t).to(device).float())

    def __matmul__(self, other):
        # torch (N,3): other @ R^T + t in fp32 (motion_util.py:323-328).  The reference tests hasattr(other, "device"), which
        # numpy >= 2 arrays also have; is_cuda is torch-only.
        if hasattr(other, "is_cuda"):
            assert other.ndim == 2 and other.size(1) == 3
            import torch
            if other.is_cuda and other.dtype == torch.float32:
                from . import ext                          # one small kernel instead of a cuBLAS GEMM + add (see csrc/preprocess.cu)
                return ext.transform_points(other.contiguous(), self.q.rotation_matrix, self.t)
            R, t = self.torch_matrices(other.device)
            return other @ R.t() + t.unsqueeze(0)
        if isinstance(other, Isometry):
            return self.dot(other)
        other = np.asarray(other)
        if other.ndim == 1:
            return self.q.rotate(other) + self.t
        return other @ self.q.rotation_matrix.T + self.t[np.newaxis, :]
