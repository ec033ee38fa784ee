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

    # The 3x3 / 4-vector algebra below runs on Python floats (IEEE doubles, like numpy's float64) instead of small numpy
    # arrays: a frame needs ~10 pose operations, and as numpy calls they cost 200 us of host time -- on the critical path
    # between one frame's pose solve and the next (tools/pipeline_timeline.py); as float arithmetic ~25 us.
    def _unit(self):
        w, x, y, z = self.q.tolist()
        n = math.sqrt(w * w + x * x + y * y + z * z)
        return (w / n, x / n, y / n, z / n) if n > 0 else (w, x, y, z)

    def _rot9(self):
        w, x, y, z = self._unit()
        return (1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y))

    @property
    def rotation_matrix(self):
        return np.array(self._rot9(), dtype=np.float64).reshape(3, 3)

    @property
    def transformation_matrix(self):
        m = np.eye(4)
        m[:3, :3] = self.rotation_matrix
        return m

    @classmethod
    def _raw(cls, q4):
        o = cls.__new__(cls)
        o.q = np.array(q4, dtype=np.float64)
        return o

    @property
    def inverse(self):
        w, x, y, z = self.q.tolist()
        n2 = w * w + x * x + y * y + z * z
        return Quaternion._raw((w / n2, -x / n2, -y / n2, -z / n2))

    def __mul__(self, o):
        a0, a1, a2, a3 = self.q.tolist()
        b0, b1, b2, b3 = o.q.tolist()
        return Quaternion._raw((
            a0 * b0 - a1 * b1 - a2 * b2 - a3 * b3,
            a0 * b1 + a1 * b0 + a2 * b3 - a3 * b2,
            a0 * b2 - a1 * b3 + a2 * b0 + a3 * b1,
            a0 * b3 + a1 * b2 - a2 * b1 + a3 * b0))

    def _rotate3(self, v):
        r = self._rot9()
        x, y, z = v
        return (r[0] * x + r[1] * y + r[2] * z, r[3] * x + r[4] * y + r[5] * z, r[6] * x + r[7] * y + r[8] * z)

    def rotate(self, v):
        v = np.asarray(v, dtype=np.float64)
        if v.shape == (3,):
            return np.array(self._rotate3(v.tolist()), dtype=np.float64)
        return self.rotation_matrix @ v

    def __repr__(self):
        return f"Quaternion({self.q[0]!r}, {self.q[1]!r}, {self.q[2]!r}, {self.q[3]!r})"


def _quat_from_matrix(M):
    R = M[:3, :3].tolist()
    # the acceptance test of pyquaternion (allclose(R R^T, I) and isclose(det R, 1), rtol 1e-5, atol 1e-8), on floats
    ok = True
    for i in range(3):
        for j in range(3):
            v = R[i][0] * R[j][0] + R[i][1] * R[j][1] + R[i][2] * R[j][2]
            e = 1.0 if i == j else 0.0
            ok = ok and abs(v - e) <= 1e-8 + 1e-5 * e
    det = (R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0])
           + R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]))
    if not ok or not abs(det - 1.0) <= 1e-8 + 1e-5:
        raise ValueError("Matrix must be special orthogonal")
    R = _Rows(R)
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


class _Rows:
    """R[i, j] on a list of lists (keeps the four-branch formula below readable)."""
    __slots__ = ("r",)

    def __init__(self, r):
        self.r = r

    def __getitem__(self, ij):
        return self.r[ij[0]][ij[1]]


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

    @classmethod
    def _raw(cls, q, t3):
        o = cls.__new__(cls)
        o.q = q
        o.t = np.array(t3, dtype=np.float64)
        return o

    def inv(self):
        qi = self.q.inverse
        x, y, z = qi._rotate3(self.t.tolist())
        return Isometry._raw(qi, (-x, -y, -z))

    def dot(self, right):
        x, y, z = self.q._rotate3(right.t.tolist())
        a, b, c = self.t.tolist()
        return Isometry._raw(self.q * right.q, (x + a, y + b, z + c))

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
