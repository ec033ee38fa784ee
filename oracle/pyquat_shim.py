"""Minimal stand-in for the third-party ``pyquaternion`` package (un-vendored dependency of the
reference, unpinned in requirements.txt:2-9; absent from this image).  TEST INFRASTRUCTURE ONLY.

Restates pyquaternion's published algorithms for the subset of the API the reference's
``utils/motion_util.py`` uses: ``Quaternion()``, ``(matrix=)``, ``(array=)``, ``(axis=, degrees=/angle=)``,
``.rotation_matrix``, ``.transformation_matrix``, ``.rotate``, ``.inverse``, ``*``, ``.q``.
"""
import numpy as np


class Quaternion:
    def __init__(self, *args, **kw):
        if "matrix" in kw:
            self.q = self._from_matrix(np.asarray(kw["matrix"], dtype=np.float64))
        elif "array" in kw:
            self.q = np.asarray(kw["array"], dtype=np.float64).copy()
        elif "axis" in kw:
            axis = np.asarray(kw["axis"], dtype=np.float64)
            if "degrees" in kw:
                angle = np.deg2rad(kw["degrees"])
            else:
                angle = kw.get("angle", kw.get("radians", 0.0))
            axis = axis / np.linalg.norm(axis)
            self.q = np.concatenate([[np.cos(angle / 2.0)], axis * np.sin(angle / 2.0)])
        elif "real" in kw or "imaginary" in kw:
            self.q = np.concatenate([[kw.get("real", 0.0)], np.asarray(kw.get("imaginary", [0, 0, 0]), dtype=np.float64)])
        elif "vector" in kw:
            self.q = np.concatenate([[0.0], np.asarray(kw["vector"], dtype=np.float64)])
        elif len(args) == 4:
            self.q = np.asarray(args, dtype=np.float64)
        elif len(args) == 1 and isinstance(args[0], Quaternion):
            self.q = args[0].q.copy()
        elif len(args) == 1:
            self.q = np.asarray(args[0], dtype=np.float64).copy()
        else:
            self.q = np.array([1.0, 0.0, 0.0, 0.0])

    @staticmethod
    def _from_matrix(matrix):
        if matrix.shape == (4, 4):
            R = matrix[:3, :3]
        else:
            R = matrix
        if not np.allclose(R @ R.conj().T, np.eye(3), rtol=1e-5, atol=1e-8):
            raise ValueError("Matrix must be orthogonal, i.e. its transpose should be its inverse")
        if not np.isclose(np.linalg.det(R), 1.0, rtol=1e-5, atol=1e-8):
            raise ValueError("Matrix must be special orthogonal i.e. its determinant must be +1.0")
        m = R.conj().T
        if m[2, 2] < 0:
            if m[0, 0] > m[1, 1]:
                t = 1 + m[0, 0] - m[1, 1] - m[2, 2]
                q = [m[1, 2] - m[2, 1], t, m[0, 1] + m[1, 0], m[2, 0] + m[0, 2]]
            else:
                t = 1 - m[0, 0] + m[1, 1] - m[2, 2]
                q = [m[2, 0] - m[0, 2], m[0, 1] + m[1, 0], t, m[1, 2] + m[2, 1]]
        else:
            if m[0, 0] < -m[1, 1]:
                t = 1 - m[0, 0] - m[1, 1] + m[2, 2]
                q = [m[0, 1] - m[1, 0], m[2, 0] + m[0, 2], m[1, 2] + m[2, 1], t]
            else:
                t = 1 + m[0, 0] + m[1, 1] + m[2, 2]
                q = [t, m[1, 2] - m[2, 1], m[2, 0] - m[0, 2], m[0, 1] - m[1, 0]]
        q = np.array(q, dtype=np.float64)
        q *= 0.5 / np.sqrt(t)
        return q

    # -- algebra --
    def _q_matrix(self):
        w, x, y, z = self.q
        return np.array([[w, -x, -y, -z], [x, w, -z, y], [y, z, w, -x], [z, -y, x, w]])

    def _q_bar_matrix(self):
        w, x, y, z = self.q
        return np.array([[w, -x, -y, -z], [x, w, z, -y], [y, -z, w, x], [z, y, -x, w]])

    def __mul__(self, other):
        if isinstance(other, Quaternion):
            return Quaternion(array=self._q_matrix() @ other.q)
        return Quaternion(array=self.q * float(other))

    @property
    def conjugate(self):
        return Quaternion(array=np.concatenate([[self.q[0]], -self.q[1:]]))

    @property
    def inverse(self):
        ss = float(np.dot(self.q, self.q))
        return Quaternion(array=self.conjugate.q / ss)

    def _normalise(self):
        n = np.linalg.norm(self.q)
        if n > 0 and not abs(1.0 - n) < 1e-14:
            self.q = self.q / n

    @property
    def normalised(self):
        q = Quaternion(array=self.q)
        q._normalise()
        return q

    @property
    def rotation_matrix(self):
        self._normalise()
        pm = self._q_matrix() @ self._q_bar_matrix().conj().T
        return pm[1:][:, 1:]

    @property
    def transformation_matrix(self):
        t = np.eye(4)
        t[:3, :3] = self.rotation_matrix
        return t

    def rotate(self, vector):
        self._normalise()
        v = Quaternion(vector=np.asarray(vector, dtype=np.float64))
        r = self * v * self.conjugate
        return r.q[1:]

    def __repr__(self):
        return "Quaternion({!r}, {!r}, {!r}, {!r})".format(*self.q)
