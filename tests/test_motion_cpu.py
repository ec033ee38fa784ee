"""Host-side pose algebra (nerf-fusion_b200/motion.py) against the reference's own utils/motion_util.py (imported unmodified
under oracle/ref_shims.py when /root/reference is present) and against the oracle restatement (always)."""
import numpy as np
import pytest

from oracle import ref_shims
from oracle.tracker_oracle import Pose
from oracle.pyquat_shim import Quaternion as OQ
from util import pkg


def _rand_pose(rng, mod, quat):
    q = rng.randn(4); q /= np.linalg.norm(q)
    return mod(q=quat(array=q), t=rng.randn(3) * 2)


def test_isometry_matches_the_oracle_restatement():
    d = pkg()
    rng = np.random.RandomState(0)
    for _ in range(50):
        seed = rng.randint(1 << 30)
        a, b = _rand_pose(np.random.RandomState(seed), d.Isometry, d.Quaternion), _rand_pose(np.random.RandomState(seed + 1), d.Isometry, d.Quaternion)
        oa, ob = _rand_pose(np.random.RandomState(seed), Pose, OQ), _rand_pose(np.random.RandomState(seed + 1), Pose, OQ)
        for mine, ref in ((a.dot(b), oa.dot(ob)), (a.inv(), oa.inv()), (a.inv().dot(b), oa.inv().dot(ob))):
            assert np.abs(mine.q.rotation_matrix - ref.R).max() < 1e-14 and np.abs(mine.t - ref.t).max() < 1e-14
        xi = rng.randn(6) * np.array([0.05, 0.05, 0.05, 0.02, 0.02, 0.02])
        for x in (xi, xi * 1e-10, np.zeros(6)):                               # generic, first-order branch (angle ~ 0), identity
            m, r = d.Isometry.from_twist(x), Pose.from_twist(x)
            assert np.abs(m.q.rotation_matrix - r.R).max() < 1e-14 and np.abs(m.t - r.t).max() < 1e-14
        pts = rng.randn(7, 3)
        assert np.abs((a @ pts) - (pts @ oa.R.T + oa.t)).max() < 1e-13
        assert np.abs((a @ pts[0]) - (oa.R @ pts[0] + oa.t)).max() < 1e-13
        m4 = a.matrix
        back = d.Isometry.from_matrix(m4)
        assert np.abs(back.q.rotation_matrix - a.q.rotation_matrix).max() < 1e-13 and np.abs(back.t - a.t).max() < 1e-15
        assert np.abs(a.rotation.t).max() == 0.0


@pytest.mark.skipif(not ref_shims.available(), reason="needs /root/reference (build container only)")
def test_isometry_matches_the_reference_module():
    d = pkg()
    ref_shims.install()
    import importlib
    mu = importlib.import_module("utils.motion_util")
    from pyquaternion import Quaternion as RQ                                 # the shim the reference imports
    rng = np.random.RandomState(1)
    for _ in range(30):
        seed = rng.randint(1 << 30)
        a, b = _rand_pose(np.random.RandomState(seed), d.Isometry, d.Quaternion), _rand_pose(np.random.RandomState(seed + 1), d.Isometry, d.Quaternion)
        ra, rb = _rand_pose(np.random.RandomState(seed), mu.Isometry, RQ), _rand_pose(np.random.RandomState(seed + 1), mu.Isometry, RQ)
        for mine, ref in ((a.dot(b), ra.dot(rb)), (a.inv(), ra.inv())):
            assert np.abs(mine.q.rotation_matrix - ref.q.rotation_matrix).max() < 1e-14 and np.abs(mine.t - ref.t).max() < 1e-14
        xi = rng.randn(6) * 0.03
        m, r = d.Isometry.from_twist(xi), mu.Isometry.from_twist(xi)
        assert np.abs(m.q.rotation_matrix - r.q.rotation_matrix).max() < 1e-14 and np.abs(m.t - r.t).max() < 1e-14
        pts = rng.randn(5, 3)
        # (the reference's own `@` takes its torch branch for numpy >= 2 arrays -- ndarray.device exists now -- and fails there)
        assert np.abs((a @ pts) - (pts @ ra.q.rotation_matrix.T + ra.t[np.newaxis, :])).max() < 1e-13
