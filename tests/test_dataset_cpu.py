"""Frame-ingest row on the CPU: the oracle restatement and the host half of the reader against vectors produced by the
reference's own ICLNUIMSequence (oracle/make_dataset_golden.py)."""
import numpy as np
import pytest

from oracle import ops
from util import GOLD, pkg


@pytest.fixture(scope="module")
def D():
    return dict(np.load(GOLD / "dataset_golden.npz"))


def _write_sequence(D, td):
    import cv2
    (td / "rgb").mkdir(); (td / "depth").mkdir()
    (td / "groundtruth.txt").write_bytes(D["traj_text"].tobytes())
    for i in range(int(D["n_frames"])):
        assert cv2.imwrite(str(td / "depth" / f"{i}.png"), D[f"f{i}_depth_u16"])
        assert cv2.imwrite(str(td / "rgb" / f"{i}.png"), D[f"f{i}_bgr_u8"])


def test_oracle_ingest_equals_reference_reader(D):
    for i in range(int(D["n_frames"])):
        d, c = ops.ingest_frame(D[f"f{i}_depth_u16"], D[f"f{i}_bgr_u8"], 5000.0, None, bgr=True, recip=False)   # torch CPU divides
        assert np.array_equal(d, D[f"f{i}_depth"]) and np.array_equal(c, D[f"f{i}_rgb"])
        d2, _ = ops.ingest_frame(D[f"f{i}_depth_u16"], None, 5000.0, (0.5, 5.0), recip=False)
        ref = D[f"f{i}_depth"].copy(); ref[(ref < 0.5) | (ref > 5.0)] = np.nan                                 # main.py:56-57
        assert np.array_equal(np.isnan(d2), np.isnan(ref)) and np.array_equal(d2[~np.isnan(ref)], ref[~np.isnan(ref)])


def test_reader_host_half_and_trajectory(D, tmp_path):
    pytest.importorskip("cv2")
    d = pkg()
    _write_sequence(D, tmp_path)
    depth, color = d.dataset.read_raw(tmp_path / "depth" / "1.png", tmp_path / "rgb" / "1.png")
    assert depth.dtype == np.uint16 and np.array_equal(depth, D["f1_depth_u16"]) and np.array_equal(color, D["f1_bgr_u8"])
    traj = d.dataset.parse_traj_file(tmp_path / "groundtruth.txt")
    first_tq = D["first_tq"]
    first = d.Isometry(q=d.Quaternion(array=first_tq[3:]), t=np.array(first_tq[:3]))
    change = first.dot(traj[0].inv())
    for i in range(int(D["n_frames"])):
        p = change.dot(traj[i])
        assert np.abs(p.q.rotation_matrix - D[f"f{i}_gt_R"]).max() < 1e-12 and np.abs(p.t - D[f"f{i}_gt_t"]).max() < 1e-12
    with pytest.raises(FileNotFoundError):
        d.dataset.read_raw(tmp_path / "depth" / "99.png", tmp_path / "rgb" / "99.png")
