"""No GPU needed: the C-ABI library loads and exports every symbol include/difusion_b200.h declares; host-only entry
points answer; the product refuses to run without CUDA (no CPU fallback)."""
import re
from pathlib import Path

import pytest
import torch

from util import MAPPING, ROOT, ns, pkg


def _declared():
    txt = (ROOT / "include" / "difusion_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dfb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    d = pkg()
    lib = d._lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
        assert n in d._lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(d._lib.SIGNATURES) == set(names)


def test_host_only_entry_points():
    lib = pkg()._lib.load()
    assert lib.dfb_version() >= 100
    assert lib.dfb_decoder_blob_floats() == 91624 + 50688 and lib.dfb_encoder_blob_floats() == 27264 + 28160
    for n in (0, 1, 77000):
        assert lib.dfb_pcproc_ws_bytes(n) > 0 and lib.dfb_box_filter_ws_bytes(n) > 0
        assert lib.dfb_integrate_ws_bytes(n, 256000) >= n * (12 + 4 + 8 * 32)
    assert lib.dfb_decode_cubes_ws_bytes(1000, 4) >= 1000 * 64 * 8
    # argument validation happens before any CUDA call
    assert lib.dfb_unproject_depth(None, 4, 4, 1.0, 1.0, 0.0, 0.0, None, None) == -1
    assert lib.dfb_unproject_depth(None, 0, 4, 1.0, 1.0, 0.0, 0.0, None, None) == 0          # empty input is a no-op
    assert b"unproject" in lib.dfb_last_error()


def test_no_cpu_fallback(weights):
    d = pkg()
    with pytest.raises(RuntimeError):
        d.DenseIndexedMap(weights, ns(dict(MAPPING)), 29, torch.device("cpu"))
    with pytest.raises(RuntimeError):
        d.ext.unproject_depth(torch.zeros(4, 4), 1, 1, 0, 0)


def test_product_never_imports_oracle():
    for f in (ROOT / "nerf-fusion_b200").rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f
