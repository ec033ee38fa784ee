import importlib
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def dfb():
    return importlib.import_module("nerf-fusion_b200")


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


@pytest.fixture(scope="session")
def weights(dfb, golden_dir):
    return dfb.weights.load_npz(golden_dir / "weights.npz")


# Engine configurations the GPU parity tests run under.  "default" is what bench.py measures (tcgen05 decoder with hi+lo
# FP16 operands, tcgen05 encoder); "fp32" is the CUDA-core FP32 pair, kept as a cross-check.  Module-scoped so that
# module-scoped maps are built once per configuration.
ENGINES = {"default": (1, 1), "fp32": (0, 0)}


@pytest.fixture(scope="module", params=["default", "fp32"])
def engine(request):
    return request.param


@pytest.fixture()
def use_engine(engine, dfb):
    lib = dfb._lib.load()
    dec, enc = ENGINES[engine]
    lib.dfb_set_decoder_engine(dec); lib.dfb_set_encoder_engine(enc)
    yield engine
    lib.dfb_set_decoder_engine(1); lib.dfb_set_encoder_engine(1)
