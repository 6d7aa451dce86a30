import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def native():
    """The product library, built in-tree (fails loudly if it cannot be)."""
    from astc_encoder_b200 import build as _build
    _build.build()
    import astc_encoder_b200 as A
    A.lib()
    return A


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def leaf_rgba(native):
    """leaf.png as the reference loads it: RGBA8, vertically flipped (main.cpp:24-25)."""
    return native.load_image(str(GOLDEN / "leaf.png"), True)


@pytest.fixture(scope="session")
def leaf_golden(native):
    return native.load_astc(str(GOLDEN / "leaf.astc"))
