import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:      # tests/host_harness.py
    sys.path.insert(1, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


try:                                       # OpenCV must be fully imported before any test rearranges sys.path / sys.modules for the
    import cv2  # noqa: F401               # reference checkout: a first import in that state fails inside cv2.typing (order dependent)
except Exception:
    pass


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    try:                                   # the oracle is many small torch ops: 64 threads thrash
        import torch
        torch.set_num_threads(min(8, os.cpu_count() or 1))
    except Exception:
        pass


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def golden_field():
    """The seeded field every golden fixture was generated with (regenerated, not stored)."""
    import numpy as np
    import oracle
    f = np.load(os.path.join(GOLDEN, "field.npz"))
    return oracle.Field.random(int(f["seed"]), float(f["grid_scale"]))
