import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
# under test every device-planned phase window is cross-checked against the host's index arithmetic (segment.py)
os.environ.setdefault("MS_B200_VERIFY_PLAN", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the reference checkout under /root/reference")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def variants_table():
    return json.load(open(os.path.join(GOLDEN, "variants.json")))


def load_npz_u64(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)
