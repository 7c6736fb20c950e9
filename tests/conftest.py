import importlib.util
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    import handwritten_ocr_b200 as p
    return p


@pytest.fixture(scope="session")
def synth(pkg):
    from handwritten_ocr_b200 import synth as s
    return s


@pytest.fixture(scope="session")
def image_small():
    return dict(np.load(os.path.join(GOLDEN, "image_small.npz")))


@pytest.fixture(scope="session")
def image_hashes():
    with open(os.path.join(GOLDEN, "image_hashes.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def text_golden():
    with open(os.path.join(GOLDEN, "text_golden.json")) as f:
        return json.load(f)
