import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def _parity_path():
    return os.environ.get("FD_PARITY_JSON") or os.path.join(ROOT, "gpurun_out", "r02_parity.json")


@pytest.fixture
def parity_record():
    """record(name, **numbers): every measured parity error of the -m gpu suite goes into one JSON file (default
    gpurun_out/r02_parity.json, which travels back from the GPU box; the committed copy is profiles/r02_parity.json) BEFORE
    the test asserts on it, so a failing run still leaves its numbers behind."""
    import json

    def record(name, **numbers):
        path = _parity_path()
        os.makedirs(os.path.dirname(path), exist_ok=True)
        try:
            data = json.load(open(path))
        except Exception:
            data = {}
        data[name] = numbers
        with open(path, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    return record
