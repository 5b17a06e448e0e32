import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


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


@pytest.fixture(autouse=True, scope="session")
def pin_exact_scorer():
    """The parity suite compares top-K ids and scores BIT FOR BIT with the oracle, which is defined for the float32 FMA-chain
    scorer (score_mode "exact").  The product default is "auto" (tensor-core 3xTF32 scorer where the shape allows); tests
    of that scorer ask for it explicitly (tests/test_fullsort_eval_gpu.py, smoke())."""
    try:
        import recbole_fairrec_b200 as pkg
    except Exception:
        yield
        return
    old = pkg.config.DEFAULTS["score_mode"]
    pkg.config.DEFAULTS["score_mode"] = "exact"
    yield
    pkg.config.DEFAULTS["score_mode"] = old
