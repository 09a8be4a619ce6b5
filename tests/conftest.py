import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a box without CUDA or without the built library."""
    try:
        import torch
        have_cuda = torch.cuda.is_available()
    except Exception:
        have_cuda = False
    have_lib = os.path.exists(os.path.join(ROOT, "repo_b200", "librepo_b200.so"))
    if have_cuda and have_lib:
        return
    skip = pytest.mark.skip(reason="needs CUDA" if not have_cuda else "librepo_b200.so is not built")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
