import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def pkg():
    entry.build()            # csrc/, host/ and oracle/ in-tree; a no-op when the libraries are newer than their sources
    return entry.load_package()


@pytest.fixture(scope="session")
def orc():
    return entry.load_oracle()


@pytest.fixture(scope="session")
def problem(pkg):
    return pkg.load_default_problem()


@pytest.fixture(scope="session")
def oracle(orc, problem):
    return orc.Oracle(problem)


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "spain2020_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if stale) and load the product library; GPU tests call through the C ABI only."""
    entry.build_cuda()
    entry.load_package()
    from sepaihrd_b200 import capi
    return capi.load_library()
