import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session', autouse=True)
def _library_is_built():
    """The C-ABI library is built in-tree by `python -m dgpmp2_b200.build` / __graft_entry__.build(); a fresh checkout
    has none (it is git-ignored), so the first test session compiles it (nvcc, ~1.5 min, no GPU needed)."""
    from dgpmp2_b200 import build as _build
    _build.build(force=False)
