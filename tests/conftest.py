import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible, e.g. a plain
    `pytest tests` in the development container."""
    have = False
    try:
        # cheap probe through our own library; torch only if the library is not built, so that
        # GPU tests on a GPU box fail loudly (instead of being skipped) when it is missing
        from pyshepseg_b200 import _lib
        have = _lib.load().ssg_device_count() > 0
    except Exception:
        try:
            import torch
            have = torch.cuda.is_available()
        except Exception:
            have = False
    if have:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)
