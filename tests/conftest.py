import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")
    config.addinivalue_line("markers", "gpu_experimental: default-off kernel experiments; run only when selected "
                                       "explicitly with -m gpu_experimental on a B200")


def pytest_collection_modifyitems(config, items):
    import torch
    have_cuda = torch.cuda.is_available()
    want_experimental = "gpu_experimental" in (config.getoption("-m") or "")
    skip_gpu = pytest.mark.skip(reason="no CUDA device")
    skip_exp = pytest.mark.skip(reason="experimental: select with -m gpu_experimental on a B200")
    for item in items:
        if "gpu_experimental" in item.keywords:
            if not (have_cuda and want_experimental):
                item.add_marker(skip_exp)
        elif "gpu" in item.keywords and not have_cuda:
            item.add_marker(skip_gpu)
