import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'ref: needs /root/reference (build container only)')


def pytest_collection_modifyitems(config, items):
    from oracle import ref_loader
    have_ref = ref_loader.available()
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    for item in items:
        if 'ref' in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason='/root/reference not present'))
        if 'gpu' in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason='no CUDA device'))
