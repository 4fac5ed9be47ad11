import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'reference: needs the live reference at /root/reference')


def pytest_collection_modifyitems(config, items):
    from oracle import ref_loader
    have_ref = ref_loader.available()
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_gpu = False
    for item in items:
        if 'reference' in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason='live reference not present on this box'))
        if 'gpu' in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason='no CUDA device'))


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(ROOT, 'tests', 'golden')
