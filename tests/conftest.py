import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'reference: needs the reference checkout at /root/reference')


def pytest_collection_modifyitems(config, items):
    from oracle import ref_import
    if not ref_import.available():
        skip = pytest.mark.skip(reason='reference checkout not mounted')
        for item in items:
            if 'reference' in item.keywords:
                item.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    return load
