import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200)')
    # the oracle's C part is test infrastructure; build it on demand
    if not os.path.exists(os.path.join(ROOT, 'oracle', 'librvs_oracle.so')):
        subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle')],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope='session')
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False))
        return cache[name]
    return get
