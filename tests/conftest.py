"""Shared pytest configuration: the `gpu` marker, golden fixtures, the CPU oracle."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / 'tests' / 'golden'
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu on the GPU box')


def load_golden(name):
    with np.load(GOLDEN / f'{name}.npz') as d:
        return {k: d[k] for k in d.files}


@pytest.fixture(scope='session')
def orc():
    """The CPU oracle (oracle/rr_oracle.c through ctypes) -- the checker, never the product."""
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope='session')
def tab(orc):
    return orc.Tables()


@pytest.fixture(scope='session')
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
