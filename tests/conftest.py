import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / 'tests' / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_ops():
    """Cases produced by the unmodified reference CPU ops (tests/golden/make_golden.py)."""
    cases = []
    with np.load(GOLDEN / 'reference_cpu_ops.npz') as npz:
        keys = sorted({k.split('/')[0] for k in npz.keys()})
        for key in keys:
            bits, n, bf16 = (int(v) for v in npz[f'{key}/meta'])
            case = {'key': key, 'name': str(npz[f'{key}/name']), 'bits': bits, 'n': n,
                    'bf16': bool(bf16)}
            for field in ('x', 'bounds', 'levels', 'g', 'y', 'state', 'gin'):
                case[field] = npz[f'{key}/{field}']
            cases.append(case)
    return cases


@pytest.fixture(scope='session')
def golden_tables():
    with np.load(GOLDEN / 'reference_tables.npz') as npz:
        return {k: npz[k] for k in npz.keys()}


@pytest.fixture(scope='session')
def golden_linear():
    with np.load(GOLDEN / 'reference_linear.npz') as npz:
        return {k: npz[k] for k in npz.keys()}
