"""fewbit_b200 -- B200-native (sm_100a) implementation of FewBit's quantized-gradient
activation path, behind the reference's own surface: ``GELU(bits=b)`` & co, the 1-bit
ReLU family, ``functional.*``, the ``torch.ops.fewbit.*`` operators, ``RandomizedLinear``
and ``util.map_module`` / ``convert_linear``.

Loader contract (reference ``fewbit/__init__.py:17-23``): the operator library
``libfewbit.so`` sits next to this file and is loaded at import unless the environment
variable ``FEWBIT_NATIVE`` is ``0``/``no``/``false``; a failure to load is a
``RuntimeWarning`` at import time and a hard ``RuntimeError`` the moment a CUDA tensor
reaches an operator -- there is no silent fallback for CUDA tensors.
"""
from os import getenv
from pathlib import Path
from warnings import warn

import torch as _torch

NATIVE_LIBRARY = Path(__file__).with_name('libfewbit.so')
NATIVE_ERROR = None   # why the operator library is not loaded (None = loaded)

if getenv('FEWBIT_NATIVE') not in ('0', 'no', 'false'):
    try:
        _torch.ops.load_library(str(NATIVE_LIBRARY))
    except Exception as e:  # noqa: BLE001 -- mirror the reference: warn and carry on
        NATIVE_ERROR = f'{type(e).__name__}: {e}'
        warn(f'Failed to load ops library: {e}.', RuntimeWarning)
else:
    NATIVE_ERROR = 'disabled by FEWBIT_NATIVE'


def native_loaded() -> bool:
    return NATIVE_ERROR is None


from . import functional  # noqa: E402,F401
from .modules import *  # noqa: E402,F401,F403
from .modules import RandomizedLinear  # noqa: E402,F401
from .util import convert_linear, map_module  # noqa: E402,F401

__version__ = '0.1.0'
