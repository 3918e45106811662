"""Module interface (mirrors reference ``fewbit/modules/__init__.py``; additionally exports
``RandomizedLinear``, which the reference README imports but the package forgot, C-10)."""
from .activations import (  # noqa: F401
    Hardshrink, Hardsigmoid, Hardtanh, LeakyReLU, ReLU, ReLU6, Softshrink, Stepwise, Threshold)
from .activations import (  # noqa: F401
    CELU, ELU, GELU, Hardswish, LogSigmoid, Mish, SELU, Sigmoid, SiLU, Softplus, Softsign, Tanh,
    Tanhshrink)
from .linear import LinearCRS, LinearGRP, RandomizedLinear  # noqa: F401

__all__ = ('Hardshrink', 'Hardsigmoid', 'Hardtanh', 'LeakyReLU', 'ReLU', 'ReLU6', 'Softshrink',
           'Stepwise', 'Threshold', 'CELU', 'ELU', 'GELU', 'Hardswish', 'LogSigmoid', 'Mish',
           'SELU', 'Sigmoid', 'SiLU', 'Softplus', 'Softsign', 'Tanh', 'Tanhshrink', 'LinearCRS',
           'LinearGRP', 'RandomizedLinear')
