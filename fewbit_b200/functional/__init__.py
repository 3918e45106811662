"""Functional interface (mirrors reference ``fewbit/functional/__init__.py``)."""
# 1-bit piecewise activations.
from .activations import (  # noqa: F401
    hardshrink, hardsigmoid, hardtanh, leaky_relu, relu, relu6, softshrink, stepwise, threshold)
# Continuous activations with b-bit quantized gradients.
from .activations import (  # noqa: F401
    celu, elu, gelu, hardswish, logsigmoid, mish, selu, sigmoid, silu, softplus, softsign, tanh,
    tanhshrink)
from .activations import (  # noqa: F401
    CONTINOUS, CONTINUOUS, STEPWISE, StepwiseStore, expand_table, make_table, store)
# Linear layer with randomized (sketched) weight gradient.
from .linear import linear_crs, linear_grp, linear_randomized  # noqa: F401
