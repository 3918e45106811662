"""``nn.Module`` facade over :mod:`fewbit_b200.functional` (reference
``fewbit/modules/activations.py``): one class per activation with the constructor of its
``torch.nn`` namesake (minus ``inplace`` / ``approximate``) plus keyword-only ``bits``.

The generated modules hold no parameters or buffers, so swapping them into a model leaves
its ``state_dict`` unchanged.
"""
from __future__ import annotations

from inspect import Parameter, Signature
from typing import Optional, Tuple

import torch as T

from .. import functional
from ..functional.activations import SIGNATURES, expand_table, stepwise

# class name -> functional name
STEPWISE = {'Hardshrink': 'hardshrink', 'Hardsigmoid': 'hardsigmoid', 'Hardtanh': 'hardtanh',
            'LeakyReLU': 'leaky_relu', 'ReLU': 'relu', 'ReLU6': 'relu6',
            'Softshrink': 'softshrink', 'Threshold': 'threshold'}
CONTINOUS = {'CELU': 'celu', 'ELU': 'elu', 'GELU': 'gelu', 'Hardswish': 'hardswish',
             'LogSigmoid': 'logsigmoid', 'Mish': 'mish', 'SELU': 'selu', 'Sigmoid': 'sigmoid',
             'SiLU': 'silu', 'Softplus': 'softplus', 'Softsign': 'softsign', 'Tanh': 'tanh',
             'Tanhshrink': 'tanhshrink'}

__all__ = tuple(STEPWISE) + ('Stepwise', ) + tuple(CONTINOUS)


class Stepwise(T.nn.Module):
    """A custom stepwise activation (reference ``fewbit/modules/activations.py:97-134``): the
    tensors ``borders`` and ``levels`` define the derivative's constant pieces, the module computes
    the activation they integrate to and keeps only few-bit codes for backward
    (:func:`fewbit_b200.functional.stepwise`).

    :param borders: points between the pieces (with or without the two outer sentinels).
    :param levels: values of the constant pieces (at most 256, mirrored pieces included).
    :param parity: ``True`` / ``False``: the table covers ``x > shift[0]`` only and is mirrored
                   evenly / oddly about ``shift`` (see :func:`functional.expand_table`).
    :param shift: the point ``(x0, s0)`` the table is mirrored about and anchored at.
    """

    def __init__(self, borders: T.Tensor, levels: T.Tensor, parity: Optional[bool] = None,
                 shift: Optional[Tuple[float, float]] = None):
        full_borders, full_levels, anchor = expand_table(borders, levels, parity, shift)   # validates
        if borders.numel() > levels.numel():
            borders = borders[1:-1]
        super().__init__()
        self.register_buffer('borders', borders, True)
        self.register_buffer('levels', levels, True)
        # the mirrored table is derived data: rebuilt from the two buffers above when they are loaded
        self.register_buffer('_full_borders', full_borders, False)
        self.register_buffer('_full_levels', full_levels, False)
        self.parity = parity
        self.shift = shift
        self._anchor = anchor

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._full_borders, self._full_levels, self._anchor = expand_table(self.borders, self.levels, self.parity,
                                                                           self.shift)

    def forward(self, xs: T.Tensor) -> T.Tensor:
        return stepwise(xs, self._full_borders, self._full_levels, None, (self._anchor, 0.0))

    def extra_repr(self) -> str:
        return f'levels={self.levels.numel()}, parity={self.parity}, shift={self.shift}'


class BuiltInStepwiseFunction(T.nn.Module):
    """Base of the generated classes: binds constructor arguments once, forwards them to the
    functional implementation on every call (reference modules/activations.py:137-218)."""

    _impl_name: str = ''
    _signature: Signature = Signature()

    def __init__(self, *args, **kwargs):
        super().__init__()
        bound = self._signature.bind(*args, **kwargs)
        bound.apply_defaults()
        self.args, self.kwargs, self.reprs = [], {}, []
        for name, param in self._signature.parameters.items():
            value = bound.arguments[name]
            setattr(self, name, value)
            if param.kind == Parameter.KEYWORD_ONLY:
                self.kwargs[name] = value
            else:
                self.args.append(value)
            self.reprs.append(f'{name}={value}')

    def __repr__(self) -> str:
        return f'{type(self).__name__}({", ".join(self.reprs)})'

    def forward(self, xs: T.Tensor) -> T.Tensor:
        return getattr(functional, self._impl_name)(xs, *self.args, **self.kwargs)


def _make_class(cls_name: str, impl_name: str):
    params = list(SIGNATURES[impl_name]) + [
        Parameter('bits', Parameter.KEYWORD_ONLY, default=None, annotation=Optional[int])]
    sig = Signature(params)
    doc = (f'In-place, memory-saving drop-in for :class:`torch.nn.{cls_name}`.\n\n'
           f'    Args: those of :class:`torch.nn.{cls_name}` (without ``inplace`` / '
           f'``approximate``) and\n'
           f'        bits: number of bits in gradient approximation: Default: 3\n\n'
           f'    See Also:\n        :class:`torch.nn.{cls_name}` -- Original PyTorch '
           f'implementation.\n')
    def __init__(self, *args, **kwargs):
        BuiltInStepwiseFunction.__init__(self, *args, **kwargs)

    __init__.__signature__ = Signature(
        [Parameter('self', Parameter.POSITIONAL_OR_KEYWORD)] + params)
    return type(cls_name, (BuiltInStepwiseFunction, ),
                {'_impl_name': impl_name, '_signature': sig, '__doc__': doc,
                 '__module__': __name__, '__init__': __init__})


for _cls_name, _impl_name in {**STEPWISE, **CONTINOUS}.items():
    globals()[_cls_name] = _make_class(_cls_name, _impl_name)
del _cls_name, _impl_name
