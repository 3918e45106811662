"""Activation functions with few-bit gradients: table store, signature shim, dispatch.

Host-side mirror of reference ``fewbit/functional/activations.py``:

* ``StepwiseStore`` / ``store``  -- :24-86  (same API, same npz key format)
* ``<name>(input, *torch_args, bits=None, borders=None, values=None)`` for the 13 continuous
  functions -- ``dispatch.forward_call`` :145-218
* ``<name>(input, *torch_args)`` for the eight 1-bit functions -- ``load_func`` :221-249

Differences, all deliberate: signatures are written out instead of being
introspected from the installed torch (C-13); CPU tensors evaluate the *right* function
(C-2, C-3); the 1-bit functions accept and ignore ``bits`` (C-5).  CUDA tensors only ever go
to ``torch.ops.fewbit.<name>`` -- if the operator library is not loaded that is an error,
never a fallback.
"""
from __future__ import annotations

from inspect import Parameter, Signature
from pathlib import Path
from typing import Optional, Tuple, Union

import numpy as np
import torch as T
import torch.nn.functional as F

# Names kept as in the reference (including its spelling) -- they are importable there.
STEPWISE = ('hardshrink', 'hardsigmoid', 'hardtanh', 'leaky_relu', 'relu', 'relu6',
            'softshrink', 'stepwise', 'threshold')
CONTINOUS = ('celu', 'elu', 'gelu', 'hardswish', 'logsigmoid', 'mish', 'selu', 'sigmoid', 'silu',
             'softplus', 'softsign', 'tanh', 'tanhshrink')
CONTINUOUS = CONTINOUS

__all__ = STEPWISE + CONTINOUS + ('store', 'make_table', 'expand_table')

BITS_DEFAULT = 3  # reference functional/activations.py:202


class StepwiseStore:
    """Registry of stepwise gradient approximations, cached per (name, bits, device, dtype).

    ``borders`` have 2^bits + 1 entries (the two outer ones are the +-100 sentinels) and
    ``levels`` 2^bits; the operators receive ``borders[1:-1]``.  Tables are stored in float64
    and cast with one ``.to(device, dtype)`` -- the same call the reference makes, so the
    rounding of the fp32 / bf16 tables is identical by construction.
    """

    def __init__(self):
        self.STORE = {}
        self.CACHE = {}

    def __len__(self) -> int:
        return len(self.STORE)

    def __repr__(self) -> str:
        return f'{type(self).__name__}(stored={len(self.STORE)}, cached={len(self.CACHE)})'

    def add(self, name: str, bits: int, value: Tuple[T.Tensor, T.Tensor]):
        borders, values = value
        entry = (borders, values.to(borders))
        self.STORE[(name, bits)] = entry
        self.CACHE[(name, bits, T.device(borders.device), borders.dtype)] = entry

    def get(self, name: str, bits: int, device: Union[None, str, T.device] = None,
            dtype: Optional[T.dtype] = None):
        key = (name, bits, T.device(device or 'cpu'), dtype or T.float32)
        if (hit := self.CACHE.get(key)) is not None:
            return hit
        if (leaf := self.STORE.get(key[:2])) is None:
            raise KeyError(f'There is not {bits}-bit quantized gradients for activation '
                           f'function {name}.')
        hit = tuple(el.to(key[2], key[3]) for el in leaf)
        self.CACHE[key] = hit
        return hit

    def items(self, cached=False):
        yield from (self.CACHE if cached else self.STORE).items()

    def load(self, path) -> 'StepwiseStore':
        """Add every ``{func}{bits:02d}-borders`` / ``-levels`` pair of an npz file."""
        with np.load(path) as npz:
            for key in sorted({k.split('-', 1)[0] for k in npz.keys()}):
                self.add(key[:-2], int(key[-2:]),
                         (T.tensor(npz[f'{key}-borders']), T.tensor(npz[f'{key}-levels'])))
        return self


def make_table(name: str, bits: int, scale: float = 1.5) -> Tuple[T.Tensor, T.Tensor]:
    """A usable (not optimal) ``bits``-bit table for ``name``: borders at the quantiles of
    N(0, scale^2) with the +-100 sentinels, levels = the mean derivative on each interval,
    ``diff(f(borders)) / diff(borders)`` -- the optimal levels for given borders
    (reference ``fewbit/approx.py:116,132``).  The reference ships tables for 1..4 bits only
    (``tools/quantize-builtins.sh:8``).  A cheap stand-in for arbitrary functions and parameters;
    the optimal tables come from ``fewbit_b200.quantize.optimal_table`` (the 5..8-bit ones of the
    13 built-in functions are shipped in ``data/extended.npz``).  float64, same layout as the
    built-in ones; pass as ``borders=``/``values=`` or ``store.add(name, bits, ...)``.
    """
    count = (1 << bits) - 1
    probs = T.arange(1, count + 1, dtype=T.float64) / (count + 1)
    inner = scale * (2.0 ** 0.5) * T.special.erfinv(2 * probs - 1)
    borders = T.cat([T.tensor([-100.0], dtype=T.float64), inner, T.tensor([100.0], dtype=T.float64)])
    fn = getattr(F, name, None) or getattr(T, name)
    values = fn(borders)
    return borders, (values[1:] - values[:-1]) / (borders[1:] - borders[:-1])


store = StepwiseStore()
# bits 1..4: the reference's own tables, number for number (tools/make_builtin_tables.py);
# bits 5..8: optimal tables for the same objective from fewbit_b200/quantize.py
# (tools/make_extended_tables.py) -- the reference ships none, its kernels stop being useful there.
store.load(Path(__file__).resolve().parent.parent / 'data' / 'builtin.npz')
_extended = Path(__file__).resolve().parent.parent / 'data' / 'extended.npz'
if _extended.exists():
    store.load(_extended)


# ------------------------------------------------------------------ device dispatch ----

def _native_op(name: str):
    """``torch.ops.fewbit.<name>`` or a loud error -- CUDA tensors never fall back."""
    from .. import NATIVE_ERROR
    if NATIVE_ERROR is not None:
        raise RuntimeError(f'fewbit.{name}: CUDA tensor given but the operator library '
                           f'libfewbit.so is not loaded ({NATIVE_ERROR}).')
    return getattr(T.ops.fewbit, name)


class _HostStepwise(T.autograd.Function):
    """CPU tensors: exact forward, quantized backward, as plain PyTorch host code.

    Not a fallback of the CUDA path (CUDA tensors never reach this) but the CPU side of the
    reference's device switch (functional/activations.py:229-237), with the forward fixed.
    """

    @staticmethod
    def forward(ctx, input, borders, levels, impl, args):
        if borders.numel() + 1 != levels.numel():
            raise ValueError('Size of `borders` should be lesser than size of `levels` by one.')
        codes = T.searchsorted(borders.contiguous(), input.detach().contiguous())
        ctx.save_for_backward(codes.to(T.uint8), levels)
        return impl(input, *args)

    @staticmethod
    def backward(ctx, grad_output):
        codes, levels = ctx.saved_tensors
        return levels[codes.long()] * grad_output, None, None, None, None


def _dispatch(name: str, input: T.Tensor, *args):
    if input.device.type == 'cuda':
        return _native_op(name)(input, *args)
    impl = getattr(F, name, None) or getattr(T, name)
    if name in STEPWISE:
        return impl(input, *args)
    borders, levels, *rest = args
    return _HostStepwise.apply(input, borders, levels, impl, tuple(rest))


# ------------------------------------------------------------ continuous functions ----

_P = Parameter
_TENSOR = _P('input', _P.POSITIONAL_OR_KEYWORD, annotation=T.Tensor)
_QUANT = [
    _P('bits', _P.KEYWORD_ONLY, default=None, annotation=Optional[int]),
    _P('borders', _P.KEYWORD_ONLY, default=None, annotation=Optional[T.Tensor]),
    _P('values', _P.KEYWORD_ONLY, default=None, annotation=Optional[T.Tensor]),
]


def _arg(name, default=_P.empty):
    return _P(name, _P.POSITIONAL_OR_KEYWORD, default=default, annotation=float)


# Documented torch.nn.functional signatures, minus `inplace` / `approximate`, which the
# reference drops as well (functional/activations.py:156-159).
SIGNATURES = {
    'celu': [_arg('alpha', 1.0)],
    'elu': [_arg('alpha', 1.0)],
    'gelu': [], 'hardswish': [], 'logsigmoid': [], 'mish': [], 'selu': [], 'sigmoid': [],
    'silu': [],
    'softplus': [_arg('beta', 1.0), _arg('threshold', 20.0)],
    'softsign': [], 'tanh': [], 'tanhshrink': [],
    'hardshrink': [_arg('lambd', 0.5)],
    'hardsigmoid': [],
    'hardtanh': [_arg('min_val', -1.0), _arg('max_val', 1.0)],
    'leaky_relu': [_arg('negative_slope', 0.01)],
    'relu': [], 'relu6': [],
    'softshrink': [_arg('lambd', 0.5)],
    'threshold': [_arg('threshold'), _arg('value')],
}


def _make_continuous(name: str):
    sig = Signature([_TENSOR] + SIGNATURES[name] + _QUANT)

    def forward_call(*args, **kwargs):
        bound = sig.bind(*args, **kwargs)
        bound.apply_defaults()
        params = bound.arguments
        input, bits = params['input'], params['bits']
        borders, values = params['borders'], params['values']

        use_builtin = bits is not None
        use_custom = borders is not None and values is not None
        if use_builtin and use_custom:
            raise ValueError('Either `bits` or `borders` and `values` should be scpecifed '
                             'not both.')
        if use_builtin or not use_custom:
            borders, values = store.get(name, bits or BITS_DEFAULT, input.device, input.dtype)

        extra = [params[p.name] for p in SIGNATURES[name]]
        return _dispatch(name, input, borders[1:-1].to(input), values.to(input), *extra)

    forward_call.__name__ = forward_call.__qualname__ = name
    forward_call.__signature__ = sig
    forward_call.__doc__ = (
        f'In-place ``{name}`` whose backward pass keeps only ``bits``-bit codes of the input.\n\n'
        f'Same arguments as :func:`torch.nn.functional.{name}` plus keyword-only ``bits`` '
        f'(built-in table, default {BITS_DEFAULT}) or ``borders`` and ``values`` (custom table).')
    return forward_call


def _make_piecewise(name: str):
    # `bits` is accepted and ignored: the modules of the reference pass it.
    sig = Signature([_TENSOR] + SIGNATURES[name] +
                    [_P('bits', _P.KEYWORD_ONLY, default=None, annotation=Optional[int])])

    def forward_call(*args, **kwargs):
        bound = sig.bind(*args, **kwargs)
        bound.apply_defaults()
        params = bound.arguments
        return _dispatch(name, params['input'], *[params[p.name] for p in SIGNATURES[name]])

    forward_call.__name__ = forward_call.__qualname__ = name
    forward_call.__signature__ = sig
    forward_call.__doc__ = (f'In-place ``{name}`` that saves a 1-bit mask for backward. Same '
                            f'arguments as :func:`torch.nn.functional.{name}`.')
    return forward_call


def expand_table(borders: T.Tensor, levels: T.Tensor, parity: Optional[bool] = None,
                 shift: Optional[Tuple[float, float]] = None):
    """The full table of a custom stepwise function and the point it is anchored at.

    ``levels`` are the constant pieces of the derivative, ``borders`` the points between them
    (with or without the two outer ends, as in the reference's ``Stepwise`` module).  With
    ``parity`` the table describes only ``x > x0`` and is mirrored about ``(x0, s0) = shift``
    (default ``(0, 0)``): ``True`` -- even, ``s(x0 - t) = s(x0 + t)``; ``False`` -- odd,
    ``s(x0 - t) = 2 s0 - s(x0 + t)`` (GELU's derivative is odd about ``(0, 1/2)``).  Half a
    table thus stands for twice the steps (reference README.md:111-112)."""
    if borders.ndim != 1 or levels.ndim != 1:
        raise ValueError('Expected number of dimensions of `borders` and `levels` is one.')
    if borders.numel() > levels.numel():
        borders = borders[1:-1]
    if borders.numel() + 1 != levels.numel():
        raise ValueError('Size of `borders` should be lesser than size of `levels` by one.')
    x0, s0 = (float(shift[0]), float(shift[1])) if shift is not None else (0.0, 0.0)
    if parity is not None:
        if borders.numel() and float(borders.min()) <= x0:
            raise ValueError('With `parity` the table describes x > shift[0]: all borders must lie above it.')
        mirrored = levels.flip(0) if parity else 2.0 * s0 - levels.flip(0)
        centre = T.full((1, ), x0, dtype=borders.dtype, device=borders.device)
        borders = T.cat([2.0 * x0 - borders.flip(0), centre, borders])
        levels = T.cat([mirrored, levels])
    if levels.numel() > 256:
        raise ValueError('Maximal number of step limited to 256.')
    return borders, levels, x0


def _piecewise_linear(x: T.Tensor, borders: T.Tensor, levels: T.Tensor, anchor: float):
    """F(x) with F' = levels[code(x)], F continuous, F(anchor) = 0; and code(x)."""
    b, l = borders.double(), levels.double()
    rise = T.zeros(l.numel(), dtype=T.float64, device=b.device)      # rise[k] = F~(borders[k]), F~(borders[0]) = 0
    if b.numel() > 1:
        rise[1:b.numel()] = T.cumsum(l[1:b.numel()] * (b[1:] - b[:-1]), 0)
    left = T.clamp(T.arange(l.numel(), device=b.device) - 1, 0, max(b.numel() - 1, 0))
    if b.numel():
        intercept = rise[left] - l * b[left]
    else:
        intercept = T.zeros_like(l)
    anchor_t = T.tensor([anchor], dtype=b.dtype, device=b.device)
    piece = int(T.searchsorted(b, anchor_t, right=False))
    intercept = intercept - (l[piece] * anchor + intercept[piece])
    code = T.searchsorted(borders.contiguous(), x.detach(), right=False)
    value = T.addcmul(intercept.float()[code], levels.float()[code], x.float()).to(x.dtype)
    return value, code


class _StepwiseHostFunc(T.autograd.Function):
    """CPU tensors: the same activation in PyTorch ops (out of place)."""

    @staticmethod
    def forward(ctx, input, borders, levels, anchor):
        value, code = _piecewise_linear(input, borders.to(input), levels.to(input), anchor)
        ctx.save_for_backward(code.to(T.uint8), levels.to(input))
        return value

    @staticmethod
    def backward(ctx, grad_output):
        code, levels = ctx.saved_tensors
        return levels[code.long()] * grad_output, None, None, None


def stepwise(input: T.Tensor, borders: T.Tensor, levels: T.Tensor, parity: Optional[bool] = None,
             shift: Optional[Tuple[float, float]] = None) -> T.Tensor:
    """Custom-table activation (reference ``fewbit.functional.stepwise`` / operator
    ``fewbit::stepwise``, fewbit/fewbit.cc:37 -- declared there, never implemented).

    The table is the function: the result is the continuous piecewise-linear ``F`` whose slopes
    are ``levels`` between ``borders``, with ``F(shift[0]) = 0``; backward multiplies the
    incoming gradient by ``levels[code(x)]`` -- exactly ``F'``.  In place on CUDA tensors, where
    only the bit-packed codes are saved; see :func:`expand_table` for ``parity`` / ``shift``."""
    full_borders, full_levels, anchor = expand_table(borders, levels, parity, shift)
    if input.device.type == 'cuda':
        return _native_op('stepwise_anchored')(input, full_borders.to(input), full_levels.to(input), anchor)
    return _StepwiseHostFunc.apply(input, full_borders, full_levels, anchor)


for _name in CONTINOUS:
    globals()[_name] = _make_continuous(_name)
for _name in STEPWISE:
    if _name != 'stepwise':
        globals()[_name] = _make_piecewise(_name)
del _name
