"""Discrete cosine transforms on torch tensors, any device, through ``torch.fft``
(reference ``fewbit/fft.py``: ``dct`` / ``idct``, types 2 and 3, the three scipy normalisations).
Used by the ``dct`` sketch of ``RandomizedLinear`` (reference ``fewbit/functional/linear.py:113-122``).

A length-N DCT-II is the real part of one length-N complex FFT of the even/odd-folded sequence,
twisted by a quarter-sample phase (Makhoul 1980); the DCT-III is its transpose, computed by
undoing the same steps.  Conventions are scipy's (``scipy.fft.dct``):

    backward:  y[k] = 2 sum_n x[n] cos(pi k (2n + 1) / 2N)
    ortho:     backward scaled by sqrt(1/4N) for k = 0 and sqrt(1/2N) otherwise (orthonormal)
    forward:   backward / 2N
"""
from __future__ import annotations

import math
from typing import Optional

import torch as T

__all__ = ('dct', 'idct')


def _resize(x: T.Tensor, n: Optional[int]) -> T.Tensor:
    """Truncate or zero-pad the last axis to length n (scipy's meaning of `n`)."""
    if n is None or n == x.shape[-1]:
        return x
    if n < x.shape[-1]:
        return x[..., :n]
    return T.nn.functional.pad(x, (0, n - x.shape[-1]))


def _twiddle(length: int, like: T.Tensor) -> T.Tensor:
    """exp(-i pi k / 2N), k = 0..N-1, in the complex dtype that matches `like`."""
    k = T.arange(length, device=like.device, dtype=T.float64)
    return T.polar(T.ones_like(k), -math.pi * k / (2 * length)).to(T.complex128 if like.dtype == T.float64
                                                                  else T.complex64)


def _dct2_backward(x: T.Tensor) -> T.Tensor:
    length = x.shape[-1]
    folded = T.cat([x[..., 0::2], x[..., 1::2].flip(-1)], dim=-1)
    spectrum = T.fft.fft(folded, dim=-1)
    return 2 * (spectrum * _twiddle(length, x)).real


def _dct3_backward(y: T.Tensor) -> T.Tensor:
    """x[n] = y[0] + 2 sum_{k>=1} y[k] cos(pi k (2n + 1) / 2N): the transpose of _dct2_backward."""
    length = y.shape[-1]
    if length == 1:
        return y.clone()
    # V[k] = (y[k] - i y[N-k]) conj(w[k]) with y[N] := 0 rebuilds the spectrum of the folded sequence
    mirrored = T.cat([T.zeros_like(y[..., :1]), y[..., 1:].flip(-1)], dim=-1)
    spectrum = T.complex(y, -mirrored) * _twiddle(length, y).conj()
    folded = T.fft.ifft(spectrum, dim=-1).real * length
    out = T.empty_like(y)
    half = (length + 1) // 2
    out[..., 0::2] = folded[..., :half]
    out[..., 1::2] = folded[..., half:].flip(-1)
    return out


def _scale_rows(x: T.Tensor, first: float, rest: float) -> T.Tensor:
    weights = T.full((x.shape[-1], ), rest, dtype=x.dtype, device=x.device)
    weights[0] = first
    return x * weights


def _transform(x: T.Tensor, kind: int, n: Optional[int], dim: int, norm: str) -> T.Tensor:
    if norm not in ('backward', 'forward', 'ortho'):
        raise ValueError(f'Unexpected normalization regime: {norm}.')
    work = _resize(x.transpose(dim, -1), n)
    length = work.shape[-1]
    if kind == 2:
        out = _dct2_backward(work)
        if norm == 'ortho':
            out = _scale_rows(out, math.sqrt(1 / (4 * length)), math.sqrt(1 / (2 * length)))
        elif norm == 'forward':
            out = out / (2 * length)
    else:
        if norm == 'ortho':       # the orthonormal matrix's transpose: undo the row weights first
            work = _scale_rows(work, math.sqrt(1 / length), math.sqrt(1 / (2 * length)))
            out = _dct3_backward(work)
        elif norm == 'forward':
            out = _dct3_backward(work) / (2 * length)
        else:
            out = _dct3_backward(work)
    return out.transpose(dim, -1)


def dct(x: T.Tensor, type: int = 2, n: Optional[int] = None, dim: int = -1, norm: str = 'backward') -> T.Tensor:
    """Discrete cosine transform of type 2 or 3 along `dim` (scipy.fft.dct semantics)."""
    if type not in (2, 3):
        raise ValueError(f'Unexpected DCT type: {type}.')
    return _transform(x, type, n, dim, norm)


def idct(x: T.Tensor, type: int = 2, n: Optional[int] = None, dim: int = -1, norm: str = 'backward') -> T.Tensor:
    """Inverse of `dct` of the same type and norm (scipy.fft.idct semantics)."""
    if type not in (2, 3):
        raise ValueError(f'Unexpected IDCT type: {type}.')
    flipped = {'backward': 'forward', 'forward': 'backward', 'ortho': 'ortho'}
    if norm not in flipped:
        raise ValueError(f'Unexpected normalization regime: {norm}.')
    return _transform(x, 5 - type, n, dim, flipped[norm])
