"""Linear layer whose weight gradient is estimated from a random sketch of the token axis
(reference ``fewbit/functional/linear.py:69-221``, ``LinearGRPFunc``).

Forward is an exact ``F.linear``; instead of the input ``X`` (N x in) only the sketch
``X_proj = S X / P`` (P x in) is saved, with ``S`` (P x N) Gaussian or Rademacher.  Backward
regenerates the same ``S`` from the saved generator state and returns

    grad_input = G W            (exact)
    grad_weight = (S G)^T X_proj   (unbiased estimate of G^T X)
    grad_bias  = sum_n G        (exact)

Deliberate fixes: without a user generator the CUDA path draws from the
device's *default* generator (the reference builds a fresh default-seeded generator per call,
i.e. the same ``S`` in every layer and step), and ``S`` is created in the input dtype.
"""
from __future__ import annotations

import contextlib
import weakref

from typing import Literal, Optional

import torch as T
import torch.nn.functional as F

from ..fft import dct

__all__ = ('linear_crs', 'linear_grp', 'linear_randomized', 'calc_proj_dim')

MatMulType = Literal['gaussian', 'rademacher', 'dct', 'dft']


def clamp(val: int, minval: Optional[int] = None, maxval: Optional[int] = None) -> int:
    # Falsy bounds are ignored, exactly as reference functional/linear.py:17-24.
    if minval:
        val = max(minval, val)
    if maxval:
        val = min(maxval, val)
    return val


def calc_proj_dim(ndim: int, proj_dim_ratio: Optional[float], proj_dim: Optional[int],
                  proj_dim_max: Optional[int], proj_dim_min: Optional[int]) -> int:
    """P = proj_dim or int(ratio * N) or N, then clamped (reference :72-81)."""
    if proj_dim:
        result = proj_dim
    elif proj_dim_ratio:
        result = int(proj_dim_ratio * ndim)
    else:
        result = ndim
    return clamp(result, proj_dim_min, proj_dim_max)


def _default_generator(device: T.device) -> T.Generator:
    if device.type == 'cuda':
        index = device.index if device.index is not None else T.cuda.current_device()
        return T.cuda.default_generators[index]
    return T.default_generator


def _sketch_matrix(kind: str, rows: int, cols: int, generator: T.Generator, device, dtype):
    """The random matrix S (rows x cols): N(0,1) entries or +-1/2 (reference :133-146)."""
    if kind == 'gaussian':
        return T.randn((rows, cols), generator=generator, device=device, dtype=dtype)
    if kind == 'rademacher':
        return T.randint(high=2, size=(rows, cols), generator=generator, device=device,
                         dtype=dtype) - 0.5
    raise ValueError(f'Unexpected matmul type: {kind}.')


def _sampled_rows(rows: int, count: int, generator: T.Generator, device) -> T.Tensor:
    """`rows` token indices drawn uniformly with replacement (reference :114-119, :123-128)."""
    probas = T.ones(count, device=device)
    probas /= count
    return T.multinomial(input=probas, num_samples=rows, replacement=True, generator=generator)


def _transform(kind: str, view: T.Tensor, inverse: bool = False) -> T.Tensor:
    """Orthonormal cosine / Fourier transform along the token axis.  torch.fft has no bf16 / fp16
    kernels: those inputs are transformed in fp32."""
    work = view if view.dtype in (T.float32, T.float64) else view.float()
    if kind == 'dct':
        return dct(work, dim=0, norm='ortho').to(view.dtype)
    return (T.fft.ifft if inverse else T.fft.fft)(work, dim=0, norm='ortho')


def _project_input(kind: str, view: T.Tensor, rows: int, generator: T.Generator) -> T.Tensor:
    """What is saved instead of the input (reference forward, :113-146), with the very arithmetic
    of the reference.  Note its scale for the two transform sketches: rows * tokens, where an
    unbiased estimate would need tokens / rows -- kept, a drop-in must return what the reference
    returns (tests/golden/reference_linear.npz pins it); see DESIGN.md section 5."""
    if kind in ('dct', 'dft'):
        picked = _sampled_rows(rows, view.shape[0], generator, view.device)
        return (rows * view.shape[0]) * _transform(kind, view)[picked, ...]
    proj = _sketch_matrix(kind, rows, view.shape[0], generator, view.device, view.dtype)
    # E[S^T S] = P I (gaussian) or P/4 I (rademacher): scale so that E[grad_weight] = G^T X
    if kind == 'gaussian':
        return (proj @ view) / rows
    return (proj @ view) * (4 / rows)


def _project_grad(kind: str, grad_view: T.Tensor, rows: int, generator: T.Generator, like: T.Tensor) -> T.Tensor:
    """The same sketch applied to grad_output (reference backward, :178-211)."""
    if kind in ('dct', 'dft'):
        picked = _sampled_rows(rows, grad_view.shape[0], generator, grad_view.device)
        return _transform(kind, grad_view, inverse=True)[picked, :]
    return _sketch_matrix(kind, rows, grad_view.shape[0], generator, grad_view.device, like.dtype) @ grad_view


def _linear_owning_output(input_view: T.Tensor, weight: T.Tensor, bias: Optional[T.Tensor],
                          shape) -> T.Tensor:
    """``F.linear`` whose result owns its storage.  For inputs with more than two dimensions
    ``F.linear`` returns a view of a 2-D product; a view created inside a custom Function may
    not be modified in place afterwards, which is exactly what the in-place few-bit
    activations do to the output of the preceding linear layer."""
    out = input_view.new_empty(*shape[:-1], weight.shape[0])
    flat = out.view(-1, weight.shape[0])
    if bias is not None:
        T.addmm(bias, input_view, weight.t(), out=flat)
    else:
        T.mm(input_view, weight.t(), out=flat)
    return out


SKETCH_KINDS = {'gaussian': 0, 'rademacher': 1}


def _native_sketch_available(tensor: T.Tensor, kind: str) -> bool:
    """CUDA tensors with a feature count the TMA can address go to the tcgen05 kernel."""
    if tensor.device.type != 'cuda' or kind not in SKETCH_KINDS:
        return False
    from .. import NATIVE_ERROR
    if NATIVE_ERROR is not None:
        raise RuntimeError(f'fewbit.linear_grp: CUDA tensor given but the operator library '
                           f'libfewbit.so is not loaded ({NATIVE_ERROR}).')
    return tensor.shape[-1] % 8 == 0 and tensor.dtype in (T.float32, T.bfloat16)


def _native_sketch(view: T.Tensor, rows: int, seed: int, offset: int, kind: str, scale: float,
                   dtype: T.dtype = T.float32, column_sums: bool = False) -> T.Tensor:
    """scale * S @ view  with S generated inside the kernel, [rows, features] in `dtype` (fp32
    accumulation, rounded once in the kernel: no separate `.to(dtype)` pass).  Operands enter the
    tensor cores as bf16; the rounding is unbiased and three orders of magnitude below the
    sketch's own O(1/sqrt(P)) noise.  `column_sums` appends the row scale * view.sum(0)."""
    narrow = dtype == T.bfloat16
    out = T.ops.fewbit.sketch_to(view.to(T.bfloat16).contiguous(), rows, seed, offset, SKETCH_KINDS[kind],
                                 scale, narrow, column_sums)
    return out if narrow or dtype == T.float32 else out.to(dtype)


def _draw_stream(generator: T.Generator):
    """(seed, offset) identifying one sketch; advances the generator's Philox offset."""
    seed = generator.initial_seed() & 0x7FFFFFFFFFFFFFFF
    offset = generator.get_offset()
    generator.set_offset(offset + 4)
    return seed, offset


class _SharedSketch:
    """A sketch taken with ``share_sketch=True``: layers that are fed the very same tensor one
    after the other (the query / key / value projections of an attention block) reuse one ``S X``
    instead of sketching it three times.

    The record hangs off the INPUT TENSOR OBJECT (attribute ``_fewbit_shared_sketch``), so it lives
    exactly as long as that tensor does -- no process-wide table, nothing that keeps a projection
    alive after its input is gone, nothing shared between threads that work on different tensors.
    It is dropped as soon as one of its consumers runs backward: an input that is fed again in a
    later step (fixed batches, eval followed by train) is sketched afresh, so the sketch noise of
    successive optimiser steps stays independent."""
    __slots__ = ('version', 'rows', 'kind', 'projection', 'stream')
    ATTRIBUTE = '_fewbit_shared_sketch'

    @classmethod
    def find(cls, tensor: T.Tensor, rows: int, kind: str) -> Optional['_SharedSketch']:
        record = tensor.__dict__.get(cls.ATTRIBUTE)
        if (record is not None and record.version == tensor._version and record.rows == rows
                and record.kind == kind):
            return record
        return None

    @classmethod
    def drop(cls, tensor_ref) -> None:
        tensor = tensor_ref() if tensor_ref is not None else None
        if tensor is not None:
            tensor.__dict__.pop(cls.ATTRIBUTE, None)


_NO_CONTEXT = contextlib.nullcontext()


def _autocast_operands(input_view: T.Tensor, weight: T.Tensor, bias: Optional[T.Tensor]):
    """``F.linear`` would be autocast; the ``out=`` product that replaces it is not (and raises on
    mixed dtypes), so the operands are cast here the way autocast would."""
    kind = input_view.device.type
    if T.is_autocast_enabled(kind):
        dtype = T.get_autocast_dtype(kind)
        return (input_view.to(dtype), weight.to(dtype), None if bias is None else bias.to(dtype))
    return input_view, weight, bias


class LinearGRPFunc(T.autograd.Function):

    @staticmethod
    def forward(ctx, input: T.Tensor, weight: T.Tensor, bias: Optional[T.Tensor],
                proj_dim_ratio: Optional[float], proj_dim: Optional[int],
                proj_dim_max: Optional[int], proj_dim_min: Optional[int], matmul: MatMulType,
                generator: Optional[T.Generator], share_sketch: bool = False) -> T.Tensor:
        if proj_dim_ratio is None and proj_dim is None:
            raise ValueError('Either proj_dim or proj_dim_ratio should be specified.')
        if proj_dim_min and proj_dim_min <= 0:
            raise ValueError('Param proj_dim_min should be strictly positive.')
        if proj_dim_min and proj_dim_max and proj_dim_max < proj_dim_min:
            raise ValueError('Param proj_dim_min should be not greater than param proj_dim_max.')
        if matmul not in ('gaussian', 'rademacher', 'dct', 'dft'):
            raise ValueError(f'Unexpected matmul type: {matmul}.')

        generator = generator or _default_generator(input.device)
        input_view = input.reshape(-1, input.shape[-1])
        proj_features = calc_proj_dim(input_view.shape[0], proj_dim_ratio, proj_dim, proj_dim_max,
                                      proj_dim_min)
        ctx.proj_features = proj_features
        ctx.matmul = matmul
        ctx.stream = ctx.generator_state = ctx.shared_input = None
        ctx.autocast = T.get_autocast_dtype(input.device.type) if T.is_autocast_enabled(input.device.type) else None
        lhs, rhs, offset_term = (input_view, weight, bias) if ctx.autocast is None else \
            _autocast_operands(input_view, weight, bias)

        if not ctx.needs_input_grad[1]:
            # frozen weight / inference: nothing will ever read a sketch, so none is taken
            ctx.save_for_backward(None, weight, bias)
        elif (proj_features > 0 and _native_sketch_available(input_view, matmul)
              and generator.device.type == 'cuda'):
            # B200 path: S never exists in memory; (seed, offset) replaces the generator state.
            shared = _SharedSketch.find(input, proj_features, matmul) if share_sketch else None
            if shared is not None:
                input_proj, (seed, offset) = shared.projection, shared.stream
            else:
                seed, offset = _draw_stream(generator)
                scale = 1.0 / proj_features if matmul == 'gaussian' else 4.0 / proj_features
                input_proj = _native_sketch(input_view, proj_features, seed, offset, matmul, scale, input.dtype)
                if share_sketch:
                    shared = _SharedSketch()
                    shared.version, shared.rows, shared.kind = input._version, proj_features, matmul
                    shared.projection, shared.stream = input_proj, (seed, offset)
                    setattr(input, _SharedSketch.ATTRIBUTE, shared)
            if share_sketch:
                ctx.shared_input = weakref.ref(input)
            ctx.save_for_backward(input_proj, weight, bias)
            ctx.stream = (seed, offset)
        else:
            # CPU tensors, the transform sketches, feature counts the TMA cannot address, P = 0
            # (an empty sketch and a zero weight gradient, as in the reference): PyTorch ops
            ctx.generator_state = generator.get_state()
            ctx.generator_device = generator.device
            ctx.save_for_backward(_project_input(matmul, input_view, proj_features, generator), weight, bias)
        return _linear_owning_output(lhs, rhs, offset_term, input.shape)

    @staticmethod
    def backward(ctx, grad_output):
        input_proj, weight, bias = ctx.saved_tensors
        grad_input = grad_weight = grad_bias = None
        # backward runs under the autocast state of forward (a context manager only when it was on:
        # entering one costs ~10 us per layer)
        with (T.autocast(grad_output.device.type, dtype=ctx.autocast) if ctx.autocast is not None else _NO_CONTEXT):
            if ctx.needs_input_grad[0]:
                grad_input = grad_output @ weight
            if ctx.needs_input_grad[1] and ctx.stream is not None:
                if ctx.shared_input is not None:
                    _SharedSketch.drop(ctx.shared_input)
                grad_view = grad_output.reshape(-1, grad_output.shape[-1])
                if _native_sketch_available(grad_view, ctx.matmul):
                    # One kernel: S G rounded to the layer's precision, and -- when G is bf16 already, so
                    # that nothing is rounded that torch's own sum would not round -- G.sum(0) as the
                    # product of one more sketch row of ones (the bias gradient).
                    with_bias = (bias is not None and ctx.needs_input_grad[2]
                                 and grad_view.dtype == T.bfloat16 and input_proj.dtype == T.bfloat16)
                    grad_proj = _native_sketch(grad_view, ctx.proj_features, *ctx.stream, ctx.matmul, 1.0,
                                               input_proj.dtype, with_bias)
                    if with_bias:
                        grad_bias = grad_proj[ctx.proj_features].clone()     # not a view: S G is freed after the GEMM
                        grad_proj = grad_proj[:ctx.proj_features]
                else:  # feature count the TMA cannot address: same S, materialised
                    proj = T.ops.fewbit.sketch_matrix(grad_view, ctx.proj_features, grad_view.shape[0],
                                                      *ctx.stream, SKETCH_KINDS[ctx.matmul])
                    grad_proj = (proj.float() @ grad_view.float()).to(input_proj.dtype)
                # (S G)^T (S X): a small [out, P] x [P, in] product in the layer's own precision
                grad_weight = grad_proj.T @ input_proj
            elif ctx.needs_input_grad[1]:
                generator = T.Generator(ctx.generator_device)
                generator.set_state(ctx.generator_state)
                grad_view = grad_output.reshape(-1, grad_output.shape[-1])
                grad_proj = _project_grad(ctx.matmul, grad_view, ctx.proj_features, generator, input_proj)
                if ctx.matmul == 'dft':
                    # sum_p conj-free product of two spectra; its real part estimates G^T X.  (The
                    # reference drops the imaginary part of one factor first and then fails on the
                    # mixed-dtype product, functional/linear.py:213-215.)
                    grad_weight = (grad_proj.T @ input_proj).real.to(weight.dtype)
                else:
                    grad_weight = grad_proj.T @ input_proj
            if bias is not None and ctx.needs_input_grad[2] and grad_bias is None:
                grad_bias = grad_output.reshape(-1, grad_output.shape[-1]).sum(dim=0)
        if grad_weight is not None and grad_weight.dtype != weight.dtype:
            grad_weight = grad_weight.to(weight.dtype)
        if grad_bias is not None and grad_bias.dtype != bias.dtype:
            grad_bias = grad_bias.to(bias.dtype)
        return (grad_input, grad_weight, grad_bias) + (None, ) * 7


linear_grp = LinearGRPFunc.apply
linear_randomized = linear_grp


class LinearCRSFunc(T.autograd.Function):
    """Linear layer whose weight gradient comes from column-row sampling over the input FEATURES
    (reference ``fewbit/functional/linear.py:27-66``): ``nopairs`` feature indices are drawn
    uniformly with replacement, the distinct ones are kept with weight ``count * in_features /
    nopairs``, and only those columns of the input are saved.  The forward result is exact; the
    weight gradient is unbiased, zero outside the sampled columns.  Plain PyTorch on either device
    (a gather and a small product -- no kernel of its own); the indices are drawn from the CPU
    default generator like the reference does, so the same seed selects the same columns.
    """

    @staticmethod
    def forward(ctx, input: T.Tensor, weight: T.Tensor, bias: Optional[T.Tensor], nopairs: int) -> T.Tensor:
        in_features = weight.shape[1]
        draws = T.randint(0, in_features, (nopairs, )).to(weight.device)
        hits = T.bincount(draws, minlength=in_features)
        pairs = T.nonzero(hits, as_tuple=True)[0]
        scale = hits[pairs].to(T.float32) * (in_features / nopairs)
        ctx.save_for_backward(input[..., pairs] * scale.to(input.dtype), weight, bias, pairs)
        return _linear_owning_output(input.reshape(-1, in_features), weight, bias, input.shape)

    @staticmethod
    def backward(ctx, grad_output):
        input_proj, weight, bias, pairs = ctx.saved_tensors
        grad_input = grad_weight = grad_bias = None
        grad_view = grad_output.reshape(-1, grad_output.shape[-1])
        if ctx.needs_input_grad[0]:
            grad_input = grad_output @ weight
        if ctx.needs_input_grad[1]:
            grad_weight = T.zeros_like(weight)
            grad_weight[:, pairs] = grad_view.T @ input_proj.reshape(-1, input_proj.shape[-1])
        if bias is not None and ctx.needs_input_grad[2]:
            grad_bias = grad_view.sum(dim=0)
        return grad_input, grad_weight, grad_bias, None


linear_crs = LinearCRSFunc.apply
