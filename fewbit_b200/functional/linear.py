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

import weakref

from typing import Literal, Optional

import torch as T
import torch.nn.functional as F

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
    if kind in ('dct', 'dft'):
        raise NotImplementedError(f"matmul='{kind}' is outside the B200 hot path (SURVEY 2.1 #8); "
                                  "use 'gaussian' or 'rademacher'.")
    raise ValueError(f'Unexpected matmul type: {kind}.')


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


def _native_sketch(view: T.Tensor, rows: int, seed: int, offset: int, kind: str, scale: float) -> T.Tensor:
    """scale * S @ view  with S generated inside the kernel (fp32 result, [rows, features]).
    Operands enter the tensor cores as bf16 (fp32 accumulation); the rounding is unbiased and
    three orders of magnitude below the sketch's own O(1/sqrt(P)) noise."""
    return T.ops.fewbit.sketch(view.to(T.bfloat16).contiguous(), rows, seed, offset,
                               SKETCH_KINDS[kind], scale)


def _draw_stream(generator: T.Generator):
    """(seed, offset) identifying one sketch; advances the generator's Philox offset."""
    seed = generator.initial_seed() & 0x7FFFFFFFFFFFFFFF
    offset = generator.get_offset()
    generator.set_offset(offset + 4)
    return seed, offset


class _SharedSketch:
    """The last sketch taken with ``share_sketch=True``, per device: layers that are fed the very
    same tensor one after the other (the query / key / value projections of an attention block)
    reuse one ``S X`` instead of sketching it three times.  Only a weak reference
    to the input is kept -- the point of the layer is NOT to keep its input alive."""
    __slots__ = ('input', 'version', 'rows', 'kind', 'projection', 'stream')

    def matches(self, tensor: T.Tensor, rows: int, kind: str) -> bool:
        return (self.input() is tensor and self.version == tensor._version and self.rows == rows
                and self.kind == kind)


_SHARED: dict = {}


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

        generator = generator or _default_generator(input.device)
        input_view = input.reshape(-1, input.shape[-1])
        proj_features = calc_proj_dim(input_view.shape[0], proj_dim_ratio, proj_dim, proj_dim_max,
                                      proj_dim_min)

        if _native_sketch_available(input_view, matmul) and generator.device.type == 'cuda':
            # B200 path: S never exists in memory; (seed, offset) replaces the generator state.
            shared = _SHARED.get(input.device) if share_sketch else None
            if shared is not None and shared.matches(input, proj_features, matmul):
                input_proj, (seed, offset) = shared.projection, shared.stream
            else:
                seed, offset = _draw_stream(generator)
                scale = 1.0 / proj_features if matmul == 'gaussian' else 4.0 / proj_features
                input_proj = _native_sketch(input_view, proj_features, seed, offset, matmul, scale).to(input.dtype)
                if share_sketch:
                    shared = _SHARED[input.device] = _SharedSketch()
                    shared.input, shared.version = weakref.ref(input), input._version
                    shared.rows, shared.kind = proj_features, matmul
                    shared.projection, shared.stream = input_proj, (seed, offset)
            ctx.save_for_backward(input_proj, weight, bias)
            ctx.proj_features = proj_features
            ctx.matmul = matmul
            ctx.stream = (seed, offset)
            return _linear_owning_output(input_view, weight, bias, input.shape)

        generator_state = generator.get_state()
        ctx.stream = None
        proj = _sketch_matrix(matmul, proj_features, input_view.shape[0], generator, input.device,
                              input.dtype)
        # E[S^T S] = P I (gaussian) or P/4 I (rademacher): scale so that E[grad_weight] = G^T X,
        # with the very arithmetic of the reference (:137, :146).
        if matmul == 'gaussian':
            input_proj = (proj @ input_view) / proj_features
        else:
            input_proj = (proj @ input_view) * (4 / proj_features)
        del proj

        ctx.save_for_backward(input_proj, weight, bias)
        ctx.proj_features = proj_features
        ctx.matmul = matmul
        ctx.generator_state = generator_state
        ctx.generator_device = generator.device
        return _linear_owning_output(input_view, weight, bias, input.shape)

    @staticmethod
    def backward(ctx, grad_output):
        input_proj, weight, bias = ctx.saved_tensors
        grad_input = grad_weight = grad_bias = None
        if ctx.needs_input_grad[0]:
            grad_input = grad_output @ weight
        if ctx.needs_input_grad[1] and ctx.stream is not None:
            grad_view = grad_output.reshape(-1, grad_output.shape[-1])
            if _native_sketch_available(grad_view, ctx.matmul):
                grad_proj = _native_sketch(grad_view, ctx.proj_features, *ctx.stream, ctx.matmul, 1.0)
            else:  # feature count the TMA cannot address: same S, materialised
                proj = T.ops.fewbit.sketch_matrix(grad_view, ctx.proj_features, grad_view.shape[0],
                                                  *ctx.stream, SKETCH_KINDS[ctx.matmul])
                grad_proj = proj.float() @ grad_view.float()
            # (S G)^T (S X): a small [out, P] x [P, in] product in the layer's own precision
            grad_weight = grad_proj.to(input_proj.dtype).T @ input_proj
        elif ctx.needs_input_grad[1]:
            generator = T.Generator(ctx.generator_device)
            generator.set_state(ctx.generator_state)
            grad_view = grad_output.reshape(-1, grad_output.shape[-1])
            proj = _sketch_matrix(ctx.matmul, ctx.proj_features, grad_view.shape[0], generator,
                                  grad_output.device, grad_output.dtype)
            grad_weight = (proj @ grad_view).T @ input_proj
        if bias is not None and ctx.needs_input_grad[2]:
            grad_bias = grad_output.reshape(-1, grad_output.shape[-1]).sum(dim=0)
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
