"""ctypes binding of the C ABI (``include/fewbit_b200.h``, ``libfewbit_b200.so``).

This is the same boundary a non-Python host (cgo, JNI, ...) would bind; the parity tests and
``bench.py`` call the kernels through it with raw device pointers taken from torch tensors.
Every wrapper raises ``RuntimeError`` on a non-zero status -- nothing falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

# FEWBIT_B200_LIBRARY: load another build of the kernels (tuning sweeps only).
LIBRARY = Path(os.environ.get('FEWBIT_B200_LIBRARY') or Path(__file__).with_name('libfewbit_b200.so'))

F32, BF16 = 0, 1
CONTINUOUS = ('celu', 'elu', 'gelu', 'hardswish', 'logsigmoid', 'mish', 'selu', 'sigmoid', 'silu',
              'softplus', 'softsign', 'tanh', 'tanhshrink')
PIECEWISE = ('hardshrink', 'hardsigmoid', 'hardtanh', 'leaky_relu', 'relu', 'relu6', 'softshrink',
             'threshold')

# name -> (restype, argtypes): must list every symbol include/fewbit_b200.h declares
# (tests/test_abi.py checks both directions).
_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
PROTOTYPES = {
    'fewbit_abi_version': (_i, []),
    'fewbit_error_string': (C.c_char_p, [_i]),
    'fewbit_state_bytes': (C.c_size_t, [_i64, _i]),
    'fewbit_bits_for_levels': (_i, [_i]),
    'fewbit_stepwise_forward': (_i, [_i, _i, _vp, _vp, _vp, _i64, _i, _vp, _i, _d, _d, _vp]),
    'fewbit_stepwise_backward': (_i, [_i, _vp, _vp, _vp, _i64, _i, _vp, _i, _vp]),
    'fewbit_stepwise_custom_forward': (_i, [_i, _vp, _vp, _vp, _i64, _i, _vp, _i, _vp, _i, _d, _vp]),
    'fewbit_piecewise_forward': (_i, [_i, _i, _vp, _vp, _vp, _i64, _d, _d, _vp]),
    'fewbit_piecewise_backward': (_i, [_i, _i, _vp, _vp, _vp, _i64, _d, _vp]),
    'fewbit_deflate': (_i, [_vp, _vp, _i64, _i, _vp]),
    'fewbit_inflate': (_i, [_vp, _vp, _i64, _i, _vp]),
    'fewbit_stepwise_forward_host': (_i, [_i, _i, _vp, _vp, _vp, _i64, _i, _vp, _i, _d, _d, _i64]),
    'fewbit_stepwise_backward_host': (_i, [_i, _vp, _vp, _vp, _i64, _i, _vp, _i, _i64]),
    'fewbit_piecewise_forward_host': (_i, [_i, _i, _vp, _vp, _vp, _i64, _d, _d, _i64]),
    'fewbit_piecewise_backward_host': (_i, [_i, _i, _vp, _vp, _vp, _i64, _d, _i64]),
    'fewbit_sketch_workspace_bytes': (C.c_size_t, [_i64, _i, _i]),
    'fewbit_sketch_forward': (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, C.c_float, C.c_uint64, C.c_uint64, _vp]),
    'fewbit_sketch_project': (_i, [_vp, _vp, _i, _vp, _i64, _i, _i, _i, _i, C.c_float, C.c_uint64, C.c_uint64, _vp]),
    'fewbit_sketch_matrix': (_i, [_vp, _i, _i64, _i, C.c_uint64, C.c_uint64, _vp]),
    'fewbit_sketch_plan': (_i, [_i64, _i, _i, _i, _i, C.POINTER(C.c_int)]),
    'fewbit_launch_count': (_i64, []),
}

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIBRARY.exists():
            raise RuntimeError(f'{LIBRARY} is missing: build it with `python -c "import '
                               f'__graft_entry__ as g; g.build()"` or `make -C fewbit_b200/csrc`.')
        handle = C.CDLL(str(LIBRARY))
        for name, (restype, argtypes) in PROTOTYPES.items():
            if not hasattr(handle, name) and 'FEWBIT_B200_LIBRARY' in os.environ:
                continue        # tuning runs against an older build of the kernels
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = restype, argtypes
        _lib = handle
    return _lib


def check(status: int, what: str):
    if status != 0:
        raise RuntimeError(f'{what} failed ({status}): {lib().fewbit_error_string(status).decode()}')


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f'fewbit_b200: unsupported dtype {t.dtype}')


def _func(table, func) -> int:
    return table.index(func) if isinstance(func, str) else int(func)


def _stream(stream=None) -> int:
    if stream is None:
        stream = torch.cuda.current_stream()
    return stream.cuda_stream


def state_bytes(n: int, bits: int) -> int:
    return int(lib().fewbit_state_bytes(n, bits))


def bits_for_levels(nlevels: int) -> int:
    return int(lib().fewbit_bits_for_levels(nlevels))


def launch_count() -> int:
    return int(lib().fewbit_launch_count())


def new_state(x: torch.Tensor, bits: int) -> torch.Tensor:
    return torch.empty(state_bytes(x.numel(), bits), dtype=torch.uint8, device=x.device)


def stepwise_forward(func, x, y, state, bits, bounds, p0=1.0, p1=20.0, stream=None):
    """y = f(x) (y may be x), state = pack(bucketize(x, bounds), bits).  Device tensors."""
    check(lib().fewbit_stepwise_forward(_func(CONTINUOUS, func), dtype_code(x), x.data_ptr(),
                                        y.data_ptr(), state.data_ptr(), x.numel(), bits,
                                        bounds.data_ptr(), bounds.numel(), p0, p1, _stream(stream)),
          'fewbit_stepwise_forward')


def stepwise_custom_forward(x, y, state, bits, bounds, levels, anchor=0.0, stream=None):
    """y = F(x), F piecewise linear with slopes `levels` and kinks at `bounds`, F(anchor) = 0."""
    check(lib().fewbit_stepwise_custom_forward(dtype_code(x), x.data_ptr(), y.data_ptr(), state.data_ptr(),
                                               x.numel(), bits, bounds.data_ptr(), bounds.numel(),
                                               levels.data_ptr(), levels.numel(), anchor, _stream(stream)),
          'fewbit_stepwise_custom_forward')


def stepwise_backward(state, gout, gin, bits, levels, stream=None):
    check(lib().fewbit_stepwise_backward(dtype_code(gout), state.data_ptr(), gout.data_ptr(),
                                         gin.data_ptr(), gout.numel(), bits, levels.data_ptr(),
                                         levels.numel(), _stream(stream)),
          'fewbit_stepwise_backward')


def piecewise_forward(func, x, y, state, p0=0.0, p1=0.0, stream=None):
    check(lib().fewbit_piecewise_forward(_func(PIECEWISE, func), dtype_code(x), x.data_ptr(),
                                         y.data_ptr(), state.data_ptr(), x.numel(), p0, p1,
                                         _stream(stream)),
          'fewbit_piecewise_forward')


def piecewise_backward(func, state, gout, gin, p0=0.0, stream=None):
    check(lib().fewbit_piecewise_backward(_func(PIECEWISE, func), dtype_code(gout),
                                          state.data_ptr(), gout.data_ptr(), gin.data_ptr(),
                                          gout.numel(), p0, _stream(stream)),
          'fewbit_piecewise_backward')


def deflate(codes, state, bits, stream=None):
    check(lib().fewbit_deflate(codes.data_ptr(), state.data_ptr(), codes.numel(), bits,
                               _stream(stream)), 'fewbit_deflate')


def inflate(state, codes, bits, stream=None):
    check(lib().fewbit_inflate(state.data_ptr(), codes.data_ptr(), codes.numel(), bits,
                               _stream(stream)), 'fewbit_inflate')


# Host-buffer variants: x/y/gout/gin are (pinned) HOST tensors, state and tables stay on device.

def stepwise_forward_host(func, x_host, y_host, state, bits, bounds, p0=1.0, p1=20.0, chunk=0):
    check(lib().fewbit_stepwise_forward_host(_func(CONTINUOUS, func), dtype_code(x_host),
                                             x_host.data_ptr(), y_host.data_ptr(),
                                             state.data_ptr(), x_host.numel(), bits,
                                             bounds.data_ptr(), bounds.numel(), p0, p1, chunk),
          'fewbit_stepwise_forward_host')


def stepwise_backward_host(state, gout_host, gin_host, bits, levels, chunk=0):
    check(lib().fewbit_stepwise_backward_host(dtype_code(gout_host), state.data_ptr(),
                                              gout_host.data_ptr(), gin_host.data_ptr(),
                                              gout_host.numel(), bits, levels.data_ptr(),
                                              levels.numel(), chunk),
          'fewbit_stepwise_backward_host')


def piecewise_forward_host(func, x_host, y_host, state, p0=0.0, p1=0.0, chunk=0):
    check(lib().fewbit_piecewise_forward_host(_func(PIECEWISE, func), dtype_code(x_host),
                                              x_host.data_ptr(), y_host.data_ptr(),
                                              state.data_ptr(), x_host.numel(), p0, p1, chunk),
          'fewbit_piecewise_forward_host')


def piecewise_backward_host(func, state, gout_host, gin_host, p0=0.0, chunk=0):
    check(lib().fewbit_piecewise_backward_host(_func(PIECEWISE, func), dtype_code(gout_host),
                                               state.data_ptr(), gout_host.data_ptr(),
                                               gin_host.data_ptr(), gout_host.numel(), p0, chunk),
          'fewbit_piecewise_backward_host')


# RandomizedLinear projection (tcgen05): out[P, D] = scale * S[P, N] @ x[N, D], S generated in-kernel.

SKETCH_KINDS = ('gaussian', 'rademacher')


def sketch_workspace(x, rows):
    """Scratch buffer fewbit_sketch_forward wants for this shape (split-K partial sums)."""
    tokens, features = x.shape
    nbytes = int(lib().fewbit_sketch_workspace_bytes(tokens, features, rows))
    return torch.empty(max(nbytes, 16), dtype=torch.uint8, device=x.device)


def sketch_forward(x, rows, seed, offset, kind='gaussian', scale=1.0, stream=None, out=None, workspace=None):
    """x: [tokens, features] bf16 device tensor -> [rows, features] fp32.  `out` / `workspace`:
    preallocated buffers (timing loops that must contain nothing but the kernels)."""
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.is_contiguous()
    tokens, features = x.shape
    if out is None:
        out = torch.empty(rows, features, dtype=torch.float32, device=x.device)
    ws = workspace if workspace is not None else sketch_workspace(x, rows)
    check(lib().fewbit_sketch_forward(x.data_ptr(), out.data_ptr(), ws.data_ptr(), tokens, features, rows,
                                      SKETCH_KINDS.index(kind), scale, seed, offset, _stream(stream)),
          'fewbit_sketch_forward')
    return out


def sketch_project(x, rows, seed, offset, kind='gaussian', scale=1.0, out_dtype=torch.float32, column_sums=False,
                   stream=None):
    """fewbit_sketch_project: [rows (+ 1), features] in fp32 or bf16; with `column_sums` the extra last
    row is scale * x.sum(0)."""
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.is_contiguous()
    tokens, features = x.shape
    total = rows + (1 if column_sums else 0)
    out = torch.empty(total, features, dtype=out_dtype, device=x.device)
    ws = sketch_workspace(x, total)
    check(lib().fewbit_sketch_project(x.data_ptr(), out.data_ptr(), 1 if out_dtype == torch.bfloat16 else 0,
                                      ws.data_ptr(), tokens, features, rows, int(column_sums),
                                      SKETCH_KINDS.index(kind), scale, seed, offset, _stream(stream)),
          'fewbit_sketch_project')
    return out


def sketch_plan(tokens, features, rows, kind='gaussian', sms=148):
    """The launch plan of the projection kernel for a shape (host only): dict of BN, split_k, sharing, rings."""
    out = (C.c_int * 8)()
    check(lib().fewbit_sketch_plan(tokens, features, rows, SKETCH_KINDS.index(kind), sms, out), 'fewbit_sketch_plan')
    keys = ('bn', 'split_k', 'share', 'pair', 'stages_per_split', 's_slots', 's_tile_bytes', 'smem_bytes')
    return dict(zip(keys, list(out)))


def sketch_matrix(rows, cols, seed, offset, kind='gaussian', device='cuda', stream=None):
    """The sketch S itself as [rows, cols] bf16 (tests / diagnostics)."""
    s = torch.empty(rows, cols, dtype=torch.bfloat16, device=device)
    check(lib().fewbit_sketch_matrix(s.data_ptr(), rows, cols, SKETCH_KINDS.index(kind), seed, offset,
                                     _stream(stream)), 'fewbit_sketch_matrix')
    return s
