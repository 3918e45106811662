"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

ctypes/numpy front end of ``oracle/fewbit_oracle.c``, the scalar CPU restatement
of the reference's quantized-gradient activation path, plus loaders for the
unmodified reference compiled into ``oracle/_ref/``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg
may import this package.  ``fewbit_b200`` never does.

Parity status: PINNED -- see ``tests/test_oracle.py`` (reference golden vectors,
the reference's own CPU build, committed fixtures under ``tests/golden/``).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent

CONTINUOUS = ('celu', 'elu', 'gelu', 'hardswish', 'logsigmoid', 'mish', 'selu', 'sigmoid',
              'silu', 'softplus', 'softsign', 'tanh', 'tanhshrink')
PIECEWISE = ('hardshrink', 'hardsigmoid', 'hardtanh', 'leaky_relu', 'relu', 'relu6',
             'softshrink', 'threshold')

NAN_TO_ZERO, NAN_TO_LAST = 0, 1

_lib = None


def build(force: bool = False) -> Path:
    """Compile the C oracle (and, when the reference tree is mounted, oracle/_ref)."""
    target = HERE / 'liboracle.so'
    if force or not target.exists() or \
            target.stat().st_mtime < (HERE / 'fewbit_oracle.c').stat().st_mtime:
        subprocess.run(['make', '-C', str(HERE), 'oracle'], check=True, capture_output=True)
    return target


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.orc_state_bytes.restype = C.c_size_t
        _lib.orc_state_bytes.argtypes = [C.c_int64, C.c_int]
        _lib.orc_bits_for_levels.restype = C.c_int
        _lib.orc_bits_for_levels.argtypes = [C.c_int]
        _lib.orc_f32_to_bf16.restype = C.c_uint16
        _lib.orc_f32_to_bf16.argtypes = [C.c_float]
        _lib.orc_bf16_to_f32.restype = C.c_float
        _lib.orc_bf16_to_f32.argtypes = [C.c_uint16]
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


def state_bytes(n: int, bits: int) -> int:
    return int(lib().orc_state_bytes(n, bits))


def bits_for_levels(nlevels: int) -> int:
    return int(lib().orc_bits_for_levels(nlevels))


# ------------------------------------------------------------------ codec --

def deflate(codes, bits: int) -> np.ndarray:
    codes = _c(codes, np.int32).ravel()
    out = np.empty(state_bytes(codes.size, bits), np.uint8)
    lib().orc_deflate(_ptr(codes), C.c_int64(codes.size), C.c_int(bits), _ptr(out))
    return out


def inflate(state, n: int, bits: int) -> np.ndarray:
    state = _c(state, np.uint8).ravel()
    assert state.size >= state_bytes(n, bits)
    codes = np.empty(n, np.int32)
    lib().orc_inflate(_ptr(codes), C.c_int64(n), C.c_int(bits), _ptr(state))
    return codes


def deflate_numpy(codes, bits: int) -> np.ndarray:
    """Third witness: the same stream written with numpy only."""
    codes = np.asarray(codes, np.int64).ravel()
    if codes.size == 0:
        return np.zeros(0, np.uint8)
    planes = ((codes[:, None] >> np.arange(bits)) & 1).astype(np.uint8)
    return np.packbits(planes.ravel(), bitorder='little')


def inflate_numpy(state, n: int, bits: int) -> np.ndarray:
    flat = np.unpackbits(np.asarray(state, np.uint8), bitorder='little')[:n * bits]
    return (flat.reshape(n, bits).astype(np.int32) << np.arange(bits)).sum(axis=1).astype(np.int32)


# ------------------------------------------------------ bf16 as raw uint16 --

def f32_to_bf16_bits(a) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 bit patterns (uint16)."""
    u = _c(a, np.float32).view(np.uint32).astype(np.uint64)
    nan = (u & 0x7fffffff) > 0x7f800000
    r = ((u + 0x7fff + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    r[nan] = ((u[nan] >> 16) | 0x40).astype(np.uint16)
    return r


def bf16_bits_to_f32(a) -> np.ndarray:
    return (_c(a, np.uint16).astype(np.uint32) << 16).view(np.float32)


# -------------------------------------------------------------- bucketize --

def bucketize(x, bounds, nan_policy: int = NAN_TO_ZERO) -> np.ndarray:
    x = np.ascontiguousarray(x)
    codes = np.empty(x.size, np.int32)
    if x.dtype == np.uint16:  # bf16 bit patterns
        bounds = _c(bounds, np.uint16)
        lib().orc_bucketize_bf16(_ptr(x), C.c_int64(x.size), _ptr(bounds), C.c_int(bounds.size),
                                 C.c_int(nan_policy), _ptr(codes))
    else:
        x = _c(x, np.float32)
        bounds = _c(bounds, np.float32)
        lib().orc_bucketize_f32(_ptr(x), C.c_int64(x.size), _ptr(bounds), C.c_int(bounds.size),
                                C.c_int(nan_policy), _ptr(codes))
    return codes


# ------------------------------------------------- continuous activations --

def stepwise_forward(func: str, x, bounds, bits: int | None = None, p0: float = 1.0,
                     p1: float = 20.0, nan_policy: int = NAN_TO_ZERO):
    """-> (y, state).  fp32 arrays, or uint16 arrays holding bf16 bit patterns."""
    fid = CONTINUOUS.index(func)
    x = np.ascontiguousarray(x).ravel()
    bf16 = x.dtype == np.uint16
    dt = np.uint16 if bf16 else np.float32
    x = _c(x, dt)
    bounds = _c(bounds, dt)
    if bits is None:
        bits = bits_for_levels(bounds.size + 1)
    y = np.empty_like(x)
    state = np.empty(state_bytes(x.size, bits), np.uint8)
    fn = lib().orc_stepwise_forward_bf16 if bf16 else lib().orc_stepwise_forward_f32
    fn(C.c_int(fid), _ptr(x), _ptr(y), _ptr(state), C.c_int64(x.size), C.c_int(bits),
       _ptr(bounds), C.c_int(bounds.size), C.c_double(p0), C.c_double(p1), C.c_int(nan_policy))
    return y, state


def stepwise_backward(state, gout, levels, bits: int | None = None) -> np.ndarray:
    gout = np.ascontiguousarray(gout).ravel()
    bf16 = gout.dtype == np.uint16
    dt = np.uint16 if bf16 else np.float32
    gout = _c(gout, dt)
    levels = _c(levels, dt)
    state = _c(state, np.uint8)
    if bits is None:
        bits = bits_for_levels(levels.size)
    assert state.size >= state_bytes(gout.size, bits)
    gin = np.empty_like(gout)
    fn = lib().orc_stepwise_backward_bf16 if bf16 else lib().orc_stepwise_backward_f32
    fn(_ptr(state), _ptr(gout), _ptr(gin), C.c_int64(gout.size), C.c_int(bits), _ptr(levels))
    return gin


def stepwise_custom_forward(x, bounds, levels, anchor: float = 0.0, bits: int | None = None):
    """-> (y, state) of the custom-table operator `stepwise` (reference schema fewbit/fewbit.cc:37;
    no reference kernel exists -- this restates the definition in include/fewbit_b200.h):
    F continuous and piecewise linear, F' = levels[code(x)] with the reference's bucket search
    (fewbit/cuda/codec.cu:118-131), F(anchor) = 0; y = fma(levels[code], x, intercept[code]).
    fp32 arrays, or uint16 arrays holding bf16 bit patterns."""
    x = np.ascontiguousarray(x).ravel()
    bf16 = x.dtype == np.uint16
    as_f32 = bf16_bits_to_f32 if bf16 else (lambda a: _c(a, np.float32))
    b, l, xv = as_f32(bounds).astype(np.float64), as_f32(levels).astype(np.float64), as_f32(x)
    if bits is None:
        bits = bits_for_levels(l.size)
    codes = bucketize(x, bounds)
    rise = np.zeros(l.size)                       # rise[k] = F~(bounds[k]), F~(bounds[0]) = 0
    if b.size > 1:
        rise[1:b.size] = np.cumsum(l[1:b.size] * np.diff(b))
    left = np.clip(np.arange(l.size) - 1, 0, max(b.size - 1, 0))
    intercept = rise[left] - l * b[left] if b.size else np.zeros(l.size)
    piece = int(np.searchsorted(b, np.float32(anchor), side='left'))
    intercept = (intercept - (l[piece] * np.float64(np.float32(anchor)) + intercept[piece])).astype(np.float32)
    # one fused multiply-add in fp32: the product of two floats is exact in double
    y = (l[codes] * xv.astype(np.float64) + intercept[codes].astype(np.float64)).astype(np.float32)
    return (f32_to_bf16_bits(y) if bf16 else y), deflate(codes, bits)


# ------------------------------------------------------- piecewise, 1 bit --

def piecewise_forward(func: str, x, p0: float = 0.0, p1: float = 0.0):
    fid = PIECEWISE.index(func)
    x = np.ascontiguousarray(x).ravel()
    bf16 = x.dtype == np.uint16
    x = _c(x, np.uint16 if bf16 else np.float32)
    y = np.empty_like(x)
    state = np.empty(state_bytes(x.size, 1), np.uint8)
    fn = lib().orc_piecewise_forward_bf16 if bf16 else lib().orc_piecewise_forward_f32
    fn(C.c_int(fid), _ptr(x), _ptr(y), _ptr(state), C.c_int64(x.size), C.c_double(p0),
       C.c_double(p1))
    return y, state


def piecewise_backward(func: str, state, gout, p0: float = 0.0) -> np.ndarray:
    fid = PIECEWISE.index(func)
    gout = np.ascontiguousarray(gout).ravel()
    bf16 = gout.dtype == np.uint16
    gout = _c(gout, np.uint16 if bf16 else np.float32)
    state = _c(state, np.uint8)
    assert state.size >= state_bytes(gout.size, 1)
    gin = np.empty_like(gout)
    fn = lib().orc_piecewise_backward_bf16 if bf16 else lib().orc_piecewise_backward_f32
    fn(C.c_int(fid), _ptr(state), _ptr(gout), _ptr(gin), C.c_int64(gout.size), C.c_double(p0))
    return gin


# --------------------------------------------- the real reference (_ref/) --

def ref_codec():
    """ctypes handle of oracle/_ref/libref_codec.so (reference fewbit/cpu/codec.h), or None."""
    path = HERE / '_ref' / 'libref_codec.so'
    if not path.exists():
        return None
    return C.CDLL(str(path))


def ref_deflate(codes, bits: int) -> np.ndarray:
    codes = _c(codes, np.int32).ravel()
    out = np.zeros(state_bytes(codes.size, bits), np.uint8)
    ref_codec().ref_deflate_u8(_ptr(codes), C.c_int64(codes.size), C.c_int32(bits), _ptr(out))
    return out


def ref_inflate(state, n: int, bits: int) -> np.ndarray:
    state = _c(state, np.uint8)
    codes = np.zeros(n, np.int32)
    ref_codec().ref_inflate_u8(_ptr(codes), C.c_int64(n), C.c_int32(bits), _ptr(state))
    return codes


def ref_ops_path() -> Path | None:
    """oracle/_ref/libfewbit_ref.so: the reference's CPU torch ops (namespace ``fewbit``).

    It registers the same op namespace as the product library, so it must be loaded in a
    process that has NOT loaded fewbit_b200 (see oracle/ref_runner.py).
    """
    path = HERE / '_ref' / 'libfewbit_ref.so'
    return path if path.exists() else None


# ------------------------------------------------------------------ sketch entries ----
# The projection's random matrix S is OUR definition (the reference draws torch.randn; SURVEY 8c:
# "the bit pattern of S is parity unpinned"), so the oracle restates it: Philox4x32 with
# PHILOX_ROUNDS rounds (Salmon et al., SC'11: multipliers 0xD2511F53 / 0xCD9E8D57, key increments
# 0x9E3779B9 / 0xBB67AE85), key = the 64-bit seed, counter = (column block, row, offset lo, offset hi).
# Rademacher: counter column block = n // 128, bit b of output word w is entry 128 (n // 128) +
# 32 w + b, set bit = +1/2, clear bit = -1/2.  Gaussian: column block = n // 8, each output word
# gives two normals by Box-Muller on its two 16-bit halves ((h + 1/2) / 65536): low half -> radius
# sqrt(-2 ln u), high half -> angle 2 pi u; entries 2 w (cos) and 2 w + 1 (sin) of the octet.
# fewbit_b200/csrc/sketch.cu: Philox, normal_pair, sign_pair, sketch_matrix_kernel.
PHILOX_ROUNDS = 7


def philox4x32(counter: np.ndarray, seed: int, rounds: int = PHILOX_ROUNDS) -> np.ndarray:
    """counter: uint32 [..., 4] -> uint32 [..., 4]."""
    c = np.asarray(counter, dtype=np.uint64).copy()
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(rounds):
        p0 = np.uint64(0xD2511F53) * c[..., 0]
        p1 = np.uint64(0xCD9E8D57) * c[..., 2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = np.stack([hi1 ^ c[..., 1] ^ k0, lo1, hi0 ^ c[..., 3] ^ k1, lo0], axis=-1)
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c.astype(np.uint32)


def sketch_matrix(rows: int, cols: int, seed: int, offset: int, kind: str = 'rademacher') -> np.ndarray:
    """S as float64 [rows, cols]: exact for 'rademacher' (+-1/2); for 'gaussian' the ideal Box-Muller
    values before the kernel's MUFU approximations and bf16 rounding."""
    p = np.arange(rows, dtype=np.uint64)[:, None]
    lo, hi = offset & 0xFFFFFFFF, (offset >> 32) & 0xFFFFFFFF
    if kind == 'rademacher':
        blocks = np.arange((cols + 127) // 128, dtype=np.uint64)[None, :]
        counter = np.stack(np.broadcast_arrays(blocks, p, np.uint64(lo), np.uint64(hi)), axis=-1)
        words = philox4x32(counter, seed)                                   # [rows, blocks, 4]
        bits = (words[..., None] >> np.arange(32, dtype=np.uint32)) & 1      # [rows, blocks, 4, 32]
        return (bits.reshape(rows, -1)[:, :cols].astype(np.float64) - 0.5)
    octets = np.arange((cols + 7) // 8, dtype=np.uint64)[None, :]
    counter = np.stack(np.broadcast_arrays(octets, p, np.uint64(lo), np.uint64(hi)), axis=-1)
    words = philox4x32(counter, seed).astype(np.float64)                    # [rows, octets, 4]
    u_radius = ((words.astype(np.uint64) & 0xFFFF) + 0.5) / 65536.0
    u_angle = ((words.astype(np.uint64) >> 16) + 0.5) / 65536.0
    radius, angle = np.sqrt(-2.0 * np.log(u_radius)), 2.0 * np.pi * u_angle
    pairs = np.stack([radius * np.cos(angle), radius * np.sin(angle)], axis=-1)   # [rows, octets, 4, 2]
    return pairs.reshape(rows, -1)[:, :cols]
