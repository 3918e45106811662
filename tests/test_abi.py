"""The C-ABI library loads on a machine without a GPU and exports exactly what
include/fewbit_b200.h declares (no compute calls here)."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

from fewbit_b200 import native

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / 'include' / 'fewbit_b200.h'


def declared_symbols():
    text = HEADER.read_text()
    return sorted(set(re.findall(r'FEWBIT_API[^;(]*?\b(fewbit_\w+)\s*\(', text)))


def test_header_is_plain_c():
    # The boundary must be consumable from C (cgo / JNI / ctypes hosts): compile it as C11.
    src = '#include "fewbit_b200.h"\nint main(void) { return FEWBIT_B200_ABI_VERSION - 1; }\n'
    subprocess.run(['gcc', '-std=c11', '-Wall', '-Werror', '-fsyntax-only', '-I', str(HEADER.parent),
                    '-x', 'c', '-'], input=src.encode(), check=True)


def test_every_declared_symbol_is_exported_and_bound():
    symbols = declared_symbols()
    assert len(symbols) >= 15
    handle = ctypes.CDLL(str(native.LIBRARY))
    for name in symbols:
        assert hasattr(handle, name), f'{name} declared in the header but not exported'
    assert sorted(native.PROTOTYPES) == symbols  # the Python binding covers the whole header


def test_only_the_abi_is_exported():
    out = subprocess.run(['nm', '-D', '--defined-only', str(native.LIBRARY)], check=True,
                         capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if ' T ' in line}
    assert exported == set(declared_symbols())


def test_host_only_entry_points():
    lib = native.lib()
    assert lib.fewbit_abi_version() == 1
    assert native.state_bytes(0, 3) == 0
    assert native.state_bytes(1, 3) == 1
    assert native.state_bytes(128 * 128 * 3072, 3) == 18874368
    assert native.state_bytes(1 << 29, 1) == 64 << 20            # config 2: 64 MiB mask
    assert native.state_bytes(5_000_000_000, 8) == 5_000_000_000  # 64-bit sizes
    # ceil(log2(L)), 1 for L <= 2 (fewbit/cpu/gelu.cc:36)
    assert [native.bits_for_levels(n) for n in (1, 2, 3, 4, 5, 8, 9, 16, 17, 255, 256)] == \
        [1, 1, 2, 2, 3, 3, 4, 4, 5, 8, 8]
    assert lib.fewbit_error_string(0) == b'success'
    for code in (-1, -2, -3, -4):
        assert lib.fewbit_error_string(code).startswith(b'fewbit:')


def test_argument_errors_need_no_gpu():
    lib = native.lib()
    # validation happens before any CUDA call
    assert lib.fewbit_stepwise_forward(99, 0, 1, 1, 1, 8, 3, 1, 7, 1.0, 20.0, None) == -3
    assert lib.fewbit_stepwise_forward(2, 0, 1, 1, 1, 8, 9, 1, 7, 1.0, 20.0, None) == -1
    assert lib.fewbit_stepwise_forward(2, 7, 16, 16, 16, 8, 3, 16, 7, 1.0, 20.0, None) == -2
    assert lib.fewbit_stepwise_forward(2, 0, None, 16, 16, 8, 3, 16, 7, 1.0, 20.0, None) == -1
    assert lib.fewbit_stepwise_forward(2, 0, 18, 16, 16, 8, 3, 16, 7, 1.0, 20.0, None) == -4
    assert lib.fewbit_stepwise_forward(2, 0, 16, 16, 16, 8, 3, 16, 8, 1.0, 20.0, None) == -1
    assert lib.fewbit_stepwise_backward(0, 16, 16, 16, 8, 3, 16, 9, None) == -1
    assert lib.fewbit_piecewise_forward(8, 0, 16, 16, 16, 8, 0.0, 0.0, None) == -3
    # empty tensors are a no-op, not an error (the reference divides by zero-ish)
    assert lib.fewbit_stepwise_forward(2, 0, None, None, None, 0, 3, None, 7, 1.0, 20.0, None) == 0
    assert lib.fewbit_piecewise_backward(4, 1, None, None, None, 0, 0.0, None) == 0


def test_operator_library_schemas():
    """torch.ops.fewbit carries the reference's schema strings verbatim (fewbit/fewbit.cc:6-37)."""
    import torch
    import fewbit_b200
    assert fewbit_b200.native_loaded(), fewbit_b200.NATIVE_ERROR
    want = {
        'hardshrink': 'fewbit::hardshrink(Tensor(a!) self, float lambd=0.5) -> Tensor(a!)',
        'hardsigmoid': 'fewbit::hardsigmoid(Tensor(a!) self) -> Tensor(a!)',
        'hardtanh': 'fewbit::hardtanh(Tensor(a!) self, float min_val=-1., float max_val=1.) -> Tensor(a!)',
        'leaky_relu': 'fewbit::leaky_relu(Tensor(a!) self, float negative_slope=0.01) -> Tensor(a!)',
        'relu': 'fewbit::relu(Tensor(a!) self) -> Tensor(a!)',
        'relu6': 'fewbit::relu6(Tensor(a!) self) -> Tensor(a!)',
        'softshrink': 'fewbit::softshrink(Tensor(a!) self, float lambd=0.5) -> Tensor(a!)',
        'threshold': 'fewbit::threshold(Tensor(a!) self, float threshold, float value) -> Tensor(a!)',
        'celu': 'fewbit::celu(Tensor(a!) self, Tensor bounds, Tensor levels, float alpha=1.) -> Tensor(a!)',
        'elu': 'fewbit::elu(Tensor(a!) self, Tensor bounds, Tensor levels, float alpha=1.) -> Tensor(a!)',
        'softplus': 'fewbit::softplus(Tensor(a!) self, Tensor bounds, Tensor levels, float beta=1., float threshold=20.) -> Tensor(a!)',
        'stepwise': 'fewbit::stepwise(Tensor(a!) self, Tensor bounds, Tensor levels, bool? parity=None, int[2]? shift=None) -> Tensor(a!)',
        'quantize': 'fewbit::quantize(Tensor inputs, Tensor bounds) -> (Tensor, Tensor)',
        'quantize_backward': 'fewbit::quantize_backward(Tensor grads, Tensor buffer, Tensor levels) -> Tensor',
    }
    for name in ('gelu', 'hardswish', 'logsigmoid', 'mish', 'selu', 'sigmoid', 'silu', 'softsign',
                 'tanh', 'tanhshrink'):
        want[name] = f'fewbit::{name}(Tensor(a!) self, Tensor bounds, Tensor levels) -> Tensor(a!)'
    assert len(want) == 24
    for name, schema in want.items():
        assert str(getattr(torch.ops.fewbit, name).default._schema) == schema


def test_cpu_tensor_is_rejected_by_the_cuda_operators():
    """As in the reference, only gelu / quantize / quantize_backward exist for CPU tensors
    (fewbit/cpu/gelu.cc:74-76); every other operator is CUDA-only and says so."""
    import torch
    import fewbit_b200  # noqa: F401
    with pytest.raises(NotImplementedError):
        torch.ops.fewbit.relu(torch.zeros(8))
    with pytest.raises(NotImplementedError):
        torch.ops.fewbit.silu(torch.zeros(8), torch.zeros(7), torch.zeros(8))


def test_cpu_operators_reproduce_the_reference(golden_ops):
    """torch.ops.fewbit.quantize / quantize_backward / gelu on CPU tensors against the fixtures
    produced by the unmodified reference CPU ops: packed bytes, gradients and values bit for bit
    (both sides use ATen's gelu and searchsorted; the packer here is parallel, the stream the same)."""
    import numpy as np
    import torch
    import fewbit_b200  # noqa: F401
    for case in golden_ops:
        if case['bf16']:
            conv = lambda a: torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)  # noqa: E731
            raw = lambda t: t.contiguous().view(torch.int16).numpy().view(np.uint16)  # noqa: E731
        else:
            conv = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
            raw = lambda t: t.contiguous().numpy()  # noqa: E731
        x, g, bounds, levels = (conv(case[k]) for k in ('x', 'g', 'bounds', 'levels'))
        y, state = torch.ops.fewbit.quantize(x, bounds)
        gin = torch.ops.fewbit.quantize_backward(g, state, levels)
        assert np.array_equal(state.numpy(), case['state']), case['key']
        assert np.array_equal(raw(gin).view(np.uint8), case['gin'].view(np.uint8)), case['key']
        assert np.array_equal(raw(y).view(np.uint8), case['y'].view(np.uint8)), case['key']
        if case['n'] > 1 and not case['bf16']:
            leaf = x.clone().requires_grad_()
            out = torch.ops.fewbit.gelu(leaf, bounds, levels)
            out.backward(g)
            assert np.array_equal(leaf.grad.numpy().view(np.uint8), case['gin'].view(np.uint8)), case['key']


@pytest.mark.parametrize('kind', ['gaussian', 'rademacher'])
def test_projection_launch_plans_are_valid_for_every_shape(kind):
    """fewbit_sketch_plan (host only): the invariants the projection kernel relies on, over a grid of
    shapes -- MMA N a multiple of 16 within TMEM, S slots that span two stages (even stage count per
    split, splits cover all tokens), 1024-byte aligned S tiles, rings inside 227 KB, pairs only on full
    768-feature slabs."""
    from fewbit_b200 import native
    for sms in (148, 132):
        for tokens in (1, 63, 64, 65, 1000, 4100, 16384, 100000):
            for features in (8, 72, 384, 392, 512, 768, 1024, 1536, 3072, 4096):
                for rows in (1, 15, 16, 144, 145, 160, 161, 3276, 3277, 10000):
                    plan = native.sketch_plan(tokens, features, rows, kind, sms)
                    where = f'{tokens}x{features}->{rows} {kind} sms={sms}: {plan}'
                    bn, share = plan['bn'], plan['share']
                    assert 64 <= bn <= 160 and bn % 16 == 0 and 3 * bn <= 512, where
                    assert share in (1, 2) and (bn // 8) % share == 0, where
                    assert not plan['pair'] or (share == 2 and features % 768 == 0), where
                    stages = -(-max(tokens, 1) // 64)
                    assert plan['split_k'] >= 1 and plan['stages_per_split'] % 2 == 0, where
                    assert plan['split_k'] * plan['stages_per_split'] >= stages, where
                    tile_rows = -(-(bn // 2 if plan['pair'] else bn) // 8) * 8
                    assert plan['s_tile_bytes'] == tile_rows * 128 and plan['s_tile_bytes'] % 1024 == 0, where
                    assert 2 <= plan['s_slots'] <= 4, where
                    assert plan['smem_bytes'] == 3 * 49152 + plan['s_slots'] * 2 * plan['s_tile_bytes'] + 512 + 1024, where
                    assert plan['smem_bytes'] <= 232448, where
