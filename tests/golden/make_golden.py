"""Generate the committed golden fixtures by RUNNING THE REFERENCE in this container.

    python tests/golden/make_golden.py          # needs /root/reference and oracle/_ref/

Outputs (small, committed):
  tests/golden/reference_cpu_ops.npz    x / bounds / levels / g -> y, state, gin from the
                                        unmodified reference CPU ops torch.ops.fewbit.quantize
                                        and quantize_backward (fewbit/cpu/gelu.cc:7-45), run by
                                        oracle/ref_runner.py in a process of its own
  tests/golden/reference_tables.npz     store.get(name, bits) of the reference Python package
                                        (fewbit/functional/activations.py:46-61) cast to fp32 and
                                        bf16, i.e. the tables exactly as the operators receive them
  tests/golden/reference_linear.npz     LinearGRP forward/backward of the reference
                                        (fewbit/functional/linear.py:84-221) on CPU, seed 42

The GPU box has no /root/reference: tests only ever read the .npz files.
"""
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
REF = Path('/root/reference')
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402  (bf16 bit helpers only)


def run_reference(x, bounds, levels, g, bf16):
    with tempfile.TemporaryDirectory() as tmp:
        inp, out = Path(tmp) / 'in.npz', Path(tmp) / 'out.npz'
        np.savez(inp, x=x, bounds=bounds, levels=levels, g=g, bf16=np.int32(bf16))
        subprocess.run([sys.executable, str(ROOT / 'oracle' / 'ref_runner.py'), 'run', str(inp),
                        str(out)], check=True, capture_output=True)
        with np.load(out) as npz:
            return npz['y'], npz['state'], npz['gin']


def reference_tables():
    """Import the reference Python package on top of the reference's own CPU op library
    (torch 2.11 raises AttributeError, not RuntimeError, for a missing op, so the package
    cannot be imported without its schemas registered)."""
    import os
    os.environ['FEWBIT_NATIVE'] = '0'
    torch.ops.load_library(str(ROOT / 'oracle' / '_ref' / 'libfewbit_ref.so'))
    sys.path.insert(0, str(REF))
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import fewbit as ref  # the reference package
    assert Path(ref.__file__).is_relative_to(REF)
    return ref


def main():
    ref = reference_tables()
    store = ref.functional.activations.store

    # ---- tables as the operators receive them -------------------------------------
    tables = {}
    for (name, bits), _ in sorted(store.items()):
        for tag, dtype in (('f32', torch.float32), ('bf16', torch.bfloat16)):
            borders, levels = store.get(name, bits, 'cpu', dtype)
            bounds = borders[1:-1]
            if dtype == torch.bfloat16:
                bounds = bounds.view(torch.int16).numpy().view(np.uint16)
                levels = levels.view(torch.int16).numpy().view(np.uint16)
            else:
                bounds, levels = bounds.numpy(), levels.numpy()
            tables[f'{name}{bits:02d}-{tag}-bounds'] = bounds
            tables[f'{name}{bits:02d}-{tag}-levels'] = levels
    np.savez_compressed(ROOT / 'tests/golden/reference_tables.npz', **tables)

    # ---- reference CPU ops --------------------------------------------------------
    rng = np.random.default_rng(20240517)
    cases = {}
    specs = []
    for name, bits in (('gelu', 2), ('gelu', 3), ('gelu', 4), ('silu', 3), ('tanh', 4),
                       ('softplus', 2), ('mish', 4)):
        for n in (1, 7, 8, 9, 13, 64, 255, 256, 257, 1000, 4099):
            specs.append((name, bits, n, False))
    specs += [('gelu', 3, 101, False), ('gelu', 3, 23, False)]            # SURVEY App. B linspace cases
    specs += [('gelu', 3, 1000, True), ('gelu', 4, 4099, True), ('silu', 2, 513, True)]
    # benchmark/bench-roberta.py:128-136 table (== fewbit/cuda/codec_test.cu:16-24)
    specs.append(('roberta', 3, 16, False))
    for idx, (name, bits, n, bf16) in enumerate(specs):
        tag = 'bf16' if bf16 else 'f32'
        if name == 'roberta':
            bounds = np.array([-2.39798704e+00, -7.11248159e-01, -3.26290283e-01, -1.55338428e-04,
                               3.26182064e-01, 7.10855860e-01, 2.39811567e+00], np.float32)
            levels = np.array([-0.00260009, -0.08883533, 0.1251944, 0.37204148, 0.6277958,
                               0.87466175, 1.08880716, 1.00259936], np.float32)
            x = np.array([2.29811567e+00, 6.10855860e-01, 2.29811567e+00, -8.11248159e-01,
                          9.99900000e+02, -2.49798704e+00, 2.26182064e-01, -4.26290283e-01,
                          -4.26290283e-01, -1.00155338e-01, -2.49798704e+00, 2.26182064e-01,
                          6.10855860e-01, 6.10855860e-01, 2.29811567e+00, 9.99900000e+02],
                         np.float32)
            g = np.ones(16, np.float32)
        else:
            bounds = tables[f'{name}{bits:02d}-{tag}-bounds']
            levels = tables[f'{name}{bits:02d}-{tag}-levels']
            if n in (101, 23):
                x = torch.linspace(-5, 5, n).numpy()
            else:
                x = (rng.standard_normal(n) * 2).astype(np.float32)
                # plant exact hits on borders, signed zeros and infinities
                fb = oracle.bf16_bits_to_f32(bounds) if bf16 else bounds
                for k, v in enumerate(list(fb[:3]) + [0.0, -0.0, np.inf, -np.inf]):
                    if k < n:
                        x[(k * 7919) % n] = v
            g = rng.standard_normal(n).astype(np.float32)
            if bf16:
                x, g = oracle.f32_to_bf16_bits(x), oracle.f32_to_bf16_bits(g)
        y, state, gin = run_reference(x, bounds, levels, g, bf16)
        key = f'case{idx:03d}'
        cases[f'{key}/meta'] = np.array([bits, n, int(bf16)], np.int64)
        cases[f'{key}/name'] = np.array(name)
        for field, arr in (('x', x), ('bounds', bounds), ('levels', levels), ('g', g), ('y', y),
                           ('state', state), ('gin', gin)):
            cases[f'{key}/{field}'] = arr
    np.savez_compressed(ROOT / 'tests/golden/reference_cpu_ops.npz', **cases)

    # ---- reference RandomizedLinear on CPU ----------------------------------------
    from fewbit.modules.linear import LinearGRP
    lin = {}
    # 'dft' is absent: the reference's backward raises on it (real grad @ complex sketch, :213-215)
    for kind in ('gaussian', 'rademacher', 'dct'):
        for bias in (False, True):
            torch.manual_seed(42)
            layer = LinearGRP(24, 12, bias, proj_dim=16, matmul=kind)
            x = torch.randn(4, 10, 24, requires_grad=True)
            gy = torch.randn(4, 10, 12)
            torch.manual_seed(1234)  # fixes the sketch drawn inside forward
            y = layer(x)
            y.backward(gy)
            key = f'{kind}-{int(bias)}'
            lin[f'{key}/weight'] = layer.weight.detach().numpy()
            if bias:
                lin[f'{key}/bias'] = layer.bias.detach().numpy()
                lin[f'{key}/grad_bias'] = layer.bias.grad.numpy()
            lin[f'{key}/x'] = x.detach().numpy()
            lin[f'{key}/gy'] = gy.numpy()
            lin[f'{key}/y'] = y.detach().numpy()
            lin[f'{key}/grad_input'] = x.grad.numpy()
            lin[f'{key}/grad_weight'] = layer.weight.grad.numpy()
    # LinearCRS (fewbit/functional/linear.py:27-66).  The module's constructor passes proj_dim in the
    # place of `bias`, so the functional form is called on explicit parameters.
    from fewbit.functional.linear import linear_crs
    for bias in (False, True):
        torch.manual_seed(42)
        weight = torch.randn(12, 24, requires_grad=True)
        b = torch.randn(12, requires_grad=True) if bias else None
        x = torch.randn(4, 10, 24, requires_grad=True)
        gy = torch.randn(4, 10, 12)
        torch.manual_seed(1234)  # fixes the sampled columns
        y = linear_crs(x, weight, b, 16)
        y.backward(gy)
        key = f'crs-{int(bias)}'
        lin[f'{key}/weight'] = weight.detach().numpy()
        if bias:
            lin[f'{key}/bias'] = b.detach().numpy()
            lin[f'{key}/grad_bias'] = b.grad.numpy()
        lin[f'{key}/x'] = x.detach().numpy()
        lin[f'{key}/gy'] = gy.numpy()
        lin[f'{key}/y'] = y.detach().numpy()
        lin[f'{key}/grad_input'] = x.grad.numpy()
        lin[f'{key}/grad_weight'] = weight.grad.numpy()
    np.savez_compressed(ROOT / 'tests/golden/reference_linear.npz', **lin)
    print('fixtures:', len(specs), 'op cases,', len(tables), 'table arrays,', len(lin), 'linear arrays')


if __name__ == '__main__':
    main()
