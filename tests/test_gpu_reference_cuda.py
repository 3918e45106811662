"""Side by side with the reference's own CUDA kernels on the same GPU: the unmodified
fewbit/cuda/codec.cu + activation.cc, compiled for sm_100 into oracle/_ref/ (oracle/Makefile)
and run in a process of their own (oracle/ref_runner.py cuda).

Codes must agree exactly (the reference stores them in bits+1 bits, App. C-1, so they are
recovered through its backward with levels = 0, 1, 2, ...); gradients must agree bit for bit;
forward values agree to 4 ulp + 2.5e-7 (the reference evaluates e.g. gelu as x * normcdf(x),
we follow ATen's expressions).  Documented deviations: relu6 saturation (C-6) and nothing else.
"""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle
from fewbit_b200 import native
from fewbit_b200.functional import CONTINOUS, store

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / 'oracle' / '_ref' / 'libfewbit_ref_cuda.so'
DEV = 'cuda:0'
N = 8192

PIECEWISE = {'hardshrink': [0.5], 'hardsigmoid': [], 'hardtanh': [-1.0, 1.0], 'leaky_relu': [0.01],
             'relu': [], 'relu6': [], 'softshrink': [0.5], 'threshold': [1.0, 3.0]}
PARAMS = {'celu': [1.5], 'elu': [0.7], 'softplus': [2.0, 10.0]}


@pytest.fixture(scope='module')
def reference(tmp_path_factory):
    if not LIB.exists():
        pytest.skip('oracle/_ref/libfewbit_ref_cuda.so not built (needs /root/reference at build time)')
    rng = np.random.default_rng(5)
    x = np.concatenate([np.linspace(-5, 5, 101), rng.standard_normal(N - 101) * 2]).astype(np.float32)
    g = rng.standard_normal(N).astype(np.float32)
    cases = {}
    for name in CONTINOUS:
        for bits in (1, 2, 3, 4):
            borders, levels = store.get(name, bits)
            key = f'{name}-{bits}'
            cases[f'{key}/name'] = np.array(name)
            cases[f'{key}/params'] = np.array(PARAMS.get(name, []), np.float64)
            cases[f'{key}/x'], cases[f'{key}/g'] = x, g
            cases[f'{key}/bounds'] = borders[1:-1].numpy()
            cases[f'{key}/levels'] = levels.numpy()
    for name, params in PIECEWISE.items():
        cases[f'{name}/name'] = np.array(name)
        cases[f'{name}/params'] = np.array(params, np.float64)
        cases[f'{name}/x'], cases[f'{name}/g'] = x, g
    tmp = tmp_path_factory.mktemp('refcuda')
    np.savez(tmp / 'in.npz', **cases)
    proc = subprocess.run([sys.executable, str(ROOT / 'oracle' / 'ref_runner.py'), 'cuda',
                           str(tmp / 'in.npz'), str(tmp / 'out.npz')], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-2000:]
    with np.load(tmp / 'out.npz') as npz:
        return x, g, {k: npz[k] for k in npz.keys()}


def close(y, ref):
    return np.all(np.abs(y.astype(np.float64) - ref) <= 4 * np.spacing(np.abs(ref)) + 2.5e-7)


def test_reference_cuda_agrees_with_its_own_cpu_codec(reference):
    """Trust check of the GPU oracle itself: its codes equal the CPU oracle's."""
    x, _, out = reference
    borders, _ = store.get('gelu', 3)
    np.testing.assert_array_equal(out['gelu-3/codes'], oracle.bucketize(x, borders[1:-1].numpy()))


@pytest.mark.parametrize('name', CONTINOUS)
def test_continuous_against_reference_cuda(reference, name):
    x, g, out = reference
    xd, gd = torch.from_numpy(x).to(DEV), torch.from_numpy(g).to(DEV)
    p = PARAMS.get(name, []) + [1.0, 20.0]
    for bits in (1, 2, 3, 4):
        borders, levels = store.get(name, bits, DEV, torch.float32)
        bounds = borders[1:-1].contiguous()
        y, gin = torch.empty_like(xd), torch.empty_like(gd)
        state = native.new_state(xd, bits)
        native.stepwise_forward(name, xd, y, state, bits, bounds, p[0], p[1])
        native.stepwise_backward(state, gd, gin, bits, levels)
        key = f'{name}-{bits}'
        codes = oracle.inflate(state.cpu().numpy(), N, bits)
        np.testing.assert_array_equal(codes, out[f'{key}/codes'], err_msg=key)
        np.testing.assert_array_equal(gin.cpu().numpy(), out[f'{key}/gin'], err_msg=key)
        assert close(y.cpu().numpy(), out[f'{key}/y']), key
        if bits == 3:   # the reference's own test criterion, against the reference itself
            # gelu: the reference kernel (x * normcdff(x), accurate in the negative tail) is
            # itself 1.04e-6 away from F.gelu (0.5x(1 + erff), which cancels there) on these
            # points with CUDA 12.9 -- measured on B200 -- so its own 1e-6 bound cannot hold for
            # both.  We are bit-exact with F.gelu (test_gpu_ops.py); allow that distance here.
            bound = 1.5e-6 if name == 'gelu' else 1e-6
            assert np.linalg.norm(y.cpu().numpy()[:101] - out[f'{key}/y'][:101]) < bound, key


@pytest.mark.parametrize('name', sorted(PIECEWISE))
def test_piecewise_against_reference_cuda(reference, name):
    x, g, out = reference
    xd, gd = torch.from_numpy(x).to(DEV), torch.from_numpy(g).to(DEV)
    p = PIECEWISE[name] + [0.0, 0.0]
    y, gin = torch.empty_like(xd), torch.empty_like(gd)
    state = native.new_state(xd, 1)
    native.piecewise_forward(name, xd, y, state, p[0], p[1])
    native.piecewise_backward(name, state, gd, gin, p[0])
    np.testing.assert_array_equal(gin.cpu().numpy(), out[f'{name}/gin'], err_msg=name)
    want = out[f'{name}/y'].copy()
    if name == 'relu6':
        want[x >= 6.0] = 6.0      # the reference writes 1.0 there
    assert close(y.cpu().numpy(), want), name
