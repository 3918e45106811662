"""RandomizedLinear (LinearGRP): bit-exact against the reference on CPU for the same seed, and
the reference's own statistical test (fewbit/modules/linear_test.py:57-92)."""
import pytest
import torch

import fewbit_b200 as fewbit
from fewbit_b200.functional.linear import calc_proj_dim


@pytest.mark.parametrize('kind', ['gaussian', 'rademacher', 'dct'])
@pytest.mark.parametrize('bias', [False, True])
def test_bit_exact_with_reference_on_cpu(golden_linear, kind, bias):
    """Same seed -> same sketch -> same y / grad_input / grad_weight / grad_bias as
    LinearGRPFunc of the reference (tests/golden/make_golden.py ran it)."""
    key = f'{kind}-{int(bias)}'
    g = {k.split('/', 1)[1]: torch.from_numpy(v) for k, v in golden_linear.items()
         if k.startswith(key + '/')}
    layer = fewbit.LinearGRP(24, 12, bias, proj_dim=16, matmul=kind)
    with torch.no_grad():
        layer.weight.copy_(g['weight'])
        if bias:
            layer.bias.copy_(g['bias'])
    x = g['x'].clone().requires_grad_()
    torch.manual_seed(1234)
    y = layer(x)
    y.backward(g['gy'])
    torch.testing.assert_close(y.detach(), g['y'], rtol=0, atol=0)
    torch.testing.assert_close(x.grad, g['grad_input'], rtol=0, atol=0)
    if kind == 'dct':
        # same sampled rows, same scale (rows * tokens, the reference's); the cosine transform is
        # this package's own torch.fft composition (fewbit_b200/fft.py): equal to rounding
        torch.testing.assert_close(layer.weight.grad, g['grad_weight'], rtol=2e-5, atol=2e-5 * g['grad_weight'].abs().max().item())
    else:
        torch.testing.assert_close(layer.weight.grad, g['grad_weight'], rtol=0, atol=0)
    if bias:
        torch.testing.assert_close(layer.bias.grad, g['grad_bias'], rtol=0, atol=0)


def test_proj_dim_rules():
    # reference functional/linear.py:17-24, 72-81: proj_dim wins, then int(ratio*N), then N;
    # falsy clamps are ignored.
    assert calc_proj_dim(16384, 0.2, None, None, 3) == 3276
    assert calc_proj_dim(100, 0.2, 64, None, None) == 64
    assert calc_proj_dim(100, None, None, None, None) == 100
    assert calc_proj_dim(10, 0.2, None, None, 3) == 3
    assert calc_proj_dim(1000, 0.5, None, 128, None) == 128
    assert calc_proj_dim(1000, 0.5, None, 0, 0) == 500
    layer = fewbit.LinearGRP(4, 4)
    with pytest.raises(ValueError):
        layer(torch.randn(8, 4))                      # neither proj_dim nor ratio
    with pytest.raises(ValueError):
        fewbit.LinearGRP(4, 4, proj_dim=2, proj_dim_min=-1)(torch.randn(8, 4))
    with pytest.raises(ValueError):
        fewbit.LinearGRP(4, 4, proj_dim=2, proj_dim_min=5, proj_dim_max=3)(torch.randn(8, 4))
    with pytest.raises(ValueError):
        fewbit.LinearGRP(4, 4, proj_dim=2, matmul='hadamard')(torch.randn(8, 4))


def test_forward_is_exact():
    # linear_test.py:57-69
    torch.manual_seed(42)
    for bias in (False, True):
        layer = fewbit.LinearGRP(8, 4, bias, proj_dim=64)
        ref = torch.nn.Linear(8, 4, bias)
        ref.load_state_dict(layer.state_dict())
        x = torch.randn(128, 8)
        with torch.no_grad():
            rel = torch.linalg.norm(ref(x) - layer(x)) / torch.linalg.norm(ref(x))
        assert rel.item() < 1e-6


@pytest.mark.parametrize('kind', ['gaussian', 'rademacher'])
def test_weight_gradient_is_unbiased(kind):
    # linear_test.py:71-92 with fewer repeats (512 instead of 2048) and a matching bound:
    # grad_input / grad_bias exact, grad_weight within 0.2 relative error of the mean.
    torch.manual_seed(42)
    layer = fewbit.LinearGRP(256, 128, True, proj_dim=64, matmul=kind)
    ref = torch.nn.Linear(256, 128, True)
    ref.load_state_dict(layer.state_dict())
    x = torch.randn(512, 256, requires_grad=True)
    acc = torch.zeros_like(layer.weight)
    repeats = 512
    for _ in range(repeats):
        layer.zero_grad()
        x.grad = None
        y = layer(x)
        y.backward(torch.ones_like(y))
        acc += layer.weight.grad
    gi, gb = x.grad.clone(), layer.bias.grad.clone()
    x.grad = None
    z = ref(x)
    z.backward(torch.ones_like(z))
    assert (torch.linalg.norm(gi - x.grad) / torch.linalg.norm(x.grad)).item() < 1e-6
    assert (torch.linalg.norm(gb - ref.bias.grad) / torch.linalg.norm(ref.bias.grad)).item() < 1e-6
    err = torch.linalg.norm(acc / repeats - ref.weight.grad) / torch.linalg.norm(ref.weight.grad)
    assert err.item() < 0.2


@pytest.mark.parametrize('kind', ['dct', 'dft'])
def test_transform_sketches_estimate_the_weight_gradient(kind):
    """Rows of the orthonormal cosine / Fourier transform of the token axis, sampled with
    replacement (reference fewbit/functional/linear.py:113-132, 178-205).  Up to the reference's
    scale -- rows * tokens where tokens / rows would be unbiased, i.e. rows^2 too large, kept for
    parity -- the mean over many draws approaches G^T X.  The reference itself cannot run 'dft'
    (its backward multiplies a real by a complex matrix and raises); here the estimate is the real
    part of the product of the two spectra."""
    torch.manual_seed(42)
    rows = 32
    layer = fewbit.LinearGRP(24, 12, True, proj_dim=rows, matmul=kind)
    x = torch.randn(64, 24, requires_grad=True)
    gy = torch.randn(64, 12)
    exact = gy.T @ x.detach()
    acc = torch.zeros_like(layer.weight)
    repeats = 1024
    for _ in range(repeats):
        layer.zero_grad()
        layer(x).backward(gy)
        assert layer.weight.grad.dtype == torch.float32
        acc += layer.weight.grad
    err = torch.linalg.norm(acc / repeats / rows ** 2 - exact) / torch.linalg.norm(exact)
    assert err.item() < 0.15, err.item()
    torch.testing.assert_close(layer.bias.grad, gy.sum(0))


def test_no_tokens_to_sketch_and_frozen_weights():
    """int(ratio * N) == 0 (a single token at ratio 0.2) is an empty sketch and a zero weight
    gradient, as in the reference; a layer whose weight needs no gradient takes no sketch at all."""
    layer = fewbit.LinearGRP(8, 4, True, proj_dim_ratio=0.2)
    x = torch.randn(1, 8, requires_grad=True)
    y = layer(x)
    torch.testing.assert_close(y, torch.nn.functional.linear(x, layer.weight, layer.bias))
    y.sum().backward()
    assert torch.count_nonzero(layer.weight.grad) == 0 and layer.weight.grad.shape == (4, 8)
    torch.testing.assert_close(x.grad, layer.weight.sum(0, keepdim=True))
    frozen = fewbit.LinearGRP(8, 4, True, proj_dim=2)
    frozen.weight.requires_grad_(False)
    state = torch.get_rng_state()
    z = frozen(torch.randn(16, 8, requires_grad=True) * 1.0)
    z.sum().backward()
    assert frozen.weight.grad is None and frozen.bias.grad is not None
    with torch.no_grad():
        before = torch.get_rng_state()
        frozen(torch.randn(16, 8))
    del state, before


def test_autocast_matches_linear():
    """Under autocast the layer computes what F.linear computes (bf16 operands), although its
    forward product is an out= call that autocast does not intercept."""
    layer = fewbit.LinearGRP(16, 8, True, proj_dim=8)
    x = torch.randn(4, 6, 16, requires_grad=True)
    with torch.autocast('cpu', dtype=torch.bfloat16):
        y = layer(x)
        expect = torch.nn.functional.linear(x, layer.weight, layer.bias)
    assert y.dtype == expect.dtype == torch.bfloat16
    torch.testing.assert_close(y, expect)
    y.float().sum().backward()
    assert layer.weight.grad.dtype == torch.float32 and x.grad.dtype == torch.float32


def test_user_generator_is_honoured_and_replayed():
    layer = fewbit.LinearGRP(16, 8, proj_dim=8, generator=torch.Generator().manual_seed(7))
    x = torch.randn(32, 16)
    grads = []
    for _ in range(2):
        layer.generator.manual_seed(7)
        layer.zero_grad()
        layer(x).sum().backward()
        grads.append(layer.weight.grad.clone())
    torch.testing.assert_close(grads[0], grads[1], rtol=0, atol=0)
    layer.zero_grad()
    layer(x).sum().backward()                        # generator advanced: a different sketch
    assert not torch.equal(layer.weight.grad, grads[0])


def test_output_can_be_modified_in_place():
    """3-bit GELU (in place) directly after a RandomizedLinear on a 3-D input: the linear's
    output must own its storage (regression: 'view ... modified inplace' from autograd)."""
    layer = fewbit.LinearGRP(16, 32, proj_dim_ratio=0.5)
    x = torch.randn(4, 8, 16, requires_grad=True)
    y = layer(x)
    assert not y._is_view() and y.shape == (4, 8, 32)
    torch.testing.assert_close(y, torch.nn.functional.linear(x, layer.weight, layer.bias))
    z = y.relu_()                      # any in-place op on the output
    z.sum().backward()
    assert layer.weight.grad is not None and x.grad is not None


# ---- LinearCRS (column-row sampling; reference fewbit/functional/linear.py:27-66) -----------

@pytest.mark.parametrize('bias', [False, True])
def test_crs_matches_the_reference_on_cpu(golden_linear, bias):
    """Same seed -> same sampled columns -> the reference's y / grad_input / grad_weight / grad_bias
    (fixture: tests/golden/make_golden.py ran the reference's linear_crs).  The forward and
    grad_input are plain products and must agree to the bit; grad_weight is one product computed
    in a different association (reshape + matmul here, einsum there): 1e-6 relative."""
    key = f'crs-{int(bias)}'
    g = {k.split('/', 1)[1]: torch.from_numpy(v) for k, v in golden_linear.items() if k.startswith(key + '/')}
    layer = fewbit.modules.LinearCRS(24, 12, bias, proj_dim=16)
    assert (layer.bias is not None) == bias            # the reference ignores `bias` here (its slip)
    with torch.no_grad():
        layer.weight.copy_(g['weight'])
        if bias:
            layer.bias.copy_(g['bias'])
    x = g['x'].clone().requires_grad_()
    torch.manual_seed(1234)
    y = layer(x)
    y.backward(g['gy'])
    torch.testing.assert_close(y.detach(), g['y'], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(x.grad, g['grad_input'], rtol=0, atol=0)
    torch.testing.assert_close(layer.weight.grad, g['grad_weight'], rtol=1e-5, atol=1e-5)
    assert torch.equal(layer.weight.grad == 0, g['grad_weight'] == 0)      # same sampled columns
    if bias:
        torch.testing.assert_close(layer.bias.grad, g['grad_bias'], rtol=1e-6, atol=1e-6)
    assert 'proj_dim=16' in repr(layer)


def test_crs_weight_gradient_is_unbiased():
    """The reference's statistical test (fewbit/modules/linear_test.py:57-92) for LinearCRS:
    averaged over many draws the sampled gradient approaches the exact one."""
    torch.manual_seed(42)
    layer = fewbit.modules.LinearCRS(8, 4, True, proj_dim=64)
    exact = torch.nn.Linear(8, 4)
    exact.load_state_dict(layer.state_dict())
    x = torch.randn(128, 8)
    exact(x).backward(torch.ones(128, 4))
    total = torch.zeros_like(layer.weight)
    repeat = 512
    for _ in range(repeat):
        layer.zero_grad()
        layer(x).backward(torch.ones(128, 4))
        total += layer.weight.grad
    rel = torch.linalg.norm(total / repeat - exact.weight.grad) / torch.linalg.norm(exact.weight.grad)
    assert rel.item() < 0.1
    assert hasattr(fewbit.functional, 'linear_crs') and fewbit.LinearCRS is fewbit.modules.LinearCRS
