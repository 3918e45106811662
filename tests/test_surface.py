"""Host-side mirror of the reference's Python surface: names, signatures, argument handling,
error behaviour, module-tree utilities.  CPU tensors only (CUDA parity is in test_gpu_*.py)."""
import inspect

import pytest
import torch
import torch.nn.functional as F

import fewbit_b200 as fewbit
from fewbit_b200 import functional as FF
from fewbit_b200.functional import CONTINOUS, STEPWISE, store

PIECEWISE_ARGS = {'hardshrink': (0.5, ), 'hardsigmoid': (), 'hardtanh': (-1.0, 1.0),
                  'leaky_relu': (0.01, ), 'relu': (), 'relu6': (), 'softshrink': (0.5, ),
                  'threshold': (1.0, 3.0)}


def test_public_names_of_the_reference_exist():
    for name in ('CELU', 'ELU', 'GELU', 'Hardswish', 'LogSigmoid', 'Mish', 'SELU', 'Sigmoid', 'SiLU',
                 'Softplus', 'Softsign', 'Tanh', 'Tanhshrink', 'Hardshrink', 'Hardsigmoid',
                 'Hardtanh', 'LeakyReLU', 'ReLU', 'ReLU6', 'Softshrink', 'Threshold', 'Stepwise',
                 'LinearGRP', 'RandomizedLinear', 'map_module', 'functional'):
        assert hasattr(fewbit, name), name
    for name in STEPWISE + CONTINOUS + ('store', 'linear_grp', 'linear_randomized'):
        assert hasattr(FF, name), name
    assert fewbit.util.convert_linear and fewbit.util.map_module
    assert fewbit.modules.linear.RandomizedLinear is fewbit.LinearGRP


def test_functional_signatures():
    sig = inspect.signature(FF.softplus)
    assert list(sig.parameters) == ['input', 'beta', 'threshold', 'bits', 'borders', 'values']
    assert sig.parameters['beta'].default == 1.0 and sig.parameters['threshold'].default == 20.0
    assert sig.parameters['bits'].kind is inspect.Parameter.KEYWORD_ONLY
    assert list(inspect.signature(FF.gelu).parameters) == ['input', 'bits', 'borders', 'values']
    assert list(inspect.signature(FF.celu).parameters)[:2] == ['input', 'alpha']
    assert list(inspect.signature(FF.hardtanh).parameters) == ['input', 'min_val', 'max_val', 'bits']
    assert list(inspect.signature(FF.threshold).parameters) == ['input', 'threshold', 'value', 'bits']
    with pytest.raises(TypeError):  # `inplace` / `approximate` are dropped, as in the reference
        FF.gelu(torch.zeros(3), approximate='tanh')


@pytest.mark.parametrize('name', CONTINOUS)
@pytest.mark.parametrize('bits', [1, 2, 3, 4])
def test_continuous_on_cpu_tensors(name, bits):
    """Forward is the true function (reference bug C-2 fixed); backward is levels[code] * g."""
    torch.manual_seed(0)
    x = (torch.randn(257) * 2).requires_grad_()
    g = torch.randn(257)
    y = getattr(FF, name)(x, bits=bits)
    ref = (getattr(F, name, None) or getattr(torch, name))(x.detach())
    torch.testing.assert_close(y.detach(), ref, rtol=0, atol=0)
    y.backward(g)
    borders, levels = store.get(name, bits)
    want = levels[torch.searchsorted(borders[1:-1], x.detach())] * g
    torch.testing.assert_close(x.grad, want, rtol=0, atol=0)


def test_quantisation_arguments():
    x = torch.linspace(-3, 3, 50)
    borders, levels = store.get('gelu', 2)
    a = FF.gelu(x.clone().requires_grad_(), borders=borders, values=levels)
    torch.testing.assert_close(a, F.gelu(x))
    with pytest.raises(ValueError):
        FF.gelu(x, bits=2, borders=borders, values=levels)
    FF.gelu(x.clone(), bits=7)                       # 5..8-bit tables are shipped here (data/extended.npz)
    with pytest.raises(KeyError):                    # more than the 8 bits a code can hold
        FF.gelu(x, bits=9)
    # default is 3 bits (reference functional/activations.py:202)
    p = x.clone().requires_grad_()
    FF.gelu(p).sum().backward()
    b3, l3 = store.get('gelu', 3)
    torch.testing.assert_close(p.grad, l3[torch.searchsorted(b3[1:-1], x)])
    # extra parameters do not change the table
    q = x.clone().requires_grad_()
    y = FF.softplus(q, 2.0, 10.0, bits=3)
    torch.testing.assert_close(y, F.softplus(x, 2.0, 10.0))


@pytest.mark.parametrize('name', sorted(PIECEWISE_ARGS))
def test_piecewise_on_cpu_tensors(name):
    x = torch.linspace(-5, 5, 101)
    args = PIECEWISE_ARGS[name]
    torch.testing.assert_close(getattr(FF, name)(x.clone(), *args), getattr(F, name)(x, *args))
    torch.testing.assert_close(getattr(FF, name)(x.clone(), *args, bits=None),
                               getattr(F, name)(x, *args))  # `bits` accepted and ignored (C-5)


def test_custom_stepwise_tables():
    """`stepwise` / `Stepwise`: the reference declares both (fewbit/fewbit.cc:37,
    modules/activations.py:97-134) without a kernel; here the table is the function -- slopes
    `levels` between `borders`, zero at the anchor -- and its gradient is levels[code] * g."""
    borders, levels = store.get('gelu', 3)
    m = fewbit.Stepwise(borders, levels)          # strips the +-100 sentinels
    assert m.borders.numel() == 7 and m.levels.numel() == 8
    assert set(m.state_dict()) == {'borders', 'levels'}
    x = torch.linspace(-4, 4, 257, requires_grad=True)
    y = m(x)
    y.backward(torch.ones_like(y))
    codes = torch.searchsorted(m.borders, x.detach(), right=False)
    torch.testing.assert_close(x.grad, m.levels[codes], rtol=0, atol=0)
    # the 3-bit table integrates to GELU up to the quantisation of its derivative
    assert (y.detach() - F.gelu(x.detach())).abs().max().item() < 0.06 and abs(m(torch.zeros(1)).item()) == 0.0
    # continuity across every kink, and slope = level inside every piece
    eps = 1e-3
    for k, b in enumerate(m.borders.tolist()):
        lo, hi = m(torch.tensor([b - eps])), m(torch.tensor([b + eps]))
        assert abs((hi - lo).item() - eps * (m.levels[k] + m.levels[k + 1]).item()) < 1e-5
    # half a table, mirrored: GELU' is odd about (0, 1/2)
    half = fewbit.Stepwise(m.borders[4:], m.levels[4:], parity=False, shift=(0.0, 0.5))
    assert half._full_levels.numel() == 8 and half._full_borders.numel() == 7
    torch.testing.assert_close(half._full_levels[:4], 1.0 - m.levels[4:].flip(0))
    pts = torch.tensor([1.3, 2.0, 0.1])
    torch.testing.assert_close(half(-pts), half(pts) - pts)                # F(-t) = F(t) - t
    even = FF.expand_table(torch.tensor([2.0, 3.0]), torch.tensor([0.5, 0.25, 0.0]), True, (1.0, 0.0))
    assert even[0].tolist() == [-1.0, 0.0, 1.0, 2.0, 3.0] and even[1].tolist() == [0.0, 0.25, 0.5, 0.5, 0.25, 0.0] and even[2] == 1.0
    with pytest.raises(ValueError, match='above'):
        FF.expand_table(torch.tensor([-1.0, 2.0]), torch.tensor([0.5, 0.25, 0.0]), True)
    with pytest.raises(ValueError, match='256'):
        FF.expand_table(torch.arange(1, 200.0), torch.zeros(200), False)
    loaded = fewbit.Stepwise(torch.tensor([1.0]), torch.tensor([1.0, 2.0]), parity=True)
    loaded.load_state_dict({'borders': torch.tensor([2.0]), 'levels': torch.tensor([1.0, 3.0])})
    assert loaded._full_borders.tolist() == [-2.0, 0.0, 2.0] and loaded._full_levels.tolist() == [3.0, 1.0, 1.0, 3.0]
    with pytest.raises(ValueError):
        fewbit.Stepwise(torch.zeros(3), torch.zeros(8))
    with pytest.raises(ValueError):
        fewbit.Stepwise(torch.zeros(256), torch.zeros(257))
    with pytest.raises(ValueError):
        fewbit.Stepwise(torch.zeros(2, 2), torch.zeros(8))


def test_modules():
    assert repr(fewbit.GELU(bits=3)) == 'GELU(bits=3)'
    assert repr(fewbit.Softplus(2.0)) == 'Softplus(beta=2.0, threshold=20.0, bits=None)'
    assert repr(fewbit.Hardtanh(min_val=-2.0, max_val=2.0)) == 'Hardtanh(min_val=-2.0, max_val=2.0, bits=None)'
    m = fewbit.Threshold(1.0, 3.0)
    assert (m.threshold, m.value) == (1.0, 3.0)
    with pytest.raises(TypeError):
        fewbit.Threshold()                     # threshold and value are required, as in torch.nn
    with pytest.raises(TypeError):
        fewbit.GELU(approximate='tanh')
    assert list(inspect.signature(fewbit.LeakyReLU).parameters) == ['negative_slope', 'bits']
    # no parameters, no buffers: swapping activations leaves state_dict untouched
    for cls in (fewbit.GELU, fewbit.ReLU, fewbit.SiLU):
        assert len(cls().state_dict()) == 0
    x = torch.linspace(-5, 5, 101)
    torch.testing.assert_close(fewbit.GELU(bits=3)(x.clone()), F.gelu(x))
    torch.testing.assert_close(fewbit.Softplus(2.0, 10.0)(x.clone()), F.softplus(x, 2.0, 10.0))
    torch.testing.assert_close(fewbit.ReLU()(x.clone()), F.relu(x))        # reference bug C-5
    torch.testing.assert_close(fewbit.LeakyReLU(0.5)(x.clone()), F.leaky_relu(x, 0.5))
    torch.testing.assert_close(fewbit.Hardtanh(-2.0, 2.0)(x.clone()), F.hardtanh(x, -2.0, 2.0))


def test_cuda_tensors_never_fall_back(monkeypatch):
    """If the operator library is missing, a CUDA tensor is an error -- not a CPU detour."""
    monkeypatch.setattr(fewbit, 'NATIVE_ERROR', 'simulated: libfewbit.so not built')

    class FakeCuda:  # quacks like a CUDA tensor as far as the dispatcher looks
        device = torch.device('cuda', 0)
        dtype = torch.float32

    from fewbit_b200.functional import activations
    with pytest.raises(RuntimeError, match='not loaded'):
        activations._dispatch('relu', FakeCuda())
    with pytest.raises(RuntimeError, match='not loaded'):
        activations._dispatch('gelu', FakeCuda(), None, None)


# ------------------------------------------------------------------ map_module etc. ----

class Block(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.dense = torch.nn.Linear(8, 16)
        self.act = torch.nn.GELU()
        self.out = torch.nn.Linear(16, 8, bias=False)

    def forward(self, x):
        return self.out(self.act(self.dense(x)))


def make_net():
    return torch.nn.Sequential(Block(), torch.nn.Sequential(Block(), torch.nn.ReLU()))


def test_map_module_visits_post_order_with_paths():
    seen = []
    net = make_net()
    out = fewbit.map_module(net, lambda m, p: seen.append(p) or m)
    assert out is net
    assert seen == ['/0/dense', '/0/act', '/0/out', '/0', '/1/0/dense', '/1/0/act', '/1/0/out',
                    '/1/0', '/1/1', '/1', '/']   # reference util.py:176-187


def test_map_module_pattern_and_replacement():
    net = make_net()
    seen = []

    def swap(m, p):
        seen.append(p)
        return fewbit.GELU(bits=3) if isinstance(m, torch.nn.GELU) else m

    fewbit.map_module(net, swap, r'/1/')      # re.match: anchored at the start of the path
    assert seen == ['/1/0/dense', '/1/0/act', '/1/0/out', '/1/0', '/1/1']
    assert isinstance(net[1][0].act, fewbit.GELU) and isinstance(net[0].act, torch.nn.GELU)
    with pytest.raises(ValueError):
        fewbit.map_module(net, lambda m, p: None)
    root = fewbit.map_module(torch.nn.GELU(), lambda m, p: fewbit.GELU(bits=2))
    assert isinstance(root, fewbit.GELU)      # the root itself may be replaced


def test_convert_linear_shares_parameters():
    net = make_net()
    weights = [m.weight for m in net.modules() if isinstance(m, torch.nn.Linear)]
    fewbit.map_module(net, lambda m, p: fewbit.convert_linear(
        m, fewbit.RandomizedLinear, proj_dim_ratio=0.2, proj_dim_min=3))   # bench-linear.py:138-144
    layers = [m for m in net.modules() if isinstance(m, torch.nn.Linear)]
    assert all(isinstance(m, fewbit.LinearGRP) for m in layers) and len(layers) == 4
    assert all(a.data_ptr() == b.weight.data_ptr() for a, b in zip(weights, layers))
    assert layers[1].bias is None and layers[0].bias is not None
    assert layers[0].proj_dim_ratio == 0.2 and layers[0].proj_dim_min == 3
    act = torch.nn.GELU()
    assert fewbit.convert_linear(act, fewbit.RandomizedLinear, proj_dim=4) is act
    y = net(torch.randn(32, 8, requires_grad=True))
    y.sum().backward()
    assert all(m.weight.grad is not None for m in layers)


def test_memory_accounting_helpers():
    # reference util_test.py: saved-tensor accounting on a toy MLP
    x = torch.randn(64, 8, requires_grad=True)
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    y = net(x).sum()
    saved = fewbit.util.estimate_memory_usage(y, saved_only=True)
    assert saved >= (64 * 8 + 64 * 16 * 2) * 4
    with fewbit.util.memory_usage_hooks() as usage:
        net(x).sum().backward()
    assert usage.forward == usage.backward and usage.value > 0
