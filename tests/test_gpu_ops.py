"""torch.ops.fewbit / fewbit.functional / modules on CUDA tensors: the reference's own tests
restated (fewbit/functional/activations_test.py), in-place + autograd semantics, errors."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import fewbit_b200 as fewbit
import oracle
from fewbit_b200 import functional as FF
from fewbit_b200 import native
from fewbit_b200.functional import CONTINOUS, store

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def to_np(t: torch.Tensor) -> np.ndarray:
    """fp32 -> float32 array, bf16 -> uint16 bit patterns (what the oracle takes)."""
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16)
    return t.numpy()


def _loaded_native():
    assert fewbit.native_loaded(), fewbit.NATIVE_ERROR


# ---- reference tests, restated ----------------------------------------------------------

PIECEWISE = [('hardshrink', {}), ('hardshrink', {'lambd': 1.0}), ('hardsigmoid', {}), ('hardtanh', {}),
             ('hardtanh', {'min_val': -2.0, 'max_val': 2.0}), ('leaky_relu', {}),
             ('leaky_relu', {'negative_slope': 0.5}), ('relu', {}), ('relu6', {}), ('softshrink', {}),
             ('softshrink', {'lambd': 1.0})]


@pytest.mark.parametrize('name,kwargs', PIECEWISE)
def test_reference_stepwise_testcase(name, kwargs):
    # activations_test.py:17-32: value and gradient vs torch on linspace(-5, 5, 101), places=6
    _loaded_native()
    xs = torch.linspace(-5, 5, 101).to(DEV)
    gs = torch.ones_like(xs)
    ps = xs.clone().requires_grad_()
    ys = getattr(F, name)(ps, **kwargs)
    ys.backward(gs)
    qs = xs.clone().requires_grad_()
    zs = getattr(FF, name)(qs.clone(), **kwargs)
    zs.backward(gs)
    assert torch.linalg.norm(zs - ys).item() < 5e-7
    assert torch.linalg.norm(ps.grad - qs.grad).item() < 5e-7


def test_reference_threshold_testcase():
    xs = torch.linspace(-5, 5, 101).to(DEV)
    ps = xs.clone().requires_grad_()
    F.threshold(ps, 1.0, 3.0).backward(torch.ones_like(xs))
    qs = xs.clone().requires_grad_()
    zs = FF.threshold(qs.clone(), 1.0, 3.0)
    zs.backward(torch.ones_like(xs))
    assert torch.linalg.norm(zs - F.threshold(xs, 1.0, 3.0)).item() < 5e-7
    assert torch.linalg.norm(ps.grad - qs.grad).item() < 5e-7


@pytest.mark.parametrize('name', CONTINOUS)
def test_reference_continuous_testcase(name):
    # activations_test.py:79-104: forward L2 < 1e-6 on 101 points; here also the CUDA backward,
    # which the reference never compares to anything.
    xs = torch.linspace(-5, 5, 101).to(DEV)
    ref = (getattr(F, name, None) or getattr(torch, name))(xs)
    qs = xs.clone().requires_grad_()
    zs = getattr(FF, name)(qs.clone(), bits=3)
    assert torch.linalg.norm(zs - ref).item() < 1e-6
    zs.backward(torch.ones_like(zs))
    borders, levels = store.get(name, 3, DEV, torch.float32)
    want = levels[torch.searchsorted(borders[1:-1].contiguous(), xs)]
    assert torch.equal(qs.grad, want)


@pytest.mark.parametrize('name', CONTINOUS)
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_forward_matches_torch_cuda_kernels(name, dtype):
    """Secondary criterion of SURVEY App. A on N(0, 2^2) inputs: fp32 within max(4 ulp, 2.5e-7)
    of the ATen CUDA result; bf16 within one bf16 ulp."""
    torch.manual_seed(11)
    x = (torch.randn(1 << 20, device=DEV) * 2).to(dtype)
    args = {'celu': (1.5, ), 'elu': (0.7, ), 'softplus': (2.0, 10.0)}.get(name, ())
    # Yardstick: the ATen CUDA kernel in fp32 opmath, rounded once.  (In bf16 torch evaluates
    # the composites softsign / tanhshrink step by step in bf16, which is less accurate than
    # one fp32 evaluation + one rounding -- what ATen's fused kernels and ours do.)
    ref = (getattr(F, name, None) or getattr(torch, name))(x.float(), *args)
    y = getattr(FF, name)(x.clone(), *args, bits=2)
    if dtype == torch.float32:
        tol = 4 * (torch.nextafter(ref.abs(), ref.abs() + 1) - ref.abs()) + 2.5e-7
    else:
        ref = ref.to(dtype)
        tol = ref.float().abs() * 2.0 ** -7 + 1e-6
    err = (y.float() - ref.float()).abs()
    assert torch.all(err <= tol), f'{name}: max err {err.max().item():.3e}'


@pytest.mark.parametrize('name', CONTINOUS)
def test_bf16_results_are_the_fp32_exact_rounding(name):
    """The bf16 kernels use short MUFU-based evaluations; after rounding they must be
    indistinguishable from ATen's fp32 math rounded once: never more than one bf16 ulp away,
    and bit-identical for > 99 % of N(0, 2^2) inputs."""
    torch.manual_seed(17)
    x = (torch.randn(1 << 22, device=DEV) * 2).to(torch.bfloat16)
    args = {'celu': (1.5, ), 'elu': (0.7, ), 'softplus': (2.0, 10.0)}.get(name, ())
    ref = (getattr(F, name, None) or getattr(torch, name))(x.float(), *args).to(torch.bfloat16)
    y = getattr(FF, name)(x.clone(), *args, bits=3)
    same = (y.view(torch.int16) == ref.view(torch.int16)) | ((y == 0) & (ref == 0))
    assert same.float().mean().item() > 0.99, f'{name}: only {same.float().mean().item():.4f} identical'
    steps = (y.view(torch.int16).int() - ref.view(torch.int16).int()).abs()
    # beyond one ulp only where fp32 formulas themselves carry an absolute error (gelu's 1 + erf
    # cancels in the negative tail: ATen's own result is off by up to 1.5e-7 there)
    far = (steps > 1) & ((y.float() - ref.float()).abs() > 2.5e-7)
    assert not far.any(), f'{name}: {int(far.sum())} results more than one bf16 ulp away'


@pytest.mark.parametrize('name,bound', [('elu', 1.5), ('celu', 1.5), ('selu', 2.5), ('logsigmoid', 3.5),
                                        ('softplus', 3.5), ('sigmoid', 3.5), ('silu', 3.5), ('mish', 5.0)])
def test_fp32_own_expm1_and_log1p_against_float64(name, bound):
    """The fp32 ELU family, logsigmoid and softplus use own expm1 / log1p code instead of libdevice,
    sigmoid / silu / mish an own correctly rounded quotient instead of the IEEE division sequence
    (ops.cuh namespace accurate; tools/fit_fp32_math.py).  Against float64 on wide inputs --
    N(0, 2^2), N(0, 20^2), a dense sweep of [-100, 100], signed zeros, tiny and huge values, the
    exponent-range edges, NaN and infinities -- they stay within `bound` ulp (measured: 0.87, 0.87,
    1.71, 2.87, 2.78; ATen's own fp32 kernels: 1.13, 1.13, 2.19, 2.66, 2.58), never more than
    3 ulp from ATen's result, and agree with ATen on where NaN and inf come out."""
    torch.manual_seed(3)
    special = torch.tensor([0.0, -0.0, 1e-30, -1e-30, 1e-8, -1e-8, float('inf'), -float('inf'), float('nan'),
                            88.0, -88.0, -87.3, -88.7, -103.0, -104.0, -1e4, 1e4, 20.0, 20.000002], device=DEV)
    x = torch.cat([torch.randn(1 << 21, device=DEV) * 2, torch.randn(1 << 19, device=DEV) * 20,
                   torch.linspace(-100, 100, 1 << 19, device=DEV), special])
    y = getattr(FF, name)(x.clone(), bits=3)
    aten = getattr(F, name)(x)
    exact = getattr(F, name)(x.double())
    assert torch.equal(torch.isnan(y), torch.isnan(aten)) and torch.equal(torch.isinf(y), torch.isinf(aten))
    # results at the bottom of the exponent range (sigmoid / silu below x = -80) are compared absolutely
    tiny = torch.isfinite(exact) & (exact.abs() < 1e-30)
    assert (y.double() - exact).abs()[tiny].max().item() < 1e-30 if tiny.any() else True
    finite = torch.isfinite(exact) & ~tiny
    mag = exact.float().abs()
    ulp = (torch.nextafter(mag, torch.full_like(mag, float('inf'))) - mag).double()
    ours = ((y.double() - exact).abs() / ulp)[finite].max().item()
    apart = ((y.double() - aten.double()).abs() / ulp)[finite].max().item()
    assert ours <= bound, f'{name}: {ours:.2f} ulp from float64'
    assert apart <= (5.0 if name == 'mish' else 3.0), f'{name}: {apart:.2f} ulp from ATen'


@pytest.mark.parametrize('name', CONTINOUS)
@pytest.mark.parametrize('bits', [5, 6, 7, 8])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_shipped_5_to_8_bit_tables_through_the_kernels(name, bits, dtype):
    """`fewbit.GELU(bits=6)` and friends: the optimal 5..8-bit tables of data/extended.npz, cast like
    the built-in ones, give exactly levels[#{bounds < x}] * g -- including the two softsign tables
    whose borders crowd the bucket LUT (exact-search fallback) and the tables with a border on a
    jump of the derivative (selu at 0, hardswish at +-3)."""
    torch.manual_seed(bits)
    n = 300007                          # > one tile per warp of the grid for a few warps, plus a ragged tail
    x = (torch.randn(n, device=DEV) * 3).to(dtype)
    x[:5] = torch.tensor([0.0, -0.0, 3.0, -3.0, 100.0], device=DEV).to(dtype)
    g = torch.randn(n, device=DEV).to(dtype)
    leaf = x.clone().requires_grad_()
    getattr(FF, name)(leaf * 1, bits=bits).backward(g)      # the CUDA operators work in place
    borders, levels = store.get(name, bits, DEV, dtype)
    codes = torch.searchsorted(borders[1:-1].contiguous(), x)
    assert torch.equal(leaf.grad, levels[codes] * g)


# ---- operator semantics -------------------------------------------------------------------

def test_in_place_and_only_codes_are_saved():
    n = 1 << 22
    x = torch.randn(n, device=DEV)
    leaf = x.clone().requires_grad_()
    h = leaf * 1.0
    ptr = h.data_ptr()
    torch.cuda.synchronize()
    before = torch.cuda.memory_allocated()
    y = FF.gelu(h, bits=3)
    torch.cuda.synchronize()
    grown = torch.cuda.memory_allocated() - before
    assert y.data_ptr() == ptr                                  # Tensor(a!) -> Tensor(a!)
    assert torch.equal(h, y)                                    # the input now holds f(x)
    assert n * 3 // 8 <= grown <= n * 3 // 8 + (2 << 20)        # nothing but the packed codes
    assert torch.allclose(y, F.gelu(x), atol=1e-6)
    y.sum().backward()
    assert leaf.grad is not None and leaf.grad.shape == x.shape


def test_direct_operator_call_like_bench_roberta():
    # benchmark/bench-roberta.py:128-139: T.ops.fewbit.gelu(xs, bounds7, levels8), fp32 tables
    from test_oracle import BOUNDS, LEVELS
    bounds, levels = torch.from_numpy(BOUNDS).to(DEV), torch.from_numpy(LEVELS).to(DEV)
    for dtype in (torch.float32, torch.bfloat16):
        x = (torch.randn(4, 128, 3072, device=DEV) * 2).to(dtype)
        h = x.clone().requires_grad_()
        y = torch.ops.fewbit.gelu(h + 0, bounds, levels)        # tables are cast to x.dtype inside
        g = torch.randn_like(y)
        y.backward(g)
        codes = torch.searchsorted(bounds.to(dtype).float(), x.float().flatten()).view_as(x)
        want = (levels.to(dtype).float()[codes] * g.float()).to(dtype)
        assert torch.equal(h.grad, want)


def test_quantize_pair_matches_reference_cpu_ops(golden_ops):
    """torch.ops.fewbit.quantize / quantize_backward on CUDA == the reference on CPU."""
    for case in golden_ops[::5]:
        if case['bf16']:
            conv = lambda a: torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16).to(DEV)  # noqa: E731
        else:
            conv = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)  # noqa: E731
        x, g, bounds, levels = (conv(case[k]) for k in ('x', 'g', 'bounds', 'levels'))
        y, state = torch.ops.fewbit.quantize(x, bounds)
        gin = torch.ops.fewbit.quantize_backward(g, state, levels)
        assert y.data_ptr() != x.data_ptr()                     # out of place, as in the reference
        assert np.array_equal(state.cpu().numpy(), case['state']), case['key']
        raw = gin.cpu().view(torch.int16).numpy() if case['bf16'] else gin.cpu().numpy()
        assert np.array_equal(raw.view(np.uint8), case['gin'].view(np.uint8)), case['key']


def test_errors_are_loud():
    x = torch.randn(64, device=DEV)
    bounds, levels = (t.contiguous() for t in store.get('gelu', 3, DEV, torch.float32))
    bounds = bounds[1:-1].contiguous()
    with pytest.raises(RuntimeError, match='contiguous'):
        torch.ops.fewbit.relu(torch.randn(8, 8, device=DEV).t())
    with pytest.raises(RuntimeError, match='float32 or bfloat16'):
        torch.ops.fewbit.relu(x.double())
    with pytest.raises(RuntimeError, match='float32 or bfloat16'):
        torch.ops.fewbit.gelu(x.half(), bounds, levels)
    with pytest.raises(RuntimeError, match='lesser than size'):
        torch.ops.fewbit.gelu(x.clone(), bounds[:-1], levels)
    with pytest.raises(RuntimeError, match='256'):
        torch.ops.fewbit.gelu(x.clone(), torch.zeros(256, device=DEV), torch.zeros(257, device=DEV))
    with pytest.raises(RuntimeError):                           # in-place on a leaf that needs grad
        FF.gelu(torch.randn(8, device=DEV, requires_grad=True), bits=3)


def test_modules_and_piecewise_modules_on_cuda():
    x = torch.linspace(-7, 7, 1001, device=DEV)
    for cls, ref, args in ((fewbit.ReLU, F.relu, ()), (fewbit.ReLU6, F.relu6, ()),
                           (fewbit.LeakyReLU, F.leaky_relu, (0.2, )),
                           (fewbit.Hardtanh, F.hardtanh, (-2.0, 2.0)),
                           (fewbit.Hardsigmoid, F.hardsigmoid, ()),
                           (fewbit.Threshold, F.threshold, (1.0, 3.0))):
        h = x.clone().requires_grad_()
        y = cls(*args)(h + 0)                                   # reference bug C-5: modules work
        assert torch.allclose(y, ref(x, *args), atol=1e-6), cls.__name__
        y.sum().backward()
        p = x.clone().requires_grad_()
        ref(p, *args).sum().backward()
        same = x != 0 if cls is fewbit.LeakyReLU else torch.ones_like(x, dtype=torch.bool)
        assert torch.allclose(h.grad[same], p.grad[same], atol=1e-6), cls.__name__
    assert torch.equal(fewbit.ReLU6()(torch.tensor([5.0, 6.0, 7.0], device=DEV)),
                       torch.tensor([5.0, 6.0, 6.0], device=DEV))          # C-6: saturates at 6
    for bits in (1, 2, 3, 4):
        m = fewbit.SiLU(bits=bits)
        h = x.clone().requires_grad_()
        m(h + 0).sum().backward()
        borders, levels = store.get('silu', bits, DEV, torch.float32)
        assert torch.equal(h.grad, levels[torch.searchsorted(borders[1:-1].contiguous(), x)])


def test_runs_on_the_current_stream_and_shapes():
    side = torch.cuda.Stream()
    x = (torch.randn(8, 128, 768, device=DEV) * 2)
    with torch.cuda.stream(side):
        h = x.clone().requires_grad_()
        y = FF.mish(h * 1, bits=4)
        y.backward(torch.ones_like(y))
    side.synchronize()
    assert y.shape == x.shape and h.grad.shape == x.shape
    borders, levels = store.get('mish', 4, DEV, torch.float32)
    assert torch.equal(h.grad, levels[torch.searchsorted(borders[1:-1].contiguous(), x)])
    z = FF.relu(torch.empty(0, 3, device=DEV))                  # empty tensors are fine
    assert z.shape == (0, 3)


def test_oracle_agrees_through_the_operator_path():
    torch.manual_seed(5)
    x = (torch.randn(70001, device=DEV) * 2).to(torch.bfloat16)
    g = torch.randn(70001, device=DEV).to(torch.bfloat16)
    borders, levels = store.get('tanh', 4, DEV, torch.bfloat16)
    h = x.clone().requires_grad_()
    FF.tanh(h + 0, bits=4).backward(g)
    bits16 = lambda t: t.detach().cpu().contiguous().view(torch.int16).numpy().view(np.uint16)  # noqa: E731
    _, state = oracle.stepwise_forward('tanh', bits16(x), bits16(borders[1:-1]), 4)
    want = oracle.stepwise_backward(state, bits16(g), bits16(levels), 4)
    assert np.array_equal(bits16(h.grad), want)


def test_randomized_linear_on_cuda():
    torch.manual_seed(42)
    layer = fewbit.RandomizedLinear(256, 128, proj_dim=64).to(DEV)
    ref = torch.nn.Linear(256, 128).to(DEV)
    ref.load_state_dict(layer.state_dict())
    x = torch.randn(512, 256, device=DEV, requires_grad=True)
    acc = torch.zeros_like(layer.weight)
    for _ in range(1024):
        layer.zero_grad()
        x.grad = None
        y = layer(x)
        y.backward(torch.ones_like(y))
        acc += layer.weight.grad
    gi = x.grad.clone()
    x.grad = None
    z = ref(x)
    z.backward(torch.ones_like(z))
    assert (torch.linalg.norm(y - z) / torch.linalg.norm(z)).item() < 1e-5
    assert (torch.linalg.norm(gi - x.grad) / torch.linalg.norm(x.grad)).item() < 1e-5
    err = torch.linalg.norm(acc / 1024 - ref.weight.grad) / torch.linalg.norm(ref.weight.grad)
    assert err.item() < 0.1                                      # linear_test.py:88-89
    # fresh randomness per call on CUDA (reference bug C-9: same S every call)
    layer.zero_grad(); layer(x).sum().backward(); a = layer.weight.grad.clone()
    layer.zero_grad(); layer(x).sum().backward()
    assert not torch.equal(a, layer.weight.grad)


def test_kernels_are_cuda_graph_capturable():
    """No hidden synchronisation or allocation in the C ABI: a forward + backward pair can be
    captured into a CUDA graph and replayed on new data (DESIGN: streams and graphs)."""
    from fewbit_b200 import native
    n = 1 << 20
    borders, levels = store.get('gelu', 3, DEV, torch.bfloat16)
    bounds, levels = borders[1:-1].contiguous(), levels.contiguous()
    x = torch.empty(n, dtype=torch.bfloat16, device=DEV)
    g = torch.empty(n, dtype=torch.bfloat16, device=DEV)
    y, gin = torch.empty_like(x), torch.empty_like(g)
    state, mask = native.new_state(x, 3), native.new_state(x, 1)
    x.normal_(0, 2); g.normal_()
    native.stepwise_forward('gelu', x, y, state, 3, bounds)       # warm up outside the capture
    native.piecewise_forward('relu', x, y, mask)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        native.stepwise_forward('gelu', x, y, state, 3, bounds)
        native.stepwise_backward(state, g, gin, 3, levels)
        native.piecewise_forward('relu', x, y, mask)
    for seed in (1, 2):
        torch.manual_seed(seed)
        x.normal_(0, 2); g.normal_()
        graph.replay()
        torch.cuda.synchronize()
        codes = torch.searchsorted(bounds.float(), x.float())
        assert torch.equal(gin, (levels.float()[codes] * g.float()).to(torch.bfloat16))
        assert torch.equal(y, torch.relu(x))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_second_device_is_respected():
    """Device guard + current stream of the tensor's device (reference bug C-8: it launches on
    the current device's legacy stream whatever the tensor's device is)."""
    dev1 = 'cuda:1'
    x = (torch.randn(3, 1000, 64, device=dev1) * 2)
    h = x.clone().requires_grad_()
    y = FF.gelu(h * 1.0, bits=3)                   # current device stays cuda:0
    y.sum().backward()
    assert y.device == h.grad.device == torch.device(dev1)
    assert torch.allclose(y, F.gelu(x), atol=1e-6)
    borders, levels = store.get('gelu', 3, dev1, torch.float32)
    assert torch.equal(h.grad, levels[torch.searchsorted(borders[1:-1].contiguous(), x)])
    layer = fewbit.RandomizedLinear(64, 32, proj_dim_ratio=0.25).to(dev1)
    out = layer(x.requires_grad_())
    out.sum().backward()
    assert layer.weight.grad.device == torch.device(dev1) and torch.isfinite(layer.weight.grad).all()


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('bits', [1, 2, 3, 5, 8])
def test_custom_stepwise_operator_against_the_oracle(dtype, bits):
    """torch.ops.fewbit.stepwise (reference schema, fewbit/fewbit.cc:37; kernel: csrc/custom.cu):
    packed codes bit-exact, values one fused multiply-add per element, gradient levels[code] * g
    exact; ragged length, in place, through autograd."""
    torch.manual_seed(bits)
    nlevels = 1 << bits
    bounds = torch.sort(torch.randn(nlevels - 1) * 1.5).values.to(dtype)
    bounds = torch.unique_consecutive(bounds)
    while bounds.numel() < nlevels - 1:                          # bf16 rounding may merge borders
        bounds = torch.unique_consecutive(torch.sort(torch.cat([bounds, torch.randn(4).to(dtype) * 3])).values)[:nlevels - 1]
    levels = torch.randn(nlevels).to(dtype)
    n = 256 * 1024 + 77
    x = (torch.randn(n, device=DEV) * 2).to(dtype)
    g = torch.randn(n, device=DEV).to(dtype)
    leaf = x.clone().requires_grad_()
    y = torch.ops.fewbit.stepwise(leaf * 1.0, bounds.to(DEV), levels.to(DEV), None, [1, 0])
    y.backward(g)
    y_ref, state_ref = oracle.stepwise_custom_forward(to_np(x), to_np(bounds), to_np(levels), 1.0, bits)
    gin_ref = oracle.stepwise_backward(state_ref, to_np(g), to_np(levels), bits)
    state = native.new_state(x, bits)
    out = torch.empty_like(x)
    native.stepwise_custom_forward(x, out, state, bits, bounds.to(DEV), levels.to(DEV), 1.0)
    torch.cuda.synchronize()
    assert np.array_equal(state.cpu().numpy(), state_ref)
    assert np.array_equal(to_np(leaf.grad), gin_ref)
    assert torch.equal(out, y.detach())
    if dtype == torch.float32:
        err = np.abs(to_np(y.detach()).astype(np.float64) - y_ref.astype(np.float64))
        assert np.all(err <= np.spacing(np.abs(y_ref)) + 1e-7 * np.abs(y_ref).max()), err.max()
    else:
        steps = np.abs(to_np(y.detach()).astype(np.int32) - y_ref.astype(np.int32))
        assert (steps <= 1).all() and (steps == 0).mean() > 0.999


def test_stepwise_module_mirrors_half_a_table_on_cuda():
    """Stepwise(parity=...) (Python expansion, real-valued shift) and the operator's own integer
    `shift` agree; a GELU table rebuilt from its upper half reproduces fewbit.GELU's gradient."""
    borders, levels = store.get('gelu', 3, DEV, torch.float32)
    inner = borders[1:-1]
    x = torch.randn(1 << 16, device=DEV) * 2
    upper = fewbit.Stepwise(inner[4:], levels[4:], parity=False, shift=(0.0, 0.5)).to(DEV)
    a = x.clone().requires_grad_()
    upper(a * 1.0).sum().backward()
    codes = torch.searchsorted(upper._full_borders, x, right=False)
    assert torch.equal(a.grad, upper._full_levels[codes])
    even_py = fewbit.Stepwise(torch.tensor([1.0, 2.5]), torch.tensor([0.5, 0.25, 0.0]), parity=True, shift=(0.0, 0.0)).to(DEV)
    even_op = torch.ops.fewbit.stepwise(x.clone(), torch.tensor([1.0, 2.5], device=DEV), torch.tensor([0.5, 0.25, 0.0], device=DEV),
                                        True, [0, 0])
    assert torch.equal(even_py(x.clone()), even_op)
    y = even_py(x.clone())
    assert torch.equal(y[x.abs() < 1.0], 0.5 * x[x.abs() < 1.0])
