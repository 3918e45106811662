"""Parity of the CUDA path against the CPU oracle, through the C ABI (fewbit_b200/native.py).

Bars (BASELINE.json north_star): packed codes and masks bit-exact; gradients bit-exact
(the requirement is <= 1 ulp in fp32 -- a single IEEE multiply reproduces exactly -- and exact
against the fp32-product/one-rounding oracle in bf16); forward values within the tolerance
written next to each check.
"""
import numpy as np
import pytest
import torch

import oracle
from fewbit_b200 import native
from fewbit_b200.functional import CONTINOUS, make_table, store

pytestmark = pytest.mark.gpu

DEV = 'cuda:0'
SIZES = [1, 7, 8, 9, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2048 + 13, 10007, 70001]
DTYPES = {'f32': torch.float32, 'bf16': torch.bfloat16}


def to_np(t: torch.Tensor) -> np.ndarray:
    """fp32 -> float32 array, bf16 -> uint16 bit patterns (what the oracle takes)."""
    t = t.detach().cpu().contiguous()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16)
    return t.numpy()


def as_f32(a: np.ndarray) -> np.ndarray:
    return oracle.bf16_bits_to_f32(a) if a.dtype == np.uint16 else a


def table(name, bits, dtype):
    if bits <= 4:
        borders, levels = store.get(name, bits, DEV, dtype)
    else:
        borders, levels = (t.to(DEV, dtype) for t in make_table(name, bits))
    return borders[1:-1].contiguous(), levels.contiguous()


def inputs(n, bounds, dtype, seed, specials=True):
    gen = torch.Generator(device='cpu').manual_seed(seed)
    x = torch.randn(n, generator=gen) * 2
    g = torch.randn(n, generator=gen)
    if specials:  # exact hits on borders, signed zeros, infinities
        planted = list(bounds.float().cpu()[:5]) + [0.0, -0.0, float('inf'), float('-inf')]
        for k, v in enumerate(planted):
            if k < n:
                x[(k * 7919 + 3) % n] = v
    return x.to(dtype).to(DEV), g.to(dtype).to(DEV)


def check_forward_values(y, y_ref, x, dtype, what):
    """fp32: |dy| <= 4 ulp(y) + 2.5e-7 (the absolute term covers 1+erf / x-tanh(x)
    cancellation, where fp32 formulas -- ATen's included -- lose relative accuracy).
    bf16: one bf16 ulp (2^-7 relative) + 1e-6."""
    y, y_ref, x = as_f32(y), as_f32(y_ref), as_f32(x)
    ok = np.isfinite(x) & np.isfinite(y_ref)
    err = np.abs(y[ok].astype(np.float64) - y_ref[ok])
    if dtype == torch.float32:
        bound = 4 * np.spacing(np.abs(y_ref[ok])).astype(np.float64) + 2.5e-7
    else:
        bound = np.abs(y_ref[ok]).astype(np.float64) * 2.0 ** -7 + 1e-6
    assert np.all(err <= bound), f'{what}: forward value off by {err.max():.3e}'


@pytest.mark.parametrize('tag', DTYPES)
@pytest.mark.parametrize('name', CONTINOUS)
def test_continuous_vs_oracle(name, tag):
    dtype = DTYPES[tag]
    p0, p1 = {'celu': (1.5, 0.0), 'elu': (0.7, 0.0), 'softplus': (2.0, 10.0)}.get(name, (1.0, 20.0))
    for bits in range(1, 9):
        bounds, levels = table(name, bits, dtype)
        for n in (SIZES if bits in (1, 3, 8) else SIZES[::3]):
            x, g = inputs(n, bounds, dtype, seed=1000 * bits + n)
            y = torch.full_like(x, 7.0)
            state = native.new_state(x, bits)
            state.fill_(0xAA)                      # stale bytes must all be overwritten
            gin = torch.empty_like(g)
            native.stepwise_forward(name, x, y, state, bits, bounds, p0, p1)
            native.stepwise_backward(state, g, gin, bits, levels)
            torch.cuda.synchronize()
            y_ref, state_ref = oracle.stepwise_forward(name, to_np(x), to_np(bounds), bits, p0, p1)
            gin_ref = oracle.stepwise_backward(state_ref, to_np(g), to_np(levels), bits)
            what = f'{name}/{tag}/bits={bits}/n={n}'
            assert state.numel() == (n * bits + 7) // 8
            assert np.array_equal(state.cpu().numpy(), state_ref), f'{what}: packed codes differ'
            assert np.array_equal(to_np(gin), gin_ref), f'{what}: gradient differs'
            check_forward_values(to_np(y), y_ref, to_np(x), dtype, what)


PIECEWISE_CASES = [('hardshrink', 0.5, 0.0), ('hardshrink', 1.0, 0.0), ('hardsigmoid', 0.0, 0.0),
                   ('hardtanh', -1.0, 1.0), ('hardtanh', -2.0, 2.0), ('leaky_relu', 0.01, 0.0),
                   ('leaky_relu', 0.5, 0.0), ('relu', 0.0, 0.0), ('relu6', 0.0, 0.0),
                   ('softshrink', 0.5, 0.0), ('softshrink', 1.0, 0.0), ('threshold', 1.0, 3.0)]


@pytest.mark.parametrize('tag', DTYPES)
@pytest.mark.parametrize('name,p0,p1', PIECEWISE_CASES)
def test_piecewise_vs_oracle(name, p0, p1, tag):
    dtype = DTYPES[tag]
    edges = torch.tensor([-6.0, -3.0, -2.0, -1.0, -0.5, 0.5, 1.0, 2.0, 3.0, 6.0])
    for n in SIZES:
        x, g = inputs(n, edges, dtype, seed=77 + n)
        if n > 100:
            x[5] = float('nan')                    # NaN takes the reference's branch
            x[6:16] = edges.to(dtype).to(DEV)
        y, gin = torch.empty_like(x), torch.empty_like(g)
        state = native.new_state(x, 1)
        state.fill_(0xFF)
        native.piecewise_forward(name, x, y, state, p0, p1)
        native.piecewise_backward(name, state, g, gin, p0)
        torch.cuda.synchronize()
        y_ref, state_ref = oracle.piecewise_forward(name, to_np(x), p0, p1)
        gin_ref = oracle.piecewise_backward(name, state_ref, to_np(g), p0)
        what = f'{name}({p0},{p1})/{tag}/n={n}'
        assert np.array_equal(state.cpu().numpy(), state_ref), f'{what}: mask differs'
        assert np.array_equal(to_np(gin), gin_ref), f'{what}: gradient differs'
        if name == 'hardsigmoid':                  # (x+3)*(1/6) in fp32 vs (x+3)/6 in double
            check_forward_values(to_np(y), y_ref, to_np(x), dtype, what)
        else:
            np.testing.assert_array_equal(as_f32(to_np(y)), as_f32(y_ref), err_msg=what)


def test_golden_reference_cases(golden_ops):
    """The CUDA path reproduces what the unmodified reference CPU ops produced
    (tests/golden/reference_cpu_ops.npz): packed bytes and gradients bit for bit."""
    for case in golden_ops:
        dtype = torch.bfloat16 if case['bf16'] else torch.float32

        def dev(a):
            t = torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16) if case['bf16'] \
                else torch.from_numpy(np.ascontiguousarray(a))
            return t.to(DEV)

        x, g, bounds, levels = (dev(case[k]) for k in ('x', 'g', 'bounds', 'levels'))
        y, gin = torch.empty_like(x), torch.empty_like(g)
        state = native.new_state(x, case['bits'])
        native.stepwise_forward('gelu', x, y, state, case['bits'], bounds)
        native.stepwise_backward(state, g, gin, case['bits'], levels)
        torch.cuda.synchronize()
        assert np.array_equal(state.cpu().numpy(), case['state']), case['key']
        assert np.array_equal(to_np(gin).view(np.uint8), case['gin'].view(np.uint8)), case['key']
        y_ref = as_f32(case['y'])
        ok = np.isfinite(as_f32(case['x'])) & np.isfinite(y_ref)
        tol = (np.abs(y_ref[ok]) * 2.0 ** -7 + 1e-6) if case['bf16'] else \
            (4 * np.spacing(np.abs(y_ref[ok])) + 1e-6)   # ATen CPU gelu: see tests/test_oracle.py
        assert np.all(np.abs(as_f32(to_np(y))[ok] - y_ref[ok]) <= tol), case['key']


def test_reference_cuda_test_vector():
    # fewbit/cuda/codec_test.cu:16-24, 93-98 (TestGelu): codes ae73501ad8fa, SURVEY App. B grads
    from test_oracle import BOUNDS, INPUTS, LEVELS
    x = torch.from_numpy(INPUTS).to(DEV)
    y, gin = torch.empty_like(x), torch.empty_like(x)
    state = native.new_state(x, 3)
    native.stepwise_forward('gelu', x, y, state, 3, torch.from_numpy(BOUNDS).to(DEV))
    native.stepwise_backward(state, torch.ones_like(x), gin, 3, torch.from_numpy(LEVELS).to(DEV))
    assert state.cpu().numpy().tobytes().hex() == 'ae73501ad8fa'
    np.testing.assert_array_equal(gin.cpu().numpy(), LEVELS[[6, 5, 6, 1, 7, 0, 4, 2, 2, 3, 0, 4, 5, 5, 6, 7]])


def test_nan_maps_to_code_zero_and_short_tables():
    # NaN -> 0 like the reference CUDA BinarySearch (codec.cu:124); 5 levels -> 3 bits, 4 bounds
    bounds = torch.tensor([-1.0, 0.0, 0.5, 2.0], device=DEV)
    levels = torch.tensor([0.1, 0.2, 0.3, 0.4, 0.5], device=DEV)
    assert native.bits_for_levels(5) == 3
    x = torch.tensor([float('nan'), -5.0, -1.0, -0.5, 0.25, 1.0, 2.0, 2.5, float('nan')], device=DEV)
    y, gin = torch.empty_like(x), torch.empty_like(x)
    state = native.new_state(x, 3)
    native.stepwise_forward('tanh', x, y, state, 3, bounds)
    native.stepwise_backward(state, torch.ones_like(x), gin, 3, levels)
    codes = oracle.inflate(state.cpu().numpy(), 9, 3).tolist()
    assert codes == [0, 0, 0, 1, 2, 3, 3, 4, 0]
    np.testing.assert_array_equal(gin.cpu().numpy(), levels.cpu().numpy()[codes])


@pytest.mark.parametrize('tag', DTYPES)
@pytest.mark.parametrize('bits', [3, 4, 6, 8])
def test_crowded_and_degenerate_tables(bits, tag):
    """Tables the cell look-up cannot separate (clustered / duplicated / single borders) must
    take the exact search and still produce lower_bound codes."""
    dtype = DTYPES[tag]
    nb = (1 << bits) - 1
    gen = torch.Generator().manual_seed(bits)
    tables = {
        'clustered': torch.cat([torch.tensor([-3.0]), torch.linspace(0.5, 0.5001, nb - 2), torch.tensor([4.0])]),
        'duplicates': torch.sort(torch.randint(-3, 4, (nb, ), generator=gen).float()).values,
        'all_equal': torch.full((nb, ), 0.25),
        'single': torch.tensor([0.1]),
        'short': torch.sort(torch.randn(nb // 2 + 1, generator=gen)).values,
    }
    for label, bounds in tables.items():
        bounds = bounds.to(dtype).to(DEV)
        levels = torch.linspace(-1, 1, bounds.numel() + 1).to(dtype).to(DEV)
        x, g = inputs(4099, bounds, dtype, seed=bits)
        x[100:100 + min(bounds.numel(), 64)] = bounds[:64]
        y, gin = torch.empty_like(x), torch.empty_like(g)
        state = native.new_state(x, bits)
        native.stepwise_forward('sigmoid', x, y, state, bits, bounds)
        native.stepwise_backward(state, g, gin, bits, levels)
        y_ref, state_ref = oracle.stepwise_forward('sigmoid', to_np(x), to_np(bounds), bits)
        assert np.array_equal(state.cpu().numpy(), state_ref), f'{label}/{tag}/bits={bits}'
        assert np.array_equal(to_np(gin), oracle.stepwise_backward(state_ref, to_np(g), to_np(levels), bits))


@pytest.mark.parametrize('tag', DTYPES)
def test_in_place_unaligned_and_empty(tag):
    dtype = DTYPES[tag]
    bounds, levels = table('silu', 3, dtype)
    # in place (y is x), as the operators run
    x, g = inputs(4099, bounds, dtype, 5)
    x0 = x.clone()
    state = native.new_state(x, 3)
    native.stepwise_forward('silu', x, x, state, 3, bounds)
    native.stepwise_backward(state, g, g, 3, levels)       # gin may alias gout too
    y_ref, state_ref = oracle.stepwise_forward('silu', to_np(x0), to_np(bounds), 3)
    assert np.array_equal(state.cpu().numpy(), state_ref)
    check_forward_values(to_np(x), y_ref, to_np(x0), dtype, 'in-place')
    # pointers that are only element-aligned: the whole tensor takes the ragged path
    big, gbig = inputs(5000, bounds, dtype, 6)
    for off in (1, 3):
        xs, gs = big[off:], gbig[off:]
        assert xs.data_ptr() % 16 != 0
        ys, gi = torch.empty_like(big)[off:], torch.empty_like(big)[off:]
        st = native.new_state(xs, 3)
        native.stepwise_forward('silu', xs, ys, st, 3, bounds)
        native.stepwise_backward(st, gs, gi, 3, levels)
        y_ref, state_ref = oracle.stepwise_forward('silu', to_np(xs), to_np(bounds), 3)
        assert np.array_equal(st.cpu().numpy(), state_ref)
        assert np.array_equal(to_np(gi), oracle.stepwise_backward(state_ref, to_np(gs), to_np(levels), 3))
        check_forward_values(to_np(ys), y_ref, to_np(xs), dtype, 'unaligned')
    # empty
    e = torch.empty(0, dtype=dtype, device=DEV)
    native.stepwise_forward('silu', e, e, native.new_state(e, 3), 3, bounds)
    native.piecewise_forward('relu', e, e, native.new_state(e, 1))
    native.piecewise_backward('relu', native.new_state(e, 1), e, e)


@pytest.mark.parametrize('bits', range(1, 9))
def test_standalone_codec(bits):
    for n in (1, 8, 13, 1000, 4099, 1 << 20):
        codes = torch.randint(0, 1 << bits, (n, ), dtype=torch.int32, device=DEV)
        state = torch.full((native.state_bytes(n, bits), ), 0x55, dtype=torch.uint8, device=DEV)
        native.deflate(codes, state, bits)
        back = torch.empty_like(codes)
        native.inflate(state, back, bits)
        assert np.array_equal(state.cpu().numpy(), oracle.deflate(codes.cpu().numpy(), bits))
        assert torch.equal(back, codes)


def test_host_staged_variants_match_device_variants():
    torch.manual_seed(3)
    n = 3 * 2048 * 5 + 77                               # several chunks + ragged tail
    for dtype in (torch.float32, torch.bfloat16):
        bounds, levels = table('gelu', 3, dtype)
        x = (torch.randn(n) * 2).to(dtype).pin_memory()
        g = torch.randn(n).to(dtype).pin_memory()
        y, gin = torch.empty_like(x).pin_memory(), torch.empty_like(g).pin_memory()
        state = torch.empty(native.state_bytes(n, 3), dtype=torch.uint8, device=DEV)
        native.stepwise_forward_host('gelu', x, y, state, 3, bounds, chunk=2048 * 2)
        native.stepwise_backward_host(state, g, gin, 3, levels, chunk=2048 * 2)
        xd, gd = x.to(DEV), g.to(DEV)
        yd, gid = torch.empty_like(xd), torch.empty_like(gd)
        sd = native.new_state(xd, 3)
        native.stepwise_forward('gelu', xd, yd, sd, 3, bounds)
        native.stepwise_backward(sd, gd, gid, 3, levels)
        assert torch.equal(state, sd) and torch.equal(y.to(DEV), yd) and torch.equal(gin.to(DEV), gid)
        mask = torch.empty(native.state_bytes(n, 1), dtype=torch.uint8, device=DEV)
        native.piecewise_forward_host('leaky_relu', x, y, mask, 0.01, chunk=2048 * 3)
        native.piecewise_backward_host('leaky_relu', mask, g, gin, 0.01, chunk=2048 * 3)
        md = native.new_state(xd, 1)
        native.piecewise_forward('leaky_relu', xd, yd, md, 0.01)
        native.piecewise_backward('leaky_relu', md, gd, gid, 0.01)
        assert torch.equal(mask, md) and torch.equal(y.to(DEV), yd) and torch.equal(gin.to(DEV), gid)


# --------------------------------------------------------------- BASELINE.json full sizes ----
# The scalar oracle would take minutes here; use size-independent properties instead, checked
# with independent torch ops on the GPU.

def unpack_on_gpu(state, n, bits):
    codes = torch.empty(n, dtype=torch.int32, device=DEV)
    native.inflate(state, codes, bits)
    return codes


@pytest.mark.parametrize('tag', DTYPES)
def test_full_size_gelu3(tag):
    """Config 1/3: 128 x 128 x 3072, 3-bit GELU.  encode -> decode round trip equals
    torch.searchsorted; backward equals levels[code] * g bit for bit; forward within 4 ulp /
    1 bf16 ulp of torch.nn.functional.gelu; idempotent."""
    dtype = DTYPES[tag]
    torch.manual_seed(0)
    n = 128 * 128 * 3072
    bounds, levels = table('gelu', 3, dtype)
    x = (torch.randn(n, device=DEV) * 2).to(dtype)
    g = torch.randn(n, device=DEV).to(dtype)
    y, gin = torch.empty_like(x), torch.empty_like(g)
    state = native.new_state(x, 3)
    native.stepwise_forward('gelu', x, y, state, 3, bounds)
    native.stepwise_backward(state, g, gin, 3, levels)
    assert state.numel() == 18874368
    codes = unpack_on_gpu(state, n, 3)
    want = torch.searchsorted(bounds.float(), x.float()).to(torch.int32)
    assert torch.equal(codes, want)
    assert torch.equal(gin, (levels.float()[want.long()] * g.float()).to(dtype))
    ref = torch.nn.functional.gelu(x.float())
    if dtype == torch.float32:
        tol = 4 * torch.abs(torch.nextafter(ref, ref * 2) - ref) + 2.5e-7
    else:
        tol = ref.abs() * 2.0 ** -7 + 1e-6
    assert torch.all((y.float() - ref).abs() <= tol)
    state2 = torch.empty_like(state)
    native.stepwise_forward('gelu', x, torch.empty_like(x), state2, 3, bounds)
    assert torch.equal(state, state2)
    # stand-alone deflate of the unpacked codes reproduces the stream (checksum of checksums)
    native.deflate(codes, state2.zero_(), 3)
    assert torch.equal(state, state2)


@pytest.mark.parametrize('name,p0,p1', [('relu', 0.0, 0.0), ('leaky_relu', 0.01, 0.0),
                                        ('hardtanh', -1.0, 1.0)])
def test_full_size_masks_bf16(name, p0, p1):
    """Config 2: 1 GiB bf16 tensor (n = 2^29), 1-bit masks.  The 64 MiB mask equals the
    predicate bit for bit; popcount matches; values and gradients match torch exactly."""
    torch.manual_seed(1)
    n = 1 << 29
    x = torch.empty(n, dtype=torch.bfloat16, device=DEV).normal_(0, 2)
    g = torch.empty(n, dtype=torch.bfloat16, device=DEV).normal_()
    y, gin = torch.empty_like(x), torch.empty_like(g)
    state = native.new_state(x, 1)
    native.piecewise_forward(name, x, y, state, p0, p1)
    native.piecewise_backward(name, state, g, gin, p0)
    assert state.numel() == 64 << 20
    pred = {'relu': x > 0, 'leaky_relu': x < 0, 'hardtanh': (x > p0) & (x < p1)}[name]
    weights = (2 ** torch.arange(8, device=DEV)).to(torch.uint8)
    packed = (pred.view(-1, 8).to(torch.uint8) * weights).sum(dim=1, dtype=torch.int32).to(torch.uint8)
    assert torch.equal(state, packed)
    fn = getattr(torch.nn.functional, name)
    args = {'relu': (), 'leaky_relu': (p0, ), 'hardtanh': (p0, p1)}[name]
    assert torch.equal(y, fn(x, *args))
    if name == 'leaky_relu':
        want = torch.where(pred, (g.float() * p0).to(torch.bfloat16), g)
    else:
        want = torch.where(pred, g, g * 0)           # 0 * g: keeps the sign of zero like the reference
    assert torch.equal(gin.view(torch.int16), want.view(torch.int16))


def test_more_than_2_32_elements():
    """64-bit element counts (the reference is capped by uint32_t)."""
    n = (1 << 32) + 4099
    x = torch.empty(n, dtype=torch.bfloat16, device=DEV)
    x[:1 << 20].normal_(0, 2)
    x[1 << 20:] = x[:1 << 20].repeat((n >> 20) + 1)[:n - (1 << 20)]
    x[-4099:].normal_(0, 2)
    state = native.new_state(x, 1)
    native.piecewise_forward('relu', x, x, state)
    assert state.numel() == (n + 7) // 8
    start = (state.numel() - 1025) * 8                 # the last 1025 bytes straddle 2^32
    assert start < (1 << 32) < n
    codes = oracle.inflate(state[-1025:].cpu().numpy(), n - start, 1)
    got = torch.from_numpy(codes.astype(np.bool_)).to(DEV)
    assert torch.equal(got, x[start:] > 0)             # relu(x) > 0  <=>  x > 0
    head = oracle.inflate(state[:4096].cpu().numpy(), 8 * 4096, 1)
    assert torch.equal(torch.from_numpy(head.astype(np.bool_)).to(DEV), x[:8 * 4096] > 0)
    assert int(torch.count_nonzero(x[start:] < 0)) == 0
