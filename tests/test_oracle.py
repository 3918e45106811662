"""Pin the CPU oracle (oracle/fewbit_oracle.c) before anything is allowed to trust it:
against the reference tests' own golden vectors, against the unmodified reference compiled
into oracle/_ref/, against committed fixtures produced by running the reference, and against
an independent numpy packer."""
import hashlib

import numpy as np
import pytest
import torch

import oracle

# fewbit/cuda/codec_test.cu:16-24 == approx_test.py:23-32 == benchmark/bench-roberta.py:128-136
BOUNDS = np.array([-2.39798704e+00, -7.11248159e-01, -3.26290283e-01, -1.55338428e-04,
                   +3.26182064e-01, +7.10855860e-01, +2.39811567e+00], np.float32)
LEVELS = np.array([-2.600090e-03, -8.883533e-02, 1.251944e-01, 3.720415e-01, +6.277958e-01,
                   +8.746618e-01, 1.088807e+00, 1.002599e+00], np.float32)
# fewbit/cuda/codec_test.cu:93-98
INPUTS = np.array([2.29811567e+00, 6.10855860e-01, 2.29811567e+00, -8.11248159e-01,
                   9.99900000e+02, -2.49798704e+00, 2.26182064e-01, -4.26290283e-01,
                   -4.26290283e-01, -1.00155338e-01, -2.49798704e+00, 2.26182064e-01,
                   6.10855860e-01, 6.10855860e-01, 2.29811567e+00, 9.99900000e+02], np.float32)
# fewbit/cuda/codec_test.cu:62-64
CODES = [6, 5, 6, 1, 7, 0, 4, 2, 2, 3, 0, 4, 5, 5, 6, 7]


def test_codec_example_of_reference_cpu_test():
    # fewbit/cpu/codec_test.cc:9-22: {0,1,4,7} at 3 bits round-trips; bytes 08 0f
    packed = oracle.deflate([0, 1, 4, 7], 3)
    assert packed.tobytes().hex() == '080f'
    assert oracle.inflate(packed, 4, 3).tolist() == [0, 1, 4, 7]


def test_codec_block_vector_of_reference_cuda_test():
    packed = oracle.deflate(CODES, 3)
    assert packed.tobytes().hex() == 'ae73501ad8fa'
    assert oracle.inflate(packed, 16, 3).tolist() == CODES


def test_gelu_vector_of_reference_cuda_test():
    # TestGelu, fewbit/cuda/codec_test.cu:92-147: codes, packed bytes and gradients.
    codes = oracle.bucketize(INPUTS, BOUNDS)
    assert codes.tolist() == CODES
    y, state = oracle.stepwise_forward('gelu', INPUTS, BOUNDS)
    assert state.tobytes().hex() == 'ae73501ad8fa'
    gin = oracle.stepwise_backward(state, np.ones(16, np.float32), LEVELS)
    want = [1.088807, 0.8746618, 1.088807, -0.08883533, 1.002599, -0.00260009, 0.6277958,
            0.1251944, 0.1251944, 0.3720415, -0.00260009, 0.6277958, 0.8746618, 0.8746618,
            1.088807, 1.002599]
    np.testing.assert_array_equal(gin, np.array(want, np.float32))
    np.testing.assert_allclose(y, torch.nn.functional.gelu(torch.from_numpy(INPUTS)).numpy(),
                               rtol=3e-7, atol=1e-7)


def test_linspace_vectors(golden_tables):
    # built-in gelu 3-bit on linspace(-5, 5, {101, 23}).
    bounds = golden_tables['gelu03-f32-bounds']
    x = torch.linspace(-5, 5, 101).numpy()
    codes = oracle.bucketize(x, bounds)
    runs = [(int(v), int(c)) for v, c in zip(*np.unique(codes, return_counts=True))]
    assert runs == [(0, 26), (1, 17), (2, 4), (3, 4), (4, 3), (5, 4), (6, 17), (7, 26)]
    packed = oracle.deflate(codes, 3)
    assert packed.size == 38
    assert hashlib.sha256(packed.tobytes()).hexdigest()[:16] == '7bbd4c7d3f87ab2d'
    x = torch.linspace(-5, 5, 23).numpy()
    codes = oracle.bucketize(x, bounds)
    assert codes.tolist() == [0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 3, 5, 6, 6, 6, 6, 7, 7, 7, 7, 7, 7]
    assert oracle.deflate(codes, 3).tobytes().hex() == '0000248956dbfeff1f'


@pytest.mark.parametrize('bits', range(1, 9))
def test_codec_random_roundtrip_like_reference(bits):
    # fewbit/cpu/codec_test.cc:24-51: 256 random codes per width; plus ragged lengths.
    rng = np.random.default_rng(42 + bits)
    for n in (0, 1, 7, 8, 9, 255, 256, 257, 1000, 4099):
        codes = rng.integers(0, 1 << bits, n, dtype=np.int32)
        packed = oracle.deflate(codes, bits)
        assert packed.size == (n * bits + 7) // 8 == oracle.state_bytes(n, bits)
        np.testing.assert_array_equal(oracle.inflate(packed, n, bits), codes)
        np.testing.assert_array_equal(packed, oracle.deflate_numpy(codes, bits))
        np.testing.assert_array_equal(oracle.inflate_numpy(packed, n, bits), codes)
        if n and (n * bits) % 8:  # pad bits of the last byte are zero
            assert packed[-1] >> ((n * bits) % 8) == 0


@pytest.mark.skipif(oracle.ref_codec() is None, reason='oracle/_ref/libref_codec.so not built')
@pytest.mark.parametrize('bits', range(1, 9))
def test_codec_against_reference_build(bits):
    # The real fewbit::Deflate / Inflate (fewbit/cpu/codec.h) through oracle/_ref.
    rng = np.random.default_rng(7 * bits)
    for n in (1, 5, 8, 13, 64, 1000, 4099, 65536):
        codes = rng.integers(0, 1 << bits, n, dtype=np.int32)
        ref = oracle.ref_deflate(codes, bits)
        np.testing.assert_array_equal(oracle.deflate(codes, bits), ref)
        np.testing.assert_array_equal(oracle.inflate(ref, n, bits), oracle.ref_inflate(ref, n, bits))


def test_against_reference_cpu_ops(golden_ops):
    """x -> (state, gin) of the unmodified reference ops (fewbit/cpu/gelu.cc) == oracle."""
    assert len(golden_ops) >= 80
    for case in golden_ops:
        y, state = oracle.stepwise_forward('gelu', case['x'], case['bounds'], case['bits'],
                                           nan_policy=oracle.NAN_TO_LAST)
        assert np.array_equal(state, case['state']), case['key']
        gin = oracle.stepwise_backward(case['state'], case['g'], case['levels'], case['bits'])
        assert np.array_equal(gin.view(np.uint8), case['gin'].view(np.uint8)), case['key']
        if case['bf16']:
            ours, ref = oracle.bf16_bits_to_f32(y), oracle.bf16_bits_to_f32(case['y'])
            finite = np.isfinite(ref)
            # double -> bf16 vs float-math -> bf16: at most one bf16 ulp apart
            assert np.all(np.abs(ours[finite] - ref[finite]) <=
                          np.maximum(np.abs(ref[finite]) * 2.0 ** -7, 1e-6)), case['key']
        else:
            finite = np.isfinite(case['y'])
            # ATen's fp32 gelu is 0.5*x*(1+erf(x/sqrt 2)): in the negative tail 1+erf cancels and
            # carries an absolute error of a few 1e-7 that the double-precision oracle lacks.
            np.testing.assert_allclose(y[finite], case['y'][finite], rtol=5e-7, atol=1e-6,
                                       err_msg=case['key'])


def test_bucketize_semantics():
    # ties go down, signed zeros are equal, infinities saturate, NaN policy.
    b = np.array([-1.0, 0.0, 2.0], np.float32)
    x = np.array([-np.inf, -1.0, np.nextafter(np.float32(-1), np.float32(0)), -0.0, 0.0,
                  1e-45, 2.0, np.nextafter(np.float32(2), np.float32(3)), np.inf, np.nan],
                 np.float32)
    assert oracle.bucketize(x, b).tolist() == [0, 0, 1, 1, 1, 2, 2, 3, 3, 0]
    assert oracle.bucketize(x, b, oracle.NAN_TO_LAST).tolist()[-1] == 3
    want = torch.searchsorted(torch.from_numpy(b), torch.from_numpy(x[:-1])).tolist()
    assert oracle.bucketize(x[:-1], b).tolist() == want


def test_bf16_helpers_match_torch():
    rng = np.random.default_rng(3)
    v = np.concatenate([rng.standard_normal(4096).astype(np.float32) * 3,
                        np.array([0.0, -0.0, np.inf, -np.inf, 1e-40, 3.3895314e38], np.float32)])
    want = torch.from_numpy(v).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    np.testing.assert_array_equal(oracle.f32_to_bf16_bits(v), want)
    got = np.array([oracle.lib().orc_f32_to_bf16(float(t)) for t in v[:512]], np.uint16)
    np.testing.assert_array_equal(got, want[:512])
    np.testing.assert_array_equal(oracle.bf16_bits_to_f32(want),
                                  torch.from_numpy(v).to(torch.bfloat16).float().numpy())


@pytest.mark.parametrize('func', oracle.CONTINUOUS)
def test_forward_values_follow_torch(func):
    # The reference's own criterion (functional/activations_test.py:88-89): L2 distance to
    # torch.nn.functional on linspace(-5, 5, 101) below 1e-6.
    x = torch.linspace(-5, 5, 101)
    fn = getattr(torch.nn.functional, func, None) or getattr(torch, func)
    y, _ = oracle.stepwise_forward(func, x.numpy(), BOUNDS)
    # float64 torch is the yardstick (exact for all but selu, whose two constants are rounded
    # to fp32 like the device does, codec.cu:591-594); torch's own fp32 CPU kernels are up to
    # 1.9e-6 (gelu) away from it, so they are only a loose cross-check here.
    assert np.linalg.norm(y - fn(x.double()).float().numpy()) < (2e-6 if func == 'selu' else 1e-7)
    assert np.linalg.norm(y - fn(x).numpy()) < 4e-6


@pytest.mark.parametrize('func,args', [
    ('hardshrink', (0.5, )), ('hardshrink', (1.0, )), ('hardsigmoid', ()), ('hardtanh', (-1.0, 1.0)),
    ('hardtanh', (-2.0, 2.0)), ('leaky_relu', (0.01, )), ('leaky_relu', (0.5, )), ('relu', ()),
    ('relu6', ()), ('softshrink', (0.5, )), ('softshrink', (1.0, )), ('threshold', (1.0, 3.0))])
def test_piecewise_like_reference_test(func, args):
    # functional/activations_test.py:17-68: value and gradient vs torch on linspace(-5, 5, 101)
    # (+ points beyond 6 so that relu6's upper branch is exercised).
    xs = torch.cat([torch.linspace(-5, 5, 101), torch.tensor([5.99, 6.0, 6.5, 100.0])])
    p = list(args) + [0.0, 0.0]
    y, state = oracle.piecewise_forward(func, xs.numpy(), p[0], p[1])
    gin = oracle.piecewise_backward(func, state, np.ones(xs.numel(), np.float32), p[0])
    q = xs.clone().requires_grad_()
    z = getattr(torch.nn.functional, func)(q, *args)
    z.backward(torch.ones_like(z))
    assert np.linalg.norm(y - z.detach().numpy()) < 5e-7
    grad = q.grad.numpy().copy()
    if func == 'leaky_relu':
        grad[xs.numpy() == 0] = 1.0  # reference passes g at exactly 0, torch slope*g (App. A)
    assert np.linalg.norm(gin - grad) < 5e-7
    assert state.size == (xs.numel() + 7) // 8


def test_empty_inputs():
    y, state = oracle.stepwise_forward('gelu', np.zeros(0, np.float32), BOUNDS)
    assert y.size == 0 and state.size == 0
    assert oracle.stepwise_backward(state, np.zeros(0, np.float32), LEVELS).size == 0
    y, state = oracle.piecewise_forward('relu', np.zeros(0, np.float32))
    assert y.size == 0 and state.size == 0


# ---- sketch entries (the projection's S is defined by this package: pinned here) ------------

def test_philox4x32_known_answers():
    """Random123's published known-answer vectors for philox4x32-10 (kat_vectors: zero, all-ones and
    pi-digit counter/key) pin the restated round function; the kernels run the same rounds, seven
    of them (oracle.PHILOX_ROUNDS)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff, ) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for counter, key, want in kat:
        got = oracle.philox4x32(np.array([counter], dtype=np.uint32), key[0] | (key[1] << 32), rounds=10)
        assert tuple(int(v) for v in got[0]) == want
    assert oracle.PHILOX_ROUNDS == 7


def test_oracle_sketch_matrix_statistics():
    r = oracle.sketch_matrix(64, 4096, 42, 4, 'rademacher')
    assert set(np.unique(r)) == {-0.5, 0.5} and abs(r.mean()) < 0.01
    g = oracle.sketch_matrix(64, 4096, 42, 4, 'gaussian')
    assert abs(g.mean()) < 0.01 and abs(g.std() - 1.0) < 0.01
    assert abs(np.corrcoef(g[:, :-1].ravel(), g[:, 1:].ravel())[0, 1]) < 0.01
    # a different offset or seed is a different matrix; the same pair reproduces it
    assert not np.array_equal(r, oracle.sketch_matrix(64, 4096, 42, 5, 'rademacher'))
    assert not np.array_equal(r, oracle.sketch_matrix(64, 4096, 43, 4, 'rademacher'))
    assert np.array_equal(r, oracle.sketch_matrix(64, 4096, 42, 4, 'rademacher'))
