"""Property tests (hypothesis) of the host-side pieces: codec layout, the CPU operators of the op
library against the oracle on arbitrary tables, and the shard rule.  CPU only."""
import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

import fewbit_b200  # noqa: F401  (registers torch.ops.fewbit)
import oracle
from fewbit_b200.sharding import shard_bounds, state_offset

settings.register_profile('fewbit', max_examples=60, deadline=None)
settings.load_profile('fewbit')


@given(bits=st.integers(1, 8), n=st.integers(0, 300), seed=st.integers(0, 2 ** 31 - 1))
def test_codec_round_trip_and_layout(bits, n, seed):
    """inflate(deflate(codes)) == codes for every length (ragged tails included), the stream has
    ceil(n bits / 8) bytes with zero pad bits, and it is the numpy LSB-first packing."""
    codes = np.random.default_rng(seed).integers(0, 1 << bits, n, dtype=np.int32)
    state = oracle.deflate(codes, bits)
    assert state.size == (n * bits + 7) // 8 == oracle.state_bytes(n, bits)
    assert np.array_equal(oracle.inflate(state, n, bits), codes)
    assert np.array_equal(state, oracle.deflate_numpy(codes, bits))
    if n * bits % 8:
        assert state[-1] >> (n * bits % 8) == 0


@given(bits=st.integers(1, 8), n=st.integers(1, 400), seed=st.integers(0, 2 ** 31 - 1), bf16=st.booleans())
def test_cpu_operators_agree_with_the_oracle_on_arbitrary_tables(bits, n, seed, bf16):
    """torch.ops.fewbit.quantize / quantize_backward on CPU tensors (csrc/torch_ops.cc) against the
    oracle for random sorted tables with ties, inputs sitting exactly on borders, signed zeros and
    infinities: the same packed bytes and the same gradients, for every bit width and length."""
    rng = np.random.default_rng(seed)
    nlevels = int(rng.integers((1 << (bits - 1)) + 1, (1 << bits) + 1)) if bits > 1 else 2
    bounds = np.sort(rng.normal(0, 2, nlevels - 1)).astype(np.float32)
    if nlevels > 3 and rng.random() < 0.3:
        bounds[1] = bounds[0]                                   # a tie
    levels = rng.normal(0, 1, nlevels).astype(np.float32)
    x = rng.normal(0, 3, n).astype(np.float32)
    x[: min(n, bounds.size)] = bounds[: min(n, bounds.size)]    # exactly on borders
    if n > 4:
        x[-4:] = [0.0, -0.0, np.inf, -np.inf]
    g = rng.normal(0, 1, n).astype(np.float32)
    dtype = torch.bfloat16 if bf16 else torch.float32
    tx, tb, tl, tg = (torch.from_numpy(a).to(dtype) for a in (x, bounds, levels, g))
    y, state = torch.ops.fewbit.quantize(tx, tb)
    gin = torch.ops.fewbit.quantize_backward(tg, state, tl)
    want_codes = oracle.bucketize(tx.float().numpy(), tb.float().numpy())
    assert np.array_equal(state.numpy(), oracle.deflate(want_codes.astype(np.int32), bits))
    want = (tl.float()[torch.from_numpy(want_codes.astype(np.int64))] * tg.float()).to(dtype)
    assert torch.equal(gin, want)
    assert torch.allclose(y.float(), torch.nn.functional.gelu(tx).float(), rtol=0, atol=0, equal_nan=True)   # gelu(-inf) is NaN


@given(n=st.integers(0, 1 << 20), world=st.integers(1, 8), bits=st.integers(1, 8),
       align=st.sampled_from([8, 256, 2048]))
def test_shards_partition_the_tensor_and_their_streams_concatenate(n, world, bits, align):
    edges = [shard_bounds(n, r, world, align) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == n
    for (b0, e0), (b1, e1) in zip(edges, edges[1:]):
        assert e0 == b1 and b0 <= e0
    for begin, end in edges[:-1]:
        assert begin % 8 == 0 and (end - begin) % align == 0
    # byte offsets of the shards' packed streams tile the whole stream
    for begin, end in edges:
        assert state_offset(begin, bits) == begin * bits // 8
    assert state_offset(edges[-1][0], bits) + oracle.state_bytes(n - edges[-1][0], bits) == oracle.state_bytes(n, bits)
