"""Saved-tensor accounting (reference fewbit/util.py:20-144): the expectations of the reference's
own fewbit/util_test.py, restated against this package, and -- on the GPU -- what the accounting
is for: a 3-bit GELU keeps the packed state instead of an activation."""
import pytest
import torch as T

import fewbit_b200 as fewbit
from fewbit_b200.util import estimate_memory_usage, memory_usage_hooks, teniter, traverse


@pytest.fixture
def models():
    T.manual_seed(0)
    plain = T.nn.Sequential(T.nn.Linear(8, 4), T.nn.Linear(4, 1))
    with_relu = T.nn.Sequential(T.nn.Linear(8, 4), T.nn.ReLU(), T.nn.Linear(4, 1))
    return plain, with_relu


def test_traverse_visits_the_graph(models):           # util_test.py:24-27
    seen = []
    traverse(models[0](T.randn(3, 8).requires_grad_()), lambda node, ten, saved: seen.append(saved))
    assert seen and any(seen) and not all(seen)


def test_teniter_counts_one_more_saved_tensor_with_relu(models):     # util_test.py:29-37
    xs = T.randn(3, 8)
    lhs = len(list(teniter(models[0](xs.requires_grad_()), False, True)))
    rhs = len(list(teniter(models[1](xs.requires_grad_()), False, True)))
    assert lhs + 1 == rhs


def test_estimate_memory_usage_counts_leaves(models):                # util_test.py:39-45
    ys = models[0](T.randn(3, 8).requires_grad_())
    assert estimate_memory_usage(ys) == 4 * (3 * 8 + 4 * 8 + 4 + 1 * 4 + 1)


def test_estimate_memory_usage_saved_only(models):                   # util_test.py:47-60
    xs = T.randn(3, 8)
    lhs = estimate_memory_usage(models[0](xs.requires_grad_()), True)
    rhs = estimate_memory_usage(models[1](xs.requires_grad_()), True)
    assert rhs - lhs == 3 * 4 * 4          # the ReLU output, one more fp32 [3, 4] tensor


def test_memory_usage_hooks(models):                                 # util_test.py:62-77
    with memory_usage_hooks() as lhs:
        xs = T.randn(3, 8)
        models[0](xs.requires_grad_())
    with memory_usage_hooks() as rhs:
        xs = T.randn(3, 8)
        ys = models[1](xs.requires_grad_())
        ys.backward(T.ones(xs.shape[0], 1))
    assert rhs.value - lhs.value == 3 * 4 * 4


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [T.float32, T.bfloat16])
def test_three_bit_gelu_trades_an_activation_for_its_packed_state(dtype):
    """A RoBERTa feed-forward block (768 -> 3072 -> GELU -> 768) on 128 x 128 tokens: with
    fewbit.GELU(bits=3) the bytes autograd saves shrink by exactly one [tokens, 3072] activation
    minus the packed 3-bit state and its 8 levels -- the in-place operator saves nothing else
    (reference cuda/activation.cc:345-363: save_for_backward({buffer, levels}))."""
    tokens, hidden, inner = 128 * 128, 768, 3072
    T.manual_seed(0)

    def block(act):
        return T.nn.Sequential(T.nn.Linear(hidden, inner), act, T.nn.Linear(inner, hidden)).to('cuda', dtype)

    def saved_bytes(model):
        with memory_usage_hooks() as usage:
            x = T.randn(tokens, hidden, device='cuda', dtype=dtype, requires_grad=True)
            y = model(x)
        y.sum().backward()
        return usage.forward, usage.backward

    vanilla_fwd, vanilla_bwd = saved_bytes(block(T.nn.GELU()))
    fewbit_fwd, fewbit_bwd = saved_bytes(block(fewbit.GELU(bits=3)))
    es = T.finfo(dtype).bits // 8
    n = tokens * inner
    state = (n * 3 + 7) // 8
    assert vanilla_fwd - fewbit_fwd == n * es - state - 8 * es
    assert vanilla_bwd - fewbit_bwd == n * es - state - 8 * es
    # and the graph walk sees the same activation disappear (C++ nodes do not expose what they save)
    x = T.randn(tokens, hidden, device='cuda', dtype=dtype, requires_grad=True)
    walked = [estimate_memory_usage(m(x), saved_only=True) for m in (block(T.nn.GELU()), block(fewbit.GELU(bits=3)))]
    assert walked[0] - walked[1] == n * es
