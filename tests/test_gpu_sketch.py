"""The tcgen05 projection of RandomizedLinear: out = scale * S @ X with S generated in-kernel.

S is never materialised by the product; `sketch_matrix` reproduces it with the same device
function so that the kernel can be checked against a plain fp32 torch matmul of the same S
(tolerance 2e-3 of the result's RMS: bf16 operands are exact in the fp32 products, only the
accumulation order differs)."""
import numpy as np
import pytest
import torch

import fewbit_b200 as fewbit
from fewbit_b200 import native

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def reference(x, rows, seed, offset, kind, scale):
    s = native.sketch_matrix(rows, x.shape[0], seed, offset, kind, DEV)
    return (s.float() @ x.float()) * scale, s


@pytest.mark.parametrize('kind', ['gaussian', 'rademacher'])
def test_sketch_matrix_statistics(kind):
    s = native.sketch_matrix(512, 4096, 1234, 8, kind, DEV).float()
    again = native.sketch_matrix(512, 4096, 1234, 8, kind, DEV).float()
    assert torch.equal(s, again)                                   # pure function of (seed, offset)
    assert not torch.equal(s, native.sketch_matrix(512, 4096, 1234, 12, kind, DEV).float())
    assert not torch.equal(s, native.sketch_matrix(512, 4096, 1235, 8, kind, DEV).float())
    # a sub-block is the same function of (p, n): tiling-independent
    assert torch.equal(s[:100, :1000], native.sketch_matrix(100, 1000, 1234, 8, kind, DEV).float())
    if kind == 'gaussian':
        assert abs(s.mean().item()) < 3e-3 and abs(s.var().item() - 1) < 1e-2
        assert abs((s ** 4).mean().item() - 3) < 0.1               # kurtosis of a normal
        assert s.abs().max().item() > 4.0                          # tails are there
    else:
        assert set(s.unique().tolist()) == {-0.5, 0.5} and abs(s.mean().item()) < 2e-3
    gram = (s @ s.T) / s.shape[1]
    var = 1.0 if kind == 'gaussian' else 0.25
    off = gram - var * torch.eye(512, device=DEV)
    assert off.abs().max().item() < 0.12 * var                     # rows uncorrelated: |.| ~ var/sqrt(4096)*5


SHAPES = [(64, 8, 16), (1000, 72, 50), (4096, 384, 160), (4100, 392, 161), (2048, 768, 333),
          (16384, 768, 3276), (3000, 3072, 600), (777, 1024, 1),
          # feature counts that are multiples of 768 run on CTA pairs (Gaussian): edge shapes of that mode
          (100, 768, 1), (64, 1536, 17), (130, 2304, 161), (8200, 768, 145),
          # S slots span two 64-token stages: odd stage counts, a single token, the two-tile push mode over many slots
          (192, 768, 20), (1, 768, 3), (5000, 512, 300), (8000, 3072, 64)]


@pytest.mark.parametrize('rows,cols,seed,offset', [(5, 300, 42, 4), (161, 4100, 2 ** 40 + 3, 2 ** 33 + 9),
                                                   (3, 128, 0, 0)])
def test_sketch_entries_match_the_oracle(rows, cols, seed, offset):
    """S entry by entry against the numpy restatement (oracle.sketch_matrix): Rademacher bit for
    bit -- that pins the counter layout (column block, row, 64-bit offset) and the 64-bit key;
    Gaussian to the accuracy of the kernel's MUFU Box-Muller plus one bf16 rounding."""
    import oracle
    got = native.sketch_matrix(rows, cols, seed, offset, 'rademacher').float().cpu().numpy()
    assert np.array_equal(got, oracle.sketch_matrix(rows, cols, seed, offset, 'rademacher'))
    got = native.sketch_matrix(rows, cols, seed, offset, 'gaussian').float().cpu().numpy()
    want = oracle.sketch_matrix(rows, cols, seed, offset, 'gaussian')
    assert np.all(np.abs(got - want) <= np.abs(want) * 2.0 ** -7 + 1e-3)


@pytest.mark.parametrize('tokens,features,rows', SHAPES)
@pytest.mark.parametrize('kind', ['gaussian', 'rademacher'])
def test_kernel_matches_matmul_with_the_same_sketch(tokens, features, rows, kind):
    torch.manual_seed(tokens + features)
    x = torch.randn(tokens, features, device=DEV).to(torch.bfloat16)
    scale = 1.0 / rows
    out = native.sketch_forward(x, rows, 42, 4, kind, scale)
    want, _ = reference(x, rows, 42, 4, kind, scale)
    assert out.shape == (rows, features) and out.dtype == torch.float32
    err = (out - want).abs().max().item()
    assert err <= 2e-3 * want.pow(2).mean().sqrt().item() + 1e-7, f'max err {err}'


def test_operator_and_errors():
    x = torch.randn(512, 64, device=DEV).to(torch.bfloat16)
    out = torch.ops.fewbit.sketch(x, 32, 7, 0, 0, 0.5)
    want, _ = reference(x, 32, 7, 0, 'gaussian', 0.5)
    assert torch.allclose(out, want, rtol=1e-3, atol=1e-3)
    with pytest.raises(RuntimeError, match='bfloat16'):
        torch.ops.fewbit.sketch(x.float(), 32, 7, 0, 0, 1.0)
    with pytest.raises(RuntimeError, match='multiple of 8'):
        torch.ops.fewbit.sketch(x[:, :60].contiguous(), 32, 7, 0, 0, 1.0)


@pytest.mark.parametrize('kind', ['gaussian', 'rademacher'])
@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_randomized_linear_uses_one_sketch_forward_and_backward(kind, dtype):
    """grad_weight = (S G)^T (S X) * scale with the SAME S in both passes."""
    torch.manual_seed(3)
    layer = fewbit.RandomizedLinear(256, 128, proj_dim=64, matmul=kind,
                                    generator=torch.Generator(DEV).manual_seed(99)).to(DEV, dtype)
    x = torch.randn(4, 128, 256, device=DEV, dtype=dtype, requires_grad=True)
    g = torch.randn(4, 128, 128, device=DEV, dtype=dtype)
    seed, offset = layer.generator.initial_seed(), layer.generator.get_offset()
    y = layer(x)
    assert layer.generator.get_offset() == offset + 4              # the stream advanced
    y.backward(g)
    s = native.sketch_matrix(64, 512, seed, offset, kind, DEV).float()
    xb, gb = x.detach().reshape(512, 256).to(torch.bfloat16).float(), g.reshape(512, 128).to(torch.bfloat16).float()
    scale = 1 / 64 if kind == 'gaussian' else 4 / 64
    x_proj = ((s @ xb) * scale).to(dtype).float()
    want = (s @ gb).T @ x_proj
    rel = (layer.weight.grad.float() - want).norm() / want.norm()
    assert rel.item() < (2e-2 if dtype == torch.bfloat16 else 1e-3)
    ref = torch.nn.functional.linear(x, layer.weight, layer.bias)
    assert torch.allclose(y, ref, rtol=1e-2, atol=1e-2)
    assert torch.allclose(x.grad.float(), (g.float().reshape(512, 128) @ layer.weight.float()).view_as(x),
                          rtol=2e-2, atol=2e-2)


def test_randomized_linear_is_unbiased_on_cuda():
    # the reference's statistical test (fewbit/modules/linear_test.py:71-92) through the kernel
    torch.manual_seed(42)
    layer = fewbit.RandomizedLinear(256, 128, proj_dim=64).to(DEV)
    ref = torch.nn.Linear(256, 128).to(DEV)
    ref.load_state_dict(layer.state_dict())
    x = torch.randn(512, 256, device=DEV, requires_grad=True)
    acc = torch.zeros_like(layer.weight)
    for _ in range(2048):
        layer.zero_grad()
        y = layer(x)
        y.backward(torch.ones_like(y))
        acc += layer.weight.grad
    z = ref(x)
    z.backward(torch.ones_like(z))
    err = torch.linalg.norm(acc / 2048 - ref.weight.grad) / torch.linalg.norm(ref.weight.grad)
    assert err.item() < 0.1


def test_layers_fed_the_same_tensor_can_share_one_sketch():
    """share_sketch=True: query / key / value style layers called one after the other
    on the very same tensor take ONE sketch -- one projection launch, one advance of the random
    stream, one saved S X -- and each weight gradient is still (S G)^T (S X) with that S.  A different
    tensor, or the same tensor modified in place, is sketched afresh."""
    torch.manual_seed(5)
    gen = torch.Generator(DEV).manual_seed(7)
    q, k, v = (fewbit.RandomizedLinear(256, 128, proj_dim=64, generator=gen, share_sketch=True).to(DEV)
               for _ in range(3))
    hidden = torch.randn(4, 128, 256, device=DEV, requires_grad=True) * 1.0
    grads = [torch.randn(4, 128, 128, device=DEV) for _ in range(3)]
    seed, offset = gen.initial_seed(), gen.get_offset()
    before = native.lib().fewbit_launch_count()
    outs = [layer(hidden) for layer in (q, k, v)]
    torch.cuda.synchronize()
    launches = native.lib().fewbit_launch_count() - before
    assert gen.get_offset() == offset + 4                         # one draw for the three layers
    sum(o.mul(g).sum() for o, g in zip(outs, grads)).backward()
    s = native.sketch_matrix(64, 512, seed, offset, 'gaussian', DEV).float()
    xb = hidden.detach().reshape(512, 256).to(torch.bfloat16).float()
    x_proj = (s @ xb) / 64
    for layer, g in zip((q, k, v), grads):
        want = (s @ g.reshape(512, 128).to(torch.bfloat16).float()).T @ x_proj
        rel = (layer.weight.grad - want).norm() / want.norm()
        assert rel.item() < 1e-3
    # three forward sketches would be >= 3 projection launches; shared: one (+ its split-K reduce)
    single = native.lib().fewbit_launch_count()
    fewbit.RandomizedLinear(256, 128, proj_dim=64, generator=gen).to(DEV)(hidden.detach())
    torch.cuda.synchronize()
    per_sketch = native.lib().fewbit_launch_count() - single
    assert launches == per_sketch, (launches, per_sketch)
    # the record lives on the input tensor and is gone once a consumer has run backward: the same
    # tensor fed again (a later step) is sketched afresh, nothing keeps the projection alive
    assert '_fewbit_shared_sketch' not in hidden.__dict__
    offset = gen.get_offset()
    q(hidden), k(hidden)
    assert gen.get_offset() == offset + 4 and '_fewbit_shared_sketch' in hidden.__dict__
    # a modified or different tensor is not a hit
    offset = gen.get_offset()
    fresh = hidden.detach().clone()
    q(fresh)
    fresh.add_(1.0)
    k(fresh)
    v(fresh.clone())
    assert gen.get_offset() == offset + 12
    # without a gradient to estimate no sketch is taken at all (inference, frozen weights)
    offset = gen.get_offset()
    with torch.no_grad():
        q(fresh)
    assert gen.get_offset() == offset


def test_single_token_and_frozen_weight_on_cuda():
    """int(0.2 * N) == 0 for N < 5 tokens used to divide by zero on the native path;
    now an empty sketch and a zero weight gradient (the reference's behaviour)."""
    layer = fewbit.RandomizedLinear(256, 128, proj_dim_ratio=0.2).to(DEV)
    x = torch.randn(1, 256, device=DEV, requires_grad=True)
    y = layer(x)
    torch.testing.assert_close(y, torch.nn.functional.linear(x, layer.weight, layer.bias))
    y.sum().backward()
    assert torch.count_nonzero(layer.weight.grad) == 0 and x.grad is not None
    layer.weight.requires_grad_(False)
    before = native.lib().fewbit_launch_count()
    layer(torch.randn(512, 256, device=DEV, requires_grad=True)).sum().backward()
    assert native.lib().fewbit_launch_count() == before       # no projection kernel ran


@pytest.mark.parametrize('kind', ['dct', 'dft'])
def test_transform_sketches_run_on_cuda(kind):
    """The dct / dft sketches are torch.fft compositions on either device (reference
    fewbit/functional/linear.py:113-132); same estimate as on the CPU for the same sampled rows."""
    torch.manual_seed(3)
    layer = fewbit.RandomizedLinear(64, 32, proj_dim=16, matmul=kind, generator=torch.Generator(DEV).manual_seed(11)).to(DEV)
    x = torch.randn(8, 16, 64, device=DEV, requires_grad=True)
    y = layer(x)
    torch.testing.assert_close(y, torch.nn.functional.linear(x, layer.weight, layer.bias), rtol=1e-4, atol=1e-4)
    y.backward(torch.randn_like(y))
    assert layer.weight.grad.shape == (32, 64) and layer.weight.grad.dtype == torch.float32
    assert torch.isfinite(layer.weight.grad).all() and x.grad is not None


@pytest.mark.parametrize('kind', ['gaussian', 'rademacher'])
@pytest.mark.parametrize('tokens,features,rows', [(16384, 768, 3276), (1000, 392, 161), (4096, 3072, 144), (130, 768, 7)])
def test_projection_with_the_following_passes_folded_in(tokens, features, rows, kind):
    """fewbit_sketch_project: the result rounded to bf16 in the kernel equals the fp32 result rounded by
    torch, and with `column_sums` one more row of S is all ones, so the last output row is
    scale * x.sum(0) -- LinearGRPFunc.backward's `.to(dtype)` and `grad_output.sum(0)` passes
    (fewbit/functional/linear.py:199-217) without their trips through memory."""
    torch.manual_seed(rows)
    x = torch.randn(tokens, features, device=DEV).to(torch.bfloat16)
    scale = 0.5
    plain = native.sketch_forward(x, rows, 9, 8, kind, scale)
    narrow = native.sketch_project(x, rows, 9, 8, kind, scale, torch.bfloat16)
    assert narrow.dtype == torch.bfloat16 and torch.equal(narrow, plain.to(torch.bfloat16))
    for dtype in (torch.float32, torch.bfloat16):
        both = native.sketch_project(x, rows, 9, 8, kind, scale, dtype, column_sums=True)
        assert both.shape == (rows + 1, features) and both.dtype == dtype
        rms = plain.pow(2).mean().sqrt().item()
        tol = 2e-3 * rms if dtype == torch.float32 else 2.0 ** -7 * plain.abs().max().item()
        assert (both[:rows].float() - plain).abs().max().item() <= tol        # the sketch rows are unchanged
        sums = x.float().sum(0) * scale
        err = (both[rows].float() - sums).abs().max().item()
        assert err <= (1e-3 if dtype == torch.float32 else 2.0 ** -7) * sums.abs().max().item() + 1e-4, err


def test_bias_gradient_comes_out_of_the_projection_kernel():
    """bf16 layer with a bias: the bias gradient is the extra row of the S G product (no separate
    reduction over grad_output) and equals grad_output.sum(0) as torch rounds it."""
    torch.manual_seed(11)
    layer = fewbit.RandomizedLinear(256, 128, proj_dim=64, generator=torch.Generator(DEV).manual_seed(5)).to(DEV, torch.bfloat16)
    x = torch.randn(8, 64, 256, device=DEV, dtype=torch.bfloat16, requires_grad=True)
    g = torch.randn(8, 64, 128, device=DEV, dtype=torch.bfloat16)
    layer(x).backward(g)
    want = g.reshape(-1, 128).sum(0)
    assert layer.bias.grad.dtype == torch.bfloat16
    assert (layer.bias.grad.float() - want.float()).abs().max().item() <= 2.0 ** -7 * want.float().abs().max().item()
    # fp32 layers keep torch's own fp32 sum (the kernel would round grad_output to bf16 first)
    layer32 = fewbit.RandomizedLinear(256, 128, proj_dim=64, generator=torch.Generator(DEV).manual_seed(5)).to(DEV)
    x32 = torch.randn(8, 64, 256, device=DEV, requires_grad=True)
    g32 = torch.randn(8, 64, 128, device=DEV)
    layer32(x32).backward(g32)
    assert torch.equal(layer32.bias.grad, g32.reshape(-1, 128).sum(0))
