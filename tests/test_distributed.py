"""N > 1 path on CPU: world_size-2 gloo.  Shards need no data-path collective: each rank packs
its own shard and the concatenation equals the single-process stream (the all_gather below is
test plumbing, not part of the path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from fewbit_b200.sharding import shard_bounds, state_offset


def test_shard_bounds_partition():
    for n in (0, 5, 2048, 4099, 1 << 20, (1 << 29) + 13):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(b % 2048 == 0 for b, _ in spans)
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)
    with pytest.raises(ValueError):
        shard_bounds(10, 0, 2, align=4)


def _worker(rank, world, port, n, bits, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import fewbit_b200 as fewbit
        torch.manual_seed(0)                      # every rank builds the same full tensor
        x = torch.randn(n) * 2
        g = torch.randn(n)
        begin, end = shard_bounds(n, rank, world)
        borders, levels = fewbit.functional.store.get('gelu', bits)
        # this rank's shard through the host path of the surface (CPU tensors)
        leaf = x[begin:end].clone().requires_grad_()
        y = fewbit.functional.gelu(leaf, bits=bits)
        y.backward(g[begin:end])
        codes = torch.searchsorted(borders[1:-1].contiguous(), x[begin:end])
        packed = torch.from_numpy(oracle.deflate(codes.numpy(), bits))
        # test plumbing only: gather the shards' packed bytes and gradients
        sizes = [shard_bounds(n, r, world) for r in range(world)]
        states, grads = [None] * world, [None] * world
        dist.all_gather_object(states, packed)
        dist.all_gather_object(grads, leaf.grad)
        if rank == 0:
            full_codes = torch.searchsorted(borders[1:-1].contiguous(), x)
            full_state = oracle.deflate(full_codes.numpy(), bits)
            cat = torch.cat(states).numpy()
            ok = np.array_equal(cat, full_state)
            for (b, _), s in zip(sizes, states):
                off = state_offset(b, bits)
                ok &= np.array_equal(full_state[off:off + s.numel()], s.numpy())
            ok &= torch.equal(torch.cat(grads), levels[full_codes] * g)
            out.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('bits', [1, 3])
def test_two_rank_sharding_gloo(bits):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    out = ctx.SimpleQueue()
    n = 3 * 4096 + 1001
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, bits, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get() is True
