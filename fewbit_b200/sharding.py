"""Batch sharding of the hot path across GPUs.

Every element is independent, and the packed stream of a concatenation is the concatenation
of the packed streams whenever each shard holds a multiple of 8 elements (8 elements <-> `bits`
whole bytes).  So N GPUs process N contiguous shards with no collective on the data path; the
packed state is consumed by the backward pass on the GPU that produced it.  ``shard_bounds``
is the partition rule for a caller that splits ONE tensor over ranks (tests/test_distributed.py,
tests/test_properties.py); bench.py scales weakly -- every rank owns a whole 1 GiB tensor -- and
needs no partition.
"""
from __future__ import annotations

from typing import Tuple


def shard_bounds(n: int, rank: int, world: int, align: int = 2048) -> Tuple[int, int]:
    """[begin, end) of `rank`'s shard: equal shares rounded to `align` elements (a multiple of
    8, so packed streams concatenate; 2048 keeps whole warp tiles), the remainder goes last."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f'bad rank {rank} of {world}')
    if align % 8:
        raise ValueError('align must be a multiple of 8 so that packed shards concatenate')
    share = (n // world) // align * align
    begin = rank * share
    end = n if rank == world - 1 else begin + share
    return begin, end


def state_offset(begin: int, bits: int) -> int:
    """Byte offset of a shard's packed codes inside the stream of the whole tensor."""
    assert begin % 8 == 0
    return begin // 8 * bits
