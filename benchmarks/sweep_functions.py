"""BASELINE.json configs[2]: bits sweep across all continuous activations and the 1-bit family,
fp32 and bf16, forward and backward, on a (128, 128, 3072) tensor per GPU.

    python benchmarks/sweep_functions.py [--bits 1,3,4,8] [--functions gelu,silu] [--dtypes bf16]
                                         [--json out.json] [--md out.md]

Under torchrun (`python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1
benchmarks/sweep_functions.py ...`) the tensor is batch-sharded: every rank sweeps the same
per-GPU shape on its own GPU at the same time (a barrier aligns each measurement, there is no
data-path collective), and rank 0 prints the aggregate GB/s = sum over ranks and the slowest
rank's fraction of peak.

GB/s = algorithmic bytes n (s + s + b/8) / CUDA-event time (12 back-to-back launches over four
rotating buffer sets so nothing is served from L2, replayed from a CUDA graph, median of 5 rounds).  Bits 1-4 use the
built-in tables, 5-8 the shipped optimal tables (`--tables synthetic`: `make_table`).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from fewbit_b200 import native  # noqa: E402
from fewbit_b200.functional import CONTINOUS, make_table, store  # noqa: E402

PIECEWISE = {'hardshrink': (0.5, 0.0), 'hardsigmoid': (0.0, 0.0), 'hardtanh': (-1.0, 1.0),
             'leaky_relu': (0.01, 0.0), 'relu': (0.0, 0.0), 'relu6': (0.0, 0.0),
             'softshrink': (0.5, 0.0), 'threshold': (1.0, 3.0)}


WORLD = int(os.environ.get('WORLD_SIZE', 1))
RANK = int(os.environ.get('RANK', 0))


def across_ranks(value):
    """(sum, min) of a per-rank number; identity on one GPU."""
    if WORLD == 1:
        return value, value
    import torch.distributed as dist
    t = torch.tensor([value], device='cuda', dtype=torch.float64)
    total, low = t.clone(), t.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM)
    dist.all_reduce(low, op=dist.ReduceOp.MIN)
    return total.item(), low.item()


def timed(fn, reps=12, rounds=5):
    """ms per launch: `reps` back-to-back launches captured in one CUDA graph (so that the host's
    launch cost, which varies from box to box and approaches the ~40 us of a kernel, stays out of the
    number), replayed `rounds` times between CUDA events; median."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn()
        side.synchronize()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    if WORLD > 1:
        torch.distributed.barrier()
    ts = []
    graph.replay()
    for _ in range(rounds):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    return statistics.median(ts)


def record(rows, name, tag, bits, f, b, peak):
    """One row: whole-job GB/s (sum over ranks) and the slowest rank's fraction of one GPU's peak."""
    (f_sum, f_min), (b_sum, b_min) = across_ranks(f), across_ranks(b)
    rows.append({'function': name, 'dtype': tag, 'bits': bits, 'fwd_GBps': f_sum, 'bwd_GBps': b_sum,
                 'fwd_frac': f_min / peak, 'bwd_frac': b_min / peak, 'n_gpus': WORLD})
    if RANK == 0:
        print(f'{name:11s} {tag:4s} b={bits}  fwd {f_sum:7.0f} GB/s ({f_min / peak:4.0%})  '
              f'bwd {b_sum:7.0f} GB/s ({b_min / peak:4.0%})', flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--bits', default='1,2,3,4,5,6,7,8')
    ap.add_argument('--functions', default=None, help='comma list (default: all)')
    ap.add_argument('--dtypes', default='f32,bf16')
    ap.add_argument('--json', default=None)
    ap.add_argument('--md', default=None)
    ap.add_argument('--tables', default='shipped', choices=('shipped', 'synthetic'),
                    help='bits 5-8: the tables the package ships (data/extended.npz) or make_table() quantile grids')
    args = ap.parse_args()
    bits_list = [int(b) for b in args.bits.split(',')]
    only = set(args.functions.split(',')) if args.functions else None
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0)))
    torch.cuda.set_device(dev)
    if WORLD > 1:
        torch.distributed.init_process_group('nccl', device_id=dev)
    torch.manual_seed(RANK)
    n = 128 * 128 * 3072
    peak = 6548.5
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        peak = json.loads(p.read_text()).get('hbm_gbs', peak)
    rows, copies = [], {}
    for tag, dtype, es in (('f32', torch.float32, 4), ('bf16', torch.bfloat16, 2)):
        if tag not in args.dtypes.split(','):
            continue
        xs = [(torch.randn(n, device=dev) * 2).to(dtype) for _ in range(4)]
        ys = [torch.empty(n, dtype=dtype, device=dev) for _ in range(4)]
        it = [0]

        def plain_copy():     # the same bytes through torch's copy kernel: the ceiling of this shape
            k = it[0] % 4
            it[0] += 1
            ys[k].copy_(xs[k])

        c_sum, c_min = across_ranks(n * 2 * es / timed(plain_copy) / 1e6)
        copies[tag] = {'GBps': c_sum, 'frac': c_min / peak}
        if RANK == 0:
            print(f'{"torch copy":11s} {tag:4s}      {c_sum:7.0f} GB/s ({c_min / peak:4.0%})   <- same bytes, same shape', flush=True)
        for name in CONTINOUS:
            if only and name not in only:
                continue
            for bits in bits_list:
                if bits <= 4 or args.tables == 'shipped':
                    borders, levels = store.get(name, bits, dev, dtype)
                else:
                    borders, levels = (t.to(dev, dtype) for t in make_table(name, bits))
                bounds, levels = borders[1:-1].contiguous(), levels.contiguous()
                states = [native.new_state(xs[0], bits) for _ in range(4)]
                nbytes = n * 2 * es + n * bits // 8

                def fwd():
                    k = it[0] % 4
                    it[0] += 1
                    native.stepwise_forward(name, xs[k], ys[k], states[k], bits, bounds)

                def bwd():
                    k = it[0] % 4
                    it[0] += 1
                    native.stepwise_backward(states[k], xs[k], ys[k], bits, levels)

                record(rows, name, tag, bits, nbytes / timed(fwd) / 1e6, nbytes / timed(bwd) / 1e6, peak)
        states = [native.new_state(xs[0], 1) for _ in range(4)]
        for name, (p0, p1) in PIECEWISE.items():
            if only and name not in only:
                continue
            nbytes = n * 2 * es + n // 8

            def fwd():
                k = it[0] % 4
                it[0] += 1
                native.piecewise_forward(name, xs[k], ys[k], states[k], p0, p1)

            def bwd():
                k = it[0] % 4
                it[0] += 1
                native.piecewise_backward(name, states[k], xs[k], ys[k], p0)

            record(rows, name, tag, 1, nbytes / timed(fwd) / 1e6, nbytes / timed(bwd) / 1e6, peak)
        del xs, ys
    if WORLD > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if RANK != 0:
        return
    if args.json:
        Path(args.json).write_text(json.dumps({'peak_GBps': peak, 'elements_per_gpu': n, 'n_gpus': WORLD,
                                               'torch_copy_same_shape': copies, 'rows': rows}, indent=1))
    if args.md:
        lines = [f'# Function sweep, {n} elements (128x128x3072) per GPU x {WORLD} B200, GB/s summed over GPUs, '
                 f'fraction = slowest GPU / {peak:.0f} GB/s (measured copy)', '',
                 'A plain `torch` copy of the same tensors (same bytes in flight, same launch ramp and tail): '
                 + ', '.join(f"{t} {c['GBps']:.0f} GB/s ({c['frac']:.0%})" for t, c in copies.items()), '',
                 '| function | dtype | bits | fwd GB/s | of peak | bwd GB/s | of peak |', '|---|---|---|---|---|---|---|']
        lines += [f"| {r['function']} | {r['dtype']} | {r['bits']} | {r['fwd_GBps']:.0f} | {r['fwd_frac']:.0%} | "
                  f"{r['bwd_GBps']:.0f} | {r['bwd_frac']:.0%} |" for r in rows]
        Path(args.md).write_text('\n'.join(lines) + '\n')


if __name__ == '__main__':
    main()
