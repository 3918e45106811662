"""BASELINE.json configs[2]: bits sweep across all continuous activations and the 1-bit family,
fp32 and bf16, forward and backward, on a (128, 128, 3072) tensor per GPU.

    python benchmarks/sweep_functions.py [--bits 1,3,4,8] [--json out.json] [--md out.md]

GB/s = algorithmic bytes n (s + s + b/8) / CUDA-event time (12 back-to-back launches over four
rotating buffer sets so nothing is served from L2, median of 5 rounds).  Bits 1-4 use the
built-in tables, 5-8 `make_table`.
"""
from __future__ import annotations

import argparse
import json
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


def timed(fn, reps=12, rounds=5):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(rounds):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--bits', default='1,2,3,4,5,6,7,8')
    ap.add_argument('--json', default=None)
    ap.add_argument('--md', default=None)
    args = ap.parse_args()
    bits_list = [int(b) for b in args.bits.split(',')]
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    n = 128 * 128 * 3072
    peak = 6548.5
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        peak = json.loads(p.read_text()).get('hbm_gbs', peak)
    rows = []
    for tag, dtype, es in (('f32', torch.float32, 4), ('bf16', torch.bfloat16, 2)):
        xs = [(torch.randn(n, device=dev) * 2).to(dtype) for _ in range(4)]
        ys = [torch.empty(n, dtype=dtype, device=dev) for _ in range(4)]
        it = [0]
        for name in CONTINOUS:
            for bits in bits_list:
                if bits <= 4:
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

                f, b = nbytes / timed(fwd) / 1e6, nbytes / timed(bwd) / 1e6
                rows.append({'function': name, 'dtype': tag, 'bits': bits, 'fwd_GBps': f, 'bwd_GBps': b,
                             'fwd_frac': f / peak, 'bwd_frac': b / peak})
                print(f'{name:11s} {tag:4s} b={bits}  fwd {f:7.0f} GB/s ({f / peak:4.0%})  bwd {b:7.0f} GB/s ({b / peak:4.0%})',
                      flush=True)
        states = [native.new_state(xs[0], 1) for _ in range(4)]
        for name, (p0, p1) in PIECEWISE.items():
            nbytes = n * 2 * es + n // 8

            def fwd():
                k = it[0] % 4
                it[0] += 1
                native.piecewise_forward(name, xs[k], ys[k], states[k], p0, p1)

            def bwd():
                k = it[0] % 4
                it[0] += 1
                native.piecewise_backward(name, states[k], xs[k], ys[k], p0)

            f, b = nbytes / timed(fwd) / 1e6, nbytes / timed(bwd) / 1e6
            rows.append({'function': name, 'dtype': tag, 'bits': 1, 'fwd_GBps': f, 'bwd_GBps': b,
                         'fwd_frac': f / peak, 'bwd_frac': b / peak})
            print(f'{name:11s} {tag:4s} b=1  fwd {f:7.0f} GB/s ({f / peak:4.0%})  bwd {b:7.0f} GB/s ({b / peak:4.0%})',
                  flush=True)
        del xs, ys
    if args.json:
        Path(args.json).write_text(json.dumps({'peak_GBps': peak, 'elements': n, 'rows': rows}, indent=1))
    if args.md:
        lines = [f'# Function sweep, {n} elements (128x128x3072), B200, peak {peak:.0f} GB/s (measured copy)', '',
                 '| function | dtype | bits | fwd GB/s | of peak | bwd GB/s | of peak |', '|---|---|---|---|---|---|---|']
        lines += [f"| {r['function']} | {r['dtype']} | {r['bits']} | {r['fwd_GBps']:.0f} | {r['fwd_frac']:.0%} | "
                  f"{r['bwd_GBps']:.0f} | {r['bwd_frac']:.0%} |" for r in rows]
        Path(args.md).write_text('\n'.join(lines) + '\n')


if __name__ == '__main__':
    main()
