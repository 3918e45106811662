"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

Runs the UNMODIFIED reference CPU ops (oracle/_ref/libfewbit_ref.so, built from
/root/reference/fewbit/fewbit.cc + fewbit/cpu/gelu.cc by oracle/Makefile) in a process of
their own.  A separate process is required because the reference registers the torch op
namespace ``fewbit`` -- the very namespace the product library registers.

    python oracle/ref_runner.py run   in.npz out.npz
        in : x, bounds, levels, g  (fp32; or uint16 bf16 bit patterns with bf16=1)
        out: y, state, gin         via torch.ops.fewbit.quantize / quantize_backward
                                   (fewbit/cpu/gelu.cc:7-31, 33-45)
    python oracle/ref_runner.py bench --n N --bits B --repeats R [--threads T]
        prints one JSON line with the median seconds of quantize and quantize_backward.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
LIB = HERE / '_ref' / 'libfewbit_ref.so'


def load():
    if not LIB.exists():
        raise SystemExit(f'{LIB} is missing: run `make -C oracle ref` where /root/reference exists')
    torch.ops.load_library(str(LIB))


def _to_torch(a: np.ndarray, bf16: bool) -> torch.Tensor:
    if bf16:
        return torch.from_numpy(a.astype(np.uint16).view(np.int16).copy()).view(torch.bfloat16)
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def _to_numpy(t: torch.Tensor) -> np.ndarray:
    if t.dtype == torch.bfloat16:
        return t.contiguous().view(torch.int16).numpy().view(np.uint16)
    return t.contiguous().numpy()


def cmd_run(args):
    load()
    with np.load(args.inp) as npz:
        bf16 = bool(int(npz['bf16'])) if 'bf16' in npz else False
        x = _to_torch(npz['x'], bf16)
        bounds = _to_torch(npz['bounds'], bf16)
        levels = _to_torch(npz['levels'], bf16)
        g = _to_torch(npz['g'], bf16)
    y, state = torch.ops.fewbit.quantize(x, bounds)
    gin = torch.ops.fewbit.quantize_backward(g, state, levels)
    np.savez(args.out, y=_to_numpy(y), state=state.numpy(), gin=_to_numpy(gin))


def synthetic_table(bits: int):
    """Same synthetic table recipe as bench.py (kept dependency-free on purpose)."""
    from statistics import NormalDist
    nd = NormalDist(0.0, 1.5)
    nb = (1 << bits) - 1
    bounds = np.array([nd.inv_cdf((i + 1) / (nb + 1)) for i in range(nb)], np.float32)
    levels = np.linspace(0.0, 1.0, nb + 1).astype(np.float32)
    return torch.from_numpy(bounds), torch.from_numpy(levels)


def cmd_bench(args):
    load()
    if args.threads:
        torch.set_num_threads(args.threads)
    torch.manual_seed(0)
    dtype = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    x = (torch.randn(args.n) * 2).to(dtype)
    g = torch.randn(args.n).to(dtype)
    bounds, levels = synthetic_table(args.bits)
    bounds, levels = bounds.to(dtype), levels.to(dtype)
    tf, tb = [], []
    for it in range(args.repeats + 1):
        t0 = time.perf_counter()
        _, state = torch.ops.fewbit.quantize(x, bounds)
        t1 = time.perf_counter()
        torch.ops.fewbit.quantize_backward(g, state, levels)
        t2 = time.perf_counter()
        if it:  # first pass is warm-up
            tf.append(t1 - t0)
            tb.append(t2 - t1)
    print(json.dumps({'n': args.n, 'bits': args.bits, 'dtype': args.dtype,
                      'threads': torch.get_num_threads(),
                      'fwd_s': float(np.median(tf)), 'bwd_s': float(np.median(tb))}))


def main(argv=None):
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest='cmd', required=True)
    r = sub.add_parser('run')
    r.add_argument('inp')
    r.add_argument('out')
    r.set_defaults(fn=cmd_run)
    b = sub.add_parser('bench')
    b.add_argument('--n', type=int, default=1 << 22)
    b.add_argument('--bits', type=int, default=3)
    b.add_argument('--repeats', type=int, default=3)
    b.add_argument('--threads', type=int, default=0)
    b.add_argument('--dtype', default='f32')
    b.set_defaults(fn=cmd_bench)
    args = ap.parse_args(argv)
    args.fn(args)


if __name__ == '__main__':
    sys.exit(main())
