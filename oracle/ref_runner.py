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
    python oracle/ref_runner.py cuda in.npz out.npz          (needs a GPU)
        runs the reference's CUDA operators (oracle/_ref/libfewbit_ref_cuda.so = unmodified
        fewbit/cuda/codec.cu + activation.cc built for sm_100) on every case of in.npz and
        stores y, the codes its kernel chose and its gradient.
    python oracle/ref_runner.py cuda-bench                   (needs a GPU)
        times the reference CUDA kernels on the benchmark shapes (fp32 only: it has no bf16).
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
LIB_CUDA = HERE / '_ref' / 'libfewbit_ref_cuda.so'


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


def load_cuda():
    if not LIB_CUDA.exists():
        raise SystemExit(f'{LIB_CUDA} is missing: run `make -C oracle ref` where /root/reference exists')
    torch.ops.load_library(str(LIB_CUDA))


def padded_view(bounds: torch.Tensor, bits: int) -> torch.Tensor:
    """The reference searches 2^(bits+1) - 1 entries (nobits = bits + 1) and so
    reads past the bounds it is given; it only works on a *view* followed by large sentinels
    (as borders[1:-1] of the store is).  Give it exactly that."""
    pad = torch.full((2 << bits, ), float('inf'), device=bounds.device, dtype=bounds.dtype)
    return torch.cat([bounds, pad])[:bounds.numel()]


def cmd_cuda(args):
    load_cuda()
    dev = torch.device('cuda:0')
    out = {}
    with np.load(args.inp, allow_pickle=False) as npz:
        keys = sorted({k.split('/')[0] for k in npz.keys()})
        for key in keys:
            name = str(npz[f'{key}/name'])
            params = [float(v) for v in npz[f'{key}/params']]
            x = torch.from_numpy(npz[f'{key}/x']).to(dev)
            g = torch.from_numpy(npz[f'{key}/g']).to(dev)
            assert x.numel() % 1024 == 0, 'reference writes out of bounds on ragged sizes (App. C-7)'
            op = getattr(torch.ops.fewbit, name)
            if f'{key}/bounds' in npz:
                bounds = torch.from_numpy(npz[f'{key}/bounds']).to(dev)
                levels = torch.from_numpy(npz[f'{key}/levels']).to(dev)
                bits = max(1, int(np.ceil(np.log2(levels.numel()))))
                view = padded_view(bounds, bits)
                leaf = x.clone().requires_grad_()
                y = op(leaf * 1.0, view, levels, *params)
                y.backward(g)
                # codes: levels = 0, 1, 2, ... and g = 1 make the gradient equal to the code
                probe = x.clone().requires_grad_()
                ramp = torch.arange(levels.numel(), device=dev, dtype=torch.float32)
                op(probe * 1.0, view, ramp, *params).backward(torch.ones_like(x))
                out[f'{key}/codes'] = probe.grad.cpu().numpy().astype(np.int32)
            else:
                leaf = x.clone().requires_grad_()
                y = op(leaf * 1.0, *params)
                y.backward(g)
            out[f'{key}/y'] = y.detach().cpu().numpy()
            out[f'{key}/gin'] = leaf.grad.cpu().numpy()
    torch.cuda.synchronize()
    np.savez(args.out, **out)


def cmd_cuda_bench(args):
    load_cuda()
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    res = {}

    def timed(fn, reps=10, rounds=5):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(rounds):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(torch.cuda.default_stream())
            for _ in range(reps):
                fn()
            b.record(torch.cuda.default_stream())   # the reference launches on the legacy stream
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / reps)
        return float(np.median(ts))

    from statistics import NormalDist  # noqa: F401
    n = 128 * 128 * 3072
    bounds = torch.tensor([-2.41658115, -0.710008025, -0.325840563, 1.06942185e-04, 0.326057166,
                           0.710240841, 2.41447878], device=dev)
    levels = torch.tensor([-1.9399123e-04, -8.8279128e-02, 0.12568383, 0.37231442, 0.62785137,
                           0.87445050, 1.0883480, 1.0001949], device=dev)
    view = padded_view(bounds, 3)
    x = torch.randn(n, device=dev) * 2
    g = torch.randn(n, device=dev)
    with torch.no_grad():
        ms = timed(lambda: torch.ops.fewbit.gelu(x, view, levels))
    # the reference packs bits+1 = 4 bits per element: count the bytes it really moves
    res['gelu3_f32_fwd'] = {'ms': ms, 'GBps_algorithmic_3bit': n * (8 + 3 / 8) / ms / 1e6,
                            'GBps_own_4bit': n * (8 + 4 / 8) / ms / 1e6}
    leaf = x.clone().requires_grad_()
    y = torch.ops.fewbit.gelu(leaf * 1.0, view, levels)
    ms = timed(lambda: torch.autograd.grad(y, leaf, g, retain_graph=True))
    res['gelu3_f32_bwd'] = {'ms': ms, 'GBps_algorithmic_3bit': n * (8 + 3 / 8) / ms / 1e6,
                            'note': 'autograd.grad: kernel + empty_like + mul by 1.0 upstream'}
    n = 1 << 28                                  # 1 GiB of fp32 (the reference has no bf16 path)
    x = torch.randn(n, device=dev) * 2
    with torch.no_grad():
        ms = timed(lambda: torch.ops.fewbit.relu(x))
    res['relu_f32_1GiB_fwd'] = {'ms': ms, 'GBps_algorithmic': n * (8 + 1 / 8) / ms / 1e6}
    print(json.dumps(res))


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
    c = sub.add_parser('cuda')
    c.add_argument('inp')
    c.add_argument('out')
    c.set_defaults(fn=cmd_cuda)
    cb = sub.add_parser('cuda-bench')
    cb.set_defaults(fn=cmd_cuda_bench)
    args = ap.parse_args(argv)
    args.fn(args)


if __name__ == '__main__':
    sys.exit(main())
