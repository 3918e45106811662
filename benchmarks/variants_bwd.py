"""A/B runs of kernel builds, backward kernels: python benchmarks/variants_bwd.py lib.so [lib2.so ...]
(one child process per library, FEWBIT_B200_LIBRARY).  GB/s of the levels backward (unpack + multiply) at the given
(dtype, bits) cells and of the ReLU mask backward, 128 x 128 x 3072 elements, 12 back-to-back launches on four buffer
sets replayed from a CUDA graph, median of 5; each result is also compared with levels[codes] * g computed by torch."""
import json
import os
import statistics
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CELLS = os.environ.get('VARIANT_CELLS', 'bf16:1,bf16:3,bf16:5,bf16:7,bf16:8,f32:3,f32:7,relu:bf16,relu:f32').split(',')


def child():
    import torch
    sys.path.insert(0, str(ROOT))
    from fewbit_b200 import native
    from fewbit_b200.functional import store
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    n = 128 * 128 * 3072
    out = {}
    for cell in CELLS:
        relu = cell.startswith('relu')
        tag, bits = (cell.split(':')[1], 1) if relu else (cell.split(':')[0], int(cell.split(':')[1]))
        dtype = torch.bfloat16 if tag == 'bf16' else torch.float32
        es = 2 if tag == 'bf16' else 4
        xs = [(torch.randn(n, device=dev) * 2).to(dtype) for _ in range(4)]
        gs = [torch.randn(n, device=dev).to(dtype) for _ in range(4)]
        gins = [torch.empty_like(t) for t in gs]
        states = [native.new_state(xs[0], bits) for _ in range(4)]
        if relu:
            for k in range(4):
                native.piecewise_forward('relu', xs[k], torch.empty_like(xs[k]), states[k], 0.0, 0.0)
            run = lambda k: native.piecewise_backward('relu', states[k], gs[k], gins[k], 0.0)    # noqa: E731
            want = torch.where(xs[0] > 0, gs[0], torch.zeros_like(gs[0]))
        else:
            borders, levels = store.get('gelu', bits, dev, dtype)
            bounds = borders[1:-1].contiguous()
            for k in range(4):
                native.stepwise_forward('gelu', xs[k].clone(), torch.empty_like(xs[k]), states[k], bits, bounds)
            run = lambda k: native.stepwise_backward(states[k], gs[k], gins[k], bits, levels)   # noqa: E731
            want = levels[torch.searchsorted(bounds.float(), xs[0].float(), right=False)] * gs[0]
        it = [0]

        def bwd():
            run(it[0] % 4)
            it[0] += 1
        for _ in range(4):
            bwd()
        torch.cuda.synchronize()
        ok = torch.equal(gins[0], want.to(dtype))
        side, graph = torch.cuda.Stream(), torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            bwd()
            side.synchronize()
            with torch.cuda.graph(graph, stream=side):
                for _ in range(12):
                    bwd()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            graph.replay()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / 12)
        gbs = n * (2 * es + bits / 8) / (statistics.median(ts) / 1e3) / 1e9
        out[cell] = f'{gbs:5.0f}({100 * gbs / 6548.5:3.0f}%){"" if ok else " WRONG"}'
        del xs, gs, gins, states
    print(json.dumps(out))


def main():
    for lib in sys.argv[1:] or [str(ROOT / 'fewbit_b200' / 'libfewbit_b200.so')]:
        env = dict(os.environ, FEWBIT_B200_LIBRARY=lib, VARIANT_CHILD='1')
        r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True, timeout=900)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-600:]
        try:
            d = json.loads(line)
            line = '  '.join(f'{k} {v}' for k, v in d.items())
        except ValueError:
            pass
        print(f'{Path(lib).name:30s} {line}', flush=True)


if __name__ == '__main__':
    child() if os.environ.get('VARIANT_CHILD') else main()
