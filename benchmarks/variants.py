"""A/B runs of kernel builds: forward GB/s of a fixed set of cells for every library given.

    make -C fewbit_b200/csrc kernels VARIANT=_x EXTRA_NVCCFLAGS=-DFEWBIT_...   # build a variant
    python benchmarks/variants.py fewbit_b200/libfewbit_b200.so fewbit_b200/libfewbit_b200_x.so ...

One child process per library (FEWBIT_B200_LIBRARY); algorithmic bytes / CUDA-event time over 12
back-to-back launches on four rotating buffer sets, median of 5 (the same convention as
benchmarks/sweep_functions.py); the fraction printed is of MEASURED_PEAKS.json's hbm_gbs.
"""
import json
import os
import statistics
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CELLS = os.environ.get('VARIANT_CELLS', 'gelu:bf16:3,gelu:bf16:5,gelu:bf16:7,gelu:bf16:8,hardswish:bf16:3,hardswish:bf16:7,'
                       'tanh:bf16:3,tanhshrink:bf16:3,selu:bf16:3,softplus:bf16:3,silu:bf16:3,gelu:bf16:1,gelu:bf16:2,'
                       'gelu:f32:3,softplus:f32:3,gelu:f32:7').split(',')


def child():
    import torch
    sys.path.insert(0, str(ROOT))
    from fewbit_b200 import native
    from fewbit_b200.functional import store
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    n = 128 * 128 * 3072
    out = {}
    bufs = {}
    for cell in CELLS:
        name, tag, bits = cell.split(':')
        bits = int(bits)
        dtype = torch.bfloat16 if tag == 'bf16' else torch.float32
        if tag not in bufs:
            bufs.clear()
            bufs[tag] = ([(torch.randn(n, device=dev) * 2).to(dtype) for _ in range(4)],
                         [torch.empty(n, dtype=dtype, device=dev) for _ in range(4)])
        xs, ys = bufs[tag]
        borders, _ = store.get(name, bits, dev, dtype)
        bounds = borders[1:-1].contiguous()
        states = [native.new_state(xs[0], bits) for _ in range(4)]
        it = [0]

        def fwd():
            k = it[0] % 4
            it[0] += 1
            native.stepwise_forward(name, xs[k], ys[k], states[k], bits, bounds)

        for _ in range(3):
            fwd()
        torch.cuda.synchronize()
        side, graph = torch.cuda.Stream(), torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            fwd()
            side.synchronize()
            with torch.cuda.graph(graph, stream=side):      # host launch cost stays out of the number
                for _ in range(12):
                    fwd()
        torch.cuda.synchronize()
        graph.replay()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            graph.replay()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / 12)
        out[cell] = (n * 2 * xs[0].element_size() + n * bits // 8) / statistics.median(ts) / 1e6
    print(json.dumps(out))


def main():
    peak = 6548.5
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        peak = json.loads(p.read_text()).get('hbm_gbs', peak)
    libs = sys.argv[1:] or [str(ROOT / 'fewbit_b200' / 'libfewbit_b200.so')]
    print(f'{"library":34s} ' + ' '.join(f'{c.replace("hardswish", "hsw").replace("tanhshrink", "tshr").replace("softplus", "splus"):>12s}' for c in CELLS))
    for lib in libs:
        env = dict(os.environ, FEWBIT_B200_LIBRARY=lib, VARIANT_CHILD='1')
        r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
        try:
            vals = json.loads(line)
            txt = ' '.join(f'{vals[c]:7.0f}({vals[c] / peak:4.0%})' for c in CELLS)
        except Exception:  # noqa: BLE001
            txt = line
        print(f'{Path(lib).name:34s} {txt}', flush=True)


if __name__ == '__main__':
    child() if os.environ.get('VARIANT_CHILD') else main()
