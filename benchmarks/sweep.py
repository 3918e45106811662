"""Tuning sweep: time the hot kernels of one or more builds of libfewbit_b200 (variants built
with `make -C fewbit_b200/csrc kernels VARIANT=_x EXTRA_NVCCFLAGS=-D...`).

    python benchmarks/sweep.py [lib.so ...]          # one child process per library / setting
Prints GB/s (algorithmic bytes / CUDA-event time over 10 back-to-back launches, median of 5).
"""
import json
import os
import statistics
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def child():
    import torch
    sys.path.insert(0, str(ROOT))
    from fewbit_b200 import native
    from fewbit_b200.functional import store
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    out = {}

    def timed(fn, nbytes, reps=10, rounds=5):
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
        return nbytes / (statistics.median(ts) / 1e3) / 1e9

    n = 1 << 29
    x = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_(0, 2)
    g = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_()
    y, gin = torch.empty_like(x), torch.empty_like(g)
    state = native.new_state(x, 1)
    nb = n * 4 + n // 8
    out['torch_copy'] = timed(lambda: y.copy_(x), n * 4)
    out['relu_bf16_fwd'] = timed(lambda: native.piecewise_forward('relu', x, y, state), nb)
    out['relu_bf16_bwd'] = timed(lambda: native.piecewise_backward('relu', state, g, gin), nb)
    out['htanh_bf16_fwd'] = timed(lambda: native.piecewise_forward('hardtanh', x, y, state, -1.0, 1.0), nb)
    del x, g, y, gin, state
    # several independent 128x128x3072 buffers cycled so that nothing stays in L2
    n = 128 * 128 * 3072
    for tag, dtype, es in (('f32', torch.float32, 4), ('bf16', torch.bfloat16, 2)):
        borders, levels = store.get('gelu', 3, dev, dtype)
        bounds = borders[1:-1].contiguous()
        bufs = [((torch.randn(n, device=dev) * 2).to(dtype), torch.empty(n, dtype=dtype, device=dev),
                 torch.empty(native.state_bytes(n, 3), dtype=torch.uint8, device=dev)) for _ in range(4)]
        nb = n * 2 * es + n * 3 // 8
        it = [0]

        def fwd():
            xx, yy, ss = bufs[it[0] % 4]
            it[0] += 1
            native.stepwise_forward('gelu', xx, yy, ss, 3, bounds)

        def bwd():
            xx, yy, ss = bufs[it[0] % 4]
            it[0] += 1
            native.stepwise_backward(ss, xx, yy, 3, levels)

        out[f'gelu3_{tag}_fwd'] = timed(fwd, nb, reps=12)
        out[f'gelu3_{tag}_bwd'] = timed(bwd, nb, reps=12)
        del bufs
    print(json.dumps(out))


def main():
    libs = sys.argv[1:] or [str(ROOT / 'fewbit_b200' / 'libfewbit_b200.so')]
    settings = os.environ.get('SWEEP_CTAS', '0').split(',')
    for lib in libs:
        for ctas in settings:
            env = dict(os.environ, FEWBIT_B200_LIBRARY=lib, SWEEP_CHILD='1')
            if ctas != '0':
                env['FEWBIT_B200_CTAS_PER_SM'] = ctas
            r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
            try:
                vals = json.loads(line)
                txt = '  '.join(f'{k}={v:7.0f}' for k, v in vals.items())
            except Exception:  # noqa: BLE001
                txt = line
            print(f'{Path(lib).name:32s} ctas={ctas:>2s}  {txt}', flush=True)


if __name__ == '__main__':
    child() if os.environ.get('SWEEP_CHILD') else main()
