"""Launch-bound regime: time per call at small tensor sizes next to torch's own elementwise kernels
(CUDA events, 200 back-to-back calls, median of 5).  Below ~4 M elements every column is the host's
call rate, not kernel time: ~8.5 us per call through the ctypes wrapper used here (fewbit_b200/native.py),
~5 us through torch's dispatcher for the torch columns; the kernels themselves take 2-3 us there."""
import statistics
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from fewbit_b200 import native  # noqa: E402
from fewbit_b200.functional import store  # noqa: E402

dev = torch.device('cuda:0')


def timed(fn, reps=200, rounds=5):
    for _ in range(20):
        fn()
    ts = []
    for _ in range(rounds):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps * 1e3)
    return statistics.median(ts)


print(f'{"n":>10s} {"dtype":5s} {"gelu3 fwd":>10s} {"gelu3 bwd":>10s} {"relu fwd":>9s} {"relu bwd":>9s} {"F.gelu":>8s} {"F.relu":>8s} {"copy":>7s}   (us per call)')
for dtype in (torch.float32, torch.bfloat16):
    borders, levels = store.get('gelu', 3, dev, dtype)
    bounds = borders[1:-1].contiguous()
    for n in (1 << 12, 1 << 16, 1 << 20, 1 << 22, 1 << 24):
        x = (torch.randn(n, device=dev) * 2).to(dtype)
        g = torch.randn(n, device=dev).to(dtype)
        y, gin = torch.empty_like(x), torch.empty_like(g)
        s3, s1 = native.new_state(x, 3), native.new_state(x, 1)
        row = [timed(lambda: native.stepwise_forward('gelu', x, y, s3, 3, bounds)),
               timed(lambda: native.stepwise_backward(s3, g, gin, 3, levels)),
               timed(lambda: native.piecewise_forward('relu', x, y, s1)),
               timed(lambda: native.piecewise_backward('relu', s1, g, gin)),
               timed(lambda: F.gelu(x)), timed(lambda: torch.relu(x)), timed(lambda: y.copy_(x))]
        print(f'{n:10d} {str(dtype)[6:]:5s} ' + ' '.join(f'{v:9.2f}' for v in row))
