"""Workload for compute-sanitizer: every kernel family on small, ragged and misaligned inputs.

    compute-sanitizer --tool memcheck  python benchmarks/sanitizer_workload.py
    compute-sanitizer --tool racecheck python benchmarks/sanitizer_workload.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from fewbit_b200 import native  # noqa: E402
from fewbit_b200.functional import make_table, store  # noqa: E402

dev = torch.device('cuda:0')
torch.manual_seed(0)
for dtype in (torch.float32, torch.bfloat16):
    for bits in (1, 3, 5, 8):
        for name in ('gelu', 'elu', 'softplus', 'tanh'):
            if bits <= 4:
                borders, levels = store.get(name, bits, dev, dtype)
            else:
                borders, levels = (t.to(dev, dtype) for t in make_table(name, bits))
            bounds, levels = borders[1:-1].contiguous(), levels.contiguous()
            for n in (1, 7, 255, 1023, 1025, 2061, 70001):
                for shift in (0, 1):
                    x = (torch.randn(n + shift, device=dev) * 2).to(dtype)[shift:]
                    g = torch.randn(n + shift, device=dev).to(dtype)[shift:]
                    y, gin = torch.empty_like(x), torch.empty_like(g)
                    state = native.new_state(x, bits)
                    native.stepwise_forward(name, x, y, state, bits, bounds)
                    native.stepwise_backward(state, g, gin, bits, levels)
    for name, (p0, p1) in {'leaky_relu': (0.01, 0.0), 'softshrink': (0.5, 0.0), 'hardtanh': (-1.0, 1.0)}.items():
        x = (torch.randn(70001, device=dev) * 2).to(dtype)
        y, gin = torch.empty_like(x), torch.empty_like(x)
        state = native.new_state(x, 1)
        native.piecewise_forward(name, x, y, state, p0, p1)
        native.piecewise_backward(name, state, x, gin, p0)
# projection kernel: plain, 1 x 2 push cluster (two feature tiles that are not a pair, D = 512), CTA pair (D = 768 and
# 1536, both kinds), and the variant that rounds to bf16 and appends the column sums
for tokens, features, rows, kind in ((1000, 72, 50, 'gaussian'), (4100, 512, 161, 'rademacher'), (1300, 512, 70, 'gaussian'),
                                     (2048, 768, 333, 'gaussian'), (700, 1536, 40, 'gaussian'), (4100, 768, 161, 'rademacher')):
    x = torch.randn(tokens, features, device=dev).to(torch.bfloat16)
    native.sketch_forward(x, rows, 7, 3, kind, 1.0 / rows)
    native.sketch_project(x, rows, 7, 3, kind, 1.0 / rows, torch.bfloat16, column_sums=True)
torch.cuda.synchronize()
print('workload done')
