"""Repeat the projection kernel on CTA-pair shapes and compare every result with a dense product of the
same S: a soak test for the cross-CTA hand-over (barriers, TMEM allocation) of pair mode."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from fewbit_b200 import native  # noqa: E402

dev = 'cuda:0'
torch.manual_seed(0)
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 300
worst = 0.0
shapes = [(2048, 768, 333), (16384, 768, 3276), (700, 1536, 40), (4100, 3072, 161), (130, 2304, 17)]
refs = {}
for i in range(rounds):
    tokens, features, rows = shapes[i % len(shapes)]
    kind = 'gaussian' if i % 7 else 'rademacher'
    key = (tokens, features)
    if key not in refs:
        refs[key] = torch.randn(tokens, features, device=dev).to(torch.bfloat16)
    x = refs[key]
    out = native.sketch_forward(x, rows, 11, i, kind, 1.0 / rows)
    s = native.sketch_matrix(rows, tokens, 11, i, kind, dev).float()
    want = (s @ x.float()) / rows
    err = ((out - want).norm() / want.norm()).item()
    worst = max(worst, err)
    assert err < 2e-3 and torch.isfinite(out).all(), (i, tokens, features, rows, kind, err)
torch.cuda.synchronize()
print(f'{rounds} launches ok, worst relative error {worst:.2e}')
