"""Timeline of one CTA of the projection kernel (set FEWBIT_B200_SKETCH_TRACE=1; diagnostics)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from fewbit_b200 import native  # noqa: E402

tokens, rows = 16384, 3276
shapes = [int(a) for a in sys.argv[1].split(',')] if len(sys.argv) > 1 else [768, 3072]
kinds = sys.argv[2].split(',') if len(sys.argv) > 2 else ['gaussian', 'rademacher']
for features in shapes:
    x = torch.randn(tokens, features, device='cuda').to(torch.bfloat16)
    for kind in kinds:
        print(features, kind, flush=True)
        for _ in range(3):
            native.sketch_forward(x, rows, 1, 0, kind, 1.0 / rows)
        torch.cuda.synchronize()
