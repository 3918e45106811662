"""Run-time knobs of the projection kernel, one process: python benchmarks/sketch_sweep.py "K=V,K=V" "K=V" ...
Each argument is one configuration of FEWBIT_B200_SKETCH_* variables (without the prefix; "-" = defaults):
SLOTS (S ring entries), BN, SPLITK, PAIR (0: no CTA pairs), CLUSTER (Cy without cta_group::2), DEBUG (1: no S
generation, 4: no MMAs; timing only).  SWEEP_FEATURES / SWEEP_ROWS choose the shapes.
Times fewbit_sketch_forward (N = 16384, P = 3276, D = 768 and 3072, both kinds) with output and workspace
preallocated, 10 back-to-back calls, median of 5, and checks every result against a matmul with the materialised S."""
import os
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from fewbit_b200 import native  # noqa: E402

PREFIX = 'FEWBIT_B200_SKETCH_'
tokens, rows = 16384, int(os.environ.get('SWEEP_ROWS', '3276'))
dev = 'cuda:0'
shapes = [int(v) for v in os.environ.get('SWEEP_FEATURES', '768,3072').split(',')]
xs = {d: torch.randn(tokens, d, device=dev).to(torch.bfloat16) for d in shapes}
smats = {k: native.sketch_matrix(rows, tokens, 1, 0, k, dev).float() for k in ('gaussian', 'rademacher')}
wants = {(d, k): (smats[k] @ xs[d].float()) / rows for d in shapes for k in smats}

for config in sys.argv[1:] or ['-']:
    for key in [k for k in os.environ if k.startswith(PREFIX)]:
        del os.environ[key]
    if config != '-':
        for item in config.split(','):
            k, v = item.split('=')
            os.environ[PREFIX + k] = v
    cells = []
    for d in shapes:
        x = xs[d]
        res = torch.empty(rows, d, dtype=torch.float32, device=dev)
        ws = native.sketch_workspace(x, rows)
        for kind in ('gaussian', 'rademacher'):
            def fn():
                native.sketch_forward(x, rows, 1, 0, kind, 1.0 / rows, out=res, workspace=ws)
            for _ in range(3):
                fn()
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(10):
                    fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) / 10)
            want = wants[(d, kind)]
            err = ((res - want).norm() / want.norm()).item()
            cells.append(f'{kind[0]}{d} {statistics.median(ts) * 1e3:6.1f} us (err {err:.1e})')
    print(f'{config:28s} ' + '  '.join(cells), flush=True)
