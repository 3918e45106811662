"""A/B of projection-kernel builds: python benchmarks/sketch_variants.py lib.so [lib2.so ...]
(one child process per library, FEWBIT_B200_LIBRARY).  Times fewbit_sketch_forward with output and
workspace preallocated, 10 back-to-back calls, median of 5; checks one result against a matmul with
the materialised S so that a fast wrong variant cannot slip through."""
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
    dev = 'cuda:0'
    tokens, rows = 16384, 3276
    out = {}
    for features in (768, 3072):
        x = torch.randn(tokens, features, device=dev).to(torch.bfloat16)
        res = torch.empty(rows, features, dtype=torch.float32, device=dev)
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
            out[f'{kind[0]}{features}_us'] = statistics.median(ts) * 1e3
        s = native.sketch_matrix(rows, tokens, 1, 0, 'rademacher', dev).float()
        want = (s @ x.float()) / rows
        out[f'err{features}'] = ((res - want).norm() / want.norm()).item()
    print(json.dumps(out))


def main():
    for lib in sys.argv[1:] or [str(ROOT / 'fewbit_b200' / 'libfewbit_b200.so')]:
        env = dict(os.environ, FEWBIT_B200_LIBRARY=lib, SKETCH_CHILD='1')
        r = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True, timeout=600)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
        print(f'{Path(lib).name:34s} {line}', flush=True)


if __name__ == '__main__':
    child() if os.environ.get('SKETCH_CHILD') else main()
