"""Projection kernel at RoBERTa-base shapes: N = 16384 tokens, P = 3276,
features 768 / 3072.  Prints time, TFLOP/s (2 P N D) against the measured bf16 peak, and the
reference's way of computing the same thing in torch (randn + matmul) for comparison."""
import json
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from fewbit_b200 import native  # noqa: E402


def timed(fn, reps=10, rounds=5):
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
    return statistics.median(ts)


def main():
    dev = 'cuda:0'
    peaks = {}
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        peaks = json.loads(p.read_text())
    peak = peaks.get('bf16_tflops', 1590.0)
    tokens, rows = 16384, 3276
    out = {}
    for features in (768, 3072):
        x = torch.randn(tokens, features, device=dev).to(torch.bfloat16)
        flops = 2.0 * rows * tokens * features
        for kind in ('gaussian', 'rademacher'):
            ms = timed(lambda: native.sketch_forward(x, rows, 1, 0, kind, 1.0 / rows))
            out[f'sketch_{kind}_D{features}'] = {'ms': ms, 'TFLOPs': flops / ms / 1e9,
                                                  'frac_of_bf16_peak': flops / ms / 1e9 / peak}
        xf = x.float()

        def torch_fp32():
            s = torch.randn(rows, tokens, device=dev)
            return (s @ xf) / rows

        def torch_bf16():
            s = torch.randn(rows, tokens, device=dev, dtype=torch.bfloat16)
            return (s @ x) / rows

        out[f'torch_randn_matmul_fp32_D{features}'] = {'ms': timed(torch_fp32, reps=3, rounds=3)}
        out[f'torch_randn_matmul_bf16_D{features}'] = {'ms': timed(torch_bf16, reps=5, rounds=3)}
        s = torch.randn(rows, tokens, device=dev, dtype=torch.bfloat16)
        out[f'cublas_bf16_matmul_only_D{features}'] = {'ms': timed(lambda: s @ x)}
    out['peak_bf16_tflops'] = peak
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
