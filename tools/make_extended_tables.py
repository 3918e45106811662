"""Optimal 5..8-bit tables for the 13 continuous activations -> fewbit_b200/data/extended.npz.

The reference ships bits 1..4 only (tools/quantize-builtins.sh:8); the kernels here take up to
8 bits, so `fewbit.GELU(bits=6)` needs tables.  Solver: fewbit_b200/quantize.py (same objective
as fewbit/approx.py).  Also reports, per table, whether the kernels' LUT bucketizer separates its
borders (one border per cell at most; otherwise the block falls back to the exact binary search).

    python tools/make_extended_tables.py [--bits 5,6,7,8]
"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from fewbit_b200 import quantize  # noqa: E402
from fewbit_b200.functional.activations import CONTINOUS  # noqa: E402

# ops.cuh, Bucketizer<T, B, false>::kCells: fp32 (byte table + borders) and bf16 (one word per cell)
CELLS_F32 = {3: 128, 4: 256, 5: 512, 6: 2048, 7: 2048, 8: 4096}
CELLS_BF16 = {b: 16 << b for b in range(3, 9)}
CELLS_BF16[6] = 2048


def bf16_round(a):
    import torch
    return torch.tensor(a, dtype=torch.float32).to(torch.bfloat16).float().numpy().astype(np.float64)


def crowded(borders, bits):
    """Would a kernel fall back to its exact search on this table (either dtype)?"""
    for inner, cells in ((borders[1:-1].astype(np.float32).astype(np.float64), CELLS_F32[bits]),
                         (bf16_round(borders[1:-1]), CELLS_BF16[bits])):
        lo, hi = inner[0], inner[-1]
        cell = np.rint(np.clip((inner - lo) / (hi - lo), 0, 1) * (cells - 1))
        if np.any(np.diff(cell) == 0):
            return True
    return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--bits', default='5,6,7,8')
    ap.add_argument('--functions', default=None, help='comma list: recompute only these, keep the other tables of the file')
    args = ap.parse_args()
    out = ROOT / 'fewbit_b200' / 'data' / 'extended.npz'
    tables = dict(np.load(out)) if args.functions and out.exists() else {}
    for name in (args.functions.split(',') if args.functions else CONTINOUS):
        for bits in (int(b) for b in args.bits.split(',')):
            borders, levels, err = quantize.optimal_table(name, bits)
            note = ''
            if crowded(borders, bits):
                # heavy-tailed derivative: confine the interior borders until the kernels' cell look-up
                # separates them (a fast table slightly off the optimum beats an optimal slow one)
                for reach in (48, 40, 32, 28, 24, 20, 18, 16, 14, 12, 10, 8, 6):
                    b2, l2, e2 = quantize.optimal_table(name, bits, reach=float(reach))
                    if not crowded(b2, bits):
                        note = f'  (interior borders confined to |x| <= {reach}: error x{e2 / err:.3f})'
                        borders, levels, err = b2, l2, e2
                        break
            tables[f'{name}{bits:02d}-borders'], tables[f'{name}{bits:02d}-levels'] = borders, levels
            print(f'{name:11s} {bits} bits  error {err:.4e}  borders [{borders[1]:+.3f}, {borders[-2]:+.3f}]  '
                  f'min gap {np.diff(borders).min():.2e}  crowded {crowded(borders, bits)}{note}', flush=True)
    np.savez_compressed(out, **tables)
    print(f'wrote {len(tables)} arrays to {out}')


if __name__ == '__main__':
    main()
