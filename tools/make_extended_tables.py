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

CELLS = {3: 128, 4: 256, 5: 512, 6: 2048, 7: 2048, 8: 2048}     # ops.cuh: Bucketizer::kCells


def crowded(borders, bits):
    inner = borders[1:-1].astype(np.float32).astype(np.float64)
    lo, hi = inner[0], inner[-1]
    cell = np.rint(np.clip((inner - lo) / (hi - lo), 0, 1) * (CELLS[bits] - 1))
    return bool(np.any(np.diff(cell) == 0))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--bits', default='5,6,7,8')
    args = ap.parse_args()
    out = ROOT / 'fewbit_b200' / 'data' / 'extended.npz'
    tables = {}
    for name in CONTINOUS:
        for bits in (int(b) for b in args.bits.split(',')):
            borders, levels, err = quantize.optimal_table(name, bits)
            tables[f'{name}{bits:02d}-borders'], tables[f'{name}{bits:02d}-levels'] = borders, levels
            print(f'{name:11s} {bits} bits  error {err:.4e}  borders [{borders[1]:+.3f}, {borders[-2]:+.3f}]  '
                  f'min gap {np.diff(borders).min():.2e}  crowded {crowded(borders, bits)}', flush=True)
    np.savez_compressed(out, **tables)
    print(f'wrote {len(tables)} arrays to {out}')


if __name__ == '__main__':
    main()
