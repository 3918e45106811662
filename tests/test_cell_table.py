"""Host-side model of the bf16 bucket look-up of the forward kernels (csrc/ops.cuh,
`Bucketizer<__nv_bfloat16, B, false>`): one 32-bit word per cell that is at once a threshold --
a float in [border, next bf16 above border) -- and a carrier of k = #{borders in earlier cells}
replicated at every code position of a packed half.

The CUDA code is what ships; this numpy restatement of its table construction pins the argument
the kernel rests on, exhaustively over all 65536 bf16 bit patterns: for every table whose borders
the cells separate, `(word & mask) + (word < x)` equals the reference's bucket search
`#{i : bounds[i] < x}` (fewbit/cuda/codec.cu:118-131), NaN -> 0 included.  The GPU parity tests
(tests/test_gpu_parity.py) check the kernels themselves against the oracle.
"""
import numpy as np
import pytest
import torch

from fewbit_b200.functional import CONTINOUS, store


def bf16_values():
    bits = np.arange(65536, dtype=np.uint32) << 16
    return bits.view(np.float32)


def to_bf16(values):
    return torch.tensor(np.asarray(values, np.float32)).to(torch.bfloat16).float().numpy()


def cell_map(bounds, cells):
    lo, hi = np.float32(bounds[0]), np.float32(bounds[-1])
    span = np.float32(hi - lo)
    if span > 0 and np.isfinite(span):
        scale = np.float32(1.0) / span
        offset = np.float32(-lo * scale)
    else:
        scale, offset = np.float32(1.0), np.float32(0.5) - lo

    def cell(x):
        with np.errstate(invalid='ignore', over='ignore'):
            t = (x.astype(np.float64) * np.float64(scale) + np.float64(offset)).astype(np.float32)
        t = np.where(np.isnan(t), np.float32(0), np.clip(t, 0, 1)).astype(np.float32)   # FFMA.SAT: NaN -> 0
        return np.rint(t.astype(np.float64) * (cells - 1)).astype(np.int64)
    return cell


def encode(border, k, bits):
    copies = 4 if bits <= 4 else 2
    low = k | (k << bits)
    if copies == 4:
        low |= low << (2 * bits)
    pattern = int(np.float32(border).view(np.uint32))
    if (pattern << 1) & 0xffffffff == 0:
        return low
    if (pattern >> 31) and low != 0:
        return ((pattern - 0x10000) & 0xffffffff) | low
    return pattern | low


def build_words(bounds, bits, cells=None):
    cells = cells or (2048 if bits == 6 else 16 << bits)
    cell = cell_map(bounds, cells)
    of_border = cell(bounds)
    words = np.zeros(cells, np.uint32)
    crowded = False
    for c in range(cells):
        k = int(np.searchsorted(of_border, c, side='left'))        # borders in earlier cells
        has = k < len(bounds) and of_border[k] == c
        crowded |= k + 1 < len(bounds) and of_border[k + 1] == c
        words[c] = encode(bounds[k] if has else np.inf, k, bits)
    return cell, words, crowded


@pytest.mark.parametrize('bits', [3, 4, 5, 6, 7, 8])
def test_word_table_reproduces_the_bucket_search_for_every_bf16_value(bits):
    x = bf16_values()
    checked = small = 0
    for name in CONTINOUS:
        borders, _ = store.get(name, bits, 'cpu', torch.bfloat16)
        bounds = borders[1:-1].float().numpy()
        if len(np.unique(bounds)) != len(bounds):
            continue        # borders that collide after the cast to bf16: the kernel's exact path
        # the kernel builds the smallest table that separates the borders: 4 << bits cells, doubling up to the
        # full size, from 5 bits on (choose_cells); at 3 and 4 bits 32 cells (one word per bank) is an A/B option
        full = 2048 if bits == 6 else 16 << bits
        cells = 32 if bits <= 4 else 4 << bits
        cell, words, crowded = build_words(bounds, bits, cells)
        small += not crowded
        while crowded and cells < full:
            cells = full if bits <= 4 else cells * 2
            cell, words, crowded = build_words(bounds, bits, cells)
        if crowded:
            continue
        checked += 1
        if cells < full:      # and the full-size table of the same borders gives the same codes
            cell_full, words_full, crowded_full = build_words(bounds, bits, full)
            assert not crowded_full, f'{name} {bits} bits: separates at {cells} cells but not at {full}'
            wf = words_full[cell_full(x)]
            with np.errstate(invalid='ignore'):
                np.testing.assert_array_equal((wf & ((1 << bits) - 1)) + (wf.view(np.float32) < x),
                                              (bounds[None, :] < x[:, None]).sum(axis=1))
        w = words[cell(x)]
        with np.errstate(invalid='ignore'):
            code = (w & ((1 << bits) - 1)) + (w.view(np.float32) < x)
            expect = (bounds[None, :] < x[:, None]).sum(axis=1)
        np.testing.assert_array_equal(code, expect, err_msg=f'{name} {bits} bits')
        # the replicated copies are what the kernel masks out for the other positions of a half
        copies = 4 if bits <= 4 else 2
        for j in range(copies):
            np.testing.assert_array_equal((w >> (bits * j)) & ((1 << bits) - 1), w & ((1 << bits) - 1))
    assert checked >= 8, f'only {checked} tables exercised the word layout at {bits} bits'
    if bits >= 5:
        assert small >= 5, f'only {small} of the {bits}-bit tables separate at the smallest cell count'
    if bits == 3:
        assert small >= 10, f'only {small} of the 3-bit tables fit the conflict-free 32-cell table'


def test_negative_zero_and_extreme_borders():
    """Borders at -0, +-tiny, the largest finite bf16 and +-inf keep the equivalence."""
    bits = 3
    bounds = to_bf16([-np.inf, -3.3895e38, -1e-40, -0.0, 9.2e-41, 1.5, 3.3895e38])
    bounds = np.unique(bounds)      # sorted, -0.0 == 0.0 collapse
    x = bf16_values()
    # a map that separates such borders does not exist; force one cell per border instead
    for i, border in enumerate(bounds):
        for k in (0, i, 7):
            word = np.array([encode(border, k, bits)], np.uint32)
            with np.errstate(invalid='ignore'):
                np.testing.assert_array_equal(word.view(np.float32) < x, border < x, err_msg=f'{border} k={k}')
            assert word[0] & 7 == k
