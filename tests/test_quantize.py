"""fewbit_b200/quantize.py: the table solver against the reference's own tables and objective
(fewbit/approx.py:61-156), the shipped 5..8-bit tables, and the CLI's npz format (fewbit/cli.py:108-124)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from fewbit_b200 import quantize
from fewbit_b200.functional import CONTINOUS, store
from fewbit_b200.functional.activations import StepwiseStore

GOLDEN = Path(__file__).parent / 'golden' / 'reference_tables.npz'


@pytest.mark.parametrize('name', ['gelu', 'silu', 'tanh', 'softplus', 'elu', 'hardswish'])
def test_solver_matches_or_beats_the_reference_tables(name):
    """Same objective as the reference's fit.  Where the reference's random-start gradient descent
    reached the optimum, the dynamic programme + Lloyd-Max polish lands on the same borders (the
    1..2-bit tables, to 1e-3); where it stopped in a local minimum (e.g. gelu 4-bit, silu 3-bit,
    elu 3..4-bit) ours is strictly better.  Never worse."""
    for bits in (1, 2, 3, 4):
        ref_borders = store.get(name, bits, 'cpu', torch.float64)[0].numpy()
        borders, levels, err = quantize.optimal_table(name, bits, candidates=1024)
        ref_err = quantize.table_error(name, ref_borders)
        assert err <= ref_err * (1 + 1e-4), (name, bits, err, ref_err)
        assert borders[0] == -100 and borders[-1] == 100 and np.all(np.diff(borders) > 0)
        # same optimum: same borders (tanh' is even, its 2-level optimum is not unique: skipped)
        if bits <= 2 and err > ref_err * (1 - 1e-4) and name in ('gelu', 'silu', 'softplus', 'elu'):
            assert np.abs(borders - ref_borders)[1:-1].max() < 2e-3, (name, bits)


def test_solver_improves_known_local_minima_of_the_reference():
    for name, bits, ratio in (('gelu', 4, 0.80), ('silu', 3, 0.70), ('elu', 4, 0.80)):
        ref_borders = store.get(name, bits, 'cpu', torch.float64)[0].numpy()
        _, _, err = quantize.optimal_table(name, bits, candidates=1024)
        assert err < ratio * quantize.table_error(name, ref_borders), (name, bits)


def test_levels_are_the_mean_derivative_and_error_falls_with_bits():
    previous = None
    for bits in (2, 4, 6):
        borders, levels, err = quantize.optimal_table('gelu', bits, candidates=1024)
        f = torch.nn.functional.gelu(torch.tensor(borders, dtype=torch.float64)).numpy()
        assert np.allclose(levels, np.diff(f) / np.diff(borders), rtol=0, atol=1e-12)
        if previous is not None:
            assert err < previous / 8          # ~1/16 per two bits (h^2 convergence), with margin
        previous = err


def test_shipped_extended_tables_are_consistent():
    """data/extended.npz (tools/make_extended_tables.py): levels are the mean derivative of their
    intervals, and each extra bit cuts the objective (about 4x for smooth derivatives)."""
    for name in CONTINOUS:
        fn = getattr(torch.nn.functional, name, None) or getattr(torch, name)
        errors = []
        for bits in (4, 5, 6, 7, 8):
            borders, levels = store.get(name, bits, 'cpu', torch.float64)
            f = fn(borders)
            assert torch.allclose(levels, (f[1:] - f[:-1]) / (borders[1:] - borders[:-1]), rtol=0, atol=1e-9), (name, bits)
            errors.append(quantize.table_error(name, borders.numpy()))
        assert all(b < a for a, b in zip(errors[1:], errors[2:])), (name, errors)     # 5 > 6 > 7 > 8
        assert errors[1] < errors[0] * 1.001, (name, errors)                          # 5-bit optimal <= reference 4-bit


def test_cli_writes_the_reference_npz_format(tmp_path, capsys):
    """`python -m fewbit_b200 quantize NOBITS SPEC -o file`: the reference's command line
    (fewbit/cli.py:168-176) and npz keys (fewbit/cli.py:108-112); an existing file is updated."""
    from fewbit_b200.__main__ import main
    out = tmp_path / 'tables.npz'
    main(['quantize', '-o', str(out), '--candidates', '512', '2', 'silu'])
    main(['quantize', '-s', '1', '-M', '10', '-o', str(out), '--candidates', '512', '1', 'torch.nn.functional:gelu'])
    assert 'saved to' in capsys.readouterr().out
    with np.load(out) as npz:
        assert sorted(npz.keys()) == ['gelu01-borders', 'gelu01-levels', 'silu02-borders', 'silu02-levels']
        assert npz['silu02-borders'].shape == (5, ) and npz['silu02-levels'].shape == (4, )
    loaded = StepwiseStore().load(out)
    borders, levels = loaded.get('silu', 2)
    assert borders.dtype == torch.float32 and borders.numel() == 5 and levels.numel() == 4
    main(['version'])
    assert 'version' in capsys.readouterr().out
