"""Built-in tables: same numbers as the reference ships, same casts as the reference makes."""
from pathlib import Path

import numpy as np
import pytest
import torch

from fewbit_b200.functional import CONTINOUS, store

REF_NPZ = Path('/root/reference/fewbit/data/builtin.npz')
OURS = Path(__file__).resolve().parents[1] / 'fewbit_b200' / 'data' / 'builtin.npz'


def test_store_is_complete():
    # 13 functions x bits 1..4: the reference's tables (tools/quantize-builtins.sh:8 of the reference);
    # bits 5..8: optimal tables from fewbit_b200/quantize.py (the reference ships none)
    assert len(store) == 104
    for name in CONTINOUS:
        for bits in range(1, 9):
            borders, levels = store.get(name, bits)
            assert borders.numel() == 2 ** bits + 1 and levels.numel() == 2 ** bits
            assert borders[0] == -100 and borders[-1] == 100
            assert torch.all(borders[1:] > borders[:-1])
    with pytest.raises(KeyError):
        store.get('gelu', 9)


def test_store_caches_per_device_and_dtype():
    a = store.get('gelu', 3, 'cpu', torch.bfloat16)
    assert a[0].dtype == torch.bfloat16 and store.get('gelu', 3, 'cpu', torch.bfloat16) is a
    assert store.get('gelu', 3)[0].dtype == torch.float32


def test_gelu3_table_of_survey_appendix_b():
    borders, levels = store.get('gelu', 3)
    np.testing.assert_array_equal(
        borders[1:-1].numpy(),
        np.array([-2.41658115, -0.710008025, -0.325840563, 1.06942185e-04, 0.326057166,
                  0.710240841, 2.41447878], np.float32))
    np.testing.assert_array_equal(
        levels.numpy(),
        np.array([-1.9399123e-04, -8.8279128e-02, 0.12568383, 0.37231442, 0.62785137, 0.87445050,
                  1.0883480, 1.0001949], np.float32))


def test_casts_match_reference_package(golden_tables):
    """fp32 / bf16 tables exactly as the reference's store hands them to the operators."""
    assert len(golden_tables) == 208
    for name in CONTINOUS:
        for bits in range(1, 5):
            for tag, dtype in (('f32', torch.float32), ('bf16', torch.bfloat16)):
                borders, levels = store.get(name, bits, 'cpu', dtype)
                bounds = borders[1:-1]
                if dtype == torch.bfloat16:
                    bounds = bounds.view(torch.int16).numpy().view(np.uint16)
                    levels = levels.view(torch.int16).numpy().view(np.uint16)
                else:
                    bounds, levels = bounds.numpy(), levels.numpy()
                key = f'{name}{bits:02d}-{tag}'
                np.testing.assert_array_equal(bounds, golden_tables[f'{key}-bounds'])
                np.testing.assert_array_equal(levels, golden_tables[f'{key}-levels'])


@pytest.mark.skipif(not REF_NPZ.exists(), reason='reference tree not mounted')
def test_npz_equals_reference_data():
    with np.load(REF_NPZ) as ref, np.load(OURS) as ours:
        assert sorted(ref.keys()) == sorted(ours.keys())
        for key in ref.keys():
            assert ref[key].dtype == ours[key].dtype == np.float64
            np.testing.assert_array_equal(ref[key], ours[key])
