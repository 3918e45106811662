"""Re-emit the built-in quantisation tables ("stored optimal boundaries").

The tables are DATA, not code: 13 functions x bits 1..4 x {borders (2^b + 1, ends at
+-100), levels (2^b)} in float64.  Bit-exact default codes require the very same numbers
the reference ships in fewbit/data/builtin.npz (loader: fewbit/functional/activations.py:69-81),
so this script reads them there and writes fewbit_b200/data/builtin.npz with the same key
format ``{func}{bits:02d}-{borders,levels}``.  tests/test_tables.py re-checks equality
whenever the reference tree is mounted.
"""
import sys
from pathlib import Path

import numpy as np

ref = Path(sys.argv[1] if len(sys.argv) > 1 else '/root/reference/fewbit/data/builtin.npz')
out = Path(__file__).resolve().parent.parent / 'fewbit_b200' / 'data' / 'builtin.npz'
with np.load(ref) as npz:
    tables = {key: np.asarray(npz[key], dtype=np.float64) for key in sorted(npz.keys())}
out.parent.mkdir(parents=True, exist_ok=True)
np.savez_compressed(out, **tables)
print(f'wrote {len(tables)} arrays to {out}')
