"""The reference arm of bench.py runs without a GPU (it times the reference's own CPU path from
oracle/_ref, or the oracle port): its JSON line must carry the keys the driver reads, and under
torchrun only rank 0 may print it."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def run_reference_arm(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    proc = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                          capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    return proc.stdout.strip().splitlines()


def test_reference_arm_prints_the_contract_line():
    lines = run_reference_arm()
    line = json.loads(lines[-1])
    assert line['impl'] == 'reference'
    if 'unavailable' in line:           # allowed by the contract, but the oracle always exists here
        raise AssertionError(line)
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in line, key
    assert line['unit'] == 'GB/s' and line['higher_is_better'] is True and line['value'] > 0
    assert line['cpu_baseline']['kind'] in ('reference', 'port') and line['cpu_baseline']['cores'] >= 1
    assert line['cpu_baseline']['value'] == line['value'] and 'sample' in line['cpu_baseline']
    assert line['e2e'] == {'value': line['value'], 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in line['config'] and 'model' not in line['config']


def test_reference_arm_is_silent_on_other_ranks():
    assert run_reference_arm({'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'}) == []
