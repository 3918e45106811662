"""Markdown table of the launches in an ncu report: `ncu -i X.ncu-rep --page raw --csv > X.csv`,
then `python tools/ncu_table.py X.csv [--traffic profiles/ncu_traffic.json]`.

One row per captured launch: duration, DRAM bytes, registers, grid, achieved warps, executed
warp instructions, issue-slot / ALU / FMA / XU / tensor pipe utilisation and DRAM throughput --
the columns DESIGN.md argues from.  With --traffic the DRAM bytes (read + write) of the mask
kernels are written to the JSON bench.py reads for `roofline.traffic`.
"""
import argparse
import csv
import json
import re

COLUMNS = [
    ('time us', 'gpu__time_duration.sum', 1.0),
    ('DRAM read MB', 'dram__bytes_read.sum', None),
    ('DRAM write MB', 'dram__bytes_write.sum', None),
    ('regs', 'launch__registers_per_thread', 1.0),
    ('grid', 'launch__grid_size', 1.0),
    ('warps act %', 'sm__warps_active.avg.pct_of_peak_sustained_active', 1.0),
    ('warp-instr M', 'smsp__inst_executed.sum', 1e-6),
    ('issue %', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 1.0),
    ('ALU %', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 1.0),
    ('FMA %', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 1.0),
    ('XU %', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 1.0),
    ('tensor %', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 1.0),
    ('DRAM %', 'dram__bytes_read.sum.pct_of_peak_sustained_elapsed+dram__bytes_write.sum.pct_of_peak_sustained_elapsed', 1.0),
]
UNIT = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'ns': 1e-3, 'us': 1.0, 'usecond': 1.0,
        'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3}


def number(text):
    try:
        return float(text.replace(',', ''))
    except ValueError:
        return float('nan')


def short(name):
    name = re.sub(r'\b(fewbit|sketch)::', '', name)
    name = re.sub(r'\(.*', '', name)
    return name.replace('void ', '')[:90]


def traffic_key(name):
    """bench.py's name of a profiled kernel (benchmarks/profile_kernels.py launches them at the
    bench's sizes), or None."""
    name = name.replace('fewbit::', '')
    tag = 'bf16' if '__nv_bfloat16' in name else 'f32'
    m = re.search(r'MaskOp<(\w+?)Fn>', name)
    if m and 'forward_tiles' in name:
        return {'Relu': 'relu', 'LeakyRelu': 'leaky_relu', 'Hardtanh': 'hardtanh'}.get(m.group(1), m.group(1).lower()) + \
            ('_forward' if tag == 'bf16' else '_f32_forward')
    if 'MaskFactorOp' in name and 'backward_tiles' in name:
        return 'relu_backward' if tag == 'bf16' else 'relu_f32_backward'     # one kernel serves every 1-bit backward
    m = re.search(r'QuantizeOp<(\w+?)Fn, (?:__nv_bfloat16|float), (\d)>', name)
    if m and 'forward_tiles' in name:
        return f'{m.group(1).lower()}{m.group(2)}_{tag}_forward'
    m = re.search(r'LevelsOp<(?:__nv_bfloat16|float), (\d)>', name)
    if m and 'backward_tiles' in name:
        return f'levels{m.group(1)}_{tag}_backward'
    if 'sketch_kernel' in name:
        return 'sketch_kernel'
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('--traffic', default=None)
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv)))
    header, units, launches = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(header)}
    print('| kernel | ' + ' | '.join(c[0] for c in COLUMNS) + ' |')
    print('|---|' + '---|' * len(COLUMNS))
    traffic = {}
    for r in launches:
        cells = []
        for _, metric, scale in COLUMNS:
            parts = metric.split('+')
            if any(m not in col for m in parts):
                cells.append('-')
                continue
            v = 0.0
            for m in parts:
                factor = UNIT.get(units[col[m]], 1.0) if scale is None or 'time' in m else 1.0
                v += number(r[col[m]]) * factor * (scale if scale is not None else 1.0)
            cells.append(f'{v:.1f}' if abs(v) < 1e4 else f'{v:.0f}')
        name = r[col['Kernel Name']]
        print(f'| `{short(name)}` | ' + ' | '.join(cells) + ' |')
        key = traffic_key(name)
        if key:
            total = sum(number(r[col[m]]) * UNIT.get(units[col[m]], 1.0) * 1e6
                        for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
            traffic.setdefault(key, total)
    if args.traffic and traffic:
        json.dump(traffic, open(args.traffic, 'w'), indent=1)


if __name__ == '__main__':
    main()
