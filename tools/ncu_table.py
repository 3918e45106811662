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
        if 'MaskOp<ReluFn>' in name or 'MaskOp<fewbit::ReluFn>' in name or 'MaskFactorOp' in name:
            key = 'relu_forward' if 'forward' in name else 'relu_backward'
            total = sum(number(r[col[m]]) * UNIT.get(units[col[m]], 1.0) * 1e6
                        for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
            traffic.setdefault(key, total)
    if args.traffic and traffic:
        json.dump(traffic, open(args.traffic, 'w'), indent=1)


if __name__ == '__main__':
    main()
