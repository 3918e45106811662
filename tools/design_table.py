"""The two tables of DESIGN.md section 4 from profiles/r02_function_sweep.json and profiles/r02_bench_n1.json:
python tools/design_table.py"""
import json
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sweep = json.load(open(ROOT / 'profiles' / 'r02_function_sweep.json'))
peak = sweep['peak_GBps']
rows = sweep['rows']
CONTINUOUS = {'celu', 'elu', 'gelu', 'hardswish', 'logsigmoid', 'mish', 'selu', 'sigmoid', 'silu', 'softplus', 'softsign',
              'tanh', 'tanhshrink'}


def span(sel, key):
    vals = [r[key] for r in rows if sel(r)]
    lo, hi = min(vals), max(vals)
    return (f'{lo / 1e3:.2f}–{hi / 1e3:.2f} TB/s' if hi - lo > 5 else f'{lo / 1e3:.2f} TB/s',
            f'{100 * lo / peak:.0f}–{100 * hi / peak:.0f} %' if hi - lo > 5 else f'{100 * lo / peak:.0f} %')


print('| Kernel (128×128×3072, back to back from a CUDA graph) | measured | of measured peak |\n|---|---|---|')
for tag in ('bf16', 'f32'):
    def cont(r, tag=tag):
        return r['dtype'] == tag and r['function'] in CONTINUOUS
    masks = [r for r in rows if r['dtype'] == tag and r['function'] not in CONTINUOUS]
    if masks:
        f = (min(r['fwd_GBps'] for r in masks), max(r['fwd_GBps'] for r in masks))
        b = (min(r['bwd_GBps'] for r in masks), max(r['bwd_GBps'] for r in masks))
        print(f'| 1-bit mask forward / backward, {tag} | {f[0] / 1e3:.2f}–{f[1] / 1e3:.2f} TB/s / {b[0] / 1e3:.2f}–{b[1] / 1e3:.2f} TB/s | '
              f'{100 * f[0] / peak:.0f}–{100 * f[1] / peak:.0f} % / {100 * b[0] / peak:.0f}–{100 * b[1] / peak:.0f} % |')
    g = next(r for r in rows if cont(r) and r['function'] == 'gelu' and r['bits'] == 3)
    print(f"| 3-bit GELU forward / backward, {tag} | {g['fwd_GBps'] / 1e3:.2f} / {g['bwd_GBps'] / 1e3:.2f} TB/s | "
          f"{100 * g['fwd_frac']:.0f} % / {100 * g['bwd_frac']:.0f} % |")
    for label, bits in (('1–4 bits', (1, 2, 3, 4)), ('5–6 bits', (5, 6)), ('7 bits', (7,)), ('8 bits', (8,))):
        a, b = span(lambda r: cont(r) and r['bits'] in bits, 'fwd_GBps')
        print(f'| forward, all 13 functions, {label}, {tag} | {a} | {b} |')
    a, b = span(cont, 'bwd_GBps')
    print(f'| backward, all functions and bit widths, {tag} | {a} | {b} |')
    below = sum(1 for r in rows if cont(r) and r['fwd_frac'] < 0.8)
    print(f'<!-- {tag}: {below} of {sum(1 for r in rows if cont(r))} forward cells below 80 % -->')
copy = sweep.get('torch_copy_same_shape')
print(f"| plain `torch` copy of the same 128×128×3072 tensors | fp32 {copy['f32']['GBps'] / 1e3:.2f} TB/s, bf16 {copy['bf16']['GBps'] / 1e3:.2f} TB/s | "
      f"{100 * copy['f32']['frac']:.0f} % / {100 * copy['bf16']['frac']:.0f} % |")

bench = json.loads(open(ROOT / 'profiles' / 'r02_bench_n1.json').read().strip().splitlines()[-1])
print('\n| Kernel (`bench.py` `rooflines`, driver-run convention) | measured | of measured peak |\n|---|---|---|')
for r in bench['rooflines']:
    unit = r.get('unit', 'GB/s')
    how = 'cold groups' if ('gelu3' in r['kernel'] or 'sketch' in r['kernel']) else '1 GiB bf16, events in the timed region'
    print(f"| `{r['kernel']}` ({how}) | {r['achieved']:.0f} {unit} | {100 * r['frac']:.1f} % |")
