"""Few-bit tables for any bit width: the optimal piecewise-constant approximation of f'.

The reference fits its tables with `fewbit.approx.approximate` (fewbit/approx.py:61-156): random
initial borders, then alternating "levels = mean of f' on each interval" and a gradient step on
the borders, for the objective

    E(borders) = sum_i  int_{b_i}^{b_{i+1}} (f'(x) - l_i)^2 dx,   l_i = (f(b_{i+1}) - f(b_i)) / (b_{i+1} - b_i)

over the domain [-100, 100].  That iteration is fragile beyond 16 levels, which is why the
reference ships 1..4-bit tables only (tools/quantize-builtins.sh:8).  The kernels
here take up to 8 bits, so this module solves the same objective robustly:

1. exact dynamic programme over candidate borders on a sinh-spaced grid (dense around zero,
   reaching the domain ends), segment costs from prefix integrals of f' and f'^2 in float64;
2. continuous polish: Lloyd-Max updates -- each border moves to the nearby point where f'
   equals the mean of its two neighbouring levels (the stationarity condition the reference's
   gradient step aims at) -- accepted only while the objective decreases.

`python -m fewbit_b200 quantize 6 gelu -o tables.npz` (the reference's `fewbit quantize NOBITS SPEC`,
fewbit/cli.py:60-124, 168-176) writes the reference's npz key format
(`<name><bits:02d>-borders` / `-levels`, fewbit/cli.py:108-112), loadable with
`fewbit_b200.functional.store.load(path)`.  Host-side tooling: numpy + torch autograd in float64.
"""
from __future__ import annotations

import argparse
from importlib import import_module
from pathlib import Path
from typing import Callable, Tuple

import numpy as np
import torch as T

DOMAIN = (-100.0, 100.0)


def _callable(spec) -> Tuple[str, Callable]:
    """'gelu' (torch.nn.functional / torch), 'package.module:function', or a callable."""
    if callable(spec):
        return getattr(spec, '__name__', 'function'), spec
    if ':' in spec:
        module, name = spec.split(':', 1)
        return name, getattr(import_module(module), name)
    fn = getattr(T.nn.functional, spec, None) or getattr(T, spec)
    return spec, fn


def _values(func: Callable, xs: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """f(xs) and f'(xs) in float64 through autograd."""
    x = T.tensor(xs, dtype=T.float64, requires_grad=True)
    y = func(x)
    y.backward(T.ones_like(y))
    return y.detach().numpy(), x.grad.numpy()


class _Integrals:
    """Prefix integrals of f' (that is f itself) and of f'^2 on a fine sinh-spaced grid."""

    def __init__(self, func: Callable, domain=DOMAIN, points: int = 1 << 18, scale: float = 0.5):
        lo, hi = domain
        t = np.linspace(np.arcsinh(lo / scale), np.arcsinh(hi / scale), points)
        self.x = scale * np.sinh(t)
        self.x[0], self.x[-1] = lo, hi
        self.f, self.df = _values(func, self.x)
        # Jumps of f' (hardswish at +-3, selu at 0): the optimal table has a border exactly there,
        # and the prefix integrals must not smear across it.  Locate each jump by bisection and add
        # the two points that bracket it as grid nodes (and as mandatory candidates of the search).
        self.jumps = []
        for i in np.nonzero(np.abs(np.diff(self.df)) > 0.01)[0]:
            a, b = self.x[i], self.x[i + 1]
            da, db = self.df[i], self.df[i + 1]
            for _ in range(60):
                m = 0.5 * (a + b)
                dm = _values(func, np.array([m]))[1][0]
                if abs(dm - da) <= abs(dm - db):
                    a = m
                else:
                    b = m
            self.jumps += [a, b]
        if self.jumps:
            self.x = np.unique(np.concatenate([self.x, self.jumps]))
            self.f, self.df = _values(func, self.x)
        sq = self.df ** 2
        # Simpson-accurate cumulative integral of f'^2: trapezoid plus the midpoint correction
        mid = 0.5 * (self.x[1:] + self.x[:-1])
        _, dmid = _values(func, mid)
        h = np.diff(self.x)
        self.s2 = np.concatenate([[0.0], np.cumsum(h / 6.0 * (sq[:-1] + 4.0 * dmid ** 2 + sq[1:]))])
        self.func = func

    def at(self, b: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """f(b) exactly and int f'^2 up to b (interpolated between fine points)."""
        f, _ = _values(self.func, np.asarray(b, dtype=np.float64))
        return f, np.interp(b, self.x, self.s2)

    def error(self, borders: np.ndarray) -> float:
        f, s2 = self.at(borders)
        return float(np.sum(np.diff(s2) - np.diff(f) ** 2 / np.diff(borders)))


def _dynamic_programme(integ: _Integrals, levels: int, candidates: int, reach: float = np.inf) -> np.ndarray:
    """Best `levels - 1` interior borders among `candidates` grid points (plus the domain ends);
    interior borders are taken from |x| <= reach only."""
    inside = np.nonzero(np.abs(integ.x) <= reach)[0]
    idx = inside[np.linspace(0, inside.size - 1, candidates + 2).round().astype(int)]
    idx = np.unique(np.concatenate([[0, integ.x.size - 1], idx, np.searchsorted(integ.x, integ.jumps[1::2])]).astype(int))
    x, f, s2 = integ.x[idx], integ.f[idx], integ.s2[idx]
    n = x.size
    with np.errstate(divide='ignore', invalid='ignore'):
        cost = (s2[None, :] - s2[:, None]) - (f[None, :] - f[:, None]) ** 2 / (x[None, :] - x[:, None])
    cost[np.tril_indices(n)] = np.inf                      # segments run left to right
    best = cost[0].copy()                                  # one segment ending at j
    back = np.zeros((levels, n), dtype=np.int32)
    for m in range(1, levels):
        total = best[:, None] + cost                       # previous end i, new segment (i, j)
        back[m] = np.argmin(total, axis=0)
        best = total[back[m], np.arange(n)]
    cuts, j = [], n - 1
    for m in range(levels - 1, 0, -1):
        j = back[m][j]
        cuts.append(j)
    return np.concatenate([[x[0]], x[np.array(cuts[::-1], dtype=int)], [x[-1]]])


def _polish(integ: _Integrals, borders: np.ndarray, sweeps: int = 200, reach: float = np.inf) -> np.ndarray:
    """Lloyd-Max: move each border to where f' equals the mean of the neighbouring levels."""
    func = integ.func
    best, best_err = borders.copy(), integ.error(borders)
    step = 1.0
    for _ in range(sweeps):
        f, _ = integ.at(best)
        lv = np.diff(f) / np.diff(best)
        target = 0.5 * (lv[:-1] + lv[1:])
        inner = best[1:-1]
        _, d = _values(func, inner)
        eps = 1e-6 * np.maximum(1.0, np.abs(inner))
        _, d_hi = _values(func, inner + eps)
        _, d_lo = _values(func, inner - eps)
        slope = (d_hi - d_lo) / (2 * eps)                  # f''(b)
        move = np.where(np.abs(slope) > 1e-12, (target - d) / np.where(slope == 0, 1.0, slope), 0.0)
        gap = np.minimum(np.diff(best)[:-1], np.diff(best)[1:])
        move = np.clip(move, -0.45 * gap, 0.45 * gap)      # keep the order
        trial = best.copy()
        trial[1:-1] = np.clip(inner + step * move, -reach, reach)
        if np.any(np.diff(trial) <= 0):
            step *= 0.5
            if step < 1e-3:
                break
            continue
        err = integ.error(trial)
        if err < best_err * (1 - 1e-12):
            best, best_err = trial, err
        else:
            step *= 0.5
            if step < 1e-3:
                break
    return best


def optimal_table(spec, bits: int, domain=DOMAIN, candidates: int = 2048, reach: float = np.inf):
    """(borders, levels, error): float64 arrays in the built-in layout (borders include the domain
    ends) minimising the reference's objective for 2**bits levels.  `reach` confines the interior
    borders to |x| <= reach (the outer pieces then run from there to the domain ends): derivatives
    with heavy tails (softsign) otherwise place borders so far out that the kernels' uniform cell
    look-up cannot separate the dense ones in the middle."""
    if not 1 <= bits <= 8:
        raise ValueError('bits must be in 1..8')
    _, func = _callable(spec)
    integ = _Integrals(func, domain)
    borders = _polish(integ, _dynamic_programme(integ, 1 << bits, candidates, reach), reach=reach)
    f, _ = integ.at(borders)
    return borders, np.diff(f) / np.diff(borders), integ.error(borders)


def table_error(spec, borders) -> float:
    """The objective for given borders (with their optimal levels)."""
    _, func = _callable(spec)
    return _Integrals(func, (float(borders[0]), float(borders[-1]))).error(np.asarray(borders, dtype=np.float64))


def save(path, name: str, bits: int, borders, levels) -> None:
    """Add (or replace) one table in an npz file, reference key format (fewbit/cli.py:108-124)."""
    path = Path(path)
    tables = {}
    if path.exists():
        with np.load(path) as npz:
            tables = dict(npz)
    case = f'{name}{bits:02d}'
    tables[f'{case}-borders'], tables[f'{case}-levels'] = np.asarray(borders), np.asarray(levels)
    np.savez(path, **tables)


def add_arguments(ap: argparse.ArgumentParser) -> None:
    """Arguments of the reference's `fewbit quantize` (fewbit/cli.py:168-176): same positionals and
    option letters.  The iteration controls of its solver are accepted and ignored."""
    ap.add_argument('-M', '--max-iters', type=int, default=None, help='ignored (reference solver only)')
    ap.add_argument('-b', '--border-error', type=float, default=None, help='ignored (reference solver only)')
    ap.add_argument('-l', '--level-error', type=float, default=None, help='ignored (reference solver only)')
    ap.add_argument('-s', '--seed', type=int, default=None, help='ignored: the solver is deterministic')
    ap.add_argument('-o', '--output', type=Path, default=None, help='npz file to create or update')
    ap.add_argument('--candidates', type=int, default=2048, help='grid points of the dynamic programme')
    ap.add_argument('nobits', type=int, help='number of bits (1..8)')
    ap.add_argument('spec', help="activation name (torch.nn.functional) or qualified 'module:function'")


def run(args) -> None:
    name, _ = _callable(args.spec)
    borders, levels, err = optimal_table(args.spec, args.nobits, candidates=args.candidates)
    print(f'{name}, {args.nobits} bits: {levels.size} levels, L2 error of the derivative {err:.6e}')
    for i, level in enumerate(levels):
        print(f'[{i:3d}] [{borders[i]:+9.4f}, {borders[i + 1]:+9.4f}) => {level:+.6e}')
    if args.output:
        save(args.output, name, args.nobits, borders, levels)
        print(f'saved to {args.output}')


def main(argv=None):
    ap = argparse.ArgumentParser(prog='python -m fewbit_b200.quantize', description=__doc__.split('\n\n')[0])
    add_arguments(ap)
    run(ap.parse_args(argv))


if __name__ == '__main__':
    main()
