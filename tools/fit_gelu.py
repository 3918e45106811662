"""Fit and validate the single-branch GELU used by the CUDA forward kernel.

    Phi(-|x|) = 0.5 * erfc(t),  t = |x| / sqrt(2)
    erfc(t)   = s * 2^( -log2(e) * t^2 + P(s) ),   s = 1 / (1 + t/2)

P(s) = log2(erfcx(t) / s) is smooth on s in (0, 1] (Numerical-Recipes style variable); we fit
it with a degree-N polynomial (Chebyshev least squares on Chebyshev nodes, converted to the
monomial basis) and check the whole pipeline in emulated fp32 against float64.
"""
import sys

import numpy as np
from numpy.polynomial import chebyshev as C
from scipy.special import erfc, erfcx

L2E = 1.4426950408889634


def target(s):
    t = 2.0 * (1.0 / s - 1.0)
    return np.log2(erfcx(t) / s)


def fit(deg, smin=0.02):
    k = np.arange(4000)
    nodes = np.cos(np.pi * (k + 0.5) / 4000)                     # Chebyshev nodes on [-1, 1]
    s = 0.5 * (nodes + 1) * (1 - smin) + smin
    cheb = C.chebfit(nodes, target(s), deg)
    # monomial in s: substitute nodes = (2 s - (1 + smin)) / (1 - smin)
    poly_u = C.cheb2poly(cheb)
    a, b = 2.0 / (1 - smin), -(1 + smin) / (1 - smin)
    out = np.zeros(deg + 1)
    base = np.array([1.0])
    for c in poly_u:
        out[:len(base)] += c * base
        base = np.convolve(base, [b, a])
    return out                                                     # out[i] * s^i


def f32(x):
    return np.asarray(x, np.float64).astype(np.float32)


def fma(a, b, c):
    return (np.float64(1) * a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def gelu_emulated(x, coef, rcp_err=0.0, ex2_err=0.0):
    x = f32(x)
    t = (np.abs(x) * np.float32(0.70710678118654752440)).astype(np.float32)
    d = fma(t, np.full_like(t, 0.5), np.ones_like(t))
    s = (1.0 / d.astype(np.float64) * (1 + rcp_err)).astype(np.float32)
    c = f32(coef)
    p = np.full_like(s, c[-1])
    for ci in c[-2::-1]:
        p = fma(p, s, np.full_like(s, ci))
    nt = (t * np.float32(-L2E)).astype(np.float32)
    a = fma(nt, t, p)
    e = (np.exp2(a.astype(np.float64)) * (1 + ex2_err)).astype(np.float32)
    hs = (s * np.float32(0.5)).astype(np.float32)
    h = (hs * e).astype(np.float32)
    m = (x * h).astype(np.float32)
    return np.where(x < 0, m, (x - m).astype(np.float32))


def gelu_aten_emulated(x):
    from scipy.special import erf
    x = f32(x)
    z = (x * np.float32(0.70710678118654752440)).astype(np.float32)
    e = erf(z.astype(np.float64)).astype(np.float32)               # a perfectly rounded erff
    return ((x * np.float32(0.5)).astype(np.float32) * (np.float32(1) + e).astype(np.float32)).astype(np.float32)


def report(deg):
    coef = fit(deg)
    s = np.linspace(0.02, 1, 200001)
    perr = np.max(np.abs(np.polyval(coef[::-1], s) - target(s)))
    x = np.concatenate([np.linspace(-9, 9, 2_000_001), np.random.default_rng(0).standard_normal(1_000_000) * 2])
    x = f32(x).astype(np.float64)
    exact = 0.5 * x * erfc(-x / np.sqrt(2))
    worst = 0
    for re, ee in ((0, 0), (6e-8, 1.2e-7), (-6e-8, -1.2e-7), (6e-8, -1.2e-7)):
        y = gelu_emulated(x, coef, re, ee).astype(np.float64)
        ulp = np.spacing(np.abs(exact).astype(np.float32)).astype(np.float64)
        err = np.abs(y - exact)
        worst = max(worst, np.max(err / ulp))
        aerr = np.max(err)
    aten = gelu_aten_emulated(x).astype(np.float64)
    ulp = np.spacing(np.abs(exact).astype(np.float32)).astype(np.float64)
    bound = 4 * np.spacing(np.abs(aten).astype(np.float32)).astype(np.float64) + 2.5e-7
    y = gelu_emulated(x, coef).astype(np.float64)
    print(f'deg {deg}: |P err|max={perr:.2e}  worst ulp vs exact={worst:.2f}  max abs err={aerr:.2e}  '
          f'within test bound vs ATen-formula: {np.all(np.abs(y - aten) <= bound)}  '
          f'max|y-aten|={np.max(np.abs(y - aten)):.2e}')
    lin = f32(np.linspace(-5, 5, 101)).astype(np.float64)
    print('    L2 on linspace(-5,5,101) vs aten-formula:',
          np.linalg.norm(gelu_emulated(lin, coef).astype(np.float64) - gelu_aten_emulated(lin)))
    return coef


if __name__ == '__main__':
    for deg in (int(a) for a in (sys.argv[1:] or ['5', '6', '7', '8', '9', '10'])):
        coef = report(deg)
        print('    coef:', ', '.join(f'{c:.9e}f' for c in f32(coef)))


def gelu_atenlike(x, coef, rcp_err=0.0, ex2_err=0.0):
    """Variant actually used: erf = copysign(1 - erfc(|z|), x); y = (0.5 x) * (1 + erf) -- the
    last two steps are ATen's formula, so rounding/cancellation behaviour matches F.gelu."""
    x = f32(x)
    t = (np.abs(x) * np.float32(0.70710678118654752440)).astype(np.float32)
    d = fma(t, np.full_like(t, 0.5), np.ones_like(t))
    s = (1.0 / d.astype(np.float64) * (1 + rcp_err)).astype(np.float32)
    c = f32(coef)
    p = np.full_like(s, c[-1])
    for ci in c[-2::-1]:
        p = fma(p, s, np.full_like(s, ci))
    nt = (t * np.float32(-L2E)).astype(np.float32)
    a = fma(nt, t, p)
    e = (np.exp2(a.astype(np.float64)) * (1 + ex2_err)).astype(np.float32)
    w = (s * e).astype(np.float32)                       # erfc(|z|)
    mag = (np.float32(1) - w).astype(np.float32)         # |erf|
    erf_ = np.copysign(mag, x)
    return ((x * np.float32(0.5)).astype(np.float32) * (np.float32(1) + erf_).astype(np.float32)).astype(np.float32)


def report2(deg):
    coef = fit(deg)
    x = np.concatenate([np.linspace(-9, 9, 2_000_001), np.random.default_rng(0).standard_normal(1_000_000) * 2])
    x = f32(x).astype(np.float64)
    exact = 0.5 * x * erfc(-x / np.sqrt(2))
    aten = gelu_aten_emulated(x).astype(np.float64)
    bound_exact = 4 * np.spacing(np.abs(exact).astype(np.float32)).astype(np.float64) + 2.5e-7
    ok, dmax, emax = True, 0, 0
    for re, ee in ((0, 0), (6e-8, 1.2e-7), (-6e-8, -1.2e-7), (6e-8, -1.2e-7), (-6e-8, 1.2e-7)):
        y = gelu_atenlike(x, coef, re, ee).astype(np.float64)
        ok &= bool(np.all(np.abs(y - exact) <= bound_exact))
        dmax = max(dmax, np.max(np.abs(y - aten)))
        emax = max(emax, np.max(np.abs(y - exact)))
    y = gelu_atenlike(x, coef).astype(np.float64)
    same = np.mean(y == aten)
    lin = f32(np.linspace(-5, 5, 101)).astype(np.float64)
    l2 = np.linalg.norm(gelu_atenlike(lin, coef).astype(np.float64) - gelu_aten_emulated(lin))
    # bf16 view: round both to bf16 and count mismatches
    def bf(v):
        u = f32(v).view(np.uint32).astype(np.uint64)
        return ((u + 0x7fff + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    xb = (bf(x).astype(np.uint32) << 16).view(np.float32).astype(np.float64)
    yb, ab = bf(gelu_atenlike(xb, coef)), bf(gelu_aten_emulated(xb))
    print(f'deg {deg}: within oracle bound (4ulp+2.5e-7): {ok}  max|y-exact|={emax:.2e}  max|y-aten|={dmax:.2e}  '
          f'bit-identical to ATen-formula: {same:.4f}  L2(linspace)={l2:.2e}  bf16 mismatches: {np.mean(yb != ab):.5f} '
          f'max bf16 ulp diff {np.max(np.abs(yb.astype(int) - ab.astype(int)))}')


if __name__ == '__main__':
    print('--- ATen-like tail ---')
    for deg in (5, 6, 7, 8, 9, 10):
        report2(deg)
