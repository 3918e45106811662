"""Coefficients and error bounds of the fp32 helper functions in fewbit_b200/csrc/ops.cuh
(namespace accurate): expm1 for arguments <= 0 (ELU family) and log1p(e) for e in (0, 1]
(softplus, logsigmoid).  Fits in float64, then replays the exact fp32 operation sequence
(fused multiply-adds emulated in float64, rounded once) and reports the worst error in ulps
against the float64 functions.

    python tools/fit_fp32_math.py
"""
import numpy as np

f32 = np.float32


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def ulps(got, want):
    want32 = want.astype(f32)
    spacing = np.spacing(np.abs(want32)).astype(np.float64)
    return np.abs(got.astype(np.float64) - want) / spacing


def cheb_fit(fn, lo, hi, degree, weight=None, n=4000):
    k = np.arange(n)
    x = 0.5 * (lo + hi) + 0.5 * (hi - lo) * np.cos(np.pi * (k + 0.5) / n)
    y = fn(x)
    w = np.ones_like(x) if weight is None else weight(x)
    v = np.vander(x, degree + 1, increasing=True)
    coef, *_ = np.linalg.lstsq(v * w[:, None], y * w, rcond=None)
    return coef


def expm1_fit():
    half = np.log(2.0) / 2
    # expm1(r) = r + r^2 Q(r),  Q = (expm1(r) - r) / r^2, fitted so that the error relative to expm1 is flat
    q = lambda r: np.where(np.abs(r) < 1e-9, 0.5 + r / 6, (np.expm1(r) - r) / (r * r))
    coef = cheb_fit(q, -half * 1.01, half * 1.01, 5, weight=lambda r: np.abs(r) + 1e-3)
    return [f32(c) for c in coef]


def expm1_replay(z, coef):
    z = np.maximum(z.astype(f32), f32(-88.0))
    magic = f32(12582912.0)
    km = fma(z, np.full_like(z, f32(1.4426950408889634)), np.full_like(z, magic))
    k = km - magic
    r = fma(k, np.full_like(z, f32(-0.693145751953125)), z)
    r = fma(k, np.full_like(z, f32(-1.428606765330187e-06)), r)
    r2 = (r.astype(np.float64) * r.astype(np.float64)).astype(f32)
    q = np.full_like(z, coef[5])
    for c in coef[4::-1]:
        q = fma(q, r, np.full_like(z, c))
    p = fma(r2, q, r)
    t = np.ldexp(np.ones_like(z), k.astype(np.int32)).astype(f32)
    tm1 = t - f32(1.0)
    return fma(t, p, tm1)


def log1p_fit():
    # log1p(e) = 2 atanh(s), s = e / (2 + e) in (0, 1/3]:  2 s + s^3 R(s^2)
    r = lambda u: np.where(u < 1e-12, 2.0 / 3, (2 * np.arctanh(np.sqrt(u)) - 2 * np.sqrt(u)) / (u * np.sqrt(u)))
    coef = cheb_fit(r, 0.0, 1.0 / 9 * 1.02, 5)
    return [f32(c) for c in coef]


def log1p_replay(e, coef, rcp_error=0.0):
    e = e.astype(f32)
    d = e + f32(2.0)
    y0 = ((1.0 / d.astype(np.float64)) * (1 + rcp_error)).astype(f32)      # MUFU.RCP: ~1 ulp
    y = fma(fma(-d, y0, np.ones_like(e)), y0, y0)                           # one Newton step
    s = (e.astype(np.float64) * y.astype(np.float64)).astype(f32)
    residual = fma(-s, e, fma(np.full_like(e, f32(-2.0)), s, e))           # e - s (2 + e); e - 2 s is exact
    s = fma(residual, y, s)                                                 # correctly rounded quotient
    u = (s.astype(np.float64) * s.astype(np.float64)).astype(f32)
    q = np.full_like(e, coef[5])
    for c in coef[4::-1]:
        q = fma(q, u, np.full_like(e, c))
    su = (s.astype(np.float64) * u.astype(np.float64)).astype(f32)
    return fma(su, q, s + s)


def main():
    ce = expm1_fit()
    z = -np.concatenate([np.logspace(-30, np.log10(88), 400000), np.linspace(0, 20, 400000)])
    err = ulps(expm1_replay(z, ce), np.expm1(z.astype(f32).astype(np.float64)))
    print('expm1  Q coefficients (r^0..r^5):', ', '.join(f'{c:.9e}f' for c in ce))
    print(f'expm1  worst error {err.max():.2f} ulp at z = {z[err.argmax()]:.6g}')
    cl = log1p_fit()
    e = np.concatenate([np.logspace(-38, 0, 400000), np.linspace(0, 1, 400000)[1:]])
    want = np.log1p(e.astype(f32).astype(np.float64))
    for rel in (0.0, 1.2e-7, -1.2e-7):
        err = ulps(log1p_replay(e, cl, rel), want)
        print(f'log1p  rcp error {rel:+.1e}: worst {err.max():.2f} ulp at e = {e[err.argmax()]:.6g}')
    print('log1p  R coefficients (u^0..u^5):', ', '.join(f'{c:.9e}f' for c in cl))


if __name__ == '__main__':
    main()
