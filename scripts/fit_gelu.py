#!/usr/bin/env python
"""Fit and error study of the GELU used by the cross-encoder's FFN-up epilogue (sentio_b200/csrc/ce_gemm.cu: gelu_erf_h2).

    erf(x / sqrt 2) ~ tanh(x (c0 + c1 x^2 + c2 x^4))      (one MUFU.TANH.F16x2 instead of a reciprocal + an exponential)

Prints the least-squares coefficients (fit of the GELU itself on |x| <= 5.5), the maximum error in exact arithmetic, and the
error of the half-precision evaluation order of the kernel (every operation rounded to fp16, tanh perturbed by the 2^-11
relative error PTX documents for tanh.approx.f16x2) next to the Abramowitz-Stegun 7.1.26 form it replaced and to the exact
function rounded to fp16.  CPU only:  python scripts/fit_gelu.py
"""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf

H = np.float16


def f(x):
    return np.asarray(x, np.float32)


def r(x):
    return np.asarray(x, np.float32).astype(H)


def fma(a, b, c):
    return r(f(a) * f(b) + f(c))


def gelu_exact(x):
    x = np.asarray(x, np.float64)
    return 0.5 * x * (1 + erf(x / np.sqrt(2)))


def fit():
    x = np.linspace(1e-3, 5.5, 20001)
    xs = np.concatenate([-x[::-1], x])

    def res(c):
        x2 = xs * xs
        return 0.5 * xs * (1 + np.tanh(xs * (c[0] + x2 * (c[1] + x2 * c[2])))) - gelu_exact(xs)

    return least_squares(res, [0.7978845608, 0.0356774, 0.0], xtol=1e-15, ftol=1e-15).x


def gelu_tanhfit_fp16(x, c, noise=True):
    x2 = np.minimum(r(f(x) * f(x)), H(36.0))
    q = fma(H(c[2]), x2, H(c[1]))
    q = fma(q, x2, H(c[0]))
    u = r(f(x) * f(q))
    t = np.tanh(f(u).astype(np.float64))
    if noise:
        t = t * (1 + np.random.default_rng(0).uniform(-2.0 ** -11, 2.0 ** -11, size=t.shape))
    t = r(t)
    hx = r(f(H(0.5)) * f(x))
    return fma(hx, t, hx)


def gelu_as_fp16(x):
    one = H(1)
    z = r(f(np.abs(x)) * f(H(0.70710678)))
    t = r(1.0 / f(fma(H(0.3275911), z, one)))
    p = fma(H(1.061405429), t, H(-1.453152027))
    for k in (1.421413741, -0.284496736, 0.254829592):
        p = fma(p, t, H(k))
    ex = r(np.exp2(f(r(f(r(f(z) * f(z))) * f(H(-1.4426950408889634))))))
    e = fma(-r(f(p) * f(t)), ex, one)
    s = np.copysign(np.abs(e), x).astype(H)
    return r(f(r(f(H(0.5)) * f(x))) * f(r(f(one) + f(s))))


def main():
    c = fit()
    xs = np.linspace(-8, 8, 400001)
    x2 = np.minimum(xs * xs, 36.0)
    exact_err = np.abs(0.5 * xs * (1 + np.tanh(xs * (c[0] + x2 * (c[1] + x2 * c[2])))) - gelu_exact(xs))
    stock = np.abs(0.5 * xs * (1 + np.tanh(0.7978845608 * xs * (1 + 0.044715 * xs * xs))) - gelu_exact(xs))
    print("coefficients c0 c1 c2:", *(f"{v:.14g}" for v in c))
    print(f"max |error| in exact arithmetic: {exact_err.max():.3e} (stock tanh-GELU: {stock.max():.3e})")
    rng = np.random.default_rng(1)
    for name, sample in (("N(0,1)", rng.normal(0, 1, 2_000_000)), ("N(0,2)", rng.normal(0, 2, 2_000_000)),
                         ("U(-6,6)", rng.uniform(-6, 6, 2_000_000))):
        x = sample.astype(H)
        ex = gelu_exact(x)
        for nm, y in (("tanh-fit fp16", gelu_tanhfit_fp16(x, c)), ("A-S 7.1.26 fp16", gelu_as_fp16(x)), ("exact -> fp16", ex.astype(H))):
            e = np.abs(y.astype(np.float64) - ex)
            print(f"{name:8s} {nm:16s} max |error| {e.max():.3e}   rms {np.sqrt((e ** 2).mean()):.3e}")


if __name__ == "__main__":
    main()
