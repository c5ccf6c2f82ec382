#!/usr/bin/env python
"""Coefficients of gelu_erf_fast (csrc/kvq_common.cuh) and its error against the exact erf GELU.

0.5*erfc(t) = 2^(t*Q(t) - 1) with Q of degree DEG - 1: weighted least squares of log2(erfc(t)) over t in [0, 4], re-weighted
towards the minimax solution of |t * (2^P(t) - erfc(t))| (the GELU's absolute error is |x| * |0.5 erfc error|).
The second half evaluates the kernel's fp32 Horner scheme (fma emulated in float64, rounded to fp32 after every step).
"""
import numpy as np
from scipy.special import erf, erfc

T, DEG = 4.0, 5          # P(t) = t * Q(t) of degree 5 (Q of degree 4); DEG = 7 was the first version (5.1e-7)


def fit():
    t = np.linspace(0, T, 80001)
    y = np.log2(erfc(t))
    a = np.stack([t ** k for k in range(1, DEG + 1)], 1)
    w = erfc(t) * np.maximum(t, 0.05)
    for _ in range(300):
        c, *_ = np.linalg.lstsq(a * w[:, None], y * w, rcond=None)
        err = np.abs((np.exp2(a @ c) - erfc(t)) * np.maximum(t, 0.05))
        w = w * (1 + 2 * err / err.max()) ** 0.5
    return c


def gelu_kernel(x, c):
    c32 = [np.float32(v) for v in c]
    x = x.astype(np.float32)
    t = np.minimum(np.abs(x) * np.float32(0.70710678118654752), np.float32(T)).astype(np.float32)
    q = np.full_like(t, c32[-1])
    for k in range(len(c32) - 2, -1, -1):
        q = (q.astype(np.float64) * t + np.float64(c32[k])).astype(np.float32)
    p = (q.astype(np.float64) * t - 1.0).astype(np.float32)
    e = np.exp2(p.astype(np.float64)).astype(np.float32)
    return (-(np.abs(x).astype(np.float64)) * e + np.maximum(x, 0)).astype(np.float32)


if __name__ == "__main__":
    c = fit()
    print(f"c1..c{DEG} =", [float(np.float32(v)) for v in c])
    x = np.linspace(-8, 8, 400001)
    exact = 0.5 * x * (1 + erf(x / np.sqrt(2)))
    got = gelu_kernel(x, c).astype(np.float64)
    d = np.abs(got - exact)
    print("max abs error %.3e at x=%.3f; max error / (|gelu| + 1e-3) = %.3e" % (d.max(), x[d.argmax()], (d / (np.abs(exact) + 1e-3)).max()))
