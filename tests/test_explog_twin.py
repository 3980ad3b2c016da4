"""NumPy twin of the branch-free exp(-x) / log(y) in csrc/bigclam_kernels.cuh (exp_neg, log_pos), with FMA
emulated in extended precision, checked against mpmath: the formulas themselves (range reduction, Estrin
polynomials, reconstruction) are accurate to <= 2 ulp on the domain the kernel evaluates them on."""
import math

import numpy as np
import pytest

mp = pytest.importorskip("mpmath")
L = np.longdouble


def fma(a, b, c):
    return (L(a) * L(b) + L(c)).astype(np.float64)


def exp_neg(x):
    magic = 6755399441055744.0
    t = fma(-x, 1.4426950408889634, magic)
    n = t - magic
    r = fma(n, -0.6931471805599453094, -x)
    r = fma(n, -2.3190468138462996e-17, r)
    r2 = r * r
    r4 = r2 * r2
    r8 = r4 * r4
    a0 = r + 1.0
    a1 = fma(0.16666666666666666, r, 0.5)
    a2 = fma(0.008333333333333333, r, 0.041666666666666664)
    a3 = fma(0.0001984126984126984, r, 0.001388888888888889)
    a4 = fma(2.7557319223985893e-06, r, 2.48015873015873e-05)
    a5 = fma(2.505210838544172e-08, r, 2.755731922398589e-07)
    a6 = fma(1.6059043836821613e-10, r, 2.08767569878681e-09)
    b0, b1, b2 = fma(a1, r2, a0), fma(a3, r2, a2), fma(a5, r2, a4)
    d0, d1 = fma(b1, r4, b0), fma(a6, r4, b2)
    p = fma(d1, r8, d0)
    return np.ldexp(p, n.astype(np.int64))


def log_pos(y):
    m, e = np.frexp(y)
    m, e = m * 2.0, e - 1
    hi = m.view(np.int64) >> 32                          # the kernel compares the high word with 0x3ff6a09e
    big = hi > 0x3ff6a09e
    m = np.where(big, m * 0.5, m)
    e = e + big
    f, d = m - 1.0, m + 1.0
    rd = (L(1.0) / L(d)).astype(np.float64)              # rcp.approx + two Newton steps ~ correctly rounded
    s = f * rd
    s = fma(fma(-d, s, f), rd, s)
    z = s * s
    z2 = z * z
    z4 = z2 * z2
    z8 = z4 * z4
    a0 = fma(0.4, z, 0.6666666666666666)
    a1 = fma(0.2222222222222222, z, 0.2857142857142857)
    a2 = fma(0.15384615384615385, z, 0.18181818181818182)
    a3 = fma(0.11764705882352941, z, 0.13333333333333333)
    a4 = fma(0.09523809523809523, z, 0.10526315789473684)
    b0, b1, b2 = fma(a1, z2, a0), fma(a3, z2, a2), fma(0.08695652173913043, z2, a4)
    q = fma(b2, z8, fma(b1, z4, b0))
    ed = e.astype(np.float64)
    inner = fma(ed, 2.3190468138462996e-17, (s * z) * q)
    return fma(ed, 0.6931471805599453094, fma(2.0, s, inner))


def test_exp_neg_and_log_pos_twin_accuracy():
    mp.mp.prec = 200
    rng = np.random.default_rng(0)
    x_lo, x_hi = -math.log(0.9999), -math.log(0.0001)            # the kernel's evaluation window
    x = np.concatenate([rng.uniform(x_lo, x_hi, 4000), 10 ** rng.uniform(-4, math.log10(x_hi), 4000)])
    x = x[(x > x_lo) & (x < x_hi)]
    ref = np.array([float(mp.exp(-mp.mpf(float(v)))) for v in x])
    assert (np.abs(exp_neg(x) - ref) / np.spacing(ref)).max() <= 2.0
    y = np.concatenate([rng.uniform(1e-4, 0.9999, 4000), 1 - 10 ** rng.uniform(-4, -0.01, 2000), 10 ** rng.uniform(-4, 0, 2000)])
    y = y[(y >= 1e-4) & (y <= 0.9999)]
    ref = np.array([float(mp.log(mp.mpf(float(v)))) for v in y])
    assert (np.abs(log_pos(y) - ref) / np.spacing(np.abs(ref))).max() <= 2.0
    # the composed edge term log(1 - exp(-x)) + x: 1 - exp(-x) cancels for small x (in the reference's own
    # formula too), which amplifies exp's last-bit error by 1/x
    t = log_pos(1.0 - exp_neg(x)) + x
    ref = np.array([float(mp.log(1 - mp.exp(-mp.mpf(float(v)))) + mp.mpf(float(v))) for v in x])
    assert (np.abs(t - ref) <= 4e-16 * np.maximum(np.abs(ref), 1.0) + 4e-16 / x).all()
