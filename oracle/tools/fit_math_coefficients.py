"""Derives the polynomial coefficients used by finmath-lib_b200/csrc/fmb_math.cuh (device exp / log).

Near-minimax fits by interpolation at Chebyshev nodes in 60-digit arithmetic (mpmath), coefficients rounded to binary64,
then the rounded polynomial's maximum error is measured in high precision.  Run: python oracle/tools/fit_math_coefficients.py
"""
import mpmath as mp

mp.mp.dps = 60


def cheb_nodes(a, b, n):
    return [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]


def fit(func, a, b, deg):
    xs = cheb_nodes(mp.mpf(a), mp.mpf(b), deg + 1)
    A = mp.matrix(deg + 1, deg + 1)
    y = mp.matrix(deg + 1, 1)
    for i, x in enumerate(xs):
        for j in range(deg + 1):
            A[i, j] = x ** j
        y[i] = func(x)
    c = mp.lu_solve(A, y)
    return [float(c[j]) for j in range(deg + 1)]


def horner(c, x):
    r = mp.mpf(0)
    for v in reversed(c):
        r = r * x + mp.mpf(v)
    return r


# exp(r) = 1 + r + r^2 * Q(r),  |r| <= ln2/2 (+ a little slack)
R = mp.log(2) / 2 * mp.mpf("1.0005")
Q = fit(lambda r: (mp.exp(r) - 1 - r) / r ** 2 if r != 0 else mp.mpf(1) / 2, -R, R, 9)
err = max(abs((1 + x + x * x * horner(Q, x)) / mp.exp(x) - 1) for x in [(-R + 2 * R * k / 4001) for k in range(4002)])
print("// exp: Q degree 9, max rel err of the rounded polynomial = %s" % mp.nstr(err, 3))
print("EXP_Q = {" + ", ".join("%.17e" % v for v in Q) + "}")

# log(1+f) = 2s + s*z*P(z),  s = f/(2+f), z = s^2, z in [0, (3-2*sqrt2)^2... ] : s in [-0.1716, 0.1716]
smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1) * mp.mpf("1.0005")
zmax = smax ** 2
P = fit(lambda z: (2 * mp.atanh(mp.sqrt(z)) / mp.sqrt(z) - 2) / z if z != 0 else mp.mpf(2) / 3, mp.mpf(0), zmax, 7)
err = 0
for k in range(1, 4001):
    s = smax * k / 4000
    z = s * s
    approx = 2 * s + s * z * horner(P, z)
    err = max(err, abs(approx / (2 * mp.atanh(s)) - 1))
print("// log: P degree 7 in z = s^2, max rel err of the rounded polynomial = %s" % mp.nstr(err, 3))
print("LOG_P = {" + ", ".join("%.17e" % v for v in P) + "}")
