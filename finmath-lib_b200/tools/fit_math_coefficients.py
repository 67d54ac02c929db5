"""Derives the polynomial coefficients used by finmath-lib_b200/csrc/fmb_math.cuh (device exp / log).

Near-minimax fits by interpolation at Chebyshev nodes in 60-digit arithmetic (mpmath), coefficients rounded to binary64,
then the rounded polynomial's maximum error is measured in high precision.  Run: python finmath-lib_b200/tools/fit_math_coefficients.py
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

# log: table-driven.  x = 2^k z, z in [0.6875, 1.375) cut into 128 intervals of equal bit-pattern length; log z = log c_i + log1p(r),
# r = z * invc_i - 1.  logc is split into a multiple of 2^-43 (k*ln2_hi + logc_hi is then exact) and the rest.  The two intervals next
# to 1 use invc = 1 (r = z - 1 exact, relative accuracy kept as x -> 1), which sets the polynomial's range to |r| <= 2^-7.
import struct

OFF_HI = 0x3fe60000


def from_hi(hi):
    return struct.unpack("<d", struct.pack("<Q", hi << 32))[0]


tab, rmax = [], mp.mpf(0)
for i in range(128):
    lo_bits = OFF_HI + (i << 13)
    a, b = mp.mpf(from_hi(lo_bits)), mp.mpf(from_hi(lo_bits + (1 << 13)))
    invc = 1.0 if i in (79, 80) else float(1 / ((a + b) / 2))
    logc = -mp.log(mp.mpf(invc))
    hi = float(mp.nint(logc * mp.mpf(2) ** 43) / mp.mpf(2) ** 43)
    tab.append((invc, hi, float(logc - mp.mpf(hi))))
    rmax = max(rmax, abs(a * mp.mpf(invc) - 1), abs(b * mp.mpf(invc) - 1))


def g(r):
    if abs(r) < mp.mpf(10) ** -15:
        return -mp.mpf(1) / 2 + r / 3 - r * r / 4
    return (mp.log1p(r) - r) / r ** 2


R = rmax * mp.mpf("1.001")
A = fit(g, -R, R, 5)
err = 0
for k in range(1, 4001):
    for sgn in (-1, 1):
        r = sgn * R * k / 4000
        err = max(err, abs((r + r * r * horner(A, r)) / mp.log1p(r) - 1))
print("// log: |r| <= %s, A degree 5, max rel err of the rounded polynomial against log1p(r) = %s" % (mp.nstr(rmax, 6), mp.nstr(err, 3)))
print("LOG_A (Horner order A5..A0) = {" + ", ".join("%.17e" % v for v in reversed(A)) + "}")
print("LOG_TAB = {   // [0,128) invc, [128,256) logc_hi, [256,384) logc_lo")
for col in range(3):
    vals = [t[col] for t in tab]
    for r0 in range(0, 128, 4):
        print("\t" + ", ".join("%.17e" % v for v in vals[r0:r0 + 4]) + ",")
print("};")
