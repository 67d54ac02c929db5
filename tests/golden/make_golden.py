"""Generates tests/golden/*.json — small known-answer fixtures that pin the oracle.

Sources of truth (all INDEPENDENT of oracle/ and of the product code):
  * MT19937: numpy's RandomState seeded with init_by_array([hi, lo]) — exactly what commons-math3 3.6.1
    MersenneTwister(long seed) does (bytecode-verified) — plus the public known-answer vector
    init_by_array{0x123,0x234,0x345,0x456} -> 1067595299 955945823 477289528 ...
  * AS241: scipy.stats.norm.ppf / mpmath (AS241 is accurate to ~1e-16, so agreement to 1e-15 relative pins the transcription).
  * tick rounding: IEEE arithmetic in Python (Math.rint == round-half-even == numpy.rint).
  * closed forms used by the reference's own tests as oracles (Black-Scholes call, T/.../MonteCarloBlackScholesModelTest.java:80).
The reference itself cannot be executed here (100 % Java, no JVM in the image): these vectors are "provisional golden vectors"
in the sense of SURVEY.md §8c, to be re-confirmed on a JVM.

Run:  python tests/golden/make_golden.py
"""
import json
import os

import mpmath as mp
import numpy as np
from scipy.stats import norm

HERE = os.path.dirname(os.path.abspath(__file__))


def words(seed, n):
    key = np.array([(seed >> 32) & 0xffffffff, seed & 0xffffffff], dtype=np.uint32)
    rs = np.random.RandomState(key)
    # RandomState.randint(..., dtype=uint32) over the full range returns successive 32-bit outputs unchanged
    return [int(v) for v in rs.randint(0, 2 ** 32, size=n, dtype=np.uint32)]


def uniforms_from_words(w):
    return [float(((w[2 * i] >> 6) << 26 | (w[2 * i + 1] >> 6)) * 2.0 ** -52) for i in range(len(w) // 2)]


def main():
    mp.mp.dps = 40
    g = {}
    kat = np.random.RandomState(np.array([0x123, 0x234, 0x345, 0x456], dtype=np.uint32))
    g["mt_kat_key"] = [0x123, 0x234, 0x345, 0x456]
    g["mt_kat_words"] = [int(v) for v in kat.randint(0, 2 ** 32, size=8, dtype=np.uint32)]
    assert g["mt_kat_words"][:3] == [1067595299, 955945823, 477289528]
    g["seeds"] = {}
    for seed in (3141, 31415, 53252, -1, -7, 2 ** 31 - 1, -2 ** 31):
        s64 = seed & 0xffffffffffffffff                      # Java int widened to long, then split into two ints
        w = words(s64, 2000)
        g["seeds"][str(seed)] = {"words_0_8": w[:8], "words_1990_2000": w[1990:2000], "uniforms_0_4": uniforms_from_words(w[:8])}
    ps = [0.5, 0.48112854170930563, 0.8554104072506368, 0.075, 0.0749999, 0.925, 0.9250001, 1e-10, 1 - 1e-10, 1e-300, 2.0 ** -52, 0.001, 0.999,
          0.3, 0.6180339887498949]
    g["icdf"] = [{"p": p, "z": float(mp.sqrt(2) * mp.erfinv(2 * mp.mpf(p) - 1)), "scipy": float(norm.ppf(p))} for p in ps]
    tick = 1.0 / (365.0 * 24.0)
    g["tick_rounding"] = [{"t": t, "rounded": float(np.rint(t / tick) * tick)} for t in (0.1, 0.5, 0.05, 0.25, 1.0 / 3.0, 0.005, 20.0, 0.001, 7.35)]
    d1 = (np.log(1 / 1.05) + (0.05 + 0.045) * 5) / (0.3 * np.sqrt(5))
    g["black_scholes_call"] = {"S0": 1.0, "r": 0.05, "sigma": 0.3, "T": 5.0, "K": 1.05,
                               "value": float(norm.cdf(d1) - 1.05 * np.exp(-0.25) * norm.cdf(d1 - 0.3 * np.sqrt(5)))}
    with open(os.path.join(HERE, "mt_as241.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", os.path.join(HERE, "mt_as241.json"))


if __name__ == "__main__":
    main()
