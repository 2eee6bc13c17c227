"""CPU check of the polynomial constants in bayhunter_b200/csrc/bh_math.cuh: the table is parsed from the
header and evaluated in exact rational arithmetic, so a mistyped digit cannot hide behind GPU rounding.
(The device sequences themselves are checked in ulps on the GPU: test_device_elementary_functions_accuracy.)"""
import os
import re
import struct
from fractions import Fraction

import pytest

mp = pytest.importorskip("mpmath")
HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bayhunter_b200", "csrc", "bh_math.cuh")


def _table():
    src = open(HDR).read()
    names = re.search(r"enum \{(.*?)K_COUNT", src, re.S).group(1)
    names = [n.strip().split("=")[0].strip() for n in re.sub(r"//.*", "", names).split(",") if n.strip()]
    body = re.search(r"kTab\[K_COUNT\] = \{(.*?)\};", src, re.S).group(1)
    vals = [float(v) for v in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", re.sub(r"//.*", "", body))]
    assert len(names) == len(vals), (len(names), len(vals))
    tab = dict(zip(names, vals))
    for name, hi in re.findall(r"#define (BH_K_\w+) BH_KHI\(0x([0-9a-fA-F]+)\)", src):
        tab[name] = struct.unpack(">d", struct.pack(">II", int(hi, 16), 0))[0]
    return tab


def test_exp_polynomial_is_accurate_to_a_fraction_of_an_ulp():
    # approximation error of 1 + r + r^2 Q(r) with the constants as the kernels use them (q9, q8 high words only, q0 = 0.5)
    t = _table()
    q = [Fraction(0.5)] + [Fraction(t["K_Q%d" % j]) for j in range(1, 8)] + [Fraction(t["BH_K_Q8"]), Fraction(t["BH_K_Q9"])]
    mp.mp.prec = 200
    a = mp.log(2) / 2
    worst = 0
    for k in range(-1000, 1001):
        r = Fraction(float(a * k / 1000))
        acc = Fraction(0)
        for c in reversed(q):
            acc = acc * r + c
        v = 1 + r + r * r * acc                         # exp_small: fma(fma(Q, r, 1), r, 1)
        ex = mp.exp(mp.mpf(r.numerator) / r.denominator)
        worst = max(worst, abs((mp.mpf(v.numerator) / v.denominator - ex) / ex))
    assert worst < 5e-17, worst                         # measured 3.9e-17 = 0.35 ulp of 1


def test_reduction_constants():
    t = _table()
    mp.mp.prec = 200
    assert t["K_MAGIC"] == 1.5 * 2.0 ** 52
    assert abs(mp.mpf(t["K_LOG2E"]) - 1 / mp.log(2)) < 2e-16
    assert abs(mp.mpf(t["K_LN2_HI"]) + mp.mpf(t["K_LN2_LO"]) - mp.log(2)) < 1e-32
    assert abs(mp.mpf(t["K_TWO_OVER_PI"]) - 2 / mp.pi) < 1e-16
    assert abs(mp.mpf(t["K_PIO2_1"]) + mp.mpf(t["K_PIO2_2"]) - mp.pi / 2) < 2e-33


def test_sincos_kernels_on_the_reduced_interval():
    t = _table()
    s = [Fraction(t["K_S%d" % j]) for j in range(1, 6)] + [Fraction(t["BH_K_S6"])]
    c = [Fraction(t["K_C%d" % j]) for j in range(1, 6)] + [Fraction(t["BH_K_C6"])]
    mp.mp.prec = 200
    ws = wc = 0
    for k in range(-500, 501):
        r = Fraction(float(mp.pi / 4 * k / 500))
        z = r * r
        ps = Fraction(0); pc = Fraction(0)
        for a, b in zip(reversed(s), reversed(c)):
            ps = ps * z + a; pc = pc * z + b
        sn = r + r * z * ps
        cs = 1 + z * (z * pc - Fraction(1, 2))
        x = mp.mpf(r.numerator) / r.denominator
        ws = max(ws, abs(mp.mpf(sn.numerator) / sn.denominator - mp.sin(x)))
        wc = max(wc, abs(mp.mpf(cs.numerator) / cs.denominator - mp.cos(x)))
    assert ws < 2e-17 and wc < 2e-17, (ws, wc)
