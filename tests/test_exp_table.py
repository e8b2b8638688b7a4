"""exp_nonpos of csrc/ptb_device.cuh (the Gaussian profile's exponential, main/device.py:378), replayed operation by
operation with correctly rounded fused multiply-adds in exact rational arithmetic: against the exact value it must stay
within 1.5 ulp on the range the digitisation uses, and its constants must be the ones the source states."""
import math
import os
import re
from decimal import Decimal, getcontext
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "pytestbeam_b200", "csrc", "ptb_device.cuh")).read()
getcontext().prec = 60


def fma(a, b, c):
    return float(Fraction(a) * Fraction(b) + Fraction(c))     # one rounding, as the hardware DFMA


def _constants():
    body = SRC[SRC.index("double exp_nonpos("):SRC.index("// exp(-(d^2)/(2 sigma^2)), second factor")]
    code = re.sub(r"//.*", "", body.split("#else")[1])
    nums = [float(x) for x in re.findall(r"(?<![\w.])-?\d+\.\d+(?:e[+-]?\d+)?", code)]
    return nums


def exp_nonpos(x, tab):
    c = _constants()
    clamp, scale, magic = c[0], c[1], c[2]
    assert (clamp, magic) == (-700.0, 6755399441055744.0)
    head, tail = c[4], c[5]
    x = max(x, clamp)
    t = fma(x, scale, magic)
    k = int(np.float64(t).view(np.int64) & 0xFFFFFFFF)
    k = k - (1 << 32) if k >= (1 << 31) else k
    kd = t - magic
    r = fma(kd, head, x)
    r = fma(kd, tail, r)
    p = fma(r, c[6], c[7])
    p = fma(p, r, c[8])
    p = fma(p, r, 0.5)
    p = fma(p, r * r, r)
    T = tab[k & 63]
    y = fma(T, p, T)
    return math.ldexp(y, k >> 6)


def test_constants_are_what_the_comments_say():
    c = _constants()
    ln2 = Decimal(2).ln()
    assert c[1] == float(64 / ln2)
    assert c[4] == -float(ln2 / 64)
    assert c[5] == -float(ln2 / 64 - Decimal(float(ln2 / 64)))
    assert (c[6], c[7], c[8]) == (1 / 120, 1 / 24, 1 / 6)


def test_within_one_and_a_half_ulp_of_the_exact_exponential():
    tab = [float(Decimal(2) ** (Decimal(j) / 64)) for j in range(64)]
    rs = np.random.RandomState(7)
    xs = np.concatenate([-rs.uniform(0, 4, 1500), -rs.uniform(0, 60, 500), -10.0 ** rs.uniform(-12, 0, 300),
                         [0.0, -0.5, -math.log(2) / 128, -math.log(2) / 64, -700.0, -699.99]])
    worst = 0.0
    for x in xs:
        got = exp_nonpos(float(x), tab)
        want = Decimal(float(x)).exp()
        ulp = Decimal(math.ulp(float(want)))
        worst = max(worst, float(abs(Decimal(got) - want) / ulp))
    assert worst < 1.5, worst


def test_clamped_below_minus_700():
    tab = [float(Decimal(2) ** (Decimal(j) / 64)) for j in range(64)]
    tiny = exp_nonpos(-1e9, tab)
    assert 0.0 < tiny < 1e-300 and tiny == exp_nonpos(-700.0, tab)
