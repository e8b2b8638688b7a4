"""The reciprocal + FMA-residual division used by the kernels (csrc/ptb_device.cuh: div_rcp) returns the correctly
rounded IEEE quotient: the recurrence is replayed here in exact rational arithmetic for the divisors the path uses."""
import math
import random
from fractions import Fraction as F


def div_rcp(a, b):
    y = 1.0 / b                                   # RN(1/b), computed once per divisor
    q0 = a * y
    r = float(F(a) - F(q0) * F(b))                # fma(-q0, b, a)
    return float(F(q0) + F(r) * F(y))             # fma(r, y, q0)


def test_constant_divisors_match_ieee_division():
    random.seed(7)
    divisors = [18.4, 50.0, 33.04, 40000.0, 3.6, 0.511, 4.0, 6.0, 7.0, 4 * math.pi * 8.854 * 1e-12 * 11.7,
                1.9999999999999998, 1.0000000000000002]
    for b in divisors:
        for i in range(4000):
            a = random.uniform(-3e4, 3e4) if i % 2 else random.lognormvariate(0, 8)
            assert div_rcp(a, b) == a / b, (a, b)
        assert div_rcp(0.0, b) == 0.0


def test_per_particle_divisors_match_ieee_division():
    random.seed(8)
    for _ in range(30000):
        r = abs(random.gauss(0, 21.0)) + 1e-9     # cluster radius
        d = random.uniform(-120, 120)             # pixel centre - track
        assert div_rcp(d * d, r * r) == (d * d) / (r * r)
        rc = r if r >= 1 else 1 / math.sqrt(2 * math.pi)
        y2 = 0.5 * (1.0 / (r * r)) if r >= 1 else 1.0 / (2.0 * (rc * rc))
        q0 = -(d * d) * y2
        rr = float(F(-(d * d)) - F(q0) * F(2.0 * (rc * rc)))
        assert float(F(q0) + F(rr) * F(y2)) == -(d * d) / (2.0 * (rc * rc))
