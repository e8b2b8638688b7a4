"""The regrouped Highland width of the kernels (csrc/ptb_device.cuh: highland_theta0_dev) against the reference's
operation order (main/hit.py:290-297), both replayed in IEEE double: they agree to a few ulp over every energy the
path can meet, for the three materials / thicknesses of the BASELINE configs.  (The CUDA implementation itself,
including its rsqrt, is covered by the GPU parity tests.)"""
import numpy as np


def reference_order(E, eps):
    p = np.sqrt(E * E - 0.511 * 0.511)
    beta = np.sqrt(1 - 1 / (1 + (E / 0.511) ** 2))
    return np.sqrt((13.6 / (beta * p)) ** 2 * eps * (1 + 0.038 * np.log(eps)) ** 2)


def regrouped(E, eps):
    hl_c = 13.6 * np.sqrt(eps * (1 + 0.038 * np.log(eps)) ** 2)
    g = E / 0.511
    g2 = g * g
    num = 1.0 + g2
    den = g2 * (E * E - 0.511 * 0.511)
    return hl_c * num * (1.0 / np.sqrt(num * den))


def test_regrouped_highland_width_within_a_few_ulp():
    rng = np.random.default_rng(3)
    E = np.concatenate([rng.uniform(0.6, 6000, 200_000), 5000 - rng.uniform(0, 0.6, 200_000),
                        100 - rng.uniform(0, 5, 100_000), rng.uniform(0.5111, 0.7, 2000)])
    for eps in (50 / 93650, 300 / 93650, 1000 / 5612):
        a, b = reference_order(E, eps), regrouped(E, eps)
        assert np.max(np.abs(a - b) / a) < 8 * np.finfo(float).eps


def test_known_widths():
    # SURVEY.md section 8(a5), computed with the reference's own function
    assert abs(regrouped(np.float64(5000.0), 50 / 93650) / 4.4853e-5 - 1) < 1e-4
    assert abs(regrouped(np.float64(5000.0), 300 / 93650) / 1.2035e-4 - 1) < 1e-4
    assert abs(regrouped(np.float64(100.0), 1000 / 5612) / 5.3647e-2 - 1) < 1e-4


def test_nan_below_the_electron_mass():
    with np.errstate(invalid="ignore"):
        assert np.isnan(regrouped(np.float64(0.4), 50 / 93650))
