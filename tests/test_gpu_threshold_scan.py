"""Threshold scan (reference main/threshold_scan.py:42-77) served by k_digitise with PTB_ZERO_RADIUS."""
import numpy as np
import pytest
from scipy import special

from oracle import cpu_oracle as co

pytestmark = pytest.mark.gpu

ITKPIX = dict(column_pitch=50, row_pitch=50, threshold=1000, thickness=300, column=400, row=384)
MIMOSA = dict(column_pitch=18.4, row_pitch=18.4, threshold=100, thickness=50, column=1152, row=576)


@pytest.mark.parametrize("dut,low,high,bins,inj,seed", [(ITKPIX, 950, 1050, 40, 200, 5), (MIMOSA, 60, 140, 32, 150, 6)])
def test_deterministic_scan_equals_oracle(dut, low, high, bins, inj, seed):
    from pytestbeam_b200.main import threshold_scan as ts
    arr = ts.create_injection_array(low, high, bins)
    want, z = co.threshold_scan(arr, inj, dut["column_pitch"], dut["row_pitch"], dut["threshold"], dut["thickness"],
                                dut["column"], dut["row"], seed)
    got = ts.create_hits(arr, inj, dut["column_pitch"], dut["row_pitch"], dut["threshold"], dut["thickness"],
                         dut["column"], dut["row"], injected_z=z)
    assert np.array_equal(got, want)


def test_production_scan_follows_the_error_function():
    from pytestbeam_b200.main import threshold_scan as ts
    inj_n = 20000
    arr, frac = ts.threshold_scan(ITKPIX, 960, 1040, 16, inj_n, seed=11)
    sigma = co.lib().orc_noise_sigma(50, 50, 300)            # 11.8 electrons
    expect = 0.5 * special.erfc((1000 - arr) / (np.sqrt(2) * sigma))
    err = np.sqrt(np.maximum(expect * (1 - expect), 1e-6) / inj_n)
    assert np.all(np.abs(frac - expect) < 5 * err + 1e-4)
    again = ts.threshold_scan(ITKPIX, 960, 1040, 16, inj_n, seed=11)[1]
    assert np.array_equal(frac, again)
