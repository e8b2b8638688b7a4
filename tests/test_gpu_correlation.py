"""ptb_correlate vs the oracle's restatement of the reference's correlation event loop."""
import numpy as np
import pytest

from oracle import cpu_oracle as co
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,n,kw", [("c1", 20000, {}), ("c4", 20000, {}), ("c1", 20000, dict(offset=100, event_size=3000)),
                                       ("c4", 5000, dict(event_numb_shift=2))])
def test_histograms_equal_oracle(name, n, kw):
    from pytestbeam_b200.main.correlation import correlation_histograms
    setup, run = H.oracle_run(name, n, 81, 82)
    hx, hy = run.hits[0], run.hits[-1]            # first and last device, as plot_default pairs them
    want_x, want_y = co.correlation(hx, hy, **kw)
    got_x, got_y = correlation_histograms(hx, hy, **kw)
    assert got_x.shape == want_x.shape and got_y.shape == want_y.shape
    assert np.array_equal(got_x, want_x) and np.array_equal(got_y, want_y)
    assert got_x.sum() == got_y.sum() > 0
