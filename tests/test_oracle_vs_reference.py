"""Live pin of the CPU oracle: the unmodified reference (baseline/_ref or /root/reference, loaded through
oracle/reference_shim.py) and the oracle run side by side under the same two seeds must agree bit for bit -- every track
array, the energy-loss table after the acceptance zeroing, every hit table.  tests/test_oracle_golden.py checks the same
against committed fixtures; this one executes the reference itself (main/hit.py:8 tracks, main/device.py:7
calculate_device_hit), with seeds and sizes the fixtures do not hold.  Skipped where no reference tree or numba exists."""
import numpy as np
import pytest

from oracle import cpu_oracle as co
from oracle import reference_shim
from pytestbeam_b200 import config as cfg

pytestmark = pytest.mark.skipif(not reference_shim.available(),
                                reason="no reference tree (baseline/_ref, /root/reference) or no numba on this box")


@pytest.mark.parametrize("name,n,seeds", [("c1", 3000, (101, 202)), ("c3", 2500, (7, 8)), ("c4", 4000, (31, 41))])
def test_oracle_equals_the_reference_run_live(name, n, seeds):
    beam, devices, materials, names = cfg.flatten(getattr(cfg, name)(n))
    tracks, eloss, hits = reference_shim.run_reference(beam, devices, materials, names, *seeds)
    run = co.run(beam, devices, materials, n, *seeds)
    for i, key in enumerate(("energy", "anglex", "angley", "x", "y", "z", "time_stamp")):
        ref = np.asarray(tracks[i])
        assert np.array_equal(ref, run.tracks[i].astype(ref.dtype)), key
    assert np.array_equal(eloss, run.eloss)
    for d, dn in enumerate(names):
        ref = hits[dn]
        assert len(ref) == len(run.hits[d]), dn
        for f in ("event_number", "frame", "column", "row", "charge"):
            assert np.array_equal(ref[f], run.hits[d][f]), (dn, f)
