"""BASELINE config 2 at full size (1e8 electrons, default telescope, production RNG) through size-independent
properties: ordering, ranges, event-number bookkeeping, and a checksum of checksums that must not depend on how the
range is cut into chunks."""
import numpy as np
import pytest

from pytestbeam_b200 import config as cfg

pytestmark = pytest.mark.gpu


def _sums(sim, n_total, chunk, seed):
    import torch
    from pytestbeam_b200.pipeline import ChunkPipeline
    from pytestbeam_b200.engine import Simulator
    D = sim.D
    pipe = ChunkPipeline(sim, chunk, seed=seed, n_buffers=1)
    pipe.set_bases(0, 0)
    acc = {"rows": np.zeros(D, dtype=np.int64), "col": np.zeros(D, dtype=np.int64), "row": np.zeros(D, dtype=np.int64),
           "event": np.zeros(D, dtype=np.int64), "accepted": 0, "pixels": 0}
    last_event = [0] * D
    for first in range(0, n_total, chunk):
        m = min(chunk, n_total - first)
        s = pipe.sets[0]
        sim.launch(first, m, seed=seed, flags=pipe.flags, slabs=s["slabs"], bases_in=pipe.bases, bases_out=pipe.bases,
                   totals=s["totals"], workspace=s["ws"], scratch_rows=pipe.rows)
        tot = Simulator.decode_totals(s["totals"], D)
        assert tot["error"] == 0
        acc["accepted"] += tot["accepted"]
        acc["pixels"] += tot["pixels"]
        for d in range(D):
            k = tot["hits"][d]
            ev = s["slabs"][d]["event"][:k]
            col = s["slabs"][d]["col"][:k].to(torch.int64) & 0xFFFF
            row = s["slabs"][d]["row"][:k].to(torch.int64) & 0xFFFF
            assert bool((ev[1:] >= ev[:-1]).all()), "event numbers must be non-decreasing (particle order)"
            assert int(ev[0]) >= last_event[d] and int(ev[0]) >= 1
            last_event[d] = int(ev[-1])
            assert int(col.min()) >= 1 and int(col.max()) <= sim.devices[d]["column"]
            assert int(row.min()) >= 1 and int(row.max()) <= sim.devices[d]["row"]
            assert bool((s["slabs"][d]["charge"][:k] >= sim.devices[d]["threshold"] * 1e-3).all())
            acc["rows"][d] += k
            acc["col"][d] += int(col.sum())
            acc["row"][d] += int(row.sum())
            acc["event"][d] += int(ev.sum())
    return acc, last_event


def test_config2_full_size_invariants():
    from pytestbeam_b200.engine import Simulator
    n = 100_000_000
    beam, devices, materials, names = cfg.flatten(cfg.c2(n))
    sim = Simulator(beam, devices, materials)
    a, last_a = _sums(sim, n, 10_000_000, seed=5)
    b, last_b = _sums(sim, n, 6_250_000, seed=5)
    for k in ("rows", "col", "row", "event"):
        assert np.array_equal(a[k], b[k]), k                     # chunking must not change a single row
    assert a["accepted"] == b["accepted"] and a["pixels"] == b["pixels"] and last_a == last_b
    # trigger fraction 0.7 (main/device.py:56) and the event-number bookkeeping (device.py:159,164)
    assert abs(a["accepted"] / n - 0.7) < 5 * np.sqrt(0.21 / n)
    for d in range(sim.D):
        assert last_a[d] <= a["accepted"] + 1
    hpp = a["rows"] / n
    assert np.all(np.abs(hpp[:6] - 1.196) < 0.02) and abs(hpp[6] - 0.189) < 0.01
    # determinism of the whole job
    c, _ = _sums(sim, n // 10, 10_000_000, seed=5)
    d, _ = _sums(sim, n // 10, 10_000_000, seed=6)
    assert not np.array_equal(c["col"], d["col"])
