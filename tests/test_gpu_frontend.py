"""The reference's entry point and its two hot-path functions, served by the CUDA library."""
import logging
import os

import numpy as np
import pytest
import yaml

from pytestbeam_b200 import config as cfg
from pytestbeam_b200 import h5min

pytestmark = pytest.mark.gpu
LOG = logging.getLogger("test")


def test_entry_point_writes_the_reference_files(tmp_path):
    from pytestbeam_b200 import pytestbeam
    from pytestbeam_b200.engine import Simulator
    setup = cfg.c4(30_000)
    setup["beam"]["seed"] = 777
    setup["saving_tracks"] = True
    setup["plotting"] = False
    (tmp_path / "setup.yml").write_text(yaml.safe_dump(setup, sort_keys=False))
    (tmp_path / "material.yml").write_text(yaml.safe_dump(cfg.default_material()))
    setup["beam"]["chunk"] = 7_000                      # optional key: five chunks, the last one short
    (tmp_path / "setup.yml").write_text(yaml.safe_dump(setup, sort_keys=False))
    pytestbeam.main(str(tmp_path))
    beam, devices, materials, names = cfg.flatten(setup)
    sim = Simulator(beam, devices, materials)
    ref = sim.run(0, 30_000, seed=777, write_tracks=True)
    out = tmp_path / "output_data"
    import json
    rec = json.loads((out / "run_record.json").read_text())
    assert rec["electrons"] == 30_000 and rec["seed"] == 777 and rec["chunk"] == 7_000 and rec["route"] == "fused"
    assert [rec["rows"][nm] for nm in names] == ref.totals["hits"] and rec["electrons_per_s"] > 0
    for d, name in enumerate(names):
        got = h5min.read_table(str(out / f"{name}_dut.h5"), "Hits")
        assert got.tobytes() == ref.hits_numpy(d).tobytes(), name
    trk = h5min.read_table(str(out / "particle_tracks.h5"), "Tracks")
    assert len(trk) == 30_000 * len(names)
    assert np.array_equal(trk["x"], ref.tracks["x"].cpu().numpy())
    assert np.array_equal(trk["time_stamp"], ref.tracks["time_stamp"].cpu().numpy().astype(np.uint16))


def test_two_call_api_equals_fused_path():
    """hit.tracks -> (touch the tuple) -> device.calculate_device_hit must give what the fused pass gives."""
    from pytestbeam_b200.main import device, hit
    n = 20_000
    setup = cfg.c3(n)
    setup["beam"]["seed"] = 4242
    beam, devices, materials, names = cfg.flatten(setup)
    lazy = hit.tracks(beam, devices, materials, LOG)
    fused = device.device_hits(beam, devices, lazy, LOG)
    assert not lazy.materialised
    tables = hit.tracks(beam, devices, materials, LOG)
    energy, ax, ay, x, y, z, t, eloss = tables               # materialises, like the reference's tuple
    assert len(x) == n * len(devices) and eloss.shape == (len(devices), n)
    assert np.all(np.diff(t.reshape(n, len(devices))[:, 0]) >= 0)
    assert np.array_equal(x[3::len(devices)], x.reshape(n, -1)[:, 3])      # the reference's stride-D slicing
    split = device.device_hits(beam, devices, tables, LOG)
    for d in range(len(devices)):
        assert split[d].tobytes() == fused[d].tobytes(), names[d]


def test_foreign_arrays_with_explicit_mask():
    from pytestbeam_b200.engine import Simulator
    from pytestbeam_b200.main import device
    from tests import helpers as H
    n = 5000
    setup, run = H.oracle_run("c1", n, 61, 62)
    sim = Simulator(*setup[:3])
    parts = device.device_hits_from_arrays(sim, run.tracks[3], run.tracks[4], run.eloss, seed=5,
                                           accepted=run.variates.accepted)
    # production noise, oracle tracks: same event numbering and the same particles hit, different charges
    for d in range(sim.D):
        assert abs(len(parts[d]) - len(run.hits[d])) < 0.1 * len(run.hits[d]) + 50
        if len(parts[d]):
            assert parts[d]["event_number"].max() <= int(run.variates.accepted.sum()) + 1


def test_optional_keys_refuse_what_is_not_built(tmp_path):
    from pytestbeam_b200 import pytestbeam
    for key, value in (("backend", "cpu"), ("precision", "f32")):
        setup = cfg.c1(100)
        setup[key] = value
        (tmp_path / "setup.yml").write_text(yaml.safe_dump(setup, sort_keys=False))
        (tmp_path / "material.yml").write_text(yaml.safe_dump(cfg.default_material()))
        with pytest.raises(ValueError, match=key):
            pytestbeam.load_job(str(tmp_path))
