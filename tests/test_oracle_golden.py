"""Pin the CPU oracle (oracle/) bit-for-bit against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py; reference entry points main/hit.py:8 and main/device.py:7)."""
import glob
import hashlib
import json
import os
import re

import numpy as np
import pytest

from oracle import cpu_oracle as co
from pytestbeam_b200 import config as cfg

TRACK_NAMES = ("energy", "anglex", "angley", "x", "y", "z", "time_stamp")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _cases(golden_dir, kind):
    return sorted(glob.glob(os.path.join(golden_dir, kind + "_*")))


@pytest.mark.parametrize("path", _cases(os.path.join(os.path.dirname(__file__), "golden"), "full"))
def test_full_outputs_bit_exact(path):
    m = re.match(r"full_(c\d)_n(\d+)_s(\d+)_(\d+)\.npz", os.path.basename(path))
    name, n, a, b = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4))
    gold = np.load(path)
    beam, devices, materials, names = cfg.flatten(getattr(cfg, name)(n))
    run = co.run(beam, devices, materials, n, a, b)
    assert list(gold["names"]) == names
    for i, nm in enumerate(TRACK_NAMES):
        ref = gold[nm]
        assert np.array_equal(ref, run.tracks[i].astype(ref.dtype)), nm
    assert np.array_equal(gold["eloss"], run.eloss)
    for d, dn in enumerate(names):
        ref = gold["hits_" + dn]
        assert ref.dtype == co.HITS_DTYPE
        assert np.array_equal(ref, run.hits[d]), dn


@pytest.mark.parametrize("path", _cases(os.path.join(os.path.dirname(__file__), "golden"), "digest"))
def test_digests_at_1e5(path):
    with open(path) as f:
        dig = json.load(f)
    n = dig["n"]
    beam, devices, materials, names = cfg.flatten(getattr(cfg, dig["config"])(n))
    run = co.run(beam, devices, materials, n, dig["seed_interp"], dig["seed_numba"])
    assert names == dig["names"]
    assert _sha(run.eloss) == dig["eloss"]
    for i, nm in enumerate(TRACK_NAMES):
        assert _sha(run.tracks[i].astype(dig["track_dtypes"][nm])) == dig["tracks"][nm], nm
    for d, dn in enumerate(names):
        assert len(run.hits[d]) == dig["hits"][dn]["n"], dn
        assert _sha(run.hits[d]) == dig["hits"][dn]["sha256"], dn


def test_helper_known_answers(golden_dir):
    with open(os.path.join(golden_dir, "helpers.json")) as f:
        kat = json.load(f)
    L = co.lib()
    fh = float.fromhex
    for thick, x0, e, want in kat["highland"]:
        assert L.orc_highland(thick, x0, fh(e)) == fh(want)
    for e, z, want in kat["cloud_sigma_terms"]:
        assert abs(0.0 + L.orc_cloud_sigma(fh(e)) * fh(z)) == fh(want)
    for x, delta, pitch, number, want in kat["pixel_index"]:
        assert L.orc_pixel_index(fh(x), delta, pitch, number) == want
    for idx, delta, pitch, number, want in kat["pixel_centre"]:
        assert L.orc_pixel_centre(idx, delta, pitch, number) == fh(want)
    for rp, cp, th, z, want in kat["noise_sigma"]:
        assert 0.0 + L.orc_noise_sigma(rp, cp, th) * fh(z) == fh(want)
    for delta, args, want in kat["moyal_pdf"]:
        a = [fh(v) for v in args]
        assert co.moyal_pdf(np.array([fh(delta)]), *a)[0] == fh(want)


def test_mt_stream_matches_numpy_randomstate():
    """Numba's njit RNG == numpy's legacy RandomState stream; the oracle replays it."""
    for seed in (0, 2, 12345, 2**32 - 1):
        mt, rs = co.MT(seed), np.random.RandomState(seed)
        for _ in range(5):
            assert mt.double() == rs.random_sample()
        for _ in range(7):   # odd count: leaves one cached value behind
            assert mt.gauss() == rs.standard_normal()
        for lam in (200000.00000000003, 5.0, 10.0, 9.99, 37.5, 0.0):
            assert mt.poisson(lam) == rs.poisson(lam)
        assert mt.gauss() == rs.standard_normal()


def test_quirk_double_wide_first_pixel():
    """int() truncates toward zero, so index 1 covers (-1, 1) pitch units."""
    L = co.lib()
    assert L.orc_pixel_index(-30000.0, 0.0, 40000.0, 1) == 1
    assert L.orc_pixel_index(-59999.0, 0.0, 40000.0, 1) == 1
    assert L.orc_pixel_index(-60001.0, 0.0, 40000.0, 1) == 0
    assert L.orc_pixel_index(20001.0, 0.0, 40000.0, 1) == 2


def test_geometry_guard():
    beam, devices, materials, _ = cfg.flatten(cfg.c1(10))
    devices[0]["z_position"] = 5
    with pytest.raises(ValueError):
        co.run(beam, devices, materials, 10, 1, 2)


def test_cluster_known_answers_f8(golden_dir):
    """Whole clusters incl. the f8 charge (the /Hits table only keeps f4), device.py:139-155, 238-303."""
    with open(os.path.join(golden_dir, "helpers.json")) as f:
        kat = json.load(f)["clusters"]
    fh = float.fromhex
    silicon = dict(cfg.SILICON)
    n_hits = 0
    for c in kat:
        pc, nc, dx, pr, nr, dy, thr, th = c["geo"]
        dev = [{"column": nc, "row": nr, "column_pitch": pc, "row_pitch": pr, "thickness": th, "z_position": 0,
                "delta_x": dx, "delta_y": dy, "threshold": thr, "trigger": "untriggered", "material": "silicon"}]
        x, y, e = np.array([fh(c["x"])]), np.array([fh(c["y"])]), np.array([[fh(c["e"])]])
        hits, q64, z, off, census = co.run_digitise_plane(dev, [silicon], 0, 1, x, y, e, np.ones(1, np.uint8),
                                                          mt=co.MT(c["seed"]))
        sig = co.lib().orc_cloud_sigma(fh(c["e"]))
        assert abs(0.0 + sig * z[0]) == fh(c["rx"]) and abs(0.0 + sig * z[1]) == fh(c["ry"])
        assert list(hits["column"]) == c["col"] and list(hits["row"]) == c["row"]
        assert [float(v) for v in q64] == [fh(v) for v in c["q"]]
        n_hits += len(q64)
    assert n_hits > 30


def test_threshold_scan_known_answers(golden_dir):
    """S-curves replayed around the reference's own calc_cluster_hits (main/threshold_scan.py:42-77)."""
    with open(os.path.join(golden_dir, "helpers.json")) as f:
        kat = json.load(f)["threshold_scan"]
    for t in kat:
        pc, pr, thr, th, nc, nr = t["geo"]
        arr = np.arange(t["low"], t["high"], (t["high"] - t["low"]) / t["bins"])
        frac, z = co.threshold_scan(arr, t["inj"], pc, pr, thr, th, nc, nr, t["seed"])
        assert np.array_equal(frac, np.array(t["frac"]))
        assert z.shape == (t["bins"], t["inj"], 2)          # two normals per injection, as in the reference
        assert 0 < frac.sum() < t["bins"]                  # a real S-curve: neither all 0 nor all 1


@pytest.mark.parametrize("path", _cases(os.path.join(os.path.dirname(__file__), "golden"), "corr"))
def test_correlation_event_loop_vs_reference(path):
    """oracle orc_eventloop vs the reference's own _eventloop_fast (main/plotting.py:730-764) on its own hit tables."""
    g = np.load(path)
    full = np.load(path.replace("corr_", "full_"))
    names = list(full["names"])
    xh, yh = co.correlation(full["hits_" + names[0]], full["hits_" + names[-1]])
    xr = np.zeros(tuple(g["x_shape"]), np.int32)
    xr[tuple(g["x_idx"])] = g["x_val"]
    yr = np.zeros(tuple(g["y_shape"]), np.int32)
    yr[tuple(g["y_idx"])] = g["y_val"]
    assert np.array_equal(xh, xr) and np.array_equal(yh, yr)
