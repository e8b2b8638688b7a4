"""Minimal HDF5 writer/reader: structure per the HDF5 File Format Specification (v0 superblock, v1 headers)."""
import struct

import numpy as np
import pytest

from pytestbeam_b200 import h5min
from pytestbeam_b200.engine import HITS_DTYPE
from pytestbeam_b200.main.hit import TRACKS_DTYPE


def _hits(n, seed=0):
    rs = np.random.RandomState(seed)
    a = np.zeros(n, dtype=HITS_DTYPE)
    a["event_number"] = np.sort(rs.randint(1, 1 << 40, n))
    a["column"] = rs.randint(1, 1153, n)
    a["row"] = rs.randint(1, 577, n)
    a["charge"] = rs.uniform(0.1, 30, n).astype(np.float32)
    return a


@pytest.mark.parametrize("n", [0, 1, 7, 100_003])
def test_hits_round_trip(tmp_path, n):
    a = _hits(n)
    path = str(tmp_path / "x_dut.h5")
    h5min.write_table(path, "Hits", a)
    b = h5min.read_table(path, "Hits")
    assert b.dtype.names == a.dtype.names and b.dtype.itemsize == 18
    assert b.tobytes() == a.tobytes()


def test_tracks_round_trip_and_uint16_wrap(tmp_path):
    n = 50
    t = np.zeros(n, dtype=TRACKS_DTYPE)
    t["energy"] = np.linspace(4999, 5000, n)
    t["time_stamp"] = (np.arange(n) * 200_000).astype(np.uint16)     # wraps like main/hit.py:456
    path = str(tmp_path / "particle_tracks.h5")
    h5min.write_table(path, "Tracks", t)
    back = h5min.read_table(path, "Tracks")
    assert back.tobytes() == t.tobytes()
    assert back["time_stamp"][1] == 200_000 % 65536


def test_on_disk_structure(tmp_path):
    a = _hits(10)
    path = str(tmp_path / "s.h5")
    h5min.write_tables(path, {"Hits": a, "Other": a[:3]})
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n"
    assert raw[8] == 0 and raw[13] == 8 and raw[14] == 8                 # superblock v0, 8-byte offsets/lengths
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)              # end-of-file address
    assert raw.count(b"TREE") >= 1 and raw.count(b"HEAP") >= 1 and raw.count(b"SNOD") >= 1
    both = h5min.read_tables(path)
    assert set(both) == {"Hits", "Other"} and len(both["Other"]) == 3
    # compound datatype message: class 6, version 1, 5 members, 18 bytes
    i = raw.index(b"event_number")
    assert raw[i - 8] == 0x16 and raw[i - 7] == 5 and struct.unpack_from("<I", raw, i - 4)[0] == 18


def test_rejects_non_tables(tmp_path):
    with pytest.raises(ValueError):
        h5min.write_table(str(tmp_path / "bad.h5"), "Hits", np.zeros((3, 3)))
