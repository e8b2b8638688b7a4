"""Minimal HDF5 writer/reader: structure per the HDF5 File Format Specification (v0 superblock, v1 headers).

The reader is pinned twice to bytes the writer never touched: a file written by libhdf5 itself (MATLAB 7.3 sample from
scipy's test data, copied to tests/golden/) and a chunked table assembled here by hand from the specification."""
import os
import struct

import numpy as np
import pytest

from pytestbeam_b200 import h5min
from pytestbeam_b200.engine import HITS_DTYPE
from pytestbeam_b200.main.hit import TRACKS_DTYPE

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _hits(n, seed=0):
    rs = np.random.RandomState(seed)
    a = np.zeros(n, dtype=HITS_DTYPE)
    a["event_number"] = np.sort(rs.randint(1, 1 << 40, n))
    a["column"] = rs.randint(1, 1153, n)
    a["row"] = rs.randint(1, 577, n)
    a["charge"] = rs.uniform(0.1, 30, n).astype(np.float32)
    return a


@pytest.mark.parametrize("n", [0, 1, 7, 511, 512, 513, 100_003])
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


def test_on_disk_structure_contiguous(tmp_path):
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


def test_table_is_chunked_extendable_and_carries_the_pytables_attributes(tmp_path):
    """What tables.Table would find: CLASS/VERSION/TITLE/FIELD_*/NROWS (main/device.py:329-340 creates a Table)."""
    a = _hits(3000)
    path = str(tmp_path / "p_dut.h5")
    h5min.write_table(path, "Hits", a, chunk_rows=1024)
    with h5min.TableReader(path) as r:
        info = r._describe("Hits")
        assert info["kind"] == "chunked" and info["chunk"] == [1024, 18]
        assert info["dims"] == [3000] and info["maxdims"] == [h5min.UNDEF]      # H5S_UNLIMITED
        chunks = r._chunks(info)
        assert [c[0] for c in chunks] == [0, 1024, 2048] and all(c[2] == 1024 * 18 for c in chunks)
        at = r.attrs("Hits")
        assert at["CLASS"] == "TABLE" and at["VERSION"] == "2.7" and at["TITLE"] == "" and at["NROWS"] == 3000
        assert [at[f"FIELD_{i}_NAME"] for i in range(5)] == list(HITS_DTYPE.names)
        assert all(at[f"FIELD_{i}_FILL"] == 0 for i in range(5))
        root = r.root_attrs()
        assert root["CLASS"] == "GROUP" and root["PYTABLES_FORMAT_VERSION"] == "2.1"
    raw = open(path, "rb").read()
    assert struct.unpack_from("<Q", raw, 40)[0] == len(raw)


def test_streaming_writer_append_reserve_and_deep_chunk_tree(tmp_path):
    """Pieces appended and rows written in place through the mapped region; > 64 chunks forces a two-level B-tree."""
    a = _hits(20_000, seed=3)
    path = str(tmp_path / "stream_dut.h5")
    with h5min.TableWriter(path, "Hits", HITS_DTYPE, chunk_rows=128) as w:
        w.append(a[:1000])
        view = w.reserve(7000)
        view[:] = a[1000:8000]
        view.flush()
        del view
        w.append(a[8000:8001])
        w.append(a[8001:])
    with h5min.TableReader(path) as r:
        info = r._describe("Hits")
        assert len(r._chunks(info)) == (20_000 + 127) // 128 == 157
        root = r._b.base + info["btree"]
        assert bytes(r._b.d[root:root + 4]) == b"TREE" and r._b.d[root + 5] == 1       # root node at level 1
        pieces = list(r.iter_rows("Hits", 1000))
        assert sum(len(p) for p in pieces) == 20_000
        assert r.read("Hits").tobytes() == a.tobytes()


def test_concatenate_tables_appends_part_files_in_order(tmp_path):
    parts, arrays = [], []
    for i, n in enumerate([1500, 0, 700]):
        a = _hits(n, seed=10 + i)
        p = str(tmp_path / f"part{i}.h5")
        h5min.write_table(p, "Hits", a, chunk_rows=512)
        parts.append(p)
        arrays.append(a)
    dest = str(tmp_path / "merged_dut.h5")
    n = h5min.concatenate_tables(dest, "Hits", parts, HITS_DTYPE, chunk_rows=600)
    assert n == 2200
    assert h5min.read_table(dest, "Hits").tobytes() == np.concatenate(arrays).tobytes()


def test_reader_on_a_file_written_by_libhdf5():
    """tests/golden/libhdf5_sample_matlab73.mat: scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat, written by MATLAB 7.4
    through libhdf5 (512-byte user block, v0 superblock with a base address, v1 object header, local heap, group B-tree,
    float datatype, v2 contiguous layout, a string attribute).  scipy's own expectation for it: arange(0, 2.1 pi, pi/4)."""
    path = os.path.join(GOLDEN, "libhdf5_sample_matlab73.mat")
    with h5min.TableReader(path) as r:
        assert list(r.objects()) == ["testdouble"]
        assert r.attrs("testdouble") == {"MATLAB_class": "double"}
        got = r.read("testdouble")
    assert got.shape == (9, 1) and got.dtype == np.dtype("<f8")
    assert np.array_equal(got.ravel(), np.arange(0, np.pi * 2.1, np.pi / 4))


def test_reader_on_a_chunked_table_assembled_by_hand(tmp_path):
    """A file built here, byte by byte from the HDF5 File Format Specification (no h5min writer code involved): one
    chunked rank-1 compound dataset "Hits" of 5 rows in chunks of 2 rows, the chunks stored out of order and indexed by
    a two-level v1 B-tree; one scalar int64 attribute."""
    UNDEF = 0xFFFFFFFFFFFFFFFF
    rows = _hits(5, seed=7)
    chunk_rows, isz = 2, 18

    def pad8(b):
        return b + b"\0" * (-len(b) % 8)

    def msg(t, body):
        body = pad8(body)
        return struct.pack("<HHBBBB", t, len(body), 0, 0, 0, 0) + body

    # datatype message: compound (class 6, version 1), 5 members, 18 bytes
    def member(name, off, enc):
        return pad8(name + b"\0") + struct.pack("<I", off) + b"\0" * 28 + enc
    i64 = bytes([0x10, 0x08, 0, 0]) + struct.pack("<IHH", 8, 0, 64)
    u16 = bytes([0x10, 0x00, 0, 0]) + struct.pack("<IHH", 2, 0, 16)
    f32 = bytes([0x11, 0x20, 31, 0]) + struct.pack("<IHHBBBBI", 4, 0, 32, 23, 8, 0, 23, 127)
    dtype = bytes([0x16, 5, 0, 0]) + struct.pack("<I", isz)
    dtype += member(b"event_number", 0, i64) + member(b"frame", 8, u16) + member(b"column", 10, u16)
    dtype += member(b"row", 12, u16) + member(b"charge", 14, f32)
    space = bytes([1, 1, 1, 0, 0, 0, 0, 0]) + struct.pack("<QQ", 5, UNDEF)
    attr = bytes([1, 0]) + struct.pack("<HHH", 6, 12, 8) + pad8(b"NROWS\0") + pad8(i64) + bytes([1, 0, 0, 0, 0, 0, 0, 0])
    attr += struct.pack("<q", 5)

    # addresses chosen by hand
    A_ROOT, A_HEAP, A_HEAPDATA, A_GTREE, A_SNOD, A_DSET = 96, 136, 168, 208, 752, 1080
    A_CHUNK = {0: 2400, 1: 2300, 2: 2200}                      # chunk i -> address: deliberately descending
    A_LEAF0, A_LEAF1, A_TOP = 2600, 4200, 5800
    node_size = 24 + 64 * 8 + 65 * 24
    EOF = A_TOP + node_size

    layout = bytes([3, 2, 2]) + struct.pack("<QII", A_TOP, chunk_rows, isz)
    dset_msgs = [msg(1, space), msg(3, dtype), msg(5, bytes([2, 3, 0, 0])), msg(8, layout), msg(0xC, attr)]
    dset = struct.pack("<BBHII", 1, 0, len(dset_msgs), 1, sum(map(len, dset_msgs))) + b"\0" * 4 + b"".join(dset_msgs)
    assert A_DSET + len(dset) <= 2200

    sb = b"\x89HDF\r\n\x1a\n" + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, EOF, UNDEF)
    sb += struct.pack("<QQII", 0, A_ROOT, 1, 0) + struct.pack("<QQ", A_GTREE, A_HEAP)
    root = struct.pack("<BBHII", 1, 0, 1, 1, 24) + b"\0" * 4 + msg(0x11, struct.pack("<QQ", A_GTREE, A_HEAP))
    heap_data = pad8(b"\0") + pad8(b"Hits\0") + struct.pack("<QQ", 1, 16)
    heap = b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", len(heap_data), 16, A_HEAPDATA)
    gtree = b"TREE" + bytes([0, 0]) + struct.pack("<HQQ", 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, A_SNOD, 8)
    snod = b"SNOD" + bytes([1, 0]) + struct.pack("<H", 1) + struct.pack("<QQII", 8, A_DSET, 0, 0) + b"\0" * 16

    def key(nbytes, row):
        return struct.pack("<IIQQ", nbytes, 0, row, 0)

    def node(level, entries, last_key, left, right):
        b = b"TREE" + bytes([1, level]) + struct.pack("<HQQ", len(entries), left, right)
        for k, child in entries:
            b += k + struct.pack("<Q", child)
        return b + last_key

    leaf0 = node(0, [(key(36, 0), A_CHUNK[0]), (key(36, 2), A_CHUNK[1])], key(0, 4), UNDEF, A_LEAF1)
    leaf1 = node(0, [(key(36, 4), A_CHUNK[2])], key(0, 6), A_LEAF0, UNDEF)
    top = node(1, [(key(36, 0), A_LEAF0), (key(36, 4), A_LEAF1)], key(36, 6), UNDEF, UNDEF)

    blob = bytearray(EOF)
    for at, piece in [(0, sb), (A_ROOT, root), (A_HEAP, heap), (A_HEAPDATA, heap_data), (A_GTREE, gtree), (A_SNOD, snod),
                      (A_DSET, dset), (A_LEAF0, leaf0), (A_LEAF1, leaf1), (A_TOP, top)]:
        blob[at:at + len(piece)] = piece
    raw = rows.tobytes() + b"\0" * isz                         # the last chunk is stored whole: one padding row
    for i in range(3):
        blob[A_CHUNK[i]:A_CHUNK[i] + 36] = raw[36 * i:36 * (i + 1)]
    path = tmp_path / "hand.h5"
    path.write_bytes(bytes(blob))

    with h5min.TableReader(str(path)) as r:
        assert r.attrs("Hits") == {"NROWS": 5}
        got = r.read("Hits")
    assert got.dtype.names == HITS_DTYPE.names and got.dtype.itemsize == 18
    assert got.tobytes() == rows.tobytes()


def test_rejects_non_tables(tmp_path):
    with pytest.raises(ValueError):
        h5min.write_table(str(tmp_path / "bad.h5"), "Hits", np.zeros((3, 3)))


def test_extents_cover_the_rows_and_the_bounce_path_copies_the_same_bytes(tmp_path, monkeypatch):
    """TableReader.extents: the raw records of a streamed table as (offset, bytes) runs -- one run when the chunks lie
    back to back, as TableWriter lays them -- and TableWriter.append_from_file with and without copy_file_range."""
    a = _hits(5000, seed=3)
    src = str(tmp_path / "src.h5")
    h5min.write_table(src, "Hits", a, chunk_rows=512)
    with h5min.TableReader(src) as r:
        ext = r.extents("Hits")
        assert r.dtype_of("Hits") == HITS_DTYPE
    assert len(ext) == 1 and ext[0][1] == 5000 * 18
    raw = open(src, "rb").read()
    assert raw[ext[0][0]:ext[0][0] + ext[0][1]] == a.tobytes()
    for name, broken in (("kernel.h5", False), ("bounce.h5", True)):
        if broken:
            def refuse(*args, **kw):
                raise OSError(18, "cross-device link")
            monkeypatch.setattr(os, "copy_file_range", refuse, raising=False)
        dest = str(tmp_path / name)
        with h5min.TableWriter(dest, "Hits", HITS_DTYPE, chunk_rows=700) as w:
            w.append(a[:10])
            fd = os.open(src, os.O_RDONLY)
            try:
                w.append_from_file(fd, ext)
            finally:
                os.close(fd)
            assert w.nrows == 5010
        assert h5min.read_table(dest, "Hits").tobytes() == np.concatenate([a[:10], a]).tobytes()
    with h5min.TableWriter(str(tmp_path / "bad.h5"), "Hits", HITS_DTYPE) as w:
        fd = os.open(src, os.O_RDONLY)
        try:
            with pytest.raises(ValueError):
                w.append_from_file(fd, [(ext[0][0], 19)])            # not a whole number of records
        finally:
            os.close(fd)
