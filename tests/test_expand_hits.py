"""ptb_expand_hits / ptb_expand_hits_narrow (host code of libptb_b200, no GPU needed): both compressed-row forms of a hit
table (PtbCompactOut in include/ptb.h) back into the reference's 18-byte records (main/device.py:20-28, 156-165)."""
import numpy as np
import pytest

from pytestbeam_b200 import native
from pytestbeam_b200.engine import HITS_DTYPE, PtbError, expand_compact


def _case(n, seed, max_count=5, p_accept=0.7):
    rs = np.random.RandomState(seed)
    counts = rs.randint(0, max_count + 1, n).astype(np.uint8)
    counts[rs.uniform(size=n) < 0.3] = 0
    acc = rs.uniform(size=n) < p_accept
    words = np.zeros((n + 31) // 32, dtype=np.uint32)
    for j in range(32):
        sel = acc[j::32]
        words[:len(sel)] |= sel.astype(np.uint32) << np.uint32(j)
    rows = int(counts.sum())
    col = rs.randint(1, 1153, rows).astype(np.uint64)
    row = rs.randint(1, 577, rows).astype(np.uint64)
    q = rs.uniform(0.1, 30, rows).astype(np.float32)
    hits = col | (row << np.uint64(16)) | (q.view(np.uint32).astype(np.uint64) << np.uint64(32))
    return counts, acc, words, hits, (col, row, q)


def _plane(form, counts, words, hits, cols, plane=3):
    """The host arrays of one plane in the wide or the narrow form (engine.plane_of)."""
    if form == "wide":
        return {"narrow": False, "hits": hits, "counts": counts, "accepted": words}
    col, row, q = cols
    pos = (col | (row << np.uint64(12))).astype(np.uint32)
    c = np.minimum(np.concatenate([counts, np.zeros(len(counts) & 1, np.uint8)]), 15)
    # counts of 15 and more stand on the escape list, in no particular order, among entries of other planes
    big = np.flatnonzero(counts >= 15)
    esc = (counts[big].astype(np.uint64) << np.uint64(48)) | (np.uint64(plane) << np.uint64(40)) | big.astype(np.uint64)
    other = (np.uint64(77) << np.uint64(48)) | (np.uint64(plane + 1) << np.uint64(40)) | big.astype(np.uint64)
    esc = np.random.RandomState(0).permutation(np.concatenate([esc, other]))
    return {"narrow": True, "charge": q, "pos_lo": (pos & 0xFFFF).astype(np.uint16), "pos_hi": (pos >> 16).astype(np.uint8),
            "counts": (c[0::2] | (c[1::2] << 4)).astype(np.uint8), "accepted": words, "escapes": esc, "plane": plane}


def _packed_plane(counts, words, cols, plane=2, exp_min=123, col_bits=11, row_base=777):
    """The host arrays of one plane in the packed form: 48-bit rows (f32 mantissa | exponent - exp_min << 23 | column << 27
    | row << (27 + col_bits)) as three u16 arrays, nibble counts, the row layout; on the escape list the counts of 15
    and more and the charges beyond the format's 16 binades (the rows then keep exponent 15), among entries that belong
    to other planes."""
    base = _plane("narrow", counts, words, None, cols, plane=plane)
    col, row, q = cols
    bits = q.view(np.uint32).astype(np.uint64)
    off = (bits >> np.uint64(23)) - np.uint64(exp_min)
    big = np.flatnonzero(off > 15)
    esc = (np.uint64(1) << np.uint64(63)) | ((big.astype(np.uint64) + np.uint64(row_base)) << np.uint64(32)) | bits[big]
    other = (np.uint64(1) << np.uint64(63)) | ((big.astype(np.uint64) + np.uint64(row_base + len(q))) << np.uint64(32)) | np.uint64(5)
    off = np.minimum(off, 15)
    w = (bits & np.uint64(0x7FFFFF)) | (off << np.uint64(23)) | (col << np.uint64(27)) | (row << np.uint64(27 + col_bits))
    layout = native.PtbPackedLayout()
    layout.exp_min[plane], layout.col_bits[plane] = exp_min, col_bits
    return {"form": "packed", "narrow": True, "word0": (w & np.uint64(0xFFFF)).astype(np.uint16),
            "word1": ((w >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.uint16), "word2": (w >> np.uint64(32)).astype(np.uint16),
            "counts": base["counts"], "accepted": words, "plane": plane, "layout": layout, "row_base": row_base,
            "escapes": np.random.RandomState(1).permutation(np.concatenate([base["escapes"], esc, other]))}


def _short(p):
    """The same plane with its last row cut off (counts then promise one row too many)."""
    keys = {"packed": ("word0", "word1", "word2"), "narrow": ("charge", "pos_lo", "pos_hi")}.get(p.get("form"), None)
    keys = keys or (("charge", "pos_lo", "pos_hi") if p["narrow"] else ("hits",))
    return dict(p, **{k: p[k][:-1] for k in keys})


def _expected(counts, acc, cols, event_base):
    # event number of particle k: 1 + accepted particles before k (main/device.py:159,164-165)
    ev_k = event_base + 1 + np.concatenate([[0], np.cumsum(acc)[:-1]]) if len(acc) else np.zeros(0, np.int64)
    out = np.zeros(int(counts.sum()), dtype=HITS_DTYPE)
    out["event_number"] = np.repeat(ev_k, counts)
    out["column"], out["row"], out["charge"] = cols
    return out


@pytest.mark.parametrize("form", ["wide", "narrow"])
@pytest.mark.parametrize("n,threads", [(0, 0), (1, 1), (31, 0), (32, 0), (33, 2), (1000, 0), (600_001, 0), (600_001, 5)])
def test_expand_matches_the_reference_numbering(n, threads, form):
    lib = native.load()
    counts, acc, words, hits, cols = _case(n, seed=n + 1, max_count=40 if form == "narrow" else 5)
    got = expand_compact(lib, _plane(form, counts, words, hits, cols), n, event_base=12345, threads=threads)
    want = _expected(counts, acc, cols, 12345)
    assert got.tobytes() == want.tobytes()


@pytest.mark.parametrize("form", ["wide", "narrow"])
def test_bits_beyond_the_range_do_not_count_and_mismatch_is_refused(form):
    lib = native.load()
    n = 70
    counts, acc, words, hits, cols = _case(n, seed=5)
    words = words.copy()
    words[-1] |= np.uint32(0xFFFFFFFF) << np.uint32(n - 64)      # garbage above particle n-1
    plane = _plane(form, counts, words, hits, cols)
    got = expand_compact(lib, plane, n, event_base=0)
    assert got.tobytes() == _expected(counts, acc, cols, 0).tobytes()
    with pytest.raises(PtbError):
        expand_compact(lib, _short(plane), n, event_base=0)


@pytest.mark.parametrize("form", ["wide", "narrow"])
def test_expand_into_a_caller_buffer(form):
    lib = native.load()
    counts, acc, words, hits, cols = _case(5000, seed=9)
    dest = np.zeros(len(hits) + 10, dtype=HITS_DTYPE)
    view = dest[3:3 + len(hits)]
    expand_compact(lib, _plane(form, counts, words, hits, cols), 5000, event_base=7, out=view)
    assert view.tobytes() == _expected(counts, acc, cols, 7).tobytes()
    assert dest[:3].tobytes() == bytes(54) and dest[3 + len(hits):].tobytes() == bytes(7 * 18)


def test_narrow_count_of_15_without_an_escape_entry_is_refused():
    lib = native.load()
    counts, acc, words, hits, cols = _case(500, seed=3, max_count=40)
    plane = _plane("narrow", counts, words, hits, cols)
    assert (counts >= 15).any()
    with pytest.raises(PtbError):
        expand_compact(lib, dict(plane, escapes=plane["escapes"][:0]), 500, event_base=0)


@pytest.mark.parametrize("form", ["wide", "narrow"])
@pytest.mark.parametrize("n,threads", [(0, 0), (33, 2), (70_000, 0), (600_001, 5)])
def test_expand_into_a_file_equals_expand_into_memory(tmp_path, form, n, threads):
    """ptb_expand_hits*_fd: the records written block by block at a byte offset of a file are the records of the
    in-memory expansion; bytes before the offset stay untouched; a bad descriptor is an error, not a crash."""
    import os
    from pytestbeam_b200.engine import expand_compact_to_fd
    lib = native.load()
    counts, acc, words, hits, cols = _case(n, seed=n + 11, max_count=40 if form == "narrow" else 5)
    plane = _plane(form, counts, words, hits, cols)
    want = expand_compact(lib, plane, n, event_base=99)
    path = tmp_path / "rows.bin"
    path.write_bytes(b"\xaa" * 1000)
    fd = os.open(str(path), os.O_RDWR)
    try:
        assert expand_compact_to_fd(lib, plane, n, 99, fd, 1000, threads=threads) == len(want)
    finally:
        os.close(fd)
    data = path.read_bytes()
    assert data[:1000] == b"\xaa" * 1000 and data[1000:] == want.tobytes()
    if len(want):
        with pytest.raises(PtbError):
            expand_compact_to_fd(lib, plane, n, 99, 10_000, 0)      # no such descriptor


@pytest.mark.parametrize("n,threads", [(0, 0), (1, 1), (33, 2), (1000, 0), (600_001, 5)])
def test_packed_rows_expand_to_the_same_records(tmp_path, n, threads):
    """ptb_expand_hits_packed / _packed_fd: the 48-bit rows (charges between 2^-4 and 2^5 ke above a threshold of 2^-4,
    11 column bits) give the records of the wide form, in memory and in a file; a short row array is refused."""
    import os
    from pytestbeam_b200.engine import expand_compact_to_fd
    lib = native.load()
    counts, acc, words, hits, cols = _case(n, seed=n + 21, max_count=40)
    col, row, q = cols
    q[::97] *= np.float32(3000.0)                      # some charges far beyond the 16 binades above 2^-4
    q[5::1013] = np.float32(2048.0 * 1.5)              # and some in the top binade itself (exponent 15, no escape entry)
    plane = _packed_plane(counts, words, cols)
    want = _expected(counts, acc, cols, 4242)
    assert expand_compact(lib, plane, n, event_base=4242, threads=threads).tobytes() == want.tobytes()
    path = tmp_path / "rows.bin"
    fd = os.open(str(path), os.O_RDWR | os.O_CREAT)
    try:
        assert expand_compact_to_fd(lib, plane, n, 4242, fd, 0, threads=threads) == len(want)
    finally:
        os.close(fd)
    assert path.read_bytes() == want.tobytes()
    if len(want):
        with pytest.raises(PtbError):
            expand_compact(lib, _short(plane), n, event_base=0)
