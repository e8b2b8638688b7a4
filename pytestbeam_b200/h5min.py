"""Minimal HDF5 writer / reader for one-dimensional tables of a packed compound type.

The reference stores its outputs with PyTables (``main/device.py:306-340`` table ``/Hits``,
``main/hit.py:434-468`` table ``/Tracks``) and reads them back with ``tb.open_file(...).root.Hits[:]``
(``main/plotting.py:502-508``).  Neither PyTables, h5py nor libhdf5 exist on the target image, so the
file contract is honoured with a writer built directly on the HDF5 File Format Specification:

    version 0 superblock, version 1 object headers, symbol-table root group (local heap + v1 B-tree + SNOD)
    one dataset per table: rank 1, compound datatype (version 1 encoding), and either
      * chunked layout (layout message v3 class 2, v1 B-tree of raw-data chunks, unlimited maximum dimension): what
        PyTables itself writes for a Table, and what lets ``TableWriter`` stream a table to disk piece by piece; or
      * contiguous layout (``write_tables``: several small tables in one file)
    the attributes PyTables uses to recognise a Table without guessing: CLASS / VERSION / TITLE / FIELD_n_NAME /
    FIELD_n_FILL / NROWS on the dataset, CLASS / VERSION / TITLE / PYTABLES_FORMAT_VERSION on the root group

What is *not* reproduced: the blosc filter (a PyTables plug-in, filter id 32001 -- cannot be produced without it; the
chunks are stored unfiltered).  Opening these files with PyTables / h5py could not be tried offline.  The reader
below parses the structures independently of the writer; the tests pin it to a file written by libhdf5 itself (a
MATLAB 7.3 file that ships with scipy's test data) and to a chunked file assembled by hand from the specification.
"""
from __future__ import annotations

import os
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
GROUP_LEAF_K = 4
GROUP_INTERNAL_K = 16
CHUNK_K = 32            # "indexed storage internal node K": not stored in a version 0 superblock, the library assumes 32
DATA_ALIGN = 4096       # first data byte of a streamed table (page-aligned so that the rows can be memory-mapped)


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ---------------------------------------------------------------------------------------------------
# datatype / dataspace / attribute messages
# ---------------------------------------------------------------------------------------------------

def _dtype_atomic(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.byteorder == ">":
        raise ValueError("big-endian members are not supported")
    size = dt.itemsize
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00           # bit 3: signed
        head = struct.pack("<BBBBI", 0x10 | 0, bits0, 0, 0, size)   # class 0, version 1
        return head + struct.pack("<HH", 0, size * 8)
    if dt.kind == "f" and size in (4, 8):
        if size == 4:
            sign, eloc, esize, msize, bias = 31, 23, 8, 23, 127
        else:
            sign, eloc, esize, msize, bias = 63, 52, 11, 52, 1023
        head = struct.pack("<BBBBI", 0x10 | 1, 0x20, sign, 0, size)  # class 1; mantissa normalisation: implied msb
        return head + struct.pack("<HHBBBBI", 0, size * 8, eloc, esize, 0, msize, bias)
    raise ValueError(f"unsupported member type {dt}")


def _dtype_compound(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    names = dt.names
    head = struct.pack("<BBBBI", 0x10 | 6, len(names) & 0xFF, (len(names) >> 8) & 0xFF, 0, dt.itemsize)
    body = b""
    for n in names:
        sub, off = dt.fields[n][0], dt.fields[n][1]
        body += _pad8(n.encode("ascii") + b"\0")
        body += struct.pack("<IB3xI4x4I", off, 0, 0, 0, 0, 0, 0)
        body += _dtype_atomic(sub)
    return head + body


def _dtype_string(size: int) -> bytes:
    # class 3 (string), version 1; padding: null-terminated (0), character set: UTF-8 (1 << 4), as PyTables under Python 3
    return struct.pack("<BBBBI", 0x10 | 3, 0x10, 0, 0, size)


_SCALAR_SPACE = struct.pack("<BBB5x", 1, 0, 0)


def _message(mtype: int, data: bytes) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _attribute(name: str, datatype: bytes, value: bytes) -> bytes:
    """Attribute message, version 1, scalar dataspace."""
    nm = name.encode("ascii") + b"\0"
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(datatype), len(_SCALAR_SPACE))
    body += _pad8(nm) + _pad8(datatype) + _pad8(_SCALAR_SPACE) + value
    return _message(0x000C, body)


def _attr_str(name: str, value: str) -> bytes:
    raw = value.encode("utf-8") or b"\0"          # PyTables stores an empty string as one NUL
    return _attribute(name, _dtype_string(len(raw)), raw)


def _attr_num(name: str, value, dt) -> bytes:
    dt = np.dtype(dt)
    return _attribute(name, _dtype_atomic(dt), np.array(value, dtype=dt).tobytes())


def _table_attributes(dtype: np.dtype, nrows: int, title: str) -> list:
    """What tables.Table writes next to its dataset (tables/table.py: CLASS, VERSION, TITLE, FIELD_*, NROWS)."""
    msgs = [_attr_str("CLASS", "TABLE"), _attr_str("VERSION", "2.7"), _attr_str("TITLE", title)]
    for i, n in enumerate(dtype.names):
        msgs.append(_attr_str(f"FIELD_{i}_NAME", n))
    for i, n in enumerate(dtype.names):
        msgs.append(_attr_num(f"FIELD_{i}_FILL", 0, dtype.fields[n][0]))
    msgs.append(_attr_num("NROWS", nrows, "<i8"))
    return msgs


def _root_attributes() -> list:
    return [_attr_str("TITLE", ""), _attr_str("CLASS", "GROUP"), _attr_str("VERSION", "1.0"),
            _attr_str("PYTABLES_FORMAT_VERSION", "2.1")]


def _object_header(messages: list) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def _dataset_header(dtype: np.dtype, nrows: int, data_addr: int) -> bytes:
    """Contiguous layout (write_tables)."""
    nbytes = nrows * dtype.itemsize
    space = struct.pack("<BBB5xQ", 1, 1, 0, nrows)
    fill = struct.pack("<BBBB", 2, 1, 0, 0)
    layout = struct.pack("<BBQQ", 3, 1, data_addr if nbytes else UNDEF, nbytes)
    return _object_header([_message(0x0001, space), _message(0x0003, _dtype_compound(dtype)),
                           _message(0x0005, fill), _message(0x0008, layout)])


def _chunked_header(dtype: np.dtype, nrows: int, chunk_rows: int, btree_addr: int, title: str, attrs: bool) -> bytes:
    space = struct.pack("<BBB5xQQ", 1, 1, 1, nrows, UNDEF)            # flags bit 0: maximum dimensions present (unlimited)
    fill = struct.pack("<BBBB", 2, 3, 0, 0)                           # incremental allocation, as for any chunked dataset
    layout = struct.pack("<BBBQII", 3, 2, 2, btree_addr, chunk_rows, dtype.itemsize)
    msgs = [_message(0x0001, space), _message(0x0003, _dtype_compound(dtype)), _message(0x0005, fill),
            _message(0x0008, layout)]
    if attrs:
        msgs += _table_attributes(dtype, nrows, title)
    return _object_header(msgs)


def _group_skeleton(names: list, header_sizes: dict, root_attrs: bool):
    """Superblock, root group header, local heap, group B-tree and symbol node for datasets `names` whose object headers
    follow in name order.  Returns (pieces: list of (address, bytes-maker(eof)), dataset header addresses, end)."""
    names = sorted(names)
    if not 1 <= len(names) <= 2 * GROUP_LEAF_K:
        raise ValueError("between 1 and 8 tables per file")
    heap_data = _pad8(b"\0")
    name_off = {}
    for n in names:
        name_off[n] = len(heap_data)
        heap_data += _pad8(n.encode("ascii") + b"\0")
    free_off = len(heap_data)
    heap_data += struct.pack("<QQ", 1, 16)      # free block: next = H5HL_FREE_NULL, size 16
    root_msgs_len = len(_object_header([_message(0x0011, b"\0" * 16)] + (_root_attributes() if root_attrs else [])))
    sizes = {"super": 96, "root_hdr": root_msgs_len, "heap": 32, "heap_data": len(heap_data),
             "btree": 24 + (2 * GROUP_INTERNAL_K + 1) * 8 + 2 * GROUP_INTERNAL_K * 8, "snod": 8 + 2 * GROUP_LEAF_K * 40}
    addr, cur = {}, 0
    for k in ("super", "root_hdr", "heap", "heap_data", "btree", "snod"):
        addr[k] = cur
        cur += (sizes[k] + 7) // 8 * 8
    hdr_addr = {}
    for n in names:
        hdr_addr[n] = cur
        cur += (header_sizes[n] + 7) // 8 * 8

    def superblock(eof):
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, GROUP_LEAF_K, GROUP_INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, addr["root_hdr"], 1, 0) + struct.pack("<QQ", addr["btree"], addr["heap"])
        assert len(sb) == 96
        return sb

    root_hdr = _object_header([_message(0x0011, struct.pack("<QQ", addr["btree"], addr["heap"]))] +
                              (_root_attributes() if root_attrs else []))
    heap = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, addr["heap_data"])
    btree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF)
    btree += struct.pack("<QQQ", 0, addr["snod"], name_off[names[-1]])
    btree += b"\0" * (sizes["btree"] - len(btree))
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for n in names:
        snod += struct.pack("<QQII16x", name_off[n], hdr_addr[n], 0, 0)
    snod += b"\0" * (sizes["snod"] - len(snod))
    fixed = [(addr["root_hdr"], root_hdr), (addr["heap"], heap), (addr["heap_data"], heap_data),
             (addr["btree"], btree), (addr["snod"], snod)]
    return superblock, fixed, hdr_addr, cur


# ---------------------------------------------------------------------------------------------------
# writers
# ---------------------------------------------------------------------------------------------------

def write_tables(path: str, tables: dict) -> None:
    """Write ``{name: structured ndarray}`` as 1-D compound datasets (contiguous layout) of the root group."""
    arrays = {n: np.ascontiguousarray(tables[n]) for n in tables}
    for n, a in arrays.items():
        if a.ndim != 1 or a.dtype.names is None:
            raise ValueError(f"table {n!r} must be a one-dimensional structured array")
    names = sorted(arrays)
    sizes = {n: len(_dataset_header(arrays[n].dtype, len(arrays[n]), 0)) for n in names}
    superblock, fixed, hdr_addr, cur = _group_skeleton(names, sizes, root_attrs=False)
    data_addr = {}
    for n in names:
        data_addr[n] = cur
        cur += (arrays[n].nbytes + 7) // 8 * 8
    eof = cur
    with open(path, "wb") as f:
        def put(at, blob):
            f.seek(at)
            f.write(blob)
        put(0, superblock(eof))
        for at, blob in fixed:
            put(at, blob)
        for n in names:
            put(hdr_addr[n], _dataset_header(arrays[n].dtype, len(arrays[n]), data_addr[n]))
            f.seek(data_addr[n])
            arrays[n].tofile(f)
        f.truncate(eof)


def _chunk_key(nbytes: int, row: int) -> bytes:
    # chunk size, filter mask, offset of the chunk in every dimension plus the element-size dimension
    return struct.pack("<IIQQ", nbytes, 0, row, 0)


def _chunk_btree(chunk_addr, n_chunks: int, chunk_rows: int, chunk_bytes: int, at: int):
    """v1 B-tree (node type 1) over chunks 0..n_chunks-1 (chunk i starts at row i*chunk_rows, at chunk_addr(i)), laid out
    from file address `at`.  Returns (root address, bytes)."""
    key_size, fanout = 24, 2 * CHUNK_K
    node_size = 24 + fanout * 8 + (fanout + 1) * key_size
    level, blob = 0, b""
    items = [(i * chunk_rows, chunk_addr(i)) for i in range(n_chunks)]     # (first row, child address)
    end_row = n_chunks * chunk_rows
    while True:
        nodes = [items[i:i + fanout] for i in range(0, len(items), fanout)] or [[]]
        addrs = [at + len(blob) + j * node_size for j in range(len(nodes))]
        for j, ent in enumerate(nodes):
            left = addrs[j - 1] if j > 0 else UNDEF
            right = addrs[j + 1] if j + 1 < len(nodes) else UNDEF
            b = b"TREE" + struct.pack("<BBHQQ", 1, level, len(ent), left, right)
            for row, child in ent:
                b += _chunk_key(chunk_bytes, row) + struct.pack("<Q", child)
            last = nodes[j + 1][0][0] if j + 1 < len(nodes) else end_row
            b += _chunk_key(0 if level == 0 else chunk_bytes, last)
            blob += b + b"\0" * (node_size - len(b))
        if len(nodes) == 1:
            return addrs[0], blob
        items = [(ent[0][0], addrs[j]) for j, ent in enumerate(nodes)]
        level += 1


class TableWriter:
    """Streams one table into ``path`` as a chunked, extendable HDF5 dataset: ``append`` rows (or ``reserve`` a
    writable memory-mapped view of the next rows and fill it in place), then ``close``.

    The chunks of the single dataset sit back to back from a page-aligned address, so row r lives at
    ``data_start + r * itemsize``; the chunk B-tree and the final row count are written by ``close``."""

    def __init__(self, path: str, name: str, dtype, chunk_rows: int = 65536, title: str = "", attrs: bool = True):
        self.path, self.name, self.dtype = path, name, np.dtype(dtype)
        if self.dtype.names is None:
            raise ValueError("a table needs a compound (structured) dtype")
        self.chunk_rows, self.title, self.attrs = int(chunk_rows), title, attrs
        if self.chunk_rows < 1 or self.chunk_rows * self.dtype.itemsize >= 1 << 32:
            raise ValueError("chunk size out of range")
        hdr = len(_chunked_header(self.dtype, 0, self.chunk_rows, UNDEF, title, attrs))
        self._superblock, self._fixed, hdr_addr, end = _group_skeleton([name], {name: hdr}, root_attrs=attrs)
        self._hdr_addr = hdr_addr[name]
        self.data_start = (end + DATA_ALIGN - 1) // DATA_ALIGN * DATA_ALIGN
        self.nrows = 0
        self._f = open(path, "w+b")
        self._closed = False

    def _row_addr(self, r):
        return self.data_start + r * self.dtype.itemsize

    def append(self, records: np.ndarray) -> None:
        a = np.ascontiguousarray(records, dtype=self.dtype)
        if a.ndim != 1:
            raise ValueError("rows must be one-dimensional")
        if len(a):
            # positioned writes on the descriptor (no shared file position, the GIL is released meanwhile): writers of
            # different files can be fed from different threads
            buf, off, fd = memoryview(a).cast("B"), self._row_addr(self.nrows), self._f.fileno()
            while len(buf):
                done = os.pwrite(fd, buf[:1 << 30], off)
                buf, off = buf[done:], off + done
            self.nrows += len(a)

    def append_with(self, n: int, fill) -> None:
        """Append n rows that `fill(fd, byte offset)` writes into the file itself (positioned writes; it returns the
        number of rows it wrote, which must be n)."""
        n = int(n)
        if n:
            self._f.flush()
            if fill(self._f.fileno(), self._row_addr(self.nrows)) != n:
                raise ValueError("the rows written do not match the rows announced")
            self.nrows += n

    def append_from_file(self, src_fd: int, extents: list) -> None:
        """Append rows that already lie in another file as raw records: extents = [(offset, bytes)] in row order
        (TableReader.extents).  The bytes move inside the kernel (copy_file_range) where the platform allows it --
        nothing passes through this process -- and through a bounce buffer otherwise."""
        isz, fd = self.dtype.itemsize, self._f.fileno()
        self._f.flush()
        for off, nbytes in extents:
            if nbytes % isz:
                raise ValueError("extent is not a whole number of records")
            dst, left = self._row_addr(self.nrows), nbytes
            while left:
                try:
                    done = os.copy_file_range(src_fd, fd, min(left, 1 << 30), off, dst)
                except (OSError, AttributeError):
                    done = 0
                if done <= 0:                                  # other file system, old kernel: read + positioned write
                    buf = os.pread(src_fd, min(left, 1 << 24), off)
                    if not buf:
                        raise ValueError("source file ends inside an extent")
                    done = os.pwrite(fd, buf, dst)
                off, dst, left = off + done, dst + done, left - done
            self.nrows += nbytes // isz

    def reserve(self, n: int) -> np.ndarray:
        """A writable view of rows [nrows, nrows+n) of the file (numpy memmap); the rows count as appended."""
        n = int(n)
        if n == 0:
            return np.empty(0, dtype=self.dtype)
        self._f.flush()
        off = self._row_addr(self.nrows)
        os.ftruncate(self._f.fileno(), off + n * self.dtype.itemsize)
        self.nrows += n
        return np.memmap(self.path, dtype=self.dtype, mode="r+", offset=off, shape=(n,))

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        f, isz = self._f, self.dtype.itemsize
        chunk_bytes = self.chunk_rows * isz
        n_chunks = (self.nrows + self.chunk_rows - 1) // self.chunk_rows
        data_end = self.data_start + n_chunks * chunk_bytes          # chunks are stored whole: the last one is padded
        f.flush()
        os.ftruncate(f.fileno(), data_end)
        root, tree = UNDEF, b""                                      # no chunk allocated: no index either
        if n_chunks:
            root, tree = _chunk_btree(lambda i: self.data_start + i * chunk_bytes, n_chunks, self.chunk_rows,
                                      chunk_bytes, data_end)
        f.seek(data_end)
        f.write(tree)
        eof = data_end + len(tree)
        f.seek(0)
        f.write(self._superblock(eof))
        for at, blob in self._fixed:
            f.seek(at)
            f.write(blob)
        f.seek(self._hdr_addr)
        f.write(_chunked_header(self.dtype, self.nrows, self.chunk_rows, root, self.title, self.attrs))
        f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def default_chunk_rows(nrows: int) -> int:
    return int(min(65536, max(512, nrows)))


def write_table(path: str, name: str, table: np.ndarray, chunk_rows: int | None = None, title: str = "") -> None:
    """One table as a chunked, extendable dataset with the PyTables Table attributes."""
    table = np.ascontiguousarray(table)
    if table.ndim != 1 or table.dtype.names is None:
        raise ValueError(f"table {name!r} must be a one-dimensional structured array")
    with TableWriter(path, name, table.dtype, chunk_rows or default_chunk_rows(len(table)), title) as w:
        w.append(table)


def concatenate_tables(dest: str, name: str, parts: list, dtype, chunk_rows: int = 65536, title: str = "") -> int:
    """Append the table `name` of every file in `parts`, in order, to a new file `dest`: the parts' raw records are
    copied extent by extent, file to file (no part is read into this process, let alone held as a whole).  Returns the
    number of rows."""
    with TableWriter(dest, name, dtype, chunk_rows, title) as w:
        for p in parts:
            with TableReader(p) as r:
                if r.dtype_of(name) != w.dtype:
                    raise ValueError(f"{p}: table {name!r} has another record type")
                fd = os.open(p, os.O_RDONLY)
                try:
                    w.append_from_file(fd, r.extents(name))
                finally:
                    os.close(fd)
        return w.nrows


# ---------------------------------------------------------------------------------------------------
# reader (independent walk of the on-disk structures)
# ---------------------------------------------------------------------------------------------------

class _Buf:
    def __init__(self, data, base=0):
        self.d, self.base = data, base

    def u(self, off, n):
        return int.from_bytes(bytes(self.d[off:off + n]), "little")

    def b(self, off):
        return int(self.d[off])

    def find0(self, start):
        i = start
        while int(self.d[i]) != 0:
            i += 1
        return i


def _parse_datatype(b: _Buf, off: int):
    """Return (numpy dtype, encoded length)."""
    cls, ver = b.b(off) & 0x0F, b.b(off) >> 4
    bits = bytes(b.d[off + 1:off + 4])
    size = b.u(off + 4, 4)
    if cls == 0:
        signed = bool(bits[0] & 0x08)
        if bits[0] & 1:
            raise ValueError("big-endian integer")
        return np.dtype(("<i" if signed else "<u") + str(size)), 12
    if cls == 1:
        if bits[0] & 1:
            raise ValueError("big-endian float")
        return np.dtype("<f" + str(size)), 20
    if cls == 3:
        return np.dtype("S" + str(size)), 8
    if cls == 6:
        if ver != 1:
            raise ValueError("only version 1 compound encodings are understood")
        nmemb = bits[0] | (bits[1] << 8)
        p = off + 8
        names, formats, offsets = [], [], []
        for _ in range(nmemb):
            end = b.find0(p)
            name = bytes(b.d[p:end]).decode("ascii")
            p += (end - p + 1 + 7) // 8 * 8
            moff = b.u(p, 4)
            if b.b(p + 4) != 0:
                raise ValueError("array members are not supported")
            p += 32
            sub, used = _parse_datatype(b, p)
            p += used
            names.append(name)
            formats.append(sub)
            offsets.append(moff)
        return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size}), p - off
    raise ValueError(f"datatype class {cls} not supported")


def _parse_dataspace(b: _Buf, off: int):
    if b.b(off) != 1:
        raise ValueError("only version 1 dataspaces are understood")
    rank, flags = b.b(off + 1), b.b(off + 2)
    dims = [b.u(off + 8 + 8 * i, 8) for i in range(rank)]
    maxdims = [b.u(off + 8 + 8 * (rank + i), 8) for i in range(rank)] if flags & 1 else None
    return dims, maxdims


def _read_object_header(b: _Buf, addr: int) -> list:
    """All messages of a version 1 object header (continuation blocks followed): list of (type, data offset, size)."""
    a = b.base + addr
    if b.b(a) != 1:
        raise ValueError("only version 1 object headers are understood")
    nmsg, total = b.u(a + 2, 2), b.u(a + 8, 4)
    blocks, msgs = [(a + 16, total)], []
    while blocks and len(msgs) < nmsg:
        p, size = blocks.pop(0)
        end = p + size
        while p + 8 <= end and len(msgs) < nmsg:
            mtype, msize = b.u(p, 2), b.u(p + 2, 2)
            msgs.append((mtype, p + 8, msize))
            if mtype == 0x0010:                        # continuation: (address, length) of another block of messages
                blocks.append((b.base + b.u(p + 8, 8), b.u(p + 16, 8)))
            p += 8 + msize
    return msgs


def _parse_attribute(b: _Buf, off: int):
    if b.b(off) != 1:
        raise ValueError("only version 1 attribute messages are understood")
    nlen, tlen, slen = b.u(off + 2, 2), b.u(off + 4, 2), b.u(off + 6, 2)
    p = off + 8
    name = bytes(b.d[p:p + nlen]).split(b"\0")[0].decode("ascii")
    p += (nlen + 7) // 8 * 8
    dt, _ = _parse_datatype(b, p)
    p += (tlen + 7) // 8 * 8
    dims, _ = _parse_dataspace(b, p) if slen else ([], None)
    p += (slen + 7) // 8 * 8
    count = int(np.prod(dims)) if dims else 1
    val = np.frombuffer(bytes(b.d[p:p + count * dt.itemsize]), dtype=dt, count=count)
    if dt.kind == "S":
        val = [v.split(b"\0")[0].decode("utf-8") for v in val.tolist()]
    else:
        val = val.tolist()
    return name, (val[0] if not dims else val)


class TableReader:
    """Read-only view of an HDF5 file of the kinds described at the top of this module (memory-mapped)."""

    def __init__(self, path: str):
        self.path = path
        size = os.path.getsize(path)
        self._mm = np.memmap(path, dtype=np.uint8, mode="r") if size else np.zeros(0, dtype=np.uint8)
        d = self._mm
        start = None
        for cand in [0] + [512 << i for i in range(12)]:          # the superblock sits at 0 or at a power of two >= 512
            if cand + 8 <= size and bytes(d[cand:cand + 8]) == SIGNATURE:
                start = cand
                break
        if start is None or int(d[start + 8]) != 0:
            raise ValueError("not a version-0 HDF5 file")
        if int(d[start + 13]) != 8 or int(d[start + 14]) != 8:
            raise ValueError("only 8-byte offsets and lengths are understood")
        b = _Buf(d, 0)
        base = b.u(start + 24, 8)
        self._b = b = _Buf(d, base)
        eof = b.u(start + 40, 8)
        if eof not in (size, size - base):
            raise ValueError("end-of-file address does not match the file size")
        self._root_hdr = b.u(start + 56 + 8, 8)
        self._objects = None

    def close(self):
        self._b = None
        mm, self._mm = self._mm, None
        if isinstance(mm, np.memmap):
            try:
                mm._mmap.close()
            except BufferError:            # rows handed out by iter_rows are still alive: the map goes with them
                pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    # -- group walk --------------------------------------------------------------------------------
    def objects(self) -> dict:
        """{name: object header address} of the root group."""
        if self._objects is not None:
            return self._objects
        b = self._b
        msgs = {t: off for t, off, _ in _read_object_header(b, self._root_hdr)}
        if 0x11 not in msgs:
            raise ValueError("root group has no symbol table")
        btree, heap = b.u(msgs[0x11], 8), b.u(msgs[0x11] + 8, 8)
        hp = b.base + heap
        if bytes(b.d[hp:hp + 4]) != b"HEAP":
            raise ValueError("bad local heap")
        heap_data = b.base + b.u(hp + 24, 8)

        def name_at(off):
            s = heap_data + off
            return bytes(b.d[s:b.find0(s)]).decode("ascii")

        def walk(node):
            a = b.base + node
            if bytes(b.d[a:a + 4]) == b"SNOD":
                for i in range(b.u(a + 6, 2)):
                    e = a + 8 + 40 * i
                    yield name_at(b.u(e, 8)), b.u(e + 8, 8)
                return
            if bytes(b.d[a:a + 4]) != b"TREE" or b.b(a + 4) != 0:
                raise ValueError("bad group B-tree node")
            for i in range(b.u(a + 6, 2)):
                yield from walk(b.u(a + 24 + 8 + 16 * i, 8))

        self._objects = dict(walk(btree))
        return self._objects

    def root_attrs(self) -> dict:
        return dict(_parse_attribute(self._b, off) for t, off, _ in _read_object_header(self._b, self._root_hdr)
                    if t == 0x000C)

    def attrs(self, name: str) -> dict:
        return dict(_parse_attribute(self._b, off) for t, off, _ in
                    _read_object_header(self._b, self.objects()[name]) if t == 0x000C)

    # -- datasets ----------------------------------------------------------------------------------
    def _describe(self, name: str):
        b = self._b
        m = {}
        for t, off, _ in _read_object_header(b, self.objects()[name]):
            m.setdefault(t, off)
        if not {0x1, 0x3, 0x8} <= set(m):
            raise KeyError(f"{name!r} is not a dataset")
        dims, maxdims = _parse_dataspace(b, m[0x1])
        dtype, _ = _parse_datatype(b, m[0x3])
        lay = m[0x8]
        ver = b.b(lay)
        info = {"dims": dims, "maxdims": maxdims, "dtype": dtype}
        if ver == 3:
            cls = b.b(lay + 1)
            if cls == 1:
                info.update(kind="contiguous", addr=b.u(lay + 2, 8), size=b.u(lay + 10, 8))
            elif cls == 2:
                nd = b.b(lay + 2)
                info.update(kind="chunked", btree=b.u(lay + 3, 8),
                            chunk=[b.u(lay + 11 + 4 * i, 4) for i in range(nd)])
            else:
                raise ValueError("compact layouts are not understood")
        elif ver in (1, 2):
            nd, cls = b.b(lay + 1), b.b(lay + 2)
            if cls != 1:
                raise ValueError("only contiguous version 1/2 layouts are understood")
            info.update(kind="contiguous", addr=b.u(lay + 8, 8), size=None)
        else:
            raise ValueError(f"layout message version {ver} not understood")
        return info

    def _chunks(self, info):
        """[(first row, file address, bytes)] of a chunked rank-1 dataset, in row order."""
        b, out = self._b, []
        nd = len(info["chunk"])
        key = 8 + 8 * nd

        def walk(node):
            a = b.base + node
            if bytes(b.d[a:a + 4]) != b"TREE" or b.b(a + 4) != 1:
                raise ValueError("bad chunk B-tree node")
            level, used = b.b(a + 5), b.u(a + 6, 2)
            p = a + 24
            for _ in range(used):
                nbytes, mask, row = b.u(p, 4), b.u(p + 4, 4), b.u(p + 8, 8)
                child = b.u(p + key, 8)
                if level:
                    walk(child)
                else:
                    if mask:
                        raise ValueError("filtered chunks are not understood")
                    out.append((row, b.base + child, nbytes))
                p += key + 8

        if info["btree"] != UNDEF:
            walk(info["btree"])
        return sorted(out)

    def nrows(self, name: str) -> int:
        return int(self._describe(name)["dims"][0])

    def dtype_of(self, name: str) -> np.dtype:
        return self._describe(name)["dtype"]

    def extents(self, name: str) -> list:
        """[(file offset, bytes)] of the raw records of a rank-1 dataset in row order, neighbouring chunks merged."""
        info = self._describe(name)
        if len(info["dims"]) != 1:
            raise ValueError("extents is for rank-1 datasets")
        n, isz = info["dims"][0], info["dtype"].itemsize
        if n == 0:
            return []
        if info["kind"] == "contiguous":
            return [(self._b.base + info["addr"], n * isz)]
        crow, out = info["chunk"][0], []
        if info["chunk"][-1] != isz:
            raise ValueError("chunk element size does not match the datatype")
        expect = 0
        for row, addr, nbytes in self._chunks(info):
            if nbytes != crow * isz or row != expect:
                raise ValueError("the chunks do not tile the dataspace")
            m = min(crow, n - row) * isz
            if m <= 0:
                break
            if out and out[-1][0] + out[-1][1] == addr:
                out[-1] = (out[-1][0], out[-1][1] + m)
            else:
                out.append((addr, m))
            expect = row + crow
        if expect < n:
            raise ValueError("the chunks do not cover the dataspace")
        return out

    def iter_rows(self, name: str, max_rows: int = 1 << 20):
        """The rows of a rank-1 dataset in pieces of at most max_rows (views of the mapped file, or small copies)."""
        info = self._describe(name)
        if len(info["dims"]) != 1:
            raise ValueError("iter_rows is for rank-1 datasets")
        n, dt, b = info["dims"][0], info["dtype"], self._b
        if info["kind"] == "contiguous":
            a = b.base + info["addr"]
            for r in range(0, n, max_rows):
                m = min(max_rows, n - r)
                yield np.frombuffer(b.d, dtype=dt, count=m, offset=a + r * dt.itemsize)
            return
        crow = info["chunk"][0]
        if info["chunk"][-1] != dt.itemsize:
            raise ValueError("chunk element size does not match the datatype")
        for row, addr, nbytes in self._chunks(info):
            if nbytes != crow * dt.itemsize:
                raise ValueError("chunk size does not match the layout")
            m = min(crow, n - row)
            for r in range(0, max(m, 0), max_rows):
                k = min(max_rows, m - r)
                yield np.frombuffer(b.d, dtype=dt, count=k, offset=addr + r * dt.itemsize)

    def read(self, name: str) -> np.ndarray:
        info = self._describe(name)
        dt, dims, b = info["dtype"], info["dims"], self._b
        count = int(np.prod(dims)) if dims else 1
        if info["kind"] == "contiguous":
            if info["size"] is not None and info["size"] != count * dt.itemsize:
                raise ValueError("layout size does not match dataspace x datatype")
            if count == 0:
                return np.zeros(dims, dtype=dt)
            return np.frombuffer(b.d, dtype=dt, count=count, offset=b.base + info["addr"]).reshape(dims).copy()
        out = np.zeros(dims[0], dtype=dt)
        at = 0
        for piece in self.iter_rows(name):
            out[at:at + len(piece)] = piece
            at += len(piece)
        if at != dims[0]:
            raise ValueError("the chunks do not cover the dataspace")
        return out


def read_tables(path: str) -> dict:
    """Return ``{name: ndarray}`` for every dataset of the root group."""
    with TableReader(path) as r:
        out = {}
        for name in r.objects():
            try:
                out[name] = r.read(name)
            except KeyError:
                continue
        return out


def read_table(path: str, name: str) -> np.ndarray:
    with TableReader(path) as r:
        return r.read(name)


def read_attrs(path: str, name: str) -> dict:
    with TableReader(path) as r:
        return r.attrs(name)
