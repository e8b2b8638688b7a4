"""Minimal HDF5 writer / reader for one-dimensional tables of a packed compound type.

The reference stores its outputs with PyTables (``main/device.py:306-340`` table ``/Hits``,
``main/hit.py:434-468`` table ``/Tracks``).  Neither PyTables, h5py nor libhdf5 exist on the target image
(SURVEY.md H5), so the file contract is honoured with a writer built directly on the HDF5 File Format
Specification (version 0 superblock, version 1 object headers, symbol-table groups, contiguous layout):

    root group "/"  ->  one dataset per table, rank 1, compound datatype (version 1 encoding), contiguous data

What is *not* reproduced: the blosc filter (a PyTables plug-in, filter id 32001 -- cannot be produced without
it; data are stored uncompressed) and the PyTables bookkeeping attributes.  Opening these files with PyTables /
h5py could not be verified offline; the reader below parses the same structures independently of the writer
and is what the tests use.
"""
from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
GROUP_LEAF_K = 4
GROUP_INTERNAL_K = 16


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ---------------------------------------------------------------------------------------------------
# datatype messages
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


# ---------------------------------------------------------------------------------------------------
# writer
# ---------------------------------------------------------------------------------------------------

def _message(mtype: int, data: bytes) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _object_header(messages: list) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def _dataset_header(dtype: np.dtype, nrows: int, data_addr: int) -> bytes:
    nbytes = nrows * dtype.itemsize
    space = struct.pack("<BBB5xQ", 1, 1, 0, nrows)
    fill = struct.pack("<BBBB", 2, 1, 0, 0)
    layout = struct.pack("<BBQQ", 3, 1, data_addr if nbytes else UNDEF, nbytes)
    return _object_header([_message(0x0001, space), _message(0x0003, _dtype_compound(dtype)),
                           _message(0x0005, fill), _message(0x0008, layout)])


def write_tables(path: str, tables: dict) -> None:
    """Write ``{name: structured ndarray}`` as 1-D compound datasets of the root group."""
    names = sorted(tables)
    if not 1 <= len(names) <= 2 * GROUP_LEAF_K:
        raise ValueError("between 1 and 8 tables per file")
    arrays = {n: np.ascontiguousarray(tables[n]) for n in names}
    for n, a in arrays.items():
        if a.ndim != 1 or a.dtype.names is None:
            raise ValueError(f"table {n!r} must be a one-dimensional structured array")

    # local heap data segment: "" at offset 0, then the link names, then one free block
    heap_data = _pad8(b"\0")
    name_off = {}
    for n in names:
        name_off[n] = len(heap_data)
        heap_data += _pad8(n.encode("ascii") + b"\0")
    free_off = len(heap_data)
    heap_data += struct.pack("<QQ", 1, 16)      # free block: next = H5HL_FREE_NULL, size 16

    sizes = {
        "super": 96,
        "root_hdr": 16 + 8 + 16,
        "heap": 32,
        "heap_data": len(heap_data),
        "btree": 24 + (2 * GROUP_INTERNAL_K + 1) * 8 + 2 * GROUP_INTERNAL_K * 8,
        "snod": 8 + 2 * GROUP_LEAF_K * 40,
    }
    addr, cur = {}, 0
    for k in ("super", "root_hdr", "heap", "heap_data", "btree", "snod"):
        addr[k] = cur
        cur += (sizes[k] + 7) // 8 * 8
    ds_hdr_addr, ds_hdr = {}, {}
    for n in names:       # headers first (their size does not depend on the data address)
        ds_hdr_addr[n] = cur
        cur += len(_dataset_header(arrays[n].dtype, len(arrays[n]), 0))
    data_addr = {}
    for n in names:
        data_addr[n] = cur
        cur += (arrays[n].nbytes + 7) // 8 * 8
    eof = cur
    for n in names:
        ds_hdr[n] = _dataset_header(arrays[n].dtype, len(arrays[n]), data_addr[n])

    superblock = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, GROUP_LEAF_K, GROUP_INTERNAL_K, 0)
    superblock += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    superblock += struct.pack("<QQII", 0, addr["root_hdr"], 1, 0) + struct.pack("<QQ", addr["btree"], addr["heap"])
    assert len(superblock) == 96

    root_hdr = _object_header([_message(0x0011, struct.pack("<QQ", addr["btree"], addr["heap"]))])
    heap = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, addr["heap_data"])
    btree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF)
    btree += struct.pack("<QQQ", 0, addr["snod"], name_off[names[-1]])
    btree += b"\0" * (sizes["btree"] - len(btree))
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for n in names:
        snod += struct.pack("<QQII16x", name_off[n], ds_hdr_addr[n], 0, 0)
    snod += b"\0" * (sizes["snod"] - len(snod))

    with open(path, "wb") as f:
        def put(at, blob):
            f.seek(at)
            f.write(blob)
        put(addr["super"], superblock)
        put(addr["root_hdr"], root_hdr)
        put(addr["heap"], heap)
        put(addr["heap_data"], heap_data)
        put(addr["btree"], btree)
        put(addr["snod"], snod)
        for n in names:
            put(ds_hdr_addr[n], ds_hdr[n])
            f.seek(data_addr[n])
            arrays[n].tofile(f)
        f.truncate(eof)


def write_table(path: str, name: str, table: np.ndarray) -> None:
    write_tables(path, {name: table})


# ---------------------------------------------------------------------------------------------------
# reader (independent walk of the on-disk structures)
# ---------------------------------------------------------------------------------------------------

class _Buf:
    def __init__(self, data: bytes):
        self.d = data

    def u(self, off, n):
        return int.from_bytes(self.d[off:off + n], "little")


def _parse_datatype(b: _Buf, off: int):
    """Return (numpy dtype, encoded length)."""
    cls, ver = b.d[off] & 0x0F, b.d[off] >> 4
    bits = b.d[off + 1:off + 4]
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
    if cls == 6:
        if ver != 1:
            raise ValueError("only version 1 compound encodings are understood")
        nmemb = bits[0] | (bits[1] << 8)
        p = off + 8
        names, formats, offsets = [], [], []
        for _ in range(nmemb):
            end = b.d.index(b"\0", p)
            name = b.d[p:end].decode("ascii")
            p += (end - p + 1 + 7) // 8 * 8
            moff = b.u(p, 4)
            if b.d[p + 4] != 0:
                raise ValueError("array members are not supported")
            p += 32
            sub, used = _parse_datatype(b, p)
            p += used
            names.append(name)
            formats.append(sub)
            offsets.append(moff)
        return np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size}), p - off
    raise ValueError(f"datatype class {cls} not supported")


def _read_object_header(b: _Buf, addr: int) -> dict:
    if b.d[addr] != 1:
        raise ValueError("only version 1 object headers are understood")
    nmsg, total = b.u(addr + 2, 2), b.u(addr + 8, 4)
    p, end, msgs = addr + 16, addr + 16 + total, {}
    for _ in range(nmsg):
        if p >= end:
            break
        mtype, msize = b.u(p, 2), b.u(p + 2, 2)
        msgs[mtype] = p + 8
        p += 8 + msize
    return msgs


def read_tables(path: str) -> dict:
    """Return ``{name: structured ndarray}`` for every 1-D compound dataset of the root group."""
    with open(path, "rb") as f:
        b = _Buf(f.read())
    if b.d[:8] != SIGNATURE or b.d[8] != 0:
        raise ValueError("not a version-0 HDF5 file")
    if b.d[13] != 8 or b.d[14] != 8:
        raise ValueError("only 8-byte offsets and lengths are understood")
    eof = b.u(40, 8)
    if eof != len(b.d):
        raise ValueError("end-of-file address does not match the file size")
    root_hdr = b.u(56 + 8, 8)
    msgs = _read_object_header(b, root_hdr)
    if 0x11 not in msgs:
        raise ValueError("root group has no symbol table")
    btree, heap = b.u(msgs[0x11], 8), b.u(msgs[0x11] + 8, 8)
    if b.d[heap:heap + 4] != b"HEAP":
        raise ValueError("bad local heap")
    heap_data = b.u(heap + 24, 8)

    def name_at(off):
        s = heap_data + off
        return b.d[s:b.d.index(b"\0", s)].decode("ascii")

    def walk(node):
        if b.d[node:node + 4] == b"SNOD":
            n = b.u(node + 6, 2)
            for i in range(n):
                e = node + 8 + 40 * i
                yield name_at(b.u(e, 8)), b.u(e + 8, 8)
            return
        if b.d[node:node + 4] != b"TREE" or b.d[node + 4] != 0:
            raise ValueError("bad group B-tree node")
        used = b.u(node + 6, 2)
        for i in range(used):
            yield from walk(b.u(node + 24 + 8 + 16 * i, 8))

    out = {}
    for name, hdr in walk(btree):
        m = _read_object_header(b, hdr)
        if not {0x1, 0x3, 0x8} <= set(m):
            continue
        sp = m[0x1]
        if b.d[sp] != 1 or b.d[sp + 1] != 1:
            continue
        nrows = b.u(sp + 8, 8)
        dtype, _ = _parse_datatype(b, m[0x3])
        lay = m[0x8]
        if b.d[lay] != 3 or b.d[lay + 1] != 1:
            raise ValueError("only contiguous version-3 layouts are understood")
        daddr, dsize = b.u(lay + 2, 8), b.u(lay + 10, 8)
        if dsize != nrows * dtype.itemsize:
            raise ValueError("layout size does not match dataspace x datatype")
        out[name] = np.frombuffer(b.d, dtype=dtype, count=nrows, offset=daddr if nrows else 0).copy()
    return out


def read_table(path: str, name: str) -> np.ndarray:
    return read_tables(path)[name]
