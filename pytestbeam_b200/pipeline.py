"""Chunked execution of the hot path: a particle range is cut into chunks that recycle a fixed set of device
buffers (1e10 electrons produce 1.3 TB of hits, far more than fits in HBM at once).

``ChunkPipeline.run`` keeps everything on the device (throughput measurements, digest sink);
``ChunkPipeline.run_to_host`` additionally lands every chunk's hit tables in pinned host memory on a second stream,
double-buffered so that the copy of chunk i overlaps the kernels of chunk i+1.  What travels over PCIe is the
compressed-row form of the tables (``PtbCompactOut`` in include/ptb.h): per hit the column, row and charge, per (particle,
plane) a row count, per particle its trigger bit.  Of its three variants the pipeline takes the most compressed one the
set-up fits (``Simulator.compact_form``): packed (6-byte rows in three u16 arrays, planes back to back, 4-bit counts with
larger ones on an escape list: 48 bytes per electron on the default telescope against 118 for the slabs), narrow (7-byte
rows: 55) or wide (8-byte rows, count bytes: 66).  ``HostChunk.records`` expands any of them to the reference's 18-byte
rows (``ptb_expand_hits*``), straight into the caller's buffer if it names one; ``records_to_fd`` into a file.
"""
from __future__ import annotations

import numpy as np
import torch

from . import native as N
from .engine import (_BASES_WORDS, _TOTALS_WORDS, HITS_DTYPE, PtbError, Simulator, expand_compact,
                     expand_compact_to_fd, plane_of)

_RETRY = N.PTB_DEV_SLAB_OVERFLOW | N.PTB_DEV_SCRATCH_OVERFLOW | N.PTB_DEV_COUNT_OVERFLOW | N.PTB_DEV_PACK_OVERFLOW


def bind_host_to_gpu(index: int) -> bool:
    """Pin the calling process to the CPUs next to GPU `index` (NVML's ideal affinity), so that the pinned sink buffers
    allocated afterwards land on that GPU's NUMA node: with several ranks per box the device-to-host copies otherwise
    cross the socket interconnect.  Returns False (and changes nothing) when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(int(index)))
        finally:
            pynvml.nvmlShutdown()
        return True
    except Exception:
        return False


class HostChunk:
    """One chunk's hit tables in host memory.  The arrays are views of the pipeline's pinned buffers: valid until the
    consumer returns (copy what must live longer)."""

    def __init__(self, lib, index, first, n, D, totals, event_base, time_base, compact=None, slabs=None):
        self.lib, self.index, self.first, self.n, self.D = lib, index, first, n, D
        self.totals, self.event_base, self.time_base = totals, event_base, time_base
        self.compact, self.slabs = compact, slabs

    @property
    def rows(self):
        return self.totals["hits"]

    def records(self, plane: int, out: np.ndarray | None = None, threads: int = 0) -> np.ndarray:
        """The plane's rows as the reference's records (main/device.py:20-28), written to `out` when given."""
        m = self.totals["hits"][plane]
        if self.compact is not None:
            return expand_compact(self.lib, plane_of(self.compact, plane), self.n, self.event_base, out=out,
                                  threads=threads)
        s = self.slabs[plane]
        if out is None:
            out = np.empty(m, dtype=HITS_DTYPE)
        out["event_number"] = s["event"].numpy()
        out["frame"] = 0
        out["column"] = s["col"].numpy().view(np.uint16)
        out["row"] = s["row"].numpy().view(np.uint16)
        out["charge"] = s["charge"].numpy()
        return out


    def records_to_fd(self, plane: int, fd: int, file_offset: int, threads: int = 0) -> int:
        """The plane's records written to file descriptor `fd` from byte `file_offset` on; returns the rows written."""
        if self.compact is not None:
            return expand_compact_to_fd(self.lib, plane_of(self.compact, plane), self.n, self.event_base, fd, file_offset,
                                        threads=threads)
        import os
        buf, off = memoryview(self.records(plane)).cast("B"), int(file_offset)   # a repeated chunk: slab rows
        rows = len(buf) // 18
        while len(buf):
            done = os.pwrite(fd, buf[:1 << 30], off)
            buf, off = buf[done:], off + done
        return rows


class ChunkPipeline:
    def __init__(self, sim: Simulator, chunk: int, *, n_buffers: int = 2, capacities=None, host: bool = False,
                 seed: int = 0, scratch_rows=None, compact: bool | None = None, narrow: bool | None = None,
                 form: str | None = None):
        self.sim, self.chunk, self.seed = sim, int(chunk), int(seed)
        self.compact = bool(host) if compact is None else bool(compact)
        if self.compact and not host:
            raise PtbError("the compressed-row output exists for the trip to the host: use host=True")
        # the most compressed row form the set-up fits (Simulator.compact_form), unless the caller names one
        # (form = "wide" | "narrow" | "packed"; narrow = True / False is the older spelling of "narrow" / "wide")
        if not self.compact:
            self.form = None
        elif form is not None:
            self.form = form
        elif narrow is not None:
            self.form = "narrow" if narrow else "wide"
        else:
            self.form = sim.compact_form()
        self.narrow = self.form in ("narrow", "packed")     # 4-bit counts, rows of all planes back to back
        self.flags = N.PTB_DIGITISE | N.PTB_RECYCLE_SLABS | (N.PTB_COMPACT if self.compact else 0)
        self.caps = list(capacities) if capacities is not None else sim.default_capacities(self.chunk)
        self.rows = list(scratch_rows) if scratch_rows is not None else sim.default_scratch_rows(self.chunk)
        dev, D = sim.device, sim.D
        self.sets = []
        for _ in range(n_buffers):
            s = {"ws": torch.empty(sim.workspace_bytes(self.chunk, self.caps, self.flags, self.rows),
                                   dtype=torch.uint8, device=dev),
                 "totals": torch.zeros(_TOTALS_WORDS, dtype=torch.int64, device=dev),
                 "done": torch.cuda.Event(), "copied": torch.cuda.Event()}
            if self.compact:
                s["cmp"] = sim.alloc_compact(self.chunk, self.caps, form=self.form)
                s["h_cmp"] = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in s["cmp"].items()
                              if isinstance(v, torch.Tensor)}
            else:
                s["slabs"] = sim.alloc_slabs(self.caps)
                if host:
                    s["h_slabs"] = [{k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in sl.items()}
                                    for sl in s["slabs"]]
            if host:
                s["h_totals"] = torch.zeros(_TOTALS_WORDS, dtype=torch.int64).pin_memory()
            self.sets.append(s)
        self.bases = sim.new_bases()
        self._base0 = (0, 0)
        self.copy_stream = torch.cuda.Stream(device=dev) if host else None
        # the chunk's totals (its sizes) come back on a stream of their own, so that the host learns them while the
        # previous chunk's rows are still crossing the link and the next copies queue up behind those without a gap
        self.meta_stream = torch.cuda.Stream(device=dev) if host else None
        self.retries = 0

    def set_bases(self, event: int = 0, time: int = 0):
        """Global prefixes of the first particle of this pipeline's range (multi-GPU shards)."""
        b = torch.zeros_like(self.bases, device="cpu")
        b[0], b[1] = int(event), int(time)
        self.bases.copy_(b)
        self._base0 = (int(event), int(time))

    def _launch(self, s, first, n, bases_in=None, bases_out=None):
        self.sim.launch(first, n, seed=self.seed, flags=self.flags, slabs=s.get("slabs"),
                        bases_in=self.bases if bases_in is None else bases_in,
                        bases_out=self.bases if bases_out is None else bases_out, totals=s["totals"], workspace=s["ws"],
                        scratch_rows=self.rows, compact=s.get("cmp"))

    # -- device-resident -------------------------------------------------------------------------------
    def run(self, first: int, n_chunks: int, after_chunk=None):
        """Enqueue n_chunks chunks starting at particle `first`; no synchronisation, hits stay on the device.
        after_chunk(index, set) is called right after a chunk's kernels are enqueued (same stream: digests)."""
        for i in range(n_chunks):
            s = self.sets[i % len(self.sets)]
            self._launch(s, first + i * self.chunk, self.chunk)
            if after_chunk is not None:
                after_chunk(i, s)

    def totals(self, idx: int = -1) -> dict:
        torch.cuda.synchronize(self.sim.device)
        return Simulator.decode_totals(self.sets[idx]["totals"], self.sim.D)

    def errors(self) -> int:
        """OR of the device error words of all buffer sets (each holds the word of the last chunk it served)."""
        torch.cuda.synchronize(self.sim.device)
        e = 0
        for s in self.sets:
            e |= Simulator.decode_totals(s["totals"], self.sim.D)["error"]
        return e

    # -- with host sink --------------------------------------------------------------------------------
    def _retry(self, i, first, n, ev_base, t_base, tot):
        """Chunk i overflowed a buffer (its counters are exact): once more, alone, with buffers of the sizes it
        reported, as structure-of-arrays slabs (which also covers a saturated count byte)."""
        self.retries += 1
        sim = self.sim
        b_in = sim.new_bases()
        host = torch.zeros(_BASES_WORDS, dtype=torch.int64)
        host[0], host[1] = int(ev_base), int(t_base)
        b_in.copy_(host)
        r = sim.run(first, n, seed=self.seed, bases_in=b_in, capacities=[max(h, 1) for h in tot["hits"]],
                    scratch_rows=[max(h, 1) for h in tot["reserved"]])
        slabs = [{k: v.cpu() for k, v in sl.items()} for sl in r.slabs]
        return HostChunk(sim.lib, i, first, n, sim.D, r.totals, ev_base, t_base, slabs=slabs)

    def run_to_host(self, first: int, n_chunks: int, consume=None, n_total: int | None = None):
        """Run n_chunks chunks (or, with n_total, the particles [first, first+n_total) in chunks) and land every chunk's
        hit tables in pinned host memory.

        consume(chunk: HostChunk) is called once per chunk, in order, while later chunks are already computing.
        Returns (rows, bytes) copied from the device.
        """
        if self.copy_stream is None:
            raise PtbError("pipeline was created without host buffers")
        if n_total is None:
            n_total = n_chunks * self.chunk
        n_chunks = (n_total + self.chunk - 1) // self.chunk
        sim, D = self.sim, self.sim.D
        compute = torch.cuda.current_stream(sim.device)
        rows = nbytes = 0
        ev_base, t_base = self._base0
        pending = None

        def size_of(i):
            return min(self.chunk, n_total - i * self.chunk)

        def finish(i):
            nonlocal rows, nbytes, ev_base, t_base
            s = self.sets[i % len(self.sets)]
            n_i, first_i = size_of(i), first + i * self.chunk
            with torch.cuda.stream(self.meta_stream):
                self.meta_stream.wait_event(s["done"])
                s["h_totals"].copy_(s["totals"], non_blocking=True)
                self.meta_stream.synchronize()
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(s["done"])
                tot = Simulator.decode_totals(s["h_totals"], D)
                nbytes += _TOTALS_WORDS * 8
                err = tot["error"]
                if err & ~_RETRY:
                    raise PtbError(f"device reported error bits {err:#x} in chunk {i}")
                chunk = None
                if err:
                    chunk = self._retry(i, first_i, n_i, ev_base, t_base, tot)
                elif self.compact and self.narrow:
                    c, h = s["cmp"], s["h_cmp"]
                    m, w, half = sum(tot["hits"]), (n_i + 31) // 32, (n_i + 1) // 2
                    row_keys = ("charge", "pos_lo", "pos_hi") if self.form == "narrow" else ("word0", "word1", "word2")
                    for k in row_keys:                           # the planes' rows lie back to back: one copy per array
                        h[k][:m].copy_(c[k][:m], non_blocking=True)
                    if n_i == self.chunk:
                        h["counts"].copy_(c["counts"], non_blocking=True)
                    else:                                        # a shorter last chunk: [D][(n_i+1)//2] sits at the front
                        h["counts"][:D * half].copy_(c["counts"][:D * half], non_blocking=True)
                    h["accepted"][:w].copy_(c["accepted"][:w], non_blocking=True)
                    n_esc = tot["escapes"]
                    if n_esc:
                        h["escapes"][:n_esc].copy_(c["escapes"][:n_esc], non_blocking=True)
                    nbytes += sum(h[k].element_size() for k in row_keys) * m + D * half + 4 * w + 8 * n_esc
                    s["copied"].record(self.copy_stream)
                    if consume is not None:
                        first_row = np.concatenate([[0], np.cumsum(tot["hits"])]).astype(np.int64)
                        cut = lambda k, dt: [h[k].numpy().view(dt)[first_row[d]:first_row[d + 1]] for d in range(D)]  # noqa: E731
                        host = {"form": self.form, "narrow": True,
                                "counts": h["counts"].numpy()[:D * half].reshape(D, half),
                                "escapes": h["escapes"].numpy().view(np.uint64)[:n_esc],
                                "accepted": h["accepted"].numpy().view(np.uint32)[:w]}
                        if self.form == "narrow":
                            host.update(charge=cut("charge", np.float32), pos_lo=cut("pos_lo", np.uint16),
                                        pos_hi=cut("pos_hi", np.uint8))
                        else:
                            host.update(word0=cut("word0", np.uint16), word1=cut("word1", np.uint16),
                                        word2=cut("word2", np.uint16), layout=c["layout"],
                                        row_base=[int(v) for v in first_row[:D]])
                        chunk = HostChunk(sim.lib, i, first_i, n_i, D, tot, ev_base, t_base, compact=host)
                elif self.compact:
                    c, h = s["cmp"], s["h_cmp"]
                    pf, m = c["plane_first"], sum(tot["hits"])
                    for d in range(D):                       # every plane's rows from its own region
                        h["hits"][pf[d]:pf[d] + tot["hits"][d]].copy_(c["hits"][pf[d]:pf[d] + tot["hits"][d]],
                                                                      non_blocking=True)
                    h["counts"][:D * n_i].copy_(c["counts"][:D * n_i], non_blocking=True)
                    w = (n_i + 31) // 32
                    h["accepted"][:w].copy_(c["accepted"][:w], non_blocking=True)
                    nbytes += 8 * m + D * n_i + 4 * w
                    s["copied"].record(self.copy_stream)
                    if consume is not None:
                        flat = h["hits"].numpy().view(np.uint64)
                        chunk = HostChunk(sim.lib, i, first_i, n_i, D, tot, ev_base, t_base, compact={
                            "hits": [flat[pf[d]:pf[d] + tot["hits"][d]] for d in range(D)],
                            "counts": h["counts"].numpy()[:D * n_i].reshape(D, n_i),
                            "accepted": h["accepted"].numpy().view(np.uint32)[:w]})
                else:
                    views = []
                    for d, (sl, hs) in enumerate(zip(s["slabs"], s["h_slabs"])):
                        m = tot["hits"][d]
                        v = {}
                        for k in sl:
                            hs[k][:m].copy_(sl[k][:m], non_blocking=True)
                            v[k] = hs[k][:m]
                            nbytes += m * sl[k].element_size()
                        views.append(v)
                    s["copied"].record(self.copy_stream)
                    if consume is not None:
                        chunk = HostChunk(sim.lib, i, first_i, n_i, D, tot, ev_base, t_base, slabs=views)
                if err:
                    s["copied"].record(self.copy_stream)
            rows += sum(tot["hits"])
            if consume is not None:
                s["copied"].synchronize()
                consume(chunk)
            ev_base += tot["accepted"]
            t_base += tot["time_sum"]

        for i in range(n_chunks):
            s = self.sets[i % len(self.sets)]
            if i >= len(self.sets):
                compute.wait_event(s["copied"])      # the set's previous copy must have drained
            self._launch(s, first + i * self.chunk, size_of(i))
            s["done"].record(compute)
            if pending is not None:
                finish(pending)
            pending = i
        if pending is not None:
            finish(pending)
        self.copy_stream.synchronize()
        return rows, nbytes



def shard_range(n_total: int, rank: int, world: int):
    """Contiguous index range of one rank: [r*N/R, (r+1)*N/R)."""
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi - lo


def weighted_shards(n_units: int, weights) -> list:
    """Contiguous shards of n_units units, one per rank, sized in proportion to the ranks' weights (largest-remainder
    rounding, every rank with a positive weight gets at least one unit while units last): [(first unit, units)] in rank
    order.  Index-range shards keep the reference's row order under rank-order concatenation whatever their sizes; a
    host that takes rows in faster from some GPUs than from others gives those GPUs more particles."""
    w = np.asarray(list(weights), dtype=np.float64)
    if len(w) == 0 or not np.all(np.isfinite(w)) or np.any(w < 0) or w.sum() <= 0:
        w = np.ones(max(len(w), 1))
    share = w / w.sum() * n_units
    n = np.floor(share).astype(np.int64)
    for i in np.argsort(-(share - n), kind="stable")[:int(n_units - n.sum())]:
        n[i] += 1
    for i in np.flatnonzero((n == 0) & (w > 0)):            # nobody idle while another rank has units to spare
        j = int(np.argmax(n))
        if n[j] > 1:
            n[j] -= 1
            n[i] += 1
    first = np.concatenate([[0], np.cumsum(n)[:-1]])
    return [(int(f), int(k)) for f, k in zip(first, n)]


def exclusive_prefix(values: np.ndarray, rank: int) -> np.ndarray:
    """Sum of the rows of all lower ranks (values: [world, k])."""
    return values[:rank].sum(axis=0) if rank > 0 else np.zeros(values.shape[1], dtype=values.dtype)
