"""Drop-in for the hot-path half of the reference's ``main/device.py``: same names, arguments, side effects.

``calculate_device_hit`` (reference ``main/device.py:7-88``) writes one ``<folder><name>_dut.h5`` per device with
table ``/Hits`` (event_number <i8, frame <u2, column <u2, row <u2, charge <f4; rows in particle order).

Two routes, chosen by what arrives as ``hit_data``:

* the untouched ``HitTables`` object of ``pytestbeam_b200.main.hit.tracks``: tracks and digitisation run fused on the GPU
  (``ChunkPipeline.run_to_host``: no per-chunk allocation or synchronisation, the compressed-row tables cross PCIe while
  the next chunk computes) and every chunk's rows are expanded straight into the memory-mapped dataset region of the
  output files -- the tables are never held in host memory as a whole;
* anything else that looks like the reference's 8-tuple (``main/hit.py:138``: a ``HitTables`` that has been
  materialised, the reference's own typed Lists, plain numpy arrays): ``calc_position`` alone, reading ``x =
  hit_data[3]``, ``y = hit_data[4]``, ``energy_lost = hit_data[7]`` exactly as ``main/device.py:132-138`` does.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import h5min
from .. import native as N
from ..config import SILICON
from ..engine import HITS_DTYPE, Simulator
from ..pipeline import ChunkPipeline
from .hit import CHUNK, HitTables, _seed_of


def create_hit_file(hit_data: np.ndarray, folder: str, index: str) -> str:
    """Write ``<folder><index>_dut.h5`` (reference ``main/device.py:306-340``); returns the path."""
    hit_data = np.asarray(hit_data, dtype=HITS_DTYPE)
    path = folder + index + "_dut.h5"
    h5min.write_table(path, "Hits", hit_data)
    return path


# ---------------------------------------------------------------------------------------------------
# sinks: where the rows of every chunk go
# ---------------------------------------------------------------------------------------------------

class MemorySink:
    """Keeps every plane's rows in host memory (``device_hits``)."""

    def __init__(self, D: int):
        self.parts = [[] for _ in range(D)]

    def room(self, plane: int, rows: int) -> np.ndarray:
        out = np.empty(rows, dtype=HITS_DTYPE)
        self.parts[plane].append(out)
        return out

    def done(self, plane: int, view) -> None:
        pass

    def tables(self) -> list:
        return [np.concatenate(p) if p else np.zeros(0, dtype=HITS_DTYPE) for p in self.parts]

    def close(self) -> None:
        pass


class FileSink:
    """One streamed ``/Hits`` table per device.  Every plane's rows are expanded by the library straight into its file
    (``ptb_expand_hits*_fd``: 4 MiB blocks, positioned writes) -- no staging copy of the table, no page faults on
    freshly mapped file pages; the planes of a chunk are served by one thread each (the call releases the GIL)."""

    parallel = True

    def __init__(self, paths: list):
        self.writers = [h5min.TableWriter(p, "Hits", HITS_DTYPE) for p in paths]

    def write(self, plane: int, chunk, threads: int = 0) -> None:
        self.writers[plane].append_with(chunk.rows[plane],
                                        lambda fd, off: chunk.records_to_fd(plane, fd, off, threads=threads))

    def room(self, plane: int, rows: int) -> np.ndarray:       # records that arrive ready-made (calc_position alone)
        return np.empty(rows, dtype=HITS_DTYPE)

    def done(self, plane: int, view) -> None:
        self.writers[plane].append(view)

    def close(self) -> None:
        for w in self.writers:
            w.close()


_PLANE_POOL = None      # worker threads shared by every run of the process: one per plane of the widest telescope seen


def _plane_pool(planes: int):
    global _PLANE_POOL
    from concurrent.futures import ThreadPoolExecutor
    if _PLANE_POOL is None or _PLANE_POOL._max_workers < planes:
        if _PLANE_POOL is not None:
            _PLANE_POOL.shutdown(wait=True)
        _PLANE_POOL = ThreadPoolExecutor(max_workers=planes, thread_name_prefix="ptb-plane")
    return _PLANE_POOL


def _consume_into(sink):
    """consume(chunk) for ChunkPipeline.run_to_host: every plane's rows go to the sink as the reference's records."""
    import os

    def one(chunk, d, threads):
        if hasattr(sink, "write"):                   # the sink takes the rows itself (files)
            sink.write(d, chunk, threads)
            return
        out = sink.room(d, chunk.rows[d])
        chunk.records(d, out=out, threads=threads)
        sink.done(d, out)

    def consume(chunk):
        if not getattr(sink, "parallel", False) or chunk.D == 1:
            for d in range(chunk.D):
                one(chunk, d, 0)
            return
        per = max(1, (os.cpu_count() or 1) // chunk.D)       # threads of each plane's expansion
        for f in [_plane_pool(chunk.D).submit(one, chunk, d, per) for d in range(chunk.D)]:
            f.result()
    return consume


def fused_pipeline(sim: Simulator, n: int, seed: int, chunk: int = CHUNK) -> ChunkPipeline:
    """The host pipeline stream_fused runs n particles through (its pinned buffers are the largest fixed cost of a run:
    a caller that times the run builds it beforehand)."""
    chunk = max(32, min(int(chunk), n))
    # a range that fits one chunk needs one buffer set
    return ChunkPipeline(sim, chunk, host=True, seed=seed, n_buffers=1 if n <= chunk else 2)


def stream_fused(sim: Simulator, first: int, n: int, seed: int, sink, event_base: int = 0, time_base: int = 0,
                 chunk: int = CHUNK, pipe: ChunkPipeline | None = None) -> dict:
    """Particles [first, first+n) through the fused GPU path into `sink`.  Returns the summed counters."""
    pipe = pipe or fused_pipeline(sim, n, seed, chunk)
    pipe.set_bases(event_base, time_base)
    tot = {"hits": [0] * sim.D, "accepted": 0, "time_sum": 0}
    inner = _consume_into(sink)

    def consume(c):
        inner(c)
        for d in range(sim.D):
            tot["hits"][d] += c.totals["hits"][d]
        tot["accepted"] += c.totals["accepted"]
        tot["time_sum"] += c.totals["time_sum"]

    rows, nbytes = pipe.run_to_host(first, 0, consume=consume, n_total=n)
    tot.update(rows=rows, d2h_bytes=nbytes, retries=pipe.retries)
    return tot


def stream_from_arrays(sim: Simulator, x, y, eloss, seed: int, sink, accepted=None, chunk: int = CHUNK) -> None:
    """calc_position alone: x, y flat particle-major [N*D] sequences, eloss [D][N] (after the acceptance zeroing), read
    chunk by chunk so that the reference's typed Lists are never converted as a whole."""
    D = sim.D
    n = len(x) // D
    eloss = eloss if isinstance(eloss, np.ndarray) else np.asarray(eloss, dtype=np.float64)
    dev = sim.device
    bases = None

    def piece(seq, lo, hi):
        a = seq[lo:hi] if isinstance(seq, np.ndarray) else np.fromiter((seq[i] for i in range(lo, hi)), np.float64, hi - lo)
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)

    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        tin = {"x": piece(x, lo * D, (lo + m) * D), "y": piece(y, lo * D, (lo + m) * D),
               "eloss": torch.from_numpy(np.ascontiguousarray(eloss[:, lo:lo + m], dtype=np.float64)).to(dev)}
        if accepted is not None:
            tin["accepted"] = torch.from_numpy(np.ascontiguousarray(accepted[lo:lo + m], dtype=np.uint8)).to(dev)
        r = sim.run(lo, m, seed=seed, tracks_in=tin, bases_in=bases, capacities=sim.default_capacities(m),
                    scratch_rows=sim.default_scratch_rows(m))
        bases = r.bases_out
        for d in range(D):
            rows = r.totals["hits"][d]
            out = sink.room(d, rows)
            if rows:
                packed = sim.pack_hits(r.slabs[d]).cpu().numpy()
                out.view(np.uint8).reshape(-1)[:] = packed
            sink.done(d, out)


def _is_hit_tuple(hit_data) -> bool:
    try:
        return len(hit_data) == 8 and len(hit_data[3]) == len(hit_data[4])
    except TypeError:
        return False


LAST_RUN: dict = {}     # counters of the most recent calculate_device_hit / device_hits (the front end's run record)


def _run(beam: dict, devices: list, hit_data, sink) -> None:
    n, D = int(beam["nmb_particles"]), len(devices)
    chunk = int(beam.get("chunk") or CHUNK)          # optional key of the beam block; the default suits any size
    LAST_RUN.clear()
    if isinstance(hit_data, HitTables) and not hit_data.materialised:
        tot = stream_fused(hit_data.sim, 0, n, hit_data.seed, sink, chunk=chunk)
        LAST_RUN.update(route="fused", seed=hit_data.seed, electrons=n, chunk=min(chunk, n), **tot)
        return
    if not _is_hit_tuple(hit_data):
        raise TypeError("hit_data must be the 8-tuple of hit.tracks: (energy, anglex, angley, x, y, z, time_stamp, "
                        "energy_lost), main/hit.py:138")
    if len(hit_data[3]) != n * D:
        raise ValueError(f"hit_data holds {len(hit_data[3])} track records, expected nmb_particles * devices = {n * D}")
    if isinstance(hit_data, HitTables):
        sim, seed = hit_data.sim, hit_data.seed
    else:
        # a foreign tuple: the digitisation needs the device geometry only, the material entries stay unused
        sim, seed = Simulator(beam, devices, [dict(SILICON) for _ in devices]), _seed_of(beam)
    stream_from_arrays(sim, hit_data[3], hit_data[4], hit_data[7], seed, sink, chunk=chunk)
    LAST_RUN.update(route="calc_position", seed=seed, electrons=n, chunk=min(chunk, n))


def device_hits(beam: dict, devices: list, hit_data, log=None) -> list:
    """The hit table of every device (structured arrays), i.e. calculate_device_hit without the file writing."""
    sink = MemorySink(len(devices))
    _run(beam, devices, hit_data, sink)
    return sink.tables()


def device_hits_from_arrays(sim: Simulator, x, y, eloss, seed: int, accepted=None) -> list:
    """Digitise given track positions: x, y flat particle-major [N*D], eloss [D, N] (after the acceptance zeroing)."""
    sink = MemorySink(sim.D)
    stream_from_arrays(sim, x, y, eloss, seed, sink, accepted=accepted)
    return sink.tables()


def calculate_device_hit(beam: dict, devices: list, hit_data, names: list, folder: str, log) -> None:
    """Create the HDF5 hit tables from the particle table and the device configuration (reference ``main/device.py:7``)."""
    log.info("Calculating particle hit positions")
    sink = FileSink([folder + name + "_dut.h5" for name in names])
    try:
        _run(beam, devices, hit_data, sink)
    finally:
        log.info("Saving Data...")
        sink.close()
    for name in names:
        log.info("Hit positions of %s" % name)
