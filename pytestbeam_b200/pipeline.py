"""Chunked execution of the hot path: a particle range is cut into chunks that recycle a fixed set of device
buffers (1e10 electrons produce 1.3 TB of hits, far more than fits in HBM at once).

``ChunkPipeline.run`` keeps everything on the device (throughput measurements, digest sink);
``ChunkPipeline.run_to_host`` additionally copies every chunk's hit slabs into pinned host memory on a
second stream, double-buffered so that the copy of chunk i overlaps the kernels of chunk i+1.
"""
from __future__ import annotations

import numpy as np
import torch

from . import native as N
from .engine import _TOTALS_WORDS, PtbError, Simulator


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


class ChunkPipeline:
    def __init__(self, sim: Simulator, chunk: int, *, n_buffers: int = 2, capacities=None, host: bool = False,
                 seed: int = 0, scratch_rows=None):
        self.sim, self.chunk, self.seed = sim, int(chunk), int(seed)
        self.flags = N.PTB_DIGITISE | N.PTB_RECYCLE_SLABS
        self.caps = list(capacities) if capacities is not None else sim.default_capacities(self.chunk)
        self.rows = list(scratch_rows) if scratch_rows is not None else sim.default_scratch_rows(self.chunk)
        dev = sim.device
        self.sets = []
        for _ in range(n_buffers):
            s = {"slabs": sim.alloc_slabs(self.caps),
                 "ws": torch.empty(sim.workspace_bytes(self.chunk, self.caps, self.flags, self.rows),
                                   dtype=torch.uint8, device=dev),
                 "totals": torch.zeros(_TOTALS_WORDS, dtype=torch.int64, device=dev),
                 "done": torch.cuda.Event(), "copied": torch.cuda.Event()}
            if host:
                s["h_totals"] = torch.zeros(_TOTALS_WORDS, dtype=torch.int64).pin_memory()
                s["h_slabs"] = [{k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in sl.items()}
                                for sl in s["slabs"]]
            self.sets.append(s)
        self.bases = sim.new_bases()
        self.copy_stream = torch.cuda.Stream(device=dev) if host else None

    def set_bases(self, event: int = 0, time: int = 0):
        """Global prefixes of the first particle of this pipeline's range (multi-GPU shards)."""
        b = torch.zeros_like(self.bases, device="cpu")
        b[0], b[1] = int(event), int(time)
        self.bases.copy_(b)

    def _launch(self, s, first, n):
        self.sim.launch(first, n, seed=self.seed, flags=self.flags, slabs=s["slabs"], bases_in=self.bases,
                        bases_out=self.bases, totals=s["totals"], workspace=s["ws"], scratch_rows=self.rows)

    # -- device-resident -------------------------------------------------------------------------------
    def run(self, first: int, n_chunks: int):
        """Enqueue n_chunks chunks starting at particle `first`; no synchronisation, hits stay on the device."""
        for i in range(n_chunks):
            self._launch(self.sets[i % len(self.sets)], first + i * self.chunk, self.chunk)

    def totals(self, idx: int = -1) -> dict:
        torch.cuda.synchronize(self.sim.device)
        return Simulator.decode_totals(self.sets[idx]["totals"], self.sim.D)

    # -- with host sink --------------------------------------------------------------------------------
    def run_to_host(self, first: int, n_chunks: int, consume=None):
        """Run n_chunks chunks and land every chunk's hits in pinned host memory.

        consume(chunk_index, totals: dict, host_slabs: list[dict of pinned tensors trimmed to the row counts])
        is called once per chunk while later chunks are already computing.  Returns (rows, bytes) copied.
        """
        if self.copy_stream is None:
            raise PtbError("pipeline was created without host buffers")
        compute = torch.cuda.current_stream(self.sim.device)
        rows = nbytes = 0
        pending = None

        def finish(i):
            nonlocal rows, nbytes
            s = self.sets[i % len(self.sets)]
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(s["done"])
                s["h_totals"].copy_(s["totals"], non_blocking=True)
                self.copy_stream.synchronize()
                tot = Simulator.decode_totals(s["h_totals"], self.sim.D)
                if tot["error"]:
                    raise PtbError(f"device reported error bits {tot['error']:#x} in chunk {i}")
                views = []
                for d, (sl, hs) in enumerate(zip(s["slabs"], s["h_slabs"])):
                    m = tot["hits"][d]
                    v = {}
                    for k in sl:
                        hs[k][:m].copy_(sl[k][:m], non_blocking=True)
                        v[k] = hs[k][:m]
                        nbytes += m * sl[k].element_size()
                    rows += m
                    views.append(v)
                s["copied"].record(self.copy_stream)
            nbytes += _TOTALS_WORDS * 8
            if consume is not None:
                s["copied"].synchronize()
                consume(i, tot, views)

        for i in range(n_chunks):
            s = self.sets[i % len(self.sets)]
            if i >= len(self.sets):
                compute.wait_event(s["copied"])      # the set's previous copy must have drained
            self._launch(s, first + i * self.chunk, self.chunk)
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


def exclusive_prefix(values: np.ndarray, rank: int) -> np.ndarray:
    """Sum of the rows of all lower ranks (values: [world, k])."""
    return values[:rank].sum(axis=0) if rank > 0 else np.zeros(values.shape[1], dtype=values.dtype)
