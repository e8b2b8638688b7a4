"""Multi-GPU execution: one process per GPU (torchrun), particles sharded by contiguous index range.

The path is embarrassingly parallel.  The only cross-shard state are two running counters of
the reference's sequential loops -- the accepted-event number (``main/device.py:134,164-165``) and the cumulative
time stamp (``main/hit.py:239``).  Every rank computes them for its own range with the cheap ``ptb_prefix`` kernel,
one all-gather of 2 int64 per rank (NCCL over NVLink on GPUs, gloo in the CPU tests) turns them into bases, and
from then on no rank talks to another.  Hit tables are written as per-rank part files and concatenated in rank
order, which reproduces the reference's row order.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import h5min
from .engine import HITS_DTYPE
from .pipeline import exclusive_prefix, shard_range


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def exchange_bases(accepted: int, time_sum: int, device=None):
    """(event base, time base) of this rank = sums over all lower ranks of (accepted, time_sum)."""
    rank, size = world()
    if size == 1:
        return 0, 0
    mine = torch.tensor([int(accepted), int(time_sum)], dtype=torch.int64, device=device or "cpu")
    parts = [torch.zeros_like(mine) for _ in range(size)]
    dist.all_gather(parts, mine)
    base = exclusive_prefix(torch.stack(parts).cpu().numpy(), rank)
    return int(base[0]), int(base[1])


def part_path(folder: str, name: str, rank: int) -> str:
    return f"{folder}{name}_dut.part{rank:03d}.h5"


def merge_part_files(folder: str, names: list, size: int, remove: bool = True) -> None:
    """Concatenate the per-rank part files of every device in rank order into ``<folder><name>_dut.h5``: the parts are
    appended piece by piece to a streamed table (``h5min.concatenate_tables``), none is ever read as a whole."""
    for name in names:
        parts = [part_path(folder, name, r) for r in range(size)]
        h5min.concatenate_tables(f"{folder}{name}_dut.h5", "Hits", parts, HITS_DTYPE)
        if remove:
            for p in parts:
                os.remove(p)


def run_sharded(beam: dict, devices: list, materials: list, names: list, folder: str, seed: int, log=None,
                chunk: int = 4_000_000, keep: int | None = None) -> dict:
    """Simulate this rank's index range on its GPU, stream the rows into per-rank part files, merge on rank 0.

    keep: write only the rows of particles [0, keep) to the files (BASELINE config 5: 1e10 electrons are 1.3 TB of
    rows; the rest only counts).  Returns this rank's totals."""
    from .engine import Simulator
    from .main.device import FileSink, stream_fused
    rank, size = world()
    n_total = int(beam["nmb_particles"])
    lo, n = shard_range(n_total, rank, size)
    sim = Simulator(beam, devices, materials)
    acc, tsum = sim.prefix(lo, n, seed=seed) if n else (0, 0)
    ev_base, t_base = exchange_bases(acc, tsum, device=sim.device if dist.is_initialized()
                                     and dist.get_backend() == "nccl" else None)
    n_keep = n if keep is None else max(0, min(lo + n, int(keep)) - lo)
    sink = FileSink([part_path(folder, name, rank) for name in names])
    hits = [0] * len(devices)
    try:
        if n_keep:
            tot = stream_fused(sim, lo, n_keep, seed, sink, ev_base, t_base, chunk=chunk)
            hits = tot["hits"]
    finally:
        sink.close()
    if size > 1:
        dist.barrier()
    if rank == 0:
        merge_part_files(folder, names, size)
    if size > 1:
        dist.barrier()
    return {"rank": rank, "first": lo, "n": n, "kept": n_keep, "event_base": ev_base, "time_base": t_base, "hits": hits}
