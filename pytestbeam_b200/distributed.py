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
    """Concatenate the per-rank part files of every device in rank order into ``<folder><name>_dut.h5``."""
    for name in names:
        parts = [h5min.read_table(part_path(folder, name, r), "Hits") for r in range(size)]
        table = np.concatenate(parts) if parts else np.zeros(0, dtype=HITS_DTYPE)
        h5min.write_table(f"{folder}{name}_dut.h5", "Hits", table)
        if remove:
            for r in range(size):
                os.remove(part_path(folder, name, r))


def run_sharded(beam: dict, devices: list, materials: list, names: list, folder: str, seed: int, log=None,
                chunk: int = 4_000_000) -> dict:
    """Simulate this rank's index range on its GPU, write part files, merge on rank 0.  Returns this rank's totals."""
    from .engine import Simulator
    from .main.device import _packed_rows
    rank, size = world()
    n_total = int(beam["nmb_particles"])
    lo, n = shard_range(n_total, rank, size)
    sim = Simulator(beam, devices, materials)
    acc, tsum = sim.prefix(lo, n, seed=seed) if n else (0, 0)
    ev_base, t_base = exchange_bases(acc, tsum, device=sim.device if dist.is_initialized()
                                     and dist.get_backend() == "nccl" else None)
    bases = sim.new_bases()
    host = torch.zeros_like(bases, device="cpu")
    host[0], host[1] = ev_base, t_base
    bases.copy_(host)
    parts = [[] for _ in devices]
    hits = [0] * len(devices)
    for c0 in range(lo, lo + n, chunk):
        m = min(chunk, lo + n - c0)
        r = sim.run(c0, m, seed=seed, bases_in=bases, capacities=sim.default_capacities(m),
                        scratch_rows=sim.default_scratch_rows(m))
        bases = r.bases_out
        for d in range(len(devices)):
            parts[d].append(_packed_rows(sim, r.slabs[d]))
            hits[d] += r.totals["hits"][d]
    for d, name in enumerate(names):
        table = np.concatenate(parts[d]) if parts[d] else np.zeros(0, dtype=HITS_DTYPE)
        h5min.write_table(part_path(folder, name, rank), "Hits", table)
    if size > 1:
        dist.barrier()
    if rank == 0:
        merge_part_files(folder, names, size)
    if size > 1:
        dist.barrier()
    return {"rank": rank, "first": lo, "n": n, "event_base": ev_base, "time_base": t_base, "hits": hits}
