"""Developer probe: time tracks-only, digitise-only (from table) and fused calls of ptb_simulate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytestbeam_b200 import config as cfg, native as N
from pytestbeam_b200.engine import Simulator, _TOTALS_WORDS

n = 4_194_304
beam, devices, materials, names = cfg.flatten(cfg.c1(1000))
sim = Simulator(beam, devices, materials, pool_entries=int(os.environ.get("PTB_POOL", 0)))
caps, rows = [int(n * 1.4) + 4096] * sim.D, sim.default_scratch_rows(n)
slabs = sim.alloc_slabs(caps)
tracks = sim.alloc_tracks(n)
totals = torch.zeros(_TOTALS_WORDS, dtype=torch.int64, device=sim.device)
bases = sim.new_bases()
allflags = N.PTB_DIGITISE | N.PTB_RECYCLE_SLABS | N.PTB_WRITE_TRACKS
ws = torch.empty(sim.workspace_bytes(n, caps, allflags, rows), dtype=torch.uint8, device=sim.device)

def timeit(label, flags, use_tracks):
    for _ in range(2):
        sim.launch(0, n, seed=7, flags=flags, slabs=slabs, tracks=tracks if use_tracks else None, bases_in=bases, bases_out=bases, totals=totals, workspace=ws, scratch_rows=rows)
    torch.cuda.synchronize()
    sim.timing(True)
    for _ in range(3):
        sim.launch(0, n, seed=7, flags=flags, slabs=slabs, tracks=tracks if use_tracks else None, bases_in=bases, bases_out=bases, totals=totals, workspace=ws, scratch_rows=rows)
    torch.cuda.synchronize()
    a, b, k = sim.timing_read()
    sim.timing(False)
    ms = a + b
    print(f"{label:28s} k_tracks {a / k:7.3f} ms  k_digitise {b / k:7.3f} ms  ({n / (ms / k) * 1e3:.3e} /s)")
    return ms / k

timeit("fused + write tracks", allflags, True)
f = timeit("fused", N.PTB_DIGITISE | N.PTB_RECYCLE_SLABS, False)
t = timeit("tracks only (no output)", N.PTB_RECYCLE_SLABS, False)
d = timeit("digitise from table", N.PTB_DIGITISE | N.PTB_RECYCLE_SLABS | N.PTB_TRACKS_FROM_TABLE, True)
print(f"tracks + digitise = {t + d:.3f} ms vs fused {f:.3f} ms")
