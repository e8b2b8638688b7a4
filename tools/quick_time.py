"""Developer probe: time ptb_simulate (Philox mode, fused, default telescope) for a few chunk sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytestbeam_b200 import config as cfg, native as N
from pytestbeam_b200.engine import Simulator, _TOTALS_WORDS

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
beam, devices, materials, names = cfg.flatten(getattr(cfg, name)(1000))
sim = Simulator(beam, devices, materials, pool_entries=int(os.environ.get("PTB_POOL", 0)))
hpp = sim.hits_per_particle()
print("hits/particle", hpp.round(4), "total", hpp.sum().round(3))
for n in [int(v) for v in os.environ.get("PTB_NS", "1048576,4194304,10000000").split(",")]:
    caps, rows = sim.default_capacities(n), sim.default_scratch_rows(n)
    flags = N.PTB_DIGITISE | N.PTB_RECYCLE_SLABS
    slabs = sim.alloc_slabs(caps)
    ws = torch.empty(sim.workspace_bytes(n, caps, flags, rows), dtype=torch.uint8, device=sim.device)
    totals = torch.zeros(_TOTALS_WORDS, dtype=torch.int64, device=sim.device)
    bases = sim.new_bases()
    for it in range(2):
        sim.launch(0, n, seed=7, flags=flags, slabs=slabs, bases_in=bases, bases_out=bases, totals=totals, workspace=ws, scratch_rows=rows)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for it in range(reps):
        sim.launch(it * n, n, seed=7, flags=flags, slabs=slabs, bases_in=bases, bases_out=bases, totals=totals, workspace=ws, scratch_rows=rows)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tot = sim.decode_totals(totals, sim.D)
    print(f"n={n}: {ms:.2f} ms/call -> {n / ms * 1e3:.3e} particles/s, hits {sum(tot['hits'])} err {tot['error']} "
          f"ws {ws.numel() / 1e6:.0f} MB pixels/particle {tot['pixels'] / n:.2f}")
print("fp64 peak probe: %.2f TFLOP/s" % (sim.lib.ptb_measure_fp64_peak(None) / 1e12))
