"""Developer probe for ncu: three ptb_simulate calls (Philox mode, default telescope, 4M electrons each)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytestbeam_b200 import config as cfg, native as N
from pytestbeam_b200.engine import Simulator, _TOTALS_WORDS

name = sys.argv[1] if len(sys.argv) > 1 else "c1"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 4_000_000
beam, devices, materials, names = cfg.flatten(getattr(cfg, name)(1000))
sim = Simulator(beam, devices, materials)
caps, rows = sim.default_capacities(n), sim.default_scratch_rows(n)
flags = N.PTB_DIGITISE | N.PTB_RECYCLE_SLABS
slabs = sim.alloc_slabs(caps)
ws = torch.empty(sim.workspace_bytes(n, caps, flags, rows), dtype=torch.uint8, device=sim.device)
totals = torch.zeros(_TOTALS_WORDS, dtype=torch.int64, device=sim.device)
bases = sim.new_bases()
for it in range(3):
    sim.launch(it * n, n, seed=7, flags=flags, slabs=slabs, bases_in=bases, bases_out=bases, totals=totals, workspace=ws, scratch_rows=rows)
torch.cuda.synchronize()
print(sim.decode_totals(totals, sim.D))
