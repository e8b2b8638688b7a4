"""Developer probe: per-stage times of ptb_simulate measured live with CUDA events (k_tracks, k_digitise, k_digitise_big +
scans, k_gather) against the whole step, for slab and compact output; 1e7-electron chunks of the default telescope."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytestbeam_b200 import config as cfg
from pytestbeam_b200.engine import Simulator
from pytestbeam_b200.pipeline import ChunkPipeline

chunk, K = int(float(os.environ.get("PTB_CHUNK", "1e7"))), int(os.environ.get("PTB_K", "10"))
name = sys.argv[1] if len(sys.argv) > 1 else "c1"
beam, devices, materials, names = cfg.flatten(getattr(cfg, name)(1000))
sim = Simulator(beam, devices, materials)
for timing in (False, True):
    pipe = ChunkPipeline(sim, chunk, seed=7)
    pipe.run(0, 3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sim.timing(timing)
    e0.record()
    pipe.run(0, K)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    line = f"timing events {'on ' if timing else 'off'}: {ms:.3f} ms/step -> {chunk / ms * 1e3:.4e} /s"
    if timing:
        st, n = sim.timing_read_stages()
        line += "  stages " + " ".join(f"{v / n:.3f}" for v in st) + f"  sum {sum(st) / n:.3f}"
    sim.timing(False)
    print(line, "err", pipe.errors(), flush=True)
