"""Developer probe: where the wall time of the entry point (hit.tracks + device.calculate_device_hit, files included) goes."""
import logging, os, shutil, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pytestbeam_b200 import config as cfg
from pytestbeam_b200.engine import Simulator
from pytestbeam_b200.main import device, hit
from pytestbeam_b200.pipeline import ChunkPipeline

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
chunk = int(float(sys.argv[2])) if len(sys.argv) > 2 else device.CHUNK
beam, devices, materials, names = cfg.flatten(cfg.default_setup(n))
beam["seed"] = 5
log = logging.getLogger("x")
T = time.perf_counter
t = T(); sim = Simulator(beam, devices, materials); sim.hits_per_particle(); torch.cuda.synchronize(); print(f"simulator + pilot {T() - t:.3f} s")
c = max(32, min(chunk, n))
t = T(); pipe = ChunkPipeline(sim, c, host=True, seed=5); torch.cuda.synchronize(); print(f"pipeline buffers (chunk {c:.3g}) {T() - t:.3f} s")
for rep in range(2):
    t = T(); rows, nb = pipe.run_to_host(0, 0, n_total=n); print(f"run_to_host, no consumer {T() - t:.3f} s ({nb / 1e6:.0f} MB)")
sink = device.MemorySink(sim.D)
t = T(); pipe.run_to_host(0, 0, consume=device._consume_into(sink), n_total=n); print(f"run_to_host -> expand into fresh host arrays {T() - t:.3f} s")
bufs = [np.empty(sum(len(p) for p in sink.parts[d]), dtype=device.HITS_DTYPE) for d in range(sim.D)]
class Pre:
    def __init__(s): s.at = [0] * sim.D
    def room(s, d, rows):
        v = bufs[d][s.at[d]:s.at[d] + rows]; s.at[d] += rows; return v
    def done(s, d, v): pass
t = T(); pipe.run_to_host(0, 0, consume=device._consume_into(Pre()), n_total=n); print(f"run_to_host -> expand into touched host arrays {T() - t:.3f} s")
for where in (None, "/dev/shm"):
    tmp = tempfile.mkdtemp(dir=where)
    fs = device.FileSink([f"{tmp}/{nm}_dut.h5" for nm in names])
    t = T(); pipe.run_to_host(0, 0, consume=device._consume_into(fs), n_total=n); t1 = T(); fs.close(); t2 = T()
    print(f"run_to_host -> files in {where or tempfile.gettempdir()}: {t1 - t:.3f} s + close {t2 - t1:.3f} s")
    shutil.rmtree(tmp)
t = T(); tables = hit.tracks(beam, devices, materials, log); tmp = tempfile.mkdtemp(); device.calculate_device_hit(beam, devices, tables, names, tmp + "/", log)
print(f"whole entry point {T() - t:.3f} s -> {n / (T() - t):.3e} electrons/s"); shutil.rmtree(tmp)
