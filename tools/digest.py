"""Developer tool: sha256 digests of the production-mode hit tables (and track state) of a few set-ups, to check that
a change of the kernels leaves the output bit-identical.  Run once per library build and diff the printed lines:

    PTB_LIB_PATH=pytestbeam_b200/_lib/var_base.so python tools/digest.py > a.txt; python tools/digest.py > b.txt
"""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pytestbeam_b200 import config as cfg
from pytestbeam_b200.engine import Simulator

n = int(float(os.environ.get("PTB_DIGEST_N", "2e6")))
for name in ("c1", "c3", "c4"):
    beam, devices, materials, names = cfg.flatten(getattr(cfg, name)(n))
    for pool in (0, 96):
        sim = Simulator(beam, devices, materials, pool_entries=pool)
        res = sim.run(12345, n, seed=99, write_tracks=(pool == 0))
        for d in range(sim.D):
            h = hashlib.sha256()
            for k in ("event", "col", "row", "charge"):
                h.update(res.slabs[d][k].cpu().numpy().tobytes())
            print(name, pool, names[d], res.totals["hits"][d], h.hexdigest()[:20])
        if pool == 0:
            h = hashlib.sha256()
            for k in ("x", "y", "energy", "time_stamp"):
                h.update(res.tracks[k].cpu().numpy().tobytes())
            print(name, "tracks", h.hexdigest()[:20], res.totals["pairs"], res.totals["inside"], res.totals["pixels"])
