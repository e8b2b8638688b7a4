"""Developer probe: wall time of the drop-in entry point on the reference's default set-up (2e6 electrons)."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yaml
from pytestbeam_b200 import config as cfg, pytestbeam, h5min
d = tempfile.mkdtemp()
setup = cfg.default_setup(int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000)
setup["plotting"] = False
setup["beam"]["seed"] = 1
open(os.path.join(d, "setup.yml"), "w").write(yaml.safe_dump(setup, sort_keys=False))
open(os.path.join(d, "material.yml"), "w").write(yaml.safe_dump(cfg.default_material()))
for rep in range(2):
    t = time.time()
    pytestbeam.main(d)
    print(f"run {rep}: {time.time() - t:.2f} s")
rows = sum(len(h5min.read_table(os.path.join(d, "output_data", f), "Hits")) for f in os.listdir(os.path.join(d, "output_data")))
print("rows written", rows)
