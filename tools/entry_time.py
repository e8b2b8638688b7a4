"""Developer probe: wall time of the drop-in entry point (python -m pytestbeam_b200.pytestbeam) on the reference's default
set-up: `python tools/entry_time.py [electrons=2e6] [repeats=2]`.  The seven *_dut.h5 files are written to a scratch
directory (18 bytes per hit row: 133 bytes per electron) and removed afterwards."""
import os, shutil, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import yaml
from pytestbeam_b200 import config as cfg, pytestbeam, h5min
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
d = tempfile.mkdtemp(dir=os.environ.get("PTB_ENTRY_DIR") or None)
free = shutil.disk_usage(d).free
if free < 160 * n:
    raise SystemExit(f"{d}: {free / 1e9:.1f} GB free, the files of {n:.3g} electrons need {133 * n / 1e9:.1f} GB")
try:
    setup = cfg.default_setup(n)
    setup["plotting"] = False
    setup["beam"]["seed"] = 1
    open(os.path.join(d, "setup.yml"), "w").write(yaml.safe_dump(setup, sort_keys=False))
    open(os.path.join(d, "material.yml"), "w").write(yaml.safe_dump(cfg.default_material()))
    for rep in range(reps):
        t = time.time()
        pytestbeam.main(d)
        dt = time.time() - t
        print(f"run {rep}: {dt:.2f} s -> {n / dt:.4g} electrons/s", flush=True)
    out = os.path.join(d, "output_data")
    rows = size = 0
    for f in (f for f in os.listdir(out) if f.endswith(".h5")):
        with h5min.TableReader(os.path.join(out, f)) as r:
            rows += r.nrows("Hits")
        size += os.path.getsize(os.path.join(out, f))
    print(f"rows written {rows}, {size / 1e9:.2f} GB")
finally:
    shutil.rmtree(d, ignore_errors=True)
