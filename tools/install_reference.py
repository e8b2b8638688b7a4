"""Install the unmodified reference (rpartzsch/pytestbeam) into ``baseline/_ref/`` so that it travels to the GPU box.

    python tools/install_reference.py [--source /root/reference]

``baseline/_ref/`` is git-ignored (never part of this repository's history) but not gpurun-ignored, so the benchmark
job on the GPU box can time the reference's own Numba path next to the CUDA path (bench.py ``--impl reference`` and
the ``cpu_baseline`` leg).  Two steps:

1. the contract's ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of source>``.
   The reference's ``setup.py`` calls ``find_packages()`` on a tree without any ``__init__.py``, so this installs
   metadata only (``pytestbeam-0.9.dist-info``) -- the reference is not an importable package, it is run as
   ``cd pytestbeam && python pytestbeam.py`` (its README);
2. therefore the script directory the README tells the user to run from (``pytestbeam/``: the entry script, ``main/*.py``
   and the two YAML files) is placed beside that metadata, byte for byte, which is what an installation of a script-style
   project is.  ``oracle/reference_shim.py`` imports ``main.hit`` / ``main.device`` from there.

Nothing under ``pytestbeam_b200/`` ever imports from ``baseline/``.
"""
from __future__ import annotations

import argparse
import filecmp
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TARGET = os.path.join(ROOT, "baseline", "_ref")
WANTED = (".py", ".yml")


def installed(target: str = TARGET) -> bool:
    return os.path.isfile(os.path.join(target, "pytestbeam", "main", "hit.py"))


def install(source: str = "/root/reference", target: str = TARGET, quiet: bool = True) -> str:
    """Returns a one-line description of what happened."""
    if not os.path.isfile(os.path.join(source, "pytestbeam", "main", "hit.py")):
        return f"reference tree not found at {source}: nothing installed"
    os.makedirs(target, exist_ok=True)
    note = "pip: skipped (metadata already present)"
    if not any(n.endswith(".dist-info") for n in os.listdir(target)):
        with tempfile.TemporaryDirectory() as tmp:          # the source mount is read-only; setup.py writes build/
            copy = os.path.join(tmp, "src")
            shutil.copytree(source, copy)
            cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                   "--find-links", "/opt/wheelhouse", "--target", target, copy]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            note = "pip: ok (metadata only, find_packages() is empty)" if r.returncode == 0 else \
                "pip: failed (" + (r.stdout.strip().splitlines() or ["?"])[-1][:120] + ")"
    n = 0
    for base, _dirs, files in os.walk(os.path.join(source, "pytestbeam")):
        rel = os.path.relpath(base, source)
        if rel.split(os.sep)[-1] == "output_data":
            continue
        for f in files:
            if not f.endswith(WANTED):
                continue
            dst_dir = os.path.join(target, rel)
            os.makedirs(dst_dir, exist_ok=True)
            src, dst = os.path.join(base, f), os.path.join(dst_dir, f)
            if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
                shutil.copyfile(src, dst)
            n += 1
    msg = f"{note}; {n} files of the script directory in {os.path.relpath(target, ROOT)}/pytestbeam"
    if not quiet:
        print(msg)
    return msg


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--source", default="/root/reference")
    a = ap.parse_args()
    print(install(a.source, quiet=True))
