"""Developer tool: build libptb_b200 variants that differ in -D macros (name=flags ...) into _lib/var_<name>.so.

    python tools/build_variants.py base= d8=-DPTB_DIGI_MIN_BLOCKS=8
"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytestbeam_b200 import build as b

procs = []
for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    out = os.path.join(b.OUT_DIR, f"var_{name}.so")
    cmd = [b.nvcc()] + b.NVCC_FLAGS + flags.split() + ["-Xptxas=-v", "-I", os.path.join(b.ROOT, "include"), "-o", out, b.SRC]
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, p in procs:
    txt = p.communicate()[0]
    lines = txt.splitlines()
    for i, ln in enumerate(lines):
        if "Compiling entry" in ln and ("PhiloxSourceELi128ELi128ELb0" in ln or "k_tracksINS_12Philox" in ln or "k_gather" in ln):
            print(name, ln.split("'")[1][:40], "|", lines[i + 2].strip()[14:60], "|", lines[i + 3].strip()[14:40] if i + 3 < len(lines) else "")
    if p.returncode:
        print(txt[-3000:])
        raise SystemExit(f"{name}: nvcc failed")
