"""The C ABI from plain C (examples/c_abi_example.c: CUDA runtime allocations, no Python, no torch in the process that
does the work) must produce the very tables the Python host produces for the same seed."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from pytestbeam_b200 import config as cfg

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plain_c_caller_matches_the_python_host(tmp_path):
    from pytestbeam_b200 import build
    from pytestbeam_b200.engine import Simulator
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    lib_dir = os.path.dirname(build.build())
    exe = str(tmp_path / "c_abi_example")
    subprocess.check_call([nvcc, "-x", "c", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "examples", "c_abi_example.c"), "-L", lib_dir, "-lptb_b200", "-lcudart"])
    n, seed = 200_000, 7
    env = dict(os.environ, LD_LIBRARY_PATH=lib_dir + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    rows_file = str(tmp_path / "itkpix_rows.bin")
    out = subprocess.run([exe, str(n), str(seed), rows_file], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    rows = [int(m) for m in re.findall(r"plane \d+ rows (\d+)", out.stdout)]
    beam, devices, materials, names = cfg.flatten(cfg.default_setup(n))
    res = Simulator(beam, devices, materials).run(0, n, seed=seed)
    assert rows == [int(v) for v in res.totals["hits"]]
    assert int(re.search(r"accepted (\d+)", out.stdout).group(1)) == res.totals["accepted"]
    # the first itkpix records, field by field
    want = res.hits_numpy(len(devices) - 1)[:5]
    got = re.findall(r"itkpix row \d+: event (-?\d+) frame (\d+) column (\d+) row (\d+) charge ([-0-9.e+]+)", out.stdout)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert (int(g[0]), int(g[1]), int(g[2]), int(g[3])) == (w["event_number"], w["frame"], w["column"], w["row"])
        assert abs(float(g[4]) - float(w["charge"])) < 1e-5
    # the narrow compressed rows, expanded by the library straight into a file: the last plane's records, byte for byte
    assert int(re.search(r"file rows (\d+)", out.stdout).group(1)) == res.totals["hits"][-1]
    assert open(rows_file, "rb").read() == res.hits_numpy(len(devices) - 1).tobytes()
    # and the packed ones (ptb_packed_layout, ptb_expand_hits_packed_fd)
    os.remove(rows_file)
    out = subprocess.run([exe, str(n), str(seed), rows_file, "packed"], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "packed rows" in out.stdout
    assert open(rows_file, "rb").read() == res.hits_numpy(len(devices) - 1).tobytes()
