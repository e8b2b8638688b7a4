"""The C-ABI shared library loads without a GPU and exports every symbol include/ptb.h declares."""
import ctypes
import os
import re

from pytestbeam_b200 import build, native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "ptb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ptb_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    build.build()
    lib = ctypes.CDLL(native.LIB_PATH)
    names = _declared()
    assert set(names) == set(native.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n


def test_struct_sizes_match_header_layout():
    assert ctypes.sizeof(native.PtbBeam) == 80
    assert ctypes.sizeof(native.PtbDevice) == 104
    assert ctypes.sizeof(native.PtbInjected) == 64
    assert ctypes.sizeof(native.PtbHitSlab) == 48
    assert ctypes.sizeof(native.PtbBases) == 16 + 8 * 32
    assert ctypes.sizeof(native.PtbTotals) == 8 * 32 + 8 * 5 + 8 + 8 * 32
    assert ctypes.sizeof(native.PtbCompactOut) == 24 + 8 * 33 + 40 + 24
    assert ctypes.sizeof(native.PtbRun) == 32 + 8 + 72 + 48 * 32 + 8 * 5 + 8 * 32 + 24 + 8 * 33 + 40 + 24


def test_struct_layout_agrees_with_a_c_compiler(tmp_path):
    """sizeof / offsetof of every struct as gcc sees include/ptb.h (plain C) against the ctypes mirror."""
    import subprocess
    structs = {"PtbBeam": native.PtbBeam, "PtbDevice": native.PtbDevice, "PtbOptions": native.PtbOptions,
               "PtbInjected": native.PtbInjected, "PtbTrackTable": native.PtbTrackTable,
               "PtbHitSlab": native.PtbHitSlab, "PtbCompactOut": native.PtbCompactOut, "PtbPackedLayout": native.PtbPackedLayout, "PtbBases": native.PtbBases, "PtbTotals": native.PtbTotals,
               "PtbRun": native.PtbRun}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ptb.h"', 'int main(void) {']
    for name, st in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for field, _ in st._fields_:
            lines.append(f'  printf("{name}.{field} %zu\\n", offsetof({name}, {field}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    seen = dict(ln.split() for ln in subprocess.check_output([str(exe)], text=True).splitlines())
    for name, st in structs.items():
        assert int(seen[name]) == ctypes.sizeof(st), name
        for field, _ in st._fields_:
            assert int(seen[f"{name}.{field}"]) == getattr(st, field).offset, f"{name}.{field}"


def test_launch_counter_starts_at_zero_without_gpu_work():
    lib = native.load()
    assert lib.ptb_launch_count() >= 0


def test_plain_c_example_compiles_against_the_header():
    """examples/c_abi_example.c is C (not C++): the header must be usable from a C compiler."""
    import shutil, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda_inc = "/usr/local/cuda/include"
    if shutil.which("gcc") is None or not os.path.isdir(cuda_inc):
        import pytest
        pytest.skip("gcc or the CUDA headers are not installed")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(root, "include"),
                           "-I", cuda_inc, os.path.join(root, "examples", "c_abi_example.c")])
