"""The product never touches the oracle or the reference: nothing under pytestbeam_b200/ imports them."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_oracle_or_reference_in_the_product():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "pytestbeam_b200")):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            text = open(os.path.join(base, f)).read()
            if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "/root/reference" in text \
                    or "cpu_oracle" in text or "libptb_oracle" in text:
                bad.append(os.path.join(base, f))
    assert not bad, bad


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import importlib
    import pytest
    from pytestbeam_b200 import native
    monkeypatch.setattr(native, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(native, "_lib", None)
    with pytest.raises(ImportError):
        native.load()
    importlib.reload(native)
