"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference (rpartzsch/pytestbeam).

Loads ``main.hit`` / ``main.device`` from ``baseline/_ref/pytestbeam`` (the git-ignored installation made by
``tools/install_reference.py``, which travels to the GPU box) or, failing that, from ``/root/reference/pytestbeam``
(the read-only mount that exists in the build container only), after stubbing the two third-party imports that
are absent on this image (``numba_progress`` and ``tables``).  Nothing on the product path
imports this module; it is used by ``oracle/make_golden.py`` (fixture generation, run in the build
container), by the ``-m "not gpu"`` tests that pin the oracle when a reference tree is present, and by
``bench.py``'s two CPU legs (``--impl reference`` and ``cpu_baseline``), which time the reference's own Numba path.

Reference call sites this drives: ``main/hit.py:8`` (``tracks``), ``main/device.py:7``
(``calculate_device_hit``).
"""
from __future__ import annotations

import logging
import os
import sys
import types

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CANDIDATES = (os.path.join(_ROOT, "baseline", "_ref"), "/root/reference")

_loaded = None


def reference_root():
    for root in REF_CANDIDATES:
        if os.path.isfile(os.path.join(root, "pytestbeam", "main", "hit.py")):
            return root
    return None


def available() -> bool:
    if reference_root() is None:
        return False
    try:
        import numba  # noqa: F401
    except Exception:
        return False
    return True


def load():
    """Return (hit, device, capture, seed_numba) for the untouched reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not present (baseline/_ref is made by tools/install_reference.py)")
    from numba import int64, njit
    from numba.experimental import jitclass

    @jitclass([("n", int64)])
    class _Proxy:
        def __init__(self):
            self.n = 0

        def update(self, k):
            self.n += k

    class ProgressBar:
        def __init__(self, total=None, **kw):
            self.p = _Proxy()

        def __enter__(self):
            return self.p

        def __exit__(self, *a):
            return False

    m = types.ModuleType("numba_progress")
    m.ProgressBar = ProgressBar
    sys.modules["numba_progress"] = m
    t = types.ModuleType("tables")
    t.table = object
    sys.modules["tables"] = t
    pkg = os.path.join(root, "pytestbeam")
    # the reference uses the top-level package name ``main``; keep it isolated from our own modules
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "main" or k.startswith("main.")}
    sys.path.insert(0, pkg)
    try:
        import main.device as device
        import main.hit as hit
    finally:
        sys.path.remove(pkg)
        for k in list(sys.modules):
            if k == "main" or k.startswith("main."):
                sys.modules["_ptb_ref_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)

    capture = {}

    def _capture_hit_file(table, folder, name):
        capture[name] = np.array(table)
        return folder + name + "_dut.h5"

    device.create_hit_file = _capture_hit_file

    @njit
    def seed_numba(s):
        np.random.seed(s)

    _loaded = (hit, device, capture, seed_numba)
    return _loaded


def run_reference(beam: dict, devices: list, materials: list, names: list, seed_interp: int, seed_numba: int):
    """Run the unmodified reference hot path under both seeds; return (hit_tables as ndarrays, hit tables per device)."""
    hit, device, capture, nseed = load()
    log = logging.getLogger("ptb_reference")
    capture.clear()
    np.random.seed(seed_interp)
    nseed(seed_numba)
    ht = hit.tracks(beam, devices, materials, log)
    tracks = tuple(np.asarray(list(a) if not isinstance(a, np.ndarray) else a) for a in ht[:7])
    energy_lost = np.array(ht[7])
    device.calculate_device_hit(beam, devices, ht, names, "", log)
    hits = {k: v.copy() for k, v in capture.items()}
    return tracks, energy_lost, hits


# ---------------------------------------------------------------------------------------------------------------
# timing of the reference's own hot path (bench.py's CPU legs)
# ---------------------------------------------------------------------------------------------------------------

def timed_run(beam: dict, devices: list, materials: list, names: list, seed: int):
    """One pass of the unmodified ``hit.tracks`` (main/hit.py:8) + ``device.calculate_device_hit`` (main/device.py:7)
    with the HDF5 writer replaced by an in-memory capture.  Returns (seconds, hit rows)."""
    import time
    hit, device, capture, nseed = load()
    log = logging.getLogger("ptb_reference")
    capture.clear()
    np.random.seed(seed & 0x7FFFFFFF)
    nseed((seed + 1) & 0x7FFFFFFF)
    t = time.perf_counter()
    ht = hit.tracks(beam, devices, materials, log)
    device.calculate_device_hit(beam, devices, ht, names, "", log)
    dt = time.perf_counter() - t
    rows = sum(len(v) for v in capture.values())
    capture.clear()
    return dt, rows


def warm_jit(setup_fn, flatten):
    """Compile the reference's two Numba loops with a throw-away N=2000 pass."""
    beam, devices, materials, names = flatten(setup_fn(2000))
    return timed_run(beam, devices, materials, names, 12345)[0]


def _pool_worker(args):
    setup_fn, flatten, n, seed = args
    beam, devices, materials, names = flatten(setup_fn(n))
    return timed_run(beam, devices, materials, names, seed)


class ReferencePool:
    """``workers`` forked processes that share the parent's warm JIT (the reference has no threading of its own:
    independent processes with distinct seeds are the fairest all-core figure)."""

    def __init__(self, workers: int):
        import multiprocessing as mp
        self.workers = int(workers)
        self.pool = mp.get_context("fork").Pool(self.workers) if self.workers > 1 else None

    def rate(self, setup_fn, flatten, n_each: int, seed: int):
        """(electrons/s, hit rows/s) of one round: every worker simulates n_each electrons."""
        import time
        jobs = [(setup_fn, flatten, n_each, seed + 7919 * i) for i in range(self.workers)]
        t = time.perf_counter()
        res = self.pool.map(_pool_worker, jobs, chunksize=1) if self.pool else [_pool_worker(jobs[0])]
        wall = time.perf_counter() - t
        return n_each * self.workers / wall, sum(r for _, r in res) / wall

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None
