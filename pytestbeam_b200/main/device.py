"""Drop-in for the hot-path half of the reference's ``main/device.py``: same names, arguments, side effects.

``calculate_device_hit`` (reference ``main/device.py:7-88``) writes one ``<folder><name>_dut.h5`` per device with
table ``/Hits`` (event_number <i8, frame <u2, column <u2, row <u2, charge <f4; rows in particle order).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import h5min
from .. import native as N
from ..engine import HITS_DTYPE, Simulator
from ..pipeline import ChunkPipeline
from .hit import CHUNK, HitTables, _seed_of


def create_hit_file(hit_data: np.ndarray, folder: str, index: str) -> str:
    """Write ``<folder><index>_dut.h5`` (reference ``main/device.py:306-340``); returns the path."""
    hit_data = np.asarray(hit_data, dtype=HITS_DTYPE)
    path = folder + index + "_dut.h5"
    h5min.write_table(path, "Hits", hit_data)
    return path


def _packed_rows(sim: Simulator, slab: dict) -> np.ndarray:
    packed = sim.pack_hits(slab)
    return np.frombuffer(packed.cpu().numpy().tobytes(), dtype=HITS_DTYPE)


def device_hits(beam: dict, devices: list, hit_data, log=None) -> list:
    """The hit table of every device (structured arrays), i.e. calculate_device_hit without the file writing."""
    n = int(beam["nmb_particles"])
    D = len(devices)
    parts = [[] for _ in range(D)]
    if isinstance(hit_data, HitTables) and not hit_data.materialised:
        # fused: tracks + digitisation in one pass per chunk, nothing but hit rows leaves the GPU
        sim, seed = hit_data.sim, hit_data.seed
        bases = None
        for lo in range(0, n, CHUNK):
            m = min(CHUNK, n - lo)
            r = sim.run(lo, m, seed=seed, bases_in=bases, capacities=sim.default_capacities(m),
                        scratch_rows=sim.default_scratch_rows(m))
            bases = r.bases_out
            for d in range(D):
                parts[d].append(_packed_rows(sim, r.slabs[d]))
    else:
        # calc_position alone, fed from the caller's tuple exactly as main/device.py:132-138 reads it
        if isinstance(hit_data, HitTables):
            sim, seed = hit_data.sim, hit_data.seed
        else:
            raise TypeError("hit_data must come from pytestbeam_b200.main.hit.tracks (use device_hits_from_arrays "
                            "for foreign tables)")
        x, y, eloss = np.asarray(hit_data[3]), np.asarray(hit_data[4]), np.asarray(hit_data[7])
        parts = device_hits_from_arrays(sim, x, y, eloss, seed)
        return parts
    return [np.concatenate(p) if p else np.zeros(0, dtype=HITS_DTYPE) for p in parts]


def device_hits_from_arrays(sim: Simulator, x, y, eloss, seed: int, accepted=None) -> list:
    """Digitise given track positions: x, y flat particle-major [N*D], eloss [D, N] (after the acceptance zeroing)."""
    D = sim.D
    n = len(x) // D
    parts = [[] for _ in range(D)]
    bases = None
    dev = sim.device
    for lo in range(0, n, CHUNK):
        m = min(CHUNK, n - lo)
        tin = {"x": torch.from_numpy(np.ascontiguousarray(x[lo * D:(lo + m) * D], dtype=np.float64)).to(dev),
               "y": torch.from_numpy(np.ascontiguousarray(y[lo * D:(lo + m) * D], dtype=np.float64)).to(dev),
               "eloss": torch.from_numpy(np.ascontiguousarray(eloss[:, lo:lo + m], dtype=np.float64)).to(dev)}
        if accepted is not None:
            tin["accepted"] = torch.from_numpy(np.ascontiguousarray(accepted[lo:lo + m], dtype=np.uint8)).to(dev)
        r = sim.run(lo, m, seed=seed, tracks_in=tin, bases_in=bases, capacities=sim.default_capacities(m),
                        scratch_rows=sim.default_scratch_rows(m))
        bases = r.bases_out
        for d in range(D):
            parts[d].append(_packed_rows(sim, r.slabs[d]))
    return [np.concatenate(p) if p else np.zeros(0, dtype=HITS_DTYPE) for p in parts]


def calculate_device_hit(beam: dict, devices: list, hit_data, names: list, folder: str, log) -> None:
    """Create the HDF5 hit tables from the particle table and the device configuration (reference ``main/device.py:7``)."""
    log.info("Calculating particle hit positions")
    tables = device_hits(beam, devices, hit_data, log)
    for dut, name in enumerate(names):
        log.info("Hit positions of %s" % name)
        log.info("Saving Data...")
        create_hit_file(tables[dut], folder, name)
