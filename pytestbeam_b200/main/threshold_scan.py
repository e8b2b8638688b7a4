"""Drop-in for the computational half of the reference's ``main/threshold_scan.py`` (S-curve of one device).

``create_hits`` (reference ``threshold_scan.py:42-77``) injects every charge of ``inj_array`` ``numb_inj`` times into
pixel (column, row) by calling ``calc_cluster_hits`` with both cluster radii 0: only the seed-pixel fallback
(``main/device.py:299-302``) can fire, i.e. a hit is recorded when (injected charge + noise) reaches the threshold.
Here all bins x injections trials are digitised in one pass of ``k_digitise`` (``PTB_TRACKS_FROM_TABLE`` +
``PTB_ZERO_RADIUS``) on a one-plane set-up.  Plotting stays with the reference.
"""
from __future__ import annotations

import numpy as np
import torch

from ..config import SILICON
from ..engine import Simulator
from .hit import _seed_of


def create_injection_array(low, high, bins):
    return np.arange(low, high, (high - low) / bins)


def scan_setup(column_pitch, row_pitch, threshold, thickness, column_numb, row_numb, column=0, row=0):
    """(beam, devices, materials) of the one-plane set-up the scan digitises on (reference call at :55-69)."""
    dev = {"column": int(column_numb), "row": int(row_numb), "column_pitch": column_pitch, "row_pitch": row_pitch,
           "thickness": thickness, "z_position": 0, "delta_x": column * column_pitch, "delta_y": row * row_pitch,
           "threshold": threshold, "trigger": "untriggered", "material": "silicon"}
    beam = {"energy": 5000.0, "sigma_x": 1, "sigma_y": 1, "loc_x": 0, "loc_y": 0, "x_disp": 0, "y_disp": 0,
            "x_angle": 0, "y_angle": 0, "nmb_particles": 1, "particle_rate": 5000}
    return beam, [dev], [dict(SILICON)]


def create_hits(inj_array, numb_inj, column_pitch, row_pitch, threshold, thickness, column_numb, row_numb, column=0,
                row=0, seed=None, injected_z=None):
    """Hit probability per injected charge (reference ``threshold_scan.py:42-77``).

    ``injected_z`` (tests): standard normals [bins, numb_inj, 2] in the reference's draw order (seed-pixel noise,
    box-pixel noise) for the deterministic mode; otherwise Philox with ``seed``.
    """
    inj_array = np.asarray(inj_array, dtype=np.float64)
    bins, numb_inj = len(inj_array), int(numb_inj)
    n = bins * numb_inj
    beam, devices, materials = scan_setup(column_pitch, row_pitch, threshold, thickness, column_numb, row_numb, column,
                                          row)
    sim = Simulator(beam, devices, materials)
    energy = np.repeat(inj_array * 1e-6 * 3.6, numb_inj)                 # threshold_scan.py:56
    dev = sim.device
    tin = {"x": torch.zeros(n, dtype=torch.float64, device=dev), "y": torch.zeros(n, dtype=torch.float64, device=dev),
           "eloss": torch.from_numpy(energy.reshape(1, n)).to(dev),
           "accepted": torch.ones(n, dtype=torch.uint8, device=dev)}
    inj = None
    if injected_z is not None:
        from ..engine import Injected
        z = np.zeros((n, 4))
        z[:, 2:] = np.asarray(injected_z, dtype=np.float64).reshape(n, 2)   # radius draws do not exist in the scan
        off = np.arange(n + 1, dtype=np.int64) * 4
        inj = Injected.from_numpy(dev, 0, n, dig_z=[z.reshape(-1)], dig_off=[off])
    res = sim.run(0, n, seed=_seed_of({"seed": seed}), injected=inj, tracks_in=tin, zero_radius=True,
                  capacities=[n + 64])
    events = res.slabs[0]["event"].cpu().numpy()                          # one row per trial that fired
    hits = np.bincount((events - 1) // numb_inj, minlength=bins)[:bins]
    return hits / numb_inj


def threshold_scan(dut: dict, low, high, bins, inj_numb, name=None, seed=None):
    """(injected charges, hit probabilities) of one device (reference ``threshold_scan.py:103-121`` minus the PDF)."""
    inj_array = create_injection_array(low, high, bins)
    hits = create_hits(inj_array, inj_numb, dut["column_pitch"], dut["row_pitch"], dut["threshold"], dut["thickness"],
                       dut["column"], dut["row"], column=0, row=0, seed=seed)
    return inj_array, hits
