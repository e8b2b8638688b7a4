"""Shared helpers of the GPU parity tests: run the CPU oracle and hand its recorded variates to CUDA."""
import numpy as np

from oracle import cpu_oracle as co
from pytestbeam_b200 import config as cfg


def setup(name, n):
    return cfg.flatten(getattr(cfg, name)(n))


def oracle_run(name, n, seed_interp=1, seed_numba=2):
    beam, devices, materials, names = setup(name, n)
    table = None
    if n < 500:
        # the reference's accept/reject pool is empty for tiny N (ValueError, reproduced by the
        # oracle); ragged small cases therefore take their Landau table from a larger draw
        rs = np.random.RandomState(seed_interp + 1000)
        table = co.landau_table(rs, beam, devices, materials, 4096)[:, :n]
    return (beam, devices, materials, names), co.run(beam, devices, materials, n, seed_interp, seed_numba,
                                                     table_override=table)


def injected_for(sim, run, first, n):
    from pytestbeam_b200.engine import Injected
    v = run.variates
    return Injected.from_numpy(sim.device, first, n, zstart=v.zstart, gap=v.gap, zkick=v.zkick, eloss=v.eloss,
                               accepted=v.accepted, dig_z=v.dig_z, dig_off=v.dig_off)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0
