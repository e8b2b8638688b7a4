"""Drop-in for the hot-path half of the reference's ``main/hit.py``: same names, arguments and return shapes.

``tracks`` (reference ``main/hit.py:8-138``) returns the 8-tuple ``(energy, anglex, angley, x, y, z, time_stamp,
energy_lost)``: seven flat particle-major sequences of length N*D and the ``[D, N]`` energy-loss table.  Here it is a
``HitTables`` object that behaves like that tuple but computes the arrays on the GPU only when they are touched;
handed untouched to ``device.calculate_device_hit`` it lets the digitisation run fused with the track
generation, so the N*D track records never leave the device.
"""
from __future__ import annotations

import os

import numpy as np

from .. import h5min
from ..engine import Simulator

TRACKS_DTYPE = np.dtype([("energy", float), ("x_angle", float), ("y_angle", float), ("x", float), ("y", float),
                         ("z", float), ("time_stamp", np.uint16)])      # main/hit.py:441-451
CHUNK = 4_000_000


def _seed_of(beam: dict) -> int:
    """Optional ``seed`` key of the beam block; the reference is unseeded, so the default is drawn fresh."""
    s = beam.get("seed")
    return int.from_bytes(os.urandom(8), "little") if s is None else int(s)


class HitTables:
    """The reference's ``hit_tables`` 8-tuple, materialised lazily from the GPU."""

    _NAMES = ("energy", "angle_x", "angle_y", "x", "y", "z", "time_stamp", "eloss")

    def __init__(self, sim: Simulator, n: int, seed: int, log=None):
        self.sim, self.n, self.seed, self.log = sim, int(n), int(seed), log
        self._arrays = None

    @property
    def materialised(self) -> bool:
        return self._arrays is not None

    def chunks(self, n: int | None = None, chunk: int = CHUNK):
        """The track records of the first n particles (default: all), chunk by chunk: yields (first particle, count,
        dict of numpy arrays in the layout of the tuple) without ever holding more than one chunk."""
        D, n = self.sim.D, self.n if n is None else min(int(n), self.n)
        bases = None
        for lo in range(0, n, chunk):
            m = min(chunk, n - lo)
            r = self.sim.run(lo, m, seed=self.seed, digitise=False, write_tracks=True, bases_in=bases)
            bases = r.bases_out
            piece = {k: r.tracks[k].cpu().numpy() for k in self._NAMES[:7]}
            piece["eloss"] = r.tracks["eloss"].cpu().numpy().reshape(D, m)
            yield lo, m, piece

    def head(self, n_events: int) -> tuple:
        """The 8-tuple restricted to the first n_events particles -- all that the reference's plot_default reads
        (main/plotting.py:41-76 slices the first <= 1e5 events) -- without materialising the rest."""
        D, n = self.sim.D, min(int(n_events), self.n)
        out = {k: np.empty(n * D, dtype=np.float64) for k in self._NAMES[:6]}
        out["time_stamp"] = np.empty(n * D, dtype=np.int64)
        out["eloss"] = np.empty((D, n), dtype=np.float64)
        for lo, m, piece in self.chunks(n):
            for k in self._NAMES[:7]:
                out[k][lo * D:(lo + m) * D] = piece[k]
            out["eloss"][:, lo:lo + m] = piece["eloss"]
        return tuple(out[k] for k in self._NAMES)

    def _materialise(self):
        if self._arrays is None:
            self._arrays = self.head(self.n)

    def __len__(self):
        return 8

    def __getitem__(self, i):
        self._materialise()
        return self._arrays[i]

    def __iter__(self):
        self._materialise()
        return iter(self._arrays)


def tracks(beam: dict, devices: list, materials: list, log) -> HitTables:
    """Create the telescope set-up and the particle tracks (reference ``main/hit.py:8``)."""
    log.info("Creating telescope setup")
    sim = Simulator(beam, devices, materials)
    seed = _seed_of(beam)
    log.info("Generating particle tracks")
    return HitTables(sim, beam["nmb_particles"], seed, log)


def create_output_tracks(hit_table, folder: str) -> None:
    """Write ``<folder>particle_tracks.h5`` with table ``/Tracks`` (reference ``main/hit.py:434-468``)."""
    fields = (("energy", "energy"), ("x_angle", "angle_x"), ("y_angle", "angle_y"), ("x", "x"), ("y", "y"), ("z", "z"))
    if isinstance(hit_table, HitTables) and not hit_table.materialised:
        # straight from the GPU into the file, chunk by chunk: the N*D records are never held in host memory
        with h5min.TableWriter(folder + "particle_tracks.h5", "Tracks", TRACKS_DTYPE) as w:
            for _lo, m, piece in hit_table.chunks():
                out = w.reserve(m * hit_table.sim.D)
                for field, key in fields:
                    out[field] = piece[key]
                out["time_stamp"] = piece["time_stamp"].astype(np.uint16)     # wraps, as in the reference (Q19)
        return
    numb = len(hit_table[0])
    out = np.zeros(numb, dtype=TRACKS_DTYPE)
    for i, (field, _key) in enumerate(fields):
        out[field] = np.asarray(hit_table[i])
    out["time_stamp"] = np.asarray(hit_table[6]).astype(np.uint16)     # wraps, as in the reference (Q19)
    h5min.write_table(folder + "particle_tracks.h5", "Tracks", out)
