"""Set-up and material descriptions with the schema of the reference's ``setup.yml`` / ``material.yml``.

The reference reads both files from the working directory (``pytestbeam/pytestbeam.py:13-17``); the
schema is summarised in SURVEY.md Appendix E.  The dictionaries built here have exactly that shape
(``devices`` mapping in beam order, ``beam``, ``data_output``, ``plotting``, ``saving_tracks``), so they
can be dumped to YAML for the reference or passed to this package's front end unchanged.

``c1`` .. ``c5`` are the five configurations named in BASELINE.json.
"""
from __future__ import annotations

import copy

SILICON = {"rad_length": 93650, "Z": 14, "A": 24, "rho": 2.33}
LEAD = {"rad_length": 5612, "Z": 82, "A": 207, "rho": 11.35}


def default_material() -> dict:
    """The two materials shipped by the reference (``pytestbeam/material.yml:1-11``)."""
    return {"silicon": dict(SILICON), "lead": dict(LEAD)}


def _mimosa26(z, delta_x=0, delta_y=0, trigger="untriggered"):
    return {
        "column": 1152, "row": 576, "column_pitch": 18.4, "row_pitch": 18.4, "thickness": 50,
        "z_position": z, "delta_x": delta_x, "delta_y": delta_y, "threshold": 100,
        "trigger": trigger, "material": "silicon",
    }


def _itkpix(z, delta_x=0, delta_y=0, trigger="triggered"):
    return {
        "column": 400, "row": 384, "column_pitch": 50, "row_pitch": 50, "thickness": 300,
        "z_position": z, "delta_x": delta_x, "delta_y": delta_y, "threshold": 1000,
        "trigger": trigger, "material": "silicon",
    }


def _beam(n, energy=5000.0, rate=5000):
    return {
        "energy": energy, "sigma_x": 5000, "sigma_y": 5000, "loc_x": 0, "loc_y": 0,
        "x_disp": 0.0001, "y_disp": 0.0001, "x_angle": 0, "y_angle": 0,
        "nmb_particles": int(n), "particle_rate": rate,
    }


def default_setup(nmb_particles: int = 2_000_000) -> dict:
    """EUDET telescope (6 MIMOSA26) + ITkPix reference plane, as in ``pytestbeam/setup.yml:2-109``."""
    z = [0, 30250, 61800, 202550, 225000, 250700]
    dx = [100, 0, 0, 100, 0, 0]
    devices = {f"mimosa26_p{i + 1}": _mimosa26(z[i], dx[i]) for i in range(6)}
    devices["itkpix"] = _itkpix(288150)
    return {
        "devices": devices,
        "beam": _beam(nmb_particles),
        "data_output": None,
        "plotting": True,
        "saving_tracks": False,
    }


def c1(n: int = 100_000) -> dict:
    """BASELINE config 1: default set-up, 1e5 electrons."""
    return default_setup(n)


def c2(n: int = 100_000_000) -> dict:
    """BASELINE config 2: default telescope, 1e8 electrons on one GPU."""
    return default_setup(n)


def c3(n: int = 100_000) -> dict:
    """BASELINE config 3: telescope + DUT + single-pixel lead absorber, 100 MeV."""
    s = default_setup(n)
    z = [0, 30250, 61800, 202550, 225000, 250700]
    dx = [100, 0, 0, 100, 0, 0]
    dev = {}
    for i in range(3):
        dev[f"mimosa26_p{i + 1}"] = _mimosa26(z[i], dx[i])
    dev["absorber"] = {
        "column": 1, "row": 1, "column_pitch": 40000, "row_pitch": 40000, "thickness": 1000,
        "z_position": 100000, "delta_x": 0, "delta_y": 0, "threshold": 1e9,
        "trigger": "untriggered", "material": "lead",
    }
    dev["dut"] = {
        "column": 512, "row": 512, "column_pitch": 33.04, "row_pitch": 33.04, "thickness": 100,
        "z_position": 130000, "delta_x": 0, "delta_y": 0, "threshold": 200,
        "trigger": "triggered", "material": "silicon",
    }
    for i in range(3, 6):
        dev[f"mimosa26_p{i + 1}"] = _mimosa26(z[i], dx[i])
    dev["itkpix"] = _itkpix(288150)
    s["devices"] = dev
    s["beam"] = _beam(n, energy=100.0)
    return s


def c4(n: int = 100_000) -> dict:
    """BASELINE config 4: off-centre planes, mixed trigger modes, 2e8 Hz."""
    s = default_setup(n)
    shifts = [(100, 0), (-2500, 100), (0, 2500), (6000, -100), (-100, -6000), (2500, 2500), (0, 0)]
    triggered = {"mimosa26_p2", "mimosa26_p4", "mimosa26_p6", "itkpix"}
    for (name, d), (sx, sy) in zip(s["devices"].items(), shifts):
        d["delta_x"], d["delta_y"] = sx, sy
        d["trigger"] = "triggered" if name in triggered else "untriggered"
    s["beam"] = _beam(n, rate=2e8)
    return s


def c5(n: int = 10_000_000_000) -> dict:
    """BASELINE config 5: default telescope, 1e10 electrons sharded over the GPUs of one box."""
    return default_setup(n)


def flatten(setup: dict, material: dict | None = None):
    """Return ``(beam, devices, materials, names)`` the way ``pytestbeam/pytestbeam.py:23-28`` builds them."""
    material = default_material() if material is None else material
    names = list(setup["devices"])
    devices = [copy.deepcopy(setup["devices"][k]) for k in names]
    materials = [dict(material[d["material"]]) for d in devices]
    return dict(setup["beam"]), devices, materials, names
