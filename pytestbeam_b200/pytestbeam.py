"""Entry point with the contract of the reference's ``pytestbeam/pytestbeam.py``: run from a directory that holds
``setup.yml`` and ``material.yml``; writes ``<data_output><device>_dut.h5`` (and ``particle_tracks.h5`` when
``saving_tracks`` is set).

    cd <dir with setup.yml, material.yml> && python -m pytestbeam_b200.pytestbeam
"""
from __future__ import annotations

import logging
import os

import yaml

from .main import device, hit


def main(workdir: str = ".") -> None:
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s %(message)s")
    log = logging.getLogger("pytestbeam")
    log.info("Preparing simulation")
    with open(os.path.join(workdir, "setup.yml"), "r") as file:
        setup = yaml.full_load(file)
    with open(os.path.join(workdir, "material.yml"), "r") as file:
        material = yaml.full_load(file)

    folder = setup["data_output"]
    if folder is None:
        folder = "output_data/"
    folder = os.path.join(workdir, folder) if not os.path.isabs(folder) else folder
    if folder.endswith("/"):
        os.makedirs(folder, exist_ok=True)

    names = [dev for dev in setup["devices"]]
    devices = [setup["devices"][dev] for dev in names]
    materials = [material[d["material"]] for d in devices]
    beam = setup["beam"]

    hit_tables = hit.tracks(beam, devices, materials, log)
    if setup.get("saving_tracks"):
        log.info("Saving tracks")
        hit.create_output_tracks(hit_tables, folder)
    device.calculate_device_hit(beam, devices, hit_tables, names, folder, log)
    if setup.get("plotting"):
        log.info("plotting is outside this package: feed the written *_dut.h5 files and the hit_tables tuple to the "
                 "reference's main.plotting.plot_default")


if __name__ == "__main__":
    main()
