"""Command-line front end with the contract of the reference's ``pytestbeam/pytestbeam.py``: it is started in a
directory holding ``setup.yml`` and ``material.yml`` and leaves ``<data_output><device>_dut.h5`` behind (plus
``particle_tracks.h5`` when ``saving_tracks`` is set).

    cd <dir with setup.yml, material.yml> && python -m pytestbeam_b200.pytestbeam

The flow is the reference's (``pytestbeam.py:9-37``): flatten the two YAML files, ``hit.tracks``, optionally
``hit.create_output_tracks``, ``device.calculate_device_hit``; only the two modules come from this package.
"""
from __future__ import annotations

import logging
import os
from dataclasses import dataclass

import yaml

from .main import device, hit


@dataclass
class Job:
    """What the reference's script keeps in module-level variables."""
    beam: dict
    devices: list       # beam order = mapping order of setup.yml
    materials: list     # material.yml entry of every device
    names: list
    folder: str         # string prefix of every output file, "output_data/" when data_output is empty
    saving_tracks: bool
    plotting: bool


def load_job(workdir: str = ".") -> Job:
    def read(name):
        with open(os.path.join(workdir, name), "r") as fh:
            return yaml.full_load(fh)

    setup, catalogue = read("setup.yml"), read("material.yml")
    names = list(setup["devices"])
    devices = [setup["devices"][n] for n in names]
    prefix = setup.get("data_output") or "output_data/"
    if not os.path.isabs(prefix):
        prefix = os.path.join(workdir, prefix)
    return Job(setup["beam"], devices, [catalogue[d["material"]] for d in devices], names, prefix,
               bool(setup.get("saving_tracks")), bool(setup.get("plotting")))


def run_job(job: Job, log) -> None:
    if job.folder.endswith("/"):
        os.makedirs(job.folder, exist_ok=True)
    tables = hit.tracks(job.beam, job.devices, job.materials, log)
    if job.saving_tracks:
        log.info("Saving tracks")
        hit.create_output_tracks(tables, job.folder)
    device.calculate_device_hit(job.beam, job.devices, tables, job.names, job.folder, log)
    if job.plotting:
        log.info("plots are not produced here: hand the *_dut.h5 files and the hit_tables tuple to the reference's "
                 "main.plotting.plot_default")


def main(workdir: str = ".") -> None:
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s %(message)s")
    log = logging.getLogger("pytestbeam")
    log.info("Preparing simulation")
    run_job(load_job(workdir), log)


if __name__ == "__main__":
    main()
