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
    # optional keys beyond the reference's schema; their defaults are the only values there are
    if str(setup.get("backend", "cuda")).lower() != "cuda":
        raise ValueError("backend: this package runs on CUDA (sm_100a) only; there is no CPU path")
    if str(setup.get("precision", "f64")).lower() not in ("f64", "float64", "double"):
        raise ValueError("precision: the physics arithmetic is fp64 (the reference's); no other precision is built")
    names = list(setup["devices"])
    devices = [setup["devices"][n] for n in names]
    prefix = setup.get("data_output") or "output_data/"
    if not os.path.isabs(prefix):
        prefix = os.path.join(workdir, prefix)
    return Job(setup["beam"], devices, [catalogue[d["material"]] for d in devices], names, prefix,
               bool(setup.get("saving_tracks")), bool(setup.get("plotting")))


def write_run_record(job: Job, seconds: float, log) -> str:
    """``<folder>run_record.json``: what was simulated and how fast (no counterpart in the reference, which logs
    milestones only)."""
    import json
    last = dict(device.LAST_RUN)
    hits = last.get("hits") or []
    n = int(job.beam["nmb_particles"])
    rec = {"electrons": n, "seconds": seconds, "electrons_per_s": n / seconds if seconds > 0 else None,
           "hits_per_s": sum(hits) / seconds if hits and seconds > 0 else None, "seed": last.get("seed"),
           "route": last.get("route"), "chunk": last.get("chunk"), "retried_chunks": last.get("retries"),
           "device_to_host_bytes": last.get("d2h_bytes"), "rows": dict(zip(job.names, hits)),
           "files": [f"{name}_dut.h5" for name in job.names] + (["particle_tracks.h5"] if job.saving_tracks else [])}
    path = job.folder + "run_record.json"
    with open(path, "w") as fh:
        json.dump(rec, fh, indent=1)
    log.info("%d electrons in %.2f s (%.3g electrons/s); record in %s" % (n, seconds, n / max(seconds, 1e-9), path))
    return path


def run_job(job: Job, log) -> None:
    import time
    if job.folder.endswith("/"):
        os.makedirs(job.folder, exist_ok=True)
    t0 = time.perf_counter()
    tables = hit.tracks(job.beam, job.devices, job.materials, log)
    if job.saving_tracks:
        log.info("Saving tracks")
        hit.create_output_tracks(tables, job.folder)
    device.calculate_device_hit(job.beam, job.devices, tables, job.names, job.folder, log)
    write_run_record(job, time.perf_counter() - t0, log)
    if job.plotting:
        log.info("plots are not produced here: hand the *_dut.h5 files and the hit_tables tuple to the reference's "
                 "main.plotting.plot_default")


def run_job_sharded(job: Job, log) -> None:
    """One process per GPU (torchrun): every rank simulates its index range, rank 0 merges the part files."""
    import torch
    import torch.distributed as dist

    from . import distributed
    from .main.hit import _seed_of
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        if job.folder.endswith("/"):
            os.makedirs(job.folder, exist_ok=True)
        seed = torch.tensor([_seed_of(job.beam) & (2**63 - 1)], dtype=torch.int64, device=torch.device("cuda", local))
        dist.broadcast(seed, 0)                       # an unseeded run still needs one seed for all ranks
        if job.saving_tracks and dist.get_rank() == 0:
            log.info("saving_tracks is ignored in sharded runs (the track table is a single-process feature)")
        info = distributed.run_sharded(job.beam, job.devices, job.materials, job.names, job.folder, int(seed.item()), log)
        log.info("rank %d: electrons [%d, %d), %d hit rows" % (info["rank"], info["first"], info["first"] + info["n"],
                                                              sum(info["hits"])))
    finally:
        dist.destroy_process_group()


def main(workdir: str = ".") -> None:
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(name)s %(levelname)s %(message)s")
    log = logging.getLogger("pytestbeam")
    log.info("Preparing simulation")
    job = load_job(workdir)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        run_job_sharded(job, log)
    else:
        run_job(job, log)


if __name__ == "__main__":
    main()
