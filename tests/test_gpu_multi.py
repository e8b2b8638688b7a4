"""N>1 on real GPUs (skipped on a single-GPU box): two ranks over NCCL reproduce the single-GPU hit tables."""
import os
import socket

import numpy as np
import pytest

from pytestbeam_b200 import config as cfg
from pytestbeam_b200 import h5min

pytestmark = pytest.mark.gpu


def _worker(rank, size, port, folder, n, seed):
    import torch
    import torch.distributed as dist
    from pytestbeam_b200 import distributed as D
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=size, device_id=torch.device("cuda", rank))
    try:
        beam, devices, materials, names = cfg.flatten(cfg.c4(n))
        D.run_sharded(beam, devices, materials, names, folder, seed, chunk=30_000)
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from pytestbeam_b200.engine import Simulator
    n, seed = 100_001, 31337
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    folder = str(tmp_path) + "/"
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, folder, n, seed)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    beam, devices, materials, names = cfg.flatten(cfg.c4(n))
    sim = Simulator(beam, devices, materials)
    one = sim.run(0, n, seed=seed)
    for d, name in enumerate(names):
        merged = h5min.read_table(folder + f"{name}_dut.h5", "Hits")
        assert merged.tobytes() == one.hits_numpy(d).tobytes(), name


def test_entry_point_under_torchrun(tmp_path):
    """python -m torch.distributed.run ... -m pytestbeam_b200.pytestbeam on 2 GPUs writes the single-GPU files."""
    import subprocess
    import sys

    import torch
    import yaml
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from pytestbeam_b200.engine import Simulator
    setup = cfg.c1(150_000)
    setup["beam"]["seed"] = 99
    setup["plotting"] = False
    (tmp_path / "setup.yml").write_text(yaml.safe_dump(setup, sort_keys=False))
    (tmp_path / "material.yml").write_text(yaml.safe_dump(cfg.default_material()))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", str(port), "-m",
                           "pytestbeam_b200.pytestbeam"], cwd=str(tmp_path), env=env, timeout=600)
    beam, devices, materials, names = cfg.flatten(setup)
    one = Simulator(beam, devices, materials).run(0, 150_000, seed=99)
    for d, name in enumerate(names):
        got = h5min.read_table(str(tmp_path / "output_data" / f"{name}_dut.h5"), "Hits")
        assert got.tobytes() == one.hits_numpy(d).tobytes(), name
