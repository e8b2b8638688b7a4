"""BASELINE config 5 as SURVEY.md section 8(d) defines it: the default telescope, N electrons (default 1e10) sharded by
index range over the GPUs of one box, per-rank hit slabs, HDF5 part files for the first M electrons (default 1e7),
digest sink for the rest (1e10 electrons are 1.3 TB of hit rows: they cannot be kept).

    python tools/run_c5.py [--electrons 1e10] [--keep 1e7] [--chunk 1e7] [--out DIR]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_c5.py ...

Every rank reduces each chunk's slabs on the device to integer digests per plane (rows, sum of event numbers, of
columns, of rows, of the charges' bit patterns, all modulo 2^64); the ranks' digests are summed.  Because the
Philox stream is keyed by the particle id and the event / time bases are exchanged up front, the digests of an
R-GPU run must equal those of a 1-GPU run: rank 0 prints one JSON line with them, compare the lines of two runs.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from pytestbeam_b200 import config as cfg, h5min
from pytestbeam_b200.distributed import exchange_bases, merge_part_files, part_path
from pytestbeam_b200.engine import HITS_DTYPE, Simulator
from pytestbeam_b200.pipeline import ChunkPipeline, bind_host_to_gpu, shard_range


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--electrons", type=float, default=1e10)
    ap.add_argument("--keep", type=float, default=1e7)
    ap.add_argument("--chunk", type=float, default=1e7)
    ap.add_argument("--seed", type=int, default=20261019)
    ap.add_argument("--out", default="gpurun_out/c5/")
    a = ap.parse_args()
    n_total, keep, chunk = int(a.electrons), int(a.keep), int(a.chunk)
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    bind_host_to_gpu(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    folder = a.out if a.out.endswith("/") else a.out + "/"
    os.makedirs(folder, exist_ok=True)
    beam, devices, materials, names = cfg.flatten(cfg.c5(n_total))
    sim = Simulator(beam, devices, materials)
    D = sim.D
    lo, n = shard_range(n_total, rank, world)
    pipe = ChunkPipeline(sim, chunk, seed=a.seed, n_buffers=1)

    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    acc, tsum = sim.prefix(lo, n, seed=a.seed) if (world > 1 and n) else (0, 0)
    pipe.set_bases(*exchange_bases(acc, tsum, device=sim.device))
    dig = torch.zeros(D, 5, dtype=torch.int64, device=sim.device)        # rows, event, col, row, charge bits
    parts = [[] for _ in range(D)]
    s = pipe.sets[0]
    for first in range(lo, lo + n, chunk):
        m = min(chunk, lo + n - first)
        sim.launch(first, m, seed=a.seed, flags=pipe.flags, slabs=s["slabs"], bases_in=pipe.bases, bases_out=pipe.bases,
                   totals=s["totals"], workspace=s["ws"], scratch_rows=pipe.rows)
        rows = s["totals"][:D].clone()                                   # totals.hits
        hits = rows.cpu().tolist()                                       # the one host read per chunk
        for d in range(D):
            k, sl = hits[d], s["slabs"][d]
            dig[d, 0] += k
            dig[d, 1] += sl["event"][:k].sum()
            dig[d, 2] += (sl["col"][:k].to(torch.int64) & 0xFFFF).sum()
            dig[d, 3] += (sl["row"][:k].to(torch.int64) & 0xFFFF).sum()
            dig[d, 4] += sl["charge"][:k].view(torch.int32).to(torch.int64).sum()
            if first + m <= keep:                                        # materialise the chunks of the first `keep` electrons
                packed = sim.pack_hits({key: v[:k] for key, v in sl.items()})
                parts[d].append(np.frombuffer(packed.cpu().numpy().tobytes(), dtype=HITS_DTYPE))
        tot = Simulator.decode_totals(s["totals"], D)
        if tot["error"]:
            raise SystemExit(f"device error bits {tot['error']:#x}")
    torch.cuda.synchronize()
    for d, name in enumerate(names):
        table = np.concatenate(parts[d]) if parts[d] else np.zeros(0, dtype=HITS_DTYPE)
        h5min.write_table(part_path(folder, name, rank), "Hits", table)
    if world > 1:
        dist.all_reduce(dig)
        dist.barrier()
    if rank == 0:
        merge_part_files(folder, names, world)
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    if rank == 0:
        kept = [int(h5min.read_table(f"{folder}{name}_dut.h5", "Hits").shape[0]) for name in names]
        print(json.dumps({"config": "c5", "electrons": n_total, "n_gpus": world, "chunk": chunk, "seed": a.seed,
                          "wall_s": wall, "electrons_per_s_incl_digest_and_files": n_total / wall,
                          "rows_in_hdf5": kept, "digest": {names[d]: [int(v) for v in dig[d].cpu().tolist()] for d in range(D)}}),
              flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
