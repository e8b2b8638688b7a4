#!/usr/bin/env python
"""Benchmark of the pytestbeam hot path (BASELINE.json metric: simulated electrons/s, tracked + digitised).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--chunk C] [--impl reference]

One step = one pass of the fused hot path (beam draw, 6 plane steps, Landau loss, Highland kicks, 7-plane cluster
digitisation, ordered hit collection into structure-of-arrays slabs) over one chunk of C electrons of the default
7-plane telescope, production (Philox) mode.  With the defaults (C=1e7, K=10) the timed region is BASELINE
config 2: 1e8 electrons on one B200.  N>1 (torchrun): every rank owns a contiguous index range of K*C electrons,
event-number / time bases come from one all-gather of per-rank prefixes, no data-path collective (weak scaling).

After the timed passes every run checks its own output (`shard_check`): the ranks' tables, digested with ptb_digest and
put end to end, must equal what rank 0 produces for the whole range alone; a mismatch ends the run non-zero.
--config c5 is BASELINE config 5 as one timed job (fixed total, index-range shards, HDF5 files for the first 1e7).

--impl reference times the reference's own CPU implementation on all host cores for the same workload, metric and
unit: the unmodified Numba path from baseline/_ref (tools/install_reference.py; kind "reference"), or the oracle port
(kind "port") when no reference tree is on the box.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "simulated electrons/sec (tracked+digitised)"
UNIT = "electrons/s"
WORKLOAD = "default 7-plane telescope (6 MIMOSA26 + ITkPix, triggered), production RNG"


def flops_per_particle(tot: dict, n: int, D: int) -> float:
    """Algorithmic FP64 flop count of SURVEY.md section 8(d), from the kernel's own census counters."""
    P, S, G, A, X = n, n * (D - 1), tot["pairs"], tot["inside"], tot["pixels"]
    V = 4 * P + 2 * S + D * P + 3 * G + X
    return (9 * P + 38 * S + 7 * D * P + 58 * G + 24 * A + 55 * X + 8 * V) / n


def digitise_flops_per_particle(tot: dict, n: int) -> float:
    """The terms of the SURVEY 8(d) count that belong to the digitisation kernel: 58 per digitised pair, 24 per
    in-acceptance pair, 55 per pixel evaluation, 8 per normal drawn there (3 per pair + 1 per pixel)."""
    G, A, X = tot["pairs"], tot["inside"], tot["pixels"]
    return (58 * G + 24 * A + 55 * X + 8 * (3 * G + X)) / n


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms; the samples stamped inside the timed region count."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t_begin = self.t_end = None

    def start(self):
        if os.environ.get("PTB_BENCH_NO_SAMPLER"):      # developer switch: is the sampling itself visible in the timing?
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def begin(self):
        import datetime
        self.t_begin = datetime.datetime.now()

    def end(self):
        import datetime
        self.t_end = datetime.datetime.now()

    def stop(self) -> dict:
        import datetime
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        slack = datetime.timedelta(milliseconds=25)       # one sampling period either side of the region
        inside, every, mx, reasons = [], [], None, set()
        for r in self.rows:
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f")
                clk, mx = float(r[1]), float(r[2])
            except Exception:
                continue
            every.append(clk)
            if self.t_begin and self.t_end and self.t_begin - slack <= ts <= self.t_end + slack:
                inside.append(clk)
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        use = sorted(inside or every)
        return {"sm_mhz": use[len(use) // 2] if use else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(inside), "samples_total": len(every)}


# --------------------------------------------------------------------------------------------------------------
# CPU port of the reference (oracle/) -- the only places bench.py executes anything under oracle/
# --------------------------------------------------------------------------------------------------------------

def _cpu_worker(args):
    n, seed = args
    from oracle import cpu_oracle as co
    from pytestbeam_b200 import config as cfg
    beam, devices, materials, _ = cfg.flatten(cfg.default_setup(n))
    t = time.perf_counter()
    r = co.run(beam, devices, materials, n, seed, seed + 1)
    return time.perf_counter() - t, sum(len(h) for h in r.hits)


def cpu_port_rate(sample: int, workers: int, seed: int = 1):
    """electrons/s of the CPU port on `workers` processes, each simulating `sample` electrons (distinct seeds)."""
    from oracle import cpu_oracle as co
    co.build()
    if workers == 1:
        dt, hits = _cpu_worker((sample, seed))
        return sample / dt, hits / dt
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t = time.perf_counter()
    with ctx.Pool(workers) as pool:
        res = pool.map(_cpu_worker, [(sample, seed + 17 * i) for i in range(workers)])
    wall = time.perf_counter() - t
    return sample * workers / wall, sum(h for _, h in res) / wall


def reference_legs(single_n: int, pool_n: int, workers: int, rounds: int = 1, warm_rounds: int = 0, seed: int = 1):
    """Times the reference's own Numba path (oracle/reference_shim.py, tree from baseline/_ref or /root/reference):
    one process (what the reference does: it has no threading) and `workers` forked processes with distinct seeds.
    Returns None when no reference tree / numba is present on this box."""
    from oracle import reference_shim as rs
    from pytestbeam_b200 import config as cfg
    if not rs.available():
        return None
    jit_s = rs.warm_jit(cfg.default_setup, cfg.flatten)          # throw-away N=2000 pass compiles both loops
    out = {"root": os.path.relpath(rs.reference_root(), ROOT) if rs.reference_root().startswith(ROOT)
           else rs.reference_root(), "jit_s": jit_s}
    if single_n:
        dt, rows = rs.timed_run(*cfg.flatten(cfg.default_setup(single_n)), seed)
        out["single"] = {"value": single_n / dt, "hits_per_s": rows / dt, "cores": 1, "electrons": single_n}
    if pool_n and workers > 1:
        pool = rs.ReferencePool(workers)                          # forked once: the workers inherit the warm JIT
        try:
            rates = []
            for r in range(warm_rounds + rounds):
                v = pool.rate(cfg.default_setup, cfg.flatten, pool_n, seed + 1000 * (r + 1))
                if r >= warm_rounds:
                    rates.append(v)
        finally:
            pool.close()
        out["pool"] = {"value": sum(v for v, _ in rates) / len(rates), "hits_per_s": sum(h for _, h in rates) / len(rates),
                       "cores": workers, "electrons_per_round": pool_n * workers, "rounds": len(rates),
                       "per_round": [v for v, _ in rates]}
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores.  The unmodified
    Numba path when a reference tree is present (kind "reference"), else the oracle port (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    workload = WORKLOAD + f", {args.chunk * args.steps * args.gpus:.3g} electrons"
    legs = None
    try:
        sample = args.cpu_sample or 100_000       # BASELINE config 1's size, per process and step
        legs = None if args.reference_tree == "none" else reference_legs(0, sample, workers, rounds=args.steps, warm_rounds=args.warmup) if workers > 1 else \
            reference_legs(sample, 0, 1)
    except Exception as exc:                      # a broken numba install must not take the arm down: say so, use the port
        print(f"reference tree unusable ({exc!r}); timing the oracle port instead", file=sys.stderr)
    if legs is not None:
        leg = legs.get("pool") or legs["single"]
        value, hits, kind = leg["value"], leg["hits_per_s"], "reference"
        per_step = sample * leg["cores"]
        what = (f"{leg['cores']} forked processes x {sample} electrons per step, distinct seeds, warm JIT: the unmodified "
                f"hit.tracks + device.calculate_device_hit of {legs['root']} (HDF5 writer stubbed); the reference has no "
                "threading of its own")
        port_v, port_h = cpu_port_rate(200_000, workers, seed=99)
        extra = {"port_all_cores": {"value": port_v, "hits_per_s": port_h, "cores": workers, "kind": "port",
                                    "sample": f"{workers} processes x 200000 electrons, oracle/ C + numpy port"},
                 "jit_s": legs["jit_s"]}
    else:
        sample = args.cpu_sample or 200_000
        rates = []
        for s in range(args.warmup + args.steps):
            r, h = cpu_port_rate(sample, workers, seed=1 + 1000 * s)
            if s >= args.warmup:
                rates.append((r, h))
        value = sum(r for r, _ in rates) / len(rates)
        hits = sum(h for _, h in rates) / len(rates)
        kind, per_step = "port", sample * workers
        what = (f"{workers} processes x {sample} electrons per step, distinct seeds: oracle/ C + numpy port of the "
                "reference's loops (no reference tree under baseline/_ref or /root/reference on this box)")
        extra = {}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step / value * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "hits_per_s": hits,
            "config": {"workload": workload, "electrons_per_step": per_step, "planes": 7,
                       "note": "each step is a bounded sample of the same workload (the rate does not depend on the "
                               "sample size: the reference is a sequential loop over independent electrons)"},
            "cpu_baseline": dict({"value": value, "unit": UNIT, "cores": workers, "kind": kind, "sample": what}, **extra),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=args.out, flush=True)


# --------------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------------

M64 = (1 << 64) - 1


class Job:
    """One rank's share of a benchmark job: contiguous index range [first, first + n), chunked."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from pytestbeam_b200 import config as cfg
        from pytestbeam_b200 import distributed as _dist_mod   # noqa: F401  (imported here, not inside the timed region)
        from pytestbeam_b200.engine import Simulator
        from pytestbeam_b200.pipeline import bind_host_to_gpu, shard_range
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}: launch with torchrun --nproc-per-node {args.gpus}")
        torch.cuda.set_device(self.local)
        self.numa_bound = bind_host_to_gpu(self.local) if self.world > 1 else False   # host buffers next to this rank's GPU
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.chunk, self.seed = args.chunk, args.seed
        if args.config == "c5":
            # BASELINE config 5: a fixed total (strong scaling), cut into index-range shards of whole chunks
            total = int(args.electrons)
            self.first, self.n_rank = shard_range(total // self.chunk, self.rank, self.world)
            self.first, self.n_rank = self.first * self.chunk, self.n_rank * self.chunk
            self.total = (total // self.chunk) * self.chunk
            setup_fn = cfg.c5
        else:
            self.n_rank = self.chunk * args.steps
            self.first = self.rank * self.n_rank
            self.total = self.n_rank * self.world
            setup_fn = {"default": cfg.default_setup, "c3": cfg.c3, "c4": cfg.c4}[args.config]
        self.K = self.n_rank // self.chunk
        self.setup = cfg.flatten(setup_fn(self.total))
        beam, devices, materials, self.names = self.setup
        self.sim = Simulator(beam, devices, materials)
        self.D = self.sim.D

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def bases(self):
        """(event base, time base) of this rank: its own prefix kernel + one all-gather of 2 int64 per rank."""
        from pytestbeam_b200.distributed import exchange_bases
        if self.world == 1:
            return 0, 0
        acc, tsum = self.sim.prefix(self.first, self.n_rank, seed=self.seed)
        return exchange_bases(acc, tsum, device=self.sim.device)

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.sim.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]


def digest_range(sim, first, n, chunk, seed, ev_base, t_base, pipe=None):
    """Per-plane checksums (ptb_digest) of the hit tables of particles [first, first+n), simulated in chunks with bases
    chained from (ev_base, t_base).  Returns a [D][7] list of Python ints modulo 2^64 with S2 taken relative to this
    range's first row."""
    import torch
    from pytestbeam_b200.engine import Simulator
    from pytestbeam_b200.pipeline import ChunkPipeline
    D = sim.D
    pipe = pipe or ChunkPipeline(sim, chunk, seed=seed, n_buffers=1)
    pipe.set_bases(ev_base, t_base)
    out = torch.zeros(D, 7, dtype=torch.int64, device=sim.device)
    rows = [0] * D
    s = pipe.sets[0]
    for lo in range(first, first + n, pipe.chunk):
        m = min(pipe.chunk, first + n - lo)
        pipe._launch(s, lo, m)
        tot = Simulator.decode_totals(s["totals"], D)        # one host read per chunk: this pass is not timed
        if tot["error"]:
            raise SystemExit(f"device error bits {tot['error']:#x} in the check pass")
        for d in range(D):
            sim.digest(s["slabs"][d], tot["hits"][d], rows[d], out[d])
            rows[d] += tot["hits"][d]
    torch.cuda.synchronize()
    return [[int(v) & M64 for v in row] for row in out.cpu().tolist()]


def combine_digests(per_rank):
    """Rank-local digests (S2 relative to the rank's first row) -> the digest of the concatenated table."""
    D = len(per_rank[0])
    tot = [[0] * 7 for _ in range(D)]
    for d in range(D):
        base = 0
        for dig in per_rank:
            w = dig[d]
            for j in range(6):
                tot[d][j] = (tot[d][j] + w[j]) & M64
            tot[d][6] = (tot[d][6] + w[6] + base * w[5]) & M64
            base += w[0]
    return tot


def shard_check(job, args):
    """Are the tables of the N shards, put end to end, the tables one GPU produces for the whole range?  Every rank
    digests its own range (same seed, same bases as the timed pass), rank 0 gathers and combines the digests and then
    simulates the WHOLE range [0, N*n) alone, bases chained from zero, with a different chunk size.  At N=1 the second
    pass checks the invariance under re-chunking only."""
    import torch
    sim, world, rank = job.sim, job.world, job.rank
    ev, t = job.bases()
    mine = digest_range(sim, job.first, job.n_rank, job.chunk, job.seed, ev, t)
    gathered = [mine]
    if world > 1:
        # the words are unsigned modulo 2^64; they travel as the int64 with the same bits
        me = torch.tensor([[v - (1 << 64) if v >= (1 << 63) else v for v in row] for row in mine], dtype=torch.int64,
                          device=sim.device)
        parts = [torch.zeros_like(me) for _ in range(world)]
        job.dist.all_gather(parts, me)
        gathered = [[[int(v) & M64 for v in row] for row in p.cpu().tolist()] for p in parts]
    result = None
    if rank == 0:
        sharded = combine_digests(gathered)
        limit = int(args.check_limit)
        n_check = min(job.total, limit)
        if n_check == job.total:
            alone = digest_range(sim, 0, job.total, max(32, (job.chunk * 3) // 4 + 17), job.seed, 0, 0)
            equal, how = alone == sharded, "rank 0 re-ran the whole range alone, bases chained from zero, other chunk size"
        else:
            # too large to repeat on one GPU within the time budget: compare the head of the range, which still
            # crosses every rank boundary it contains
            alone = digest_range(sim, 0, n_check, max(32, (job.chunk * 3) // 4 + 17), job.seed, 0, 0)
            equal, how = None, f"whole-range repeat skipped (> {limit:.3g} electrons)"
        result = {"equal": equal, "rows": sum(w[0] for w in sharded), "electrons": job.total, "ranks": world, "how": how,
                  "digest_words": "rows, sum event, sum column, sum row, sum charge bits, S1, S2 (order-sensitive); per plane",
                  "sharded": {job.names[d]: sharded[d] for d in range(job.D)}}
        if equal is False:
            result["alone"] = {job.names[d]: alone[d] for d in range(job.D)}
    if world > 1:
        job.dist.barrier()
    return result


def link_probe(job, nbytes: int = 1 << 29, reps: int = 4):
    """What bounds the end-to-end rate on this box: every rank copies `nbytes` from its GPU into pinned host memory --
    alone (rank 0 only: the PCIe link) and all ranks at once (what the host takes in).  GB/s; None on ranks != 0."""
    torch = job.torch
    src = torch.empty(nbytes, dtype=torch.uint8, device=job.sim.device).fill_(1)
    dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()

    def rate():
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        return reps * nbytes / (time.perf_counter() - t) / 1e9

    dst.copy_(src)
    job.sync_all()
    alone = rate() if job.rank == 0 else 0.0
    job.sync_all()
    together = rate()
    job.sync_all()
    t = torch.tensor([together], dtype=torch.float64, device=job.sim.device)
    parts = [torch.zeros_like(t) for _ in range(job.world)]
    if job.world > 1:
        job.dist.all_gather(parts, t)
    else:
        parts = [t]
    if job.rank != 0:
        return None
    per_rank = [round(float(p[0]), 1) for p in parts]
    return {"link_alone_GBps": round(alone, 1), "all_ranks_at_once_GBps": per_rank, "sum_GBps": round(sum(per_rank), 1),
            "what": "device-to-host copies of 512 MiB into pinned memory: rank 0 alone (its PCIe link), then every rank at "
                    "the same time (what the host takes in); the end-to-end rate x bytes_per_electron cannot exceed the sum"}


def entry_point_rate(n: int, seed: int):
    """The reference-shaped entry: hit.tracks + device.calculate_device_hit of the drop-in modules, files included
    (seven *_dut.h5 tables streamed to a scratch directory)."""
    import logging
    import shutil
    import tempfile
    from pytestbeam_b200 import config as cfg
    from pytestbeam_b200.main import device, hit
    beam, devices, materials, names = cfg.flatten(cfg.default_setup(n))
    beam["seed"] = seed
    log = logging.getLogger("bench-entry")
    log.setLevel(logging.WARNING)
    tmp = tempfile.mkdtemp(prefix="ptb_entry_", dir=os.environ.get("PTB_ENTRY_DIR") or None)
    try:
        t = time.perf_counter()
        tables = hit.tracks(beam, devices, materials, log)
        device.calculate_device_hit(beam, devices, tables, names, tmp + "/", log)
        dt = time.perf_counter() - t
        size = sum(os.path.getsize(os.path.join(tmp, f)) for f in os.listdir(tmp))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return {"value": n / dt, "unit": UNIT, "electrons": n, "seconds": dt, "file_bytes": size,
            "what": "hit.tracks + device.calculate_device_hit (pytestbeam_b200.main), seven streamed *_dut.h5 files written; "
                    "new Simulator, pilot runs and buffers inside the timed call, after one small untimed call in the process"}


def run_gpu(args):
    import torch

    from pytestbeam_b200.pipeline import ChunkPipeline

    job = Job(args)
    sim, D, lib = job.sim, job.D, job.sim.lib
    world, rank, chunk, seed = job.world, job.rank, job.chunk, job.seed
    K, W = job.K, args.warmup
    workload = {"default": WORKLOAD, "c5": WORKLOAD + " (BASELINE config 5: fixed total, index-range shards)"}.get(
        args.config, f"BASELINE config {args.config} (pytestbeam_b200/config.py), production RNG")
    pipe = ChunkPipeline(sim, chunk, seed=seed)
    host_pipe = ChunkPipeline(sim, chunk, seed=seed, host=True, capacities=pipe.caps, scratch_rows=pipe.rows)

    e2e_shard = {"first": job.first, "chunks": K, "policy": "equal index ranges"}

    def run_job(p, to_host):
        """The timed region: shard prefixes (N>1), then this rank's chunks."""
        if to_host:
            if world > 1:
                from pytestbeam_b200.distributed import exchange_bases
                acc, tsum = sim.prefix(e2e_shard["first"], e2e_shard["chunks"] * chunk, seed=seed)
                p.set_bases(*exchange_bases(acc, tsum, device=sim.device))
            else:
                p.set_bases(0, 0)
            return p.run_to_host(e2e_shard["first"], e2e_shard["chunks"])
        p.set_bases(*job.bases())
        p.run(job.first, K)
        return None

    sampler = ClockSampler(job.local)
    if rank == 0:
        sampler.start()          # up and sampling by the time the timed region begins

    # warm-up: W untimed steps on both paths (for N>1 including the prefix exchange, so that the NCCL communicator
    # exists before the timed region)
    if world > 1:
        job.bases()
    pipe.set_bases(0, 0)
    pipe.run(job.first, W)
    host_pipe.set_bases(0, 0)
    host_pipe.run_to_host(job.first, max(W, 2))
    job.sync_all()
    fp64_peak = lib.ptb_measure_fp64_peak(None) if rank == 0 else 0.0
    job.sync_all()

    launches0 = lib.ptb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sim.timing(True)                 # CUDA events around every k_tracks / k_digitise launch, on the launch stream
    job.sync_all()
    sampler.begin()
    e0.record()
    run_job(pipe, False)
    e1.record()
    job.sync_all()
    sampler.end()
    ms = e0.elapsed_time(e1)
    launches = lib.ptb_launch_count() - launches0
    stages, sim_launches = sim.timing_read_stages()
    trk_ms, dig_ms = stages[0], stages[1]
    sim.timing(False)
    launch_ms = dig_ms / max(sim_launches, 1)        # dominant kernel: k_digitise
    tracks_launch_ms = trk_ms / max(sim_launches, 1)
    clocks = sampler.stop() if rank == 0 else None
    err = pipe.errors()                              # every buffer set, not only the last one used
    if err:
        raise SystemExit(f"device error bits {err:#x}")
    tot = pipe.totals(0)

    # end to end through the public API: host buffers, the device->host copy of every chunk's tables inside the timed
    # region (compressed-row form, include/ptb.h PtbCompactOut; the most compressed variant the set-up fits).
    # N > 1: this pass is bound by what the host takes in over the PCIe links, and the links of one box are not alike
    # (profiles/README.md), so the same N*K chunks are cut into index-range shards in proportion to the rate every
    # rank's link reaches while all of them copy (measured here, before the clock starts); the tables do not depend
    # on where a range is cut (tests/test_gpu_production.py, shard_check)
    links = link_probe(job) if world > 1 else None
    if world > 1 and not args.equal_e2e_shards:
        try:
            from pytestbeam_b200.pipeline import weighted_shards
            rates = torch.tensor(links["all_ranks_at_once_GBps"] if rank == 0 else [0.0] * world, dtype=torch.float64,
                                 device=sim.device)
            job.dist.broadcast(rates, 0)
            shards = weighted_shards(K * world, [float(v) for v in rates.cpu()])
            e2e_shard.update(first=shards[rank][0] * chunk, chunks=shards[rank][1],
                             policy="index ranges in proportion to the measured link rates",
                             chunks_per_rank=[k for _, k in shards])
        except Exception as exc:                      # the probe is a convenience: equal shards remain correct
            print(f"link-weighted shards unavailable ({exc!r}); equal shards", file=sys.stderr)
    job.sync_all()
    t0 = time.perf_counter()
    e0.record()
    rows, d2h = run_job(host_pipe, True)
    e1.record()
    job.sync_all()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)   # wall clock: what the caller waits for
    ms, e2e_ms = job.max_over_ranks(ms, e2e_ms)
    value = job.total / (ms * 1e-3)
    hits_per_particle = sum(tot["hits"]) / chunk

    check = shard_check(job, args) if not args.no_check else None

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        F_all = flops_per_particle(tot, chunk, D)
        F = digitise_flops_per_particle(tot, chunk)    # the share of SURVEY 8(d)'s count that k_digitise executes
        alg_bytes = 18.0 * sum(tot["hits"])            # 18 B per hit row, nothing read
        traffic, ncu = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            traffic = tj["dram_bytes_per_launch"] * (chunk / tj["chunk"])
            ncu = {k: tj[k] for k in ("issue_active_pct", "lanes_active_per_instruction", "pipe_pct",
                                      "warp_instructions_per_launch") if k in tj}
            ncu["note"] = "from the committed ncu capture (profiles/), not measured in this run"
        except Exception:
            pass
        issue = None
        if ncu and "warp_instructions_per_launch" in ncu and clocks and clocks.get("sm_mhz"):
            # the roof that actually binds: warp instructions issued per second against 4 schedulers x SMs x clock
            props = torch.cuda.get_device_properties(sim.device)
            inst = ncu["warp_instructions_per_launch"] * (chunk / tj["chunk"])
            peak = 4.0 * props.multi_processor_count * clocks["sm_mhz"] * 1e6
            issue = {"achieved": inst / (launch_ms * 1e-3), "peak": peak, "unit": "warp instructions/s",
                     "frac": inst / (launch_ms * 1e-3) / peak,
                     "note": "instruction count of the committed ncu capture (profiles/traffic.json) over the live k_digitise "
                             "duration; peak = 4 issue slots per SM and clock"}
        ach_tf = F * chunk / (launch_ms * 1e-3) / 1e12
        ach_gb = alg_bytes / (launch_ms * 1e-3) / 1e9
        roof = {"bound": "fp64", "kernel": "k_digitise<PhiloxSource>", "achieved": ach_tf, "peak": fp64_peak / 1e12,
                "unit": "TFLOP/s", "frac": ach_tf / (fp64_peak / 1e12) if fp64_peak > 0 else None,
                "peak_source": "in-job DFMA probe (ptb_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
                "flops_per_electron": F, "flops_per_electron_whole_path": F_all, "launch_ms": launch_ms,
                "k_tracks_launch_ms": tracks_launch_ms, "launches_timed": sim_launches,
                "stage_ms_per_step": {"k_tracks": stages[0] / max(sim_launches, 1), "k_digitise": stages[1] / max(sim_launches, 1),
                                      "k_digitise_big+scans": stages[2] / max(sim_launches, 1),
                                      "k_gather": stages[3] / max(sim_launches, 1)},
                "share_of_step": launch_ms / (ms / K), "traffic": traffic, "traffic_over_algorithmic":
                    traffic / alg_bytes if traffic else None, "ncu": ncu, "issue": issue,
                "binding": "instruction issue (see ncu.issue_active_pct; FP64 pipe about a quarter busy): profiles/README.md",
                "hbm": {"achieved": ach_gb, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gb / hbm_peak,
                        "bytes_per_electron": alg_bytes / chunk, "peak_source": hbm_src}}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if args.config == "c5" else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload + f", {job.total:.3g} electrons", "electrons_per_step": chunk,
                           "planes": D, "hits_per_electron": hits_per_particle, "seed": seed,
                           "parallelism": f"index-range shards x{world}", "host_numa_bound": job.numa_bound,
                           "variates": "physics arithmetic in fp64; normal variates are 24-bit (fp32 special-function-unit "
                                       "Box-Muller, tails end at 6.8 sigma) except the beam position (fp64 Box-Muller on "
                                       "32-bit uniforms); Landau loss from the same 24-bit normals (ptb_device.cuh)",
                           "l2": "no inputs; 1.2 GB of hit rows written per step (> 126 MB L2)"},
                "hits_per_s": value * hits_per_particle,
                "clocks": clocks,
                "e2e": {"value": job.total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": d2h / max(e2e_shard["chunks"], 1), "rows_per_step": rows / max(e2e_shard["chunks"], 1),
                        "bytes_per_electron": d2h / max(e2e_shard["chunks"], 1) / chunk,
                        "form": {"packed": "PTB_COMPACT_PACKED", "narrow": "PTB_COMPACT_NARROW"}.get(host_pipe.form, "PTB_COMPACT"),
                        "host_ingest": links, "shards": {k: v for k, v in e2e_shard.items() if k != "first"},
                        "note": "ChunkPipeline.run_to_host: every chunk's hit tables copied to pinned host memory on a "
                                "second stream in compressed-row form (PtbCompactOut, include/ptb.h: packed = 6-byte rows "
                                "in three arrays + 4-bit counts per particle and plane + trigger mask; narrow = 7-byte "
                                "rows; wide = 8-byte rows + count bytes); ptb_expand_hits* turns it into the reference's "
                                "18-byte rows; "
                                "production mode has no per-step input arrays"},
                "gpu_launches": int(launches),
                "roofline": roof}
        if check is not None:
            line["shard_check"] = check
        if world == 1 and args.config == "default":
            if not args.no_entry:
                entry_point_rate(200_000, seed)           # imports, pilots and the page cache warm: not reported
                line["e2e_entry"] = entry_point_rate(args.entry_electrons, seed)
                # the same entry point at BASELINE config 2's size (133 bytes of file per electron), disk permitting
                import shutil
                import tempfile
                big = int(args.entry_electrons_large)
                free = shutil.disk_usage(os.environ.get("PTB_ENTRY_DIR") or tempfile.gettempdir()).free
                while big > args.entry_electrons and free < 3 * 133 * big:
                    big //= 10
                # (a scratch directory that took more than a second for the 0.27 GB of the default run is too slow a
                #  disk for 13 GB within this benchmark's minutes: the large run is left out then)
                if big > args.entry_electrons and line["e2e_entry"]["seconds"] < 1.0:
                    line["e2e_entry"]["large"] = entry_point_rate(big, seed)
            line["cpu_baseline"] = cpu_baseline_legs(args)
        print(json.dumps(line), file=args.out, flush=True)
    if world > 1:
        job.dist.barrier()
        job.dist.destroy_process_group()
    if check is not None and check["equal"] is False:
        raise SystemExit("shard_check failed: the sharded tables differ from the single-GPU tables")


def cpu_baseline_legs(args) -> dict:
    """The reference's own Numba path on this box's host cores (1 process = what the reference does; all cores as forked
    processes), with the oracle port beside it; the port alone when no reference tree is installed."""
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    port_n = args.cpu_sample or 1_000_000
    port_v, port_h = cpu_port_rate(port_n, 1)
    port = {"value": port_v, "hits_per_s": port_h, "cores": 1, "kind": "port",
            "sample": f"{port_n} electrons, single thread, oracle/ C + numpy port of the reference's loops"}
    legs = None
    try:
        if args.reference_tree != "none":
            legs = reference_legs(args.cpu_sample or 1_000_000, args.cpu_sample or 100_000, workers, rounds=2)
    except Exception as exc:
        print(f"reference tree unusable ({exc!r}); cpu_baseline falls back to the oracle port", file=sys.stderr)
    if legs is None:
        return dict(port, unit=UNIT, note="no reference tree under baseline/_ref or /root/reference on this box")
    one = legs["single"]
    out = {"value": one["value"], "unit": UNIT, "cores": 1, "kind": "reference", "hits_per_s": one["hits_per_s"],
           "sample": f"{one['electrons']} electrons, default telescope, one process, warm JIT: the unmodified hit.tracks + "
                     f"device.calculate_device_hit of {legs['root']} (HDF5 writer stubbed)",
           "jit_s": legs["jit_s"], "port": port}
    if "pool" in legs:
        pl = legs["pool"]
        out["all_cores"] = {"value": pl["value"], "hits_per_s": pl["hits_per_s"], "cores": pl["cores"], "kind": "reference",
                            "sample": f"{pl['cores']} forked processes x {pl['electrons_per_round'] // pl['cores']} "
                                      f"electrons, {pl['rounds']} rounds, distinct seeds"}
    return out


def run_c5(args):
    """BASELINE config 5 as a bench arm: the default telescope, a fixed total (default 1e10 electrons) in index-range
    shards; per-rank hit slabs; the rows of the first --keep electrons go through the host pipeline into per-rank HDF5
    part files that rank 0 concatenates, everything else is produced into the recycled device slabs and dropped
    (1.3 TB of rows).  One timed pass over the whole job; then the shard check."""
    import torch

    from pytestbeam_b200.distributed import merge_part_files, part_path
    from pytestbeam_b200.main.device import FileSink, fused_pipeline, stream_fused
    from pytestbeam_b200.pipeline import ChunkPipeline

    job = Job(args)
    sim, D, lib = job.sim, job.D, job.sim.lib
    world, rank, chunk, seed = job.world, job.rank, job.chunk, job.seed
    keep = min(int(args.keep), job.total) // chunk * chunk                 # whole chunks
    folder = args.out_dir if args.out_dir.endswith("/") else args.out_dir + "/"
    os.makedirs(folder, exist_ok=True)
    pipe = ChunkPipeline(sim, chunk, seed=seed)
    n_keep = max(0, min(job.first + job.n_rank, keep) - job.first)          # this rank's share of the kept head
    head_pipe = fused_pipeline(sim, n_keep, seed, chunk) if n_keep else None   # buffers exist before the clock starts
    sampler = ClockSampler(job.local)
    if rank == 0:
        sampler.start()
    job.bases()
    pipe.set_bases(0, 0)
    pipe.run(job.first, args.warmup)
    job.sync_all()
    launches0 = lib.ptb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    job.sync_all()
    sampler.begin()
    t0 = time.perf_counter()
    e0.record()
    ev, t = job.bases()
    sink = FileSink([part_path(folder, name, rank) for name in job.names])
    kept_rows = 0
    try:
        if n_keep:
            head = stream_fused(sim, job.first, n_keep, seed, sink, ev, t, chunk=chunk, pipe=head_pipe)
            ev, t, kept_rows = ev + head["accepted"], t + head["time_sum"], head["rows"]
    finally:
        sink.close()
    pipe.set_bases(ev, t)
    pipe.run(job.first + n_keep, (job.n_rank - n_keep) // chunk)
    e1.record()
    job.sync_all()
    if rank == 0:
        merge_part_files(folder, job.names, world)
    job.sync_all()
    wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.end()
    dev_ms = e0.elapsed_time(e1)
    launches = lib.ptb_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    err = pipe.errors()
    if err:
        raise SystemExit(f"device error bits {err:#x}")
    wall_ms, dev_ms = job.max_over_ranks(wall_ms, dev_ms)
    check = shard_check(job, args) if not args.no_check else None
    if rank == 0:
        from pytestbeam_b200 import h5min
        rows_in_files = []
        for name in job.names:
            with h5min.TableReader(f"{folder}{name}_dut.h5") as r:
                rows_in_files.append(r.nrows("Hits"))
        value = job.total / (wall_ms * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": job.K, "warmup": args.warmup,
                "ms_per_step": wall_ms / job.K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD + f" (BASELINE config 5), {job.total:.3g} electrons in index-range shards",
                           "electrons_per_step": chunk, "planes": D, "seed": seed, "parallelism": f"index-range shards x{world}",
                           "kept_electrons": keep, "sink": f"rows of the first {keep:.3g} electrons -> per-rank HDF5 part "
                           "files -> concatenated by rank 0; all other rows produced into recycled device slabs and dropped"},
                "wall_s": wall_ms * 1e-3, "device_s": dev_ms * 1e-3, "target": "1e10 electrons < 60 s on 8 B200 (BASELINE.json)",
                "rows_in_hdf5": rows_in_files, "rows_kept_this_rank": kept_rows, "clocks": clocks,
                "gpu_launches": int(launches),
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": None,
                        "note": "value IS the end-to-end figure of this arm: wall clock over prefix exchange, all chunks, "
                                "part files and the merge"}}
        if check is not None:
            line["shard_check"] = check
        print(json.dumps(line), file=args.out, flush=True)
        if not args.keep_files:
            for name in job.names:
                try:
                    os.remove(f"{folder}{name}_dut.h5")
                except OSError:
                    pass
    if world > 1:
        job.dist.barrier()
        job.dist.destroy_process_group()
    if check is not None and check["equal"] is False:
        raise SystemExit("shard_check failed: the sharded tables differ from the single-GPU tables")


def _claim_stdout():
    """Keep stdout for the one JSON line: whatever libraries print to file descriptor 1 meanwhile (NCCL's version
    banner, for one) goes to stderr.  Returns the stream to print the line to."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chunk", type=int, default=10_000_000)
    ap.add_argument("--seed", type=int, default=20261019)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="default", choices=["default", "c3", "c4", "c5"],
                    help="default = BASELINE configs 1/2; c3, c4 = the other telescopes; c5 = the sharded 1e10 target run")
    ap.add_argument("--electrons", type=float, default=1e10, help="c5: total electrons over all ranks")
    ap.add_argument("--keep", type=float, default=1e7, help="c5: electrons whose rows are written to HDF5")
    ap.add_argument("--out-dir", default=os.path.join(ROOT, "gpurun_out", "c5"), help="c5: where the HDF5 files go")
    ap.add_argument("--keep-files", action="store_true", help="c5: leave the merged *_dut.h5 files in --out-dir")
    ap.add_argument("--no-check", action="store_true", help="skip the shard check pass")
    ap.add_argument("--equal-e2e-shards", action="store_true",
                    help="N > 1: equal index ranges in the end-to-end pass too (default: in proportion to the link rates)")
    ap.add_argument("--check-limit", type=float, default=2.5e10,
                    help="largest job rank 0 repeats alone for the shard check (electrons)")
    ap.add_argument("--no-entry", action="store_true", help="skip the entry-point (files included) measurement")
    ap.add_argument("--entry-electrons", type=int, default=2_000_000, help="size of the entry-point run (setup.yml's)")
    ap.add_argument("--entry-electrons-large", type=float, default=1e8,
                    help="size of the second entry-point run (BASELINE config 2's; reduced tenfold while the scratch "
                         "directory has less than three times the files' size free)")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--reference-tree", default="auto", choices=["auto", "none"],
                    help="none: time the oracle port even when a reference tree is installed (tests)")
    args = ap.parse_args()
    args.out = _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c5":
        run_c5(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
