#!/usr/bin/env python
"""Benchmark of the pytestbeam hot path (BASELINE.json metric: simulated electrons/s, tracked + digitised).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--chunk C] [--impl reference]

One step = one pass of the fused hot path (beam draw, 6 plane steps, Landau loss, Highland kicks, 7-plane cluster
digitisation, ordered hit collection into structure-of-arrays slabs) over one chunk of C electrons of the default
7-plane telescope, production (Philox) mode.  With the defaults (C=1e7, K=10) the timed region is BASELINE
config 2: 1e8 electrons on one B200.  N>1 (torchrun): every rank owns a contiguous index range of K*C electrons,
event-number / time bases come from one all-gather of per-rank prefixes, no data-path collective (weak scaling).

--impl reference times the CPU port of the reference algorithm (oracle/, kind "port": the reference is Python+Numba
and cannot travel to the GPU box) on all host cores for the same workload, metric and unit.
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
        legs = reference_legs(0, sample, workers, rounds=args.steps, warm_rounds=args.warmup) if workers > 1 else \
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

def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from pytestbeam_b200 import config as cfg
    from pytestbeam_b200 import native as N
    from pytestbeam_b200.engine import Simulator
    from pytestbeam_b200.distributed import exchange_bases
    from pytestbeam_b200.pipeline import ChunkPipeline, bind_host_to_gpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    numa_bound = bind_host_to_gpu(local) if world > 1 else False   # host buffers next to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    K, W, chunk, seed = args.steps, args.warmup, args.chunk, args.seed
    setup_fn = {"default": cfg.default_setup, "c3": cfg.c3, "c4": cfg.c4}[args.config]
    workload = WORKLOAD if args.config == "default" else f"BASELINE config {args.config} (pytestbeam_b200/config.py), production RNG"
    beam, devices, materials, _ = cfg.flatten(setup_fn(chunk * K * world))
    sim = Simulator(beam, devices, materials)
    D = sim.D
    lib = sim.lib
    n_rank = chunk * K
    first = rank * n_rank            # contiguous index range of this rank
    pipe = ChunkPipeline(sim, chunk, seed=seed)
    host_pipe = ChunkPipeline(sim, chunk, seed=seed, host=True, capacities=pipe.caps)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def job(p, to_host):
        """The timed region: shard prefixes (N>1), then K chunks."""
        if world > 1:
            acc, tsum = sim.prefix(first, n_rank, seed=seed)          # this rank's own range only
            p.set_bases(*exchange_bases(acc, tsum, device=sim.device))  # one all-gather of 2 int64 per rank
        else:
            p.set_bases(0, 0)
        if to_host:
            return p.run_to_host(first, K)
        p.run(first, K)
        return None

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # up and sampling by the time the timed region begins

    # warm-up: W untimed steps on both paths (for N>1 including the prefix exchange, so that the NCCL communicator
    # exists before the timed region)
    if world > 1:
        exchange_bases(*sim.prefix(first, min(n_rank, 1 << 20), seed=seed), device=sim.device)
    pipe.set_bases(0, 0)
    pipe.run(first, W)
    host_pipe.set_bases(0, 0)
    host_pipe.run_to_host(first, max(W, 2))
    sync_all()

    fp64_peak = lib.ptb_measure_fp64_peak(None) if rank == 0 else 0.0
    sync_all()

    launches0 = lib.ptb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sim.timing(True)                 # CUDA events around every k_simulate launch, on the launch stream
    sync_all()
    sampler.begin()
    e0.record()
    job(pipe, False)
    e1.record()
    sync_all()
    sampler.end()
    ms = e0.elapsed_time(e1)
    launches = lib.ptb_launch_count() - launches0
    trk_ms, dig_ms, sim_launches = sim.timing_read()
    sim.timing(False)
    launch_ms = dig_ms / max(sim_launches, 1)        # dominant kernel: k_digitise
    tracks_launch_ms = trk_ms / max(sim_launches, 1)
    clocks = sampler.stop() if rank == 0 else None
    tot = pipe.totals(0)
    if tot["error"]:
        raise SystemExit(f"device error bits {tot['error']:#x}")

    # end to end through the public API: host buffers, D2H of every chunk's hits inside the timed region
    sync_all()
    t0 = time.perf_counter()
    e0.record()
    rows, d2h = job(host_pipe, True)
    e1.record()
    sync_all()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)   # wall clock: what the caller waits for

    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=sim.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = (float(v) for v in t.cpu())
    total_particles = n_rank * world
    value = total_particles / (ms * 1e-3)
    hits_per_particle = sum(tot["hits"]) / chunk

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
            ncu = {k: tj[k] for k in ("issue_active_pct", "lanes_active_per_instruction", "pipe_pct") if k in tj}
            ncu["note"] = "from the committed ncu capture (profiles/), not measured in this run"
        except Exception:
            pass
        ach_tf = F * chunk / (launch_ms * 1e-3) / 1e12
        ach_gb = alg_bytes / (launch_ms * 1e-3) / 1e9
        roof = {"bound": "fp64", "kernel": "k_digitise<PhiloxSource>", "achieved": ach_tf, "peak": fp64_peak / 1e12,
                "unit": "TFLOP/s", "frac": ach_tf / (fp64_peak / 1e12) if fp64_peak > 0 else None,
                "peak_source": "in-job DFMA probe (ptb_measure_fp64_peak); MEASURED_PEAKS.json has no FP64 figure",
                "flops_per_electron": F, "flops_per_electron_whole_path": F_all, "launch_ms": launch_ms,
                "k_tracks_launch_ms": tracks_launch_ms, "launches_timed": sim_launches,
                "share_of_step": launch_ms / (ms / K), "traffic": traffic, "ncu": ncu,
                "binding": "instruction issue (see ncu.issue_active_pct; FP64 pipe about a quarter busy): profiles/README.md",
                "hbm": {"achieved": ach_gb, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gb / hbm_peak,
                        "bytes_per_electron": alg_bytes / chunk, "peak_source": hbm_src}}
        cpu_sample = args.cpu_sample or 1_000_000
        cpu_val, cpu_hits = cpu_port_rate(cpu_sample, 1) if (world == 1 and args.config == "default") else (None, None)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload + f", {total_particles:.3g} electrons", "electrons_per_step": chunk,
                           "planes": D, "hits_per_electron": hits_per_particle, "seed": seed,
                           "parallelism": f"index-range shards x{world}", "host_numa_bound": numa_bound,
                           "l2": "no inputs; 1.2 GB of hit rows written per step (> 126 MB L2)"},
                "hits_per_s": value * hits_per_particle,
                "clocks": clocks,
                "e2e": {"value": total_particles / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": d2h / K, "rows_per_step": rows / K,
                        "note": "ChunkPipeline.run_to_host: SoA hit slabs copied to pinned host memory on a second "
                                "stream; production mode has no per-step input arrays"},
                "gpu_launches": int(launches),
                "roofline": roof}
        if cpu_val is not None:
            line["cpu_baseline"] = {"value": cpu_val, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{cpu_sample} electrons, default telescope, single thread "
                                              "(oracle/ C + numpy port of the reference's sequential loops; the "
                                              "reference itself, Python+Numba, ran 6.4-7.2e4 electrons/s on one core "
                                              "in the build container and cannot travel to this box)",
                                    "hits_per_s": cpu_hits}
        print(json.dumps(line), file=args.out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
    ap.add_argument("--config", default="default", choices=["default", "c3", "c4"],
                    help="telescope set-up (default = BASELINE configs 1/2/5; c3, c4 = the other two)")
    ap.add_argument("--cpu-sample", type=int, default=0)
    args = ap.parse_args()
    args.out = _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
