"""Host-side logic of the N>1 path on CPU: gloo, world_size 2 (shard ranges, prefix exchange, part-file merge)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from pytestbeam_b200 import h5min
from pytestbeam_b200.engine import HITS_DTYPE
from pytestbeam_b200.pipeline import exclusive_prefix, shard_range


def test_shard_ranges_tile_the_particles():
    for n, w in [(10, 1), (10, 3), (1_000_000_007, 8), (5, 8)]:
        pieces = [shard_range(n, r, w) for r in range(w)]
        assert pieces[0][0] == 0 and sum(m for _, m in pieces) == n
        for (a, m), (b, _) in zip(pieces[:-1], pieces[1:]):
            assert a + m == b
    vals = np.array([[3, 10], [4, 20], [5, 30]])
    assert list(exclusive_prefix(vals, 0)) == [0, 0] and list(exclusive_prefix(vals, 2)) == [7, 30]


def test_numa_binding_is_a_no_op_without_nvml():
    """No GPU driver here: the helper reports False and leaves the affinity alone."""
    from pytestbeam_b200.pipeline import bind_host_to_gpu
    before = os.sched_getaffinity(0)
    assert bind_host_to_gpu(0) is False
    assert os.sched_getaffinity(0) == before


def _reference_arm(*extra):
    import json, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "2000", *extra], capture_output=True, text=True, timeout=600)
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert out.returncode == 0 and len(lines) == 1, out.stderr[-500:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "electrons/s" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    return d


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference`: stdout carries exactly the JSON line.  With a reference tree on the box (baseline/_ref
    or /root/reference) the unmodified Numba path is what gets timed, the oracle port otherwise."""
    from oracle import reference_shim
    d = _reference_arm()
    assert d["cpu_baseline"]["kind"] == ("reference" if reference_shim.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1
    if d["cpu_baseline"]["kind"] == "reference":
        assert d["cpu_baseline"]["port_all_cores"]["value"] > 0


def test_bench_reference_arm_falls_back_to_the_port():
    d = _reference_arm("--reference-tree", "none")
    assert d["cpu_baseline"]["kind"] == "port"


def _worker(rank, size, port, folder, q):
    import torch.distributed as dist
    from pytestbeam_b200 import distributed as D
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        # every rank "counted" accepted particles and a time sum for its own range
        acc, tsum = 100 + rank, 1000 * (rank + 1)
        ev, t = D.exchange_bases(acc, tsum)
        rows = np.zeros(3 + rank, dtype=HITS_DTYPE)
        rows["event_number"] = ev + 1 + np.arange(len(rows))
        rows["column"], rows["row"], rows["charge"] = rank + 1, 7, 0.5
        h5min.write_table(D.part_path(folder, "plane", rank), "Hits", rows)
        dist.barrier()
        if rank == 0:
            D.merge_part_files(folder, ["plane"], size)
        dist.barrier()
        q.put((rank, ev, t))
    finally:
        dist.destroy_process_group()


def test_prefix_exchange_and_merge_world_size_2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    folder = str(tmp_path) + "/"
    procs = [ctx.Process(target=_worker, args=(r, 2, port, folder, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict((r, (ev, t)) for r, ev, t in (q.get(timeout=10) for _ in range(2)))
    assert got[0] == (0, 0) and got[1] == (100, 1000)
    merged = h5min.read_table(folder + "plane_dut.h5", "Hits")
    assert len(merged) == 3 + 4
    assert list(merged["column"]) == [1, 1, 1, 2, 2, 2, 2]                 # rank order = particle order
    assert list(merged["event_number"]) == [1, 2, 3, 101, 102, 103, 104]
    assert not os.path.exists(folder + "plane_dut.part000.h5")


def test_bench_combines_piecewise_digests_like_one_table():
    """bench.py's shard_check puts rank-local digests end to end (combine_digests): with the order-sensitive word S2 taken
    relative to every piece's first row, the combination must equal the digest of the whole table, modulo 2^64."""
    import importlib.util
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    M = (1 << 64) - 1
    rs = np.random.RandomState(3)

    def digest(cols, mix):
        """[rows, sum of four columns, S1 = sum m_i, S2 = sum (i+1) m_i] of one piece, S2 relative to its first row."""
        w = [len(mix)] + [int(sum(int(v) for v in c)) & M for c in cols]
        w.append(sum(int(m) for m in mix) & M)
        w.append(sum((i + 1) * int(m) for i, m in enumerate(mix)) & M)
        return w

    planes, cuts = 3, [0, 5, 5, 40, 64]                        # an empty piece included
    whole, pieces = [], [[] for _ in range(len(cuts) - 1)]
    for d in range(planes):
        n = cuts[-1]
        cols = [rs.randint(0, 1 << 62, n, dtype=np.int64).astype(np.uint64) for _ in range(4)]
        mix = rs.randint(0, 1 << 62, n, dtype=np.int64).astype(np.uint64) * np.uint64(4) + np.uint64(3)
        whole.append(digest(cols, mix))
        for p in range(len(cuts) - 1):
            lo, hi = cuts[p], cuts[p + 1]
            pieces[p].append(digest([c[lo:hi] for c in cols], mix[lo:hi]))
    assert bench.combine_digests(pieces) == whole


def test_weighted_shards_tile_the_units_in_proportion():
    from pytestbeam_b200.pipeline import weighted_shards
    s = weighted_shards(80, [11.6] * 4 + [18.1] * 4)
    assert [k for _, k in s] == [8, 8, 8, 8, 12, 12, 12, 12]
    assert [f for f, _ in s] == [0, 8, 16, 24, 32, 44, 56, 68]
    for n, w in ((10, [1, 1, 1]), (7, [0.0, 5.0]), (3, [9, 1, 1, 1]), (100, [57.2, 47.3]), (5, []), (4, [float("nan"), 1])):
        s = weighted_shards(n, w)
        assert sum(k for _, k in s) == n and all(k >= 0 for _, k in s)
        assert [f for f, _ in s] == [sum(k for _, k in s[:i]) for i in range(len(s))]
    assert weighted_shards(7, [0.0, 5.0]) == [(0, 0), (0, 7)]
    assert all(k >= 1 for _, k in weighted_shards(4, [100, 1, 1, 1]))
