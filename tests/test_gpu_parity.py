"""Deterministic-mode parity: CUDA kernels fed with the oracle's recorded variates vs the CPU oracle
(itself pinned bit-exactly to the unmodified reference, tests/test_oracle_golden.py) and vs the golden
fixtures produced by the reference.  IDs (event, column, row) must be bit-exact; fp64 track state within
1e-6 relative (BASELINE.json north_star); f4 charges equal up to rounding of values that differ by ulps."""
import glob
import os
import re

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

TRACK_KEYS = ("energy", "angle_x", "angle_y", "x", "y", "z")


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    return torch


def _sim(setup):
    from pytestbeam_b200.engine import Simulator
    beam, devices, materials, names = setup
    return Simulator(beam, devices, materials)


def _compare_hits(res, run, D, exact_charge_fraction=0.999):
    n_total = 0
    for d in range(D):
        got = res.hits_numpy(d)
        want = run.hits[d]
        assert len(got) == len(want), f"plane {d}: {len(got)} rows vs {len(want)}"
        assert np.array_equal(got["event_number"], want["event_number"]), f"plane {d} event_number"
        assert np.array_equal(got["column"], want["column"]), f"plane {d} column"
        assert np.array_equal(got["row"], want["row"]), f"plane {d} row"
        assert np.all(got["frame"] == 0)
        if len(want):
            same = got["charge"] == want["charge"]
            assert same.mean() >= exact_charge_fraction, f"plane {d}: only {same.mean():.5f} of f4 charges identical"
            ulp = np.abs(got["charge"].view(np.int32).astype(np.int64) - want["charge"].view(np.int32).astype(np.int64))
            assert ulp.max() <= 1, f"plane {d}: f4 charge differs by {ulp.max()} ulp"
            if "charge64" in res.slabs[d]:
                q64 = res.slabs[d]["charge64"].cpu().numpy()
                assert H.rel_err(q64, run.charge64[d]) < 1e-9
        n_total += len(want)
    return n_total


@pytest.mark.parametrize("name,n,seeds", [("c1", 20000, (1, 2)), ("c3", 20000, (3, 4)), ("c4", 20000, (5, 6)),
                                          ("c1", 1, (7, 8)), ("c1", 33, (9, 10))])
def test_fused_deterministic_vs_oracle(torch_cuda, name, n, seeds):
    setup, run = H.oracle_run(name, n, *seeds)
    sim = _sim(setup)
    D = sim.D
    res = sim.run(0, n, injected=H.injected_for(sim, run, 0, n), write_tracks=True, keep_charge64=True)
    for i, k in enumerate(TRACK_KEYS):
        got = res.tracks[k].cpu().numpy()
        assert H.rel_err(got, run.tracks[i]) < 1e-6, k
        if k in ("energy", "z"):
            assert np.array_equal(got, run.tracks[i]), k   # no transcendental on these paths
    assert np.array_equal(res.tracks["time_stamp"].cpu().numpy(), run.tracks[6])
    assert np.array_equal(res.tracks["eloss"].cpu().numpy().reshape(D, n), run.eloss)
    assert np.array_equal(res.tracks["accepted"].cpu().numpy(), run.variates.accepted)
    rows = _compare_hits(res, run, D)
    assert res.totals["hits"] == [len(h) for h in run.hits]
    assert res.totals["accepted"] == int(run.variates.accepted.sum())
    assert res.totals["pairs"] == run.census["pairs"]
    assert res.totals["inside"] == run.census["inside"]
    assert res.totals["pixels"] == run.census["pixels"]
    assert rows == sum(res.totals["hits"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "full_*.npz"))))
def test_cuda_vs_reference_golden(torch_cuda, path):
    """CUDA (injected) against the output of the unmodified reference itself (tests/golden/full_*.npz)."""
    m = re.match(r"full_(c\d)_n(\d+)_s(\d+)_(\d+)\.npz", os.path.basename(path))
    name, n, a, b = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4))
    gold = np.load(path)
    setup, run = H.oracle_run(name, n, a, b)
    sim = _sim(setup)
    res = sim.run(0, n, injected=H.injected_for(sim, run, 0, n), write_tracks=True)
    for gk, k in (("energy", "energy"), ("anglex", "angle_x"), ("angley", "angle_y"), ("x", "x"), ("y", "y")):
        assert H.rel_err(res.tracks[k].cpu().numpy(), gold[gk]) < 1e-6, k
    assert np.array_equal(res.tracks["time_stamp"].cpu().numpy(), gold["time_stamp"])
    for d, dn in enumerate(setup[3]):
        got, want = res.hits_numpy(d), gold["hits_" + dn]
        for f in ("event_number", "frame", "column", "row"):
            assert np.array_equal(got[f], want[f]), (dn, f)
        if len(want):
            ulp = np.abs(got["charge"].view(np.int32).astype(np.int64) - want["charge"].view(np.int32).astype(np.int64))
            assert ulp.max() <= 1


@pytest.mark.parametrize("name", ["c1", "c4"])
def test_shard_invariance_deterministic(torch_cuda, name):
    """Index-range shards with chained bases reproduce the unsharded hit tables byte for byte."""
    n = 10007
    setup, run = H.oracle_run(name, n, 21, 22)
    sim = _sim(setup)
    whole = sim.run(0, n, injected=H.injected_for(sim, run, 0, n), write_tracks=True)
    for cuts in ([0, 5000, n], [0, 31, 4097, 4098, n]):
        bases = None
        parts, tparts = [], []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            r = sim.run(lo, hi - lo, injected=H.injected_for(sim, run, lo, hi - lo), write_tracks=True,
                        bases_in=bases)
            bases = r.bases_out
            parts.append([r.hits_numpy(d) for d in range(sim.D)])
            tparts.append({k: v.cpu().numpy() for k, v in r.tracks.items()})
        for d in range(sim.D):
            cat = np.concatenate([p[d] for p in parts])
            assert cat.tobytes() == whole.hits_numpy(d).tobytes(), (cuts, d)
        for k in ("energy", "x", "y", "time_stamp"):
            cat = np.concatenate([t[k] for t in tparts])
            assert np.array_equal(cat, whole.tracks[k].cpu().numpy()), k


def test_digitise_from_track_table(torch_cuda):
    """PTB_TRACKS_FROM_TABLE = calc_position alone (main/device.py:91-168), fed from a hit_tables tuple."""
    import torch
    n = 15000
    setup, run = H.oracle_run("c4", n, 31, 32)
    sim = _sim(setup)
    D = sim.D
    dev = sim.device
    tracks_in = {"x": torch.from_numpy(run.tracks[3]).to(dev), "y": torch.from_numpy(run.tracks[4]).to(dev),
                 "eloss": torch.from_numpy(np.ascontiguousarray(run.eloss)).to(dev),
                 "accepted": torch.from_numpy(run.variates.accepted).to(dev)}
    from pytestbeam_b200.engine import Injected
    inj = Injected.from_numpy(dev, 0, n, dig_z=run.variates.dig_z, dig_off=run.variates.dig_off)
    res = sim.run(0, n, injected=inj, tracks_in=tracks_in, keep_charge64=True)
    _compare_hits(res, run, D, exact_charge_fraction=1.0)   # same x, y, e bits in: charges must be identical up to exp()


def test_prefix_and_pack(torch_cuda):
    n = 5000
    setup, run = H.oracle_run("c1", n, 41, 42)
    sim = _sim(setup)
    inj = H.injected_for(sim, run, 0, n)
    acc, tsum = sim.prefix(0, n, injected=inj)
    assert acc == int(run.variates.accepted.sum())
    assert tsum == int(run.variates.gap.sum())
    res = sim.run(0, n, injected=inj)
    for d in range(sim.D):
        packed = sim.pack_hits(res.slabs[d]).cpu().numpy()
        assert packed.tobytes() == res.hits_numpy(d).tobytes()
        assert packed.tobytes() == run.hits[d].tobytes() or len(run.hits[d]) > 0


def test_geometry_is_refused_like_the_oracle(torch_cuda):
    from pytestbeam_b200.engine import Simulator
    beam, devices, materials, _ = H.setup("c1", 10)
    devices[0]["z_position"] = 5
    with pytest.raises(ValueError):
        Simulator(beam, devices, materials)


def test_fine_pitch_planes_exercise_pool_rounds_and_overflow_retries(torch_cuda):
    """4 um pixels and a zero threshold: ~100 hits per particle on one plane.  The profile pool no longer holds a
    whole group (it is served in rounds), and undersized slabs / scratch rows are reported with the exact need."""
    from pytestbeam_b200 import config as cfg
    from oracle import cpu_oracle as co
    from pytestbeam_b200.engine import Simulator
    n = 3000
    setup = cfg.c1(n)
    fine = setup["devices"]["mimosa26_p3"]
    fine.update(column=5000, row=2500, column_pitch=4.0, row_pitch=4.0, threshold=0)
    setup["devices"]["mimosa26_p5"].update(column=3000, row=3000, column_pitch=6.0, row_pitch=7.0, threshold=1)
    beam, devices, materials, names = cfg.flatten(setup)
    run = co.run(beam, devices, materials, n, 71, 72)
    assert len(run.hits[2]) > 20 * n
    sim = Simulator(beam, devices, materials)
    res = sim.run(0, n, injected=H.injected_for(sim, run, 0, n), keep_charge64=True,
                  capacities=[64 * n] * len(devices))
    assert len(run.hits[2]) / (n / 32) > 512   # hundreds of hits per group and plane
    _compare_hits(res, run, len(devices))
    assert all(r >= h for r, h in zip(res.totals["reserved"], res.totals["hits"]))
    tiny = Simulator(beam, devices, materials, pool_entries=96)   # more rounds per group
    res3 = tiny.run(0, n, injected=H.injected_for(tiny, run, 0, n), capacities=[64 * n] * len(devices))
    for d in range(len(devices)):
        assert res3.hits_numpy(d).tobytes() == res.hits_numpy(d).tobytes()
    assert res.totals["pixels"] == run.census["pixels"]
    # undersized slabs (and with them the default scratch rows): the device reports the exact need, the host retries
    res2 = sim.run(0, n, injected=H.injected_for(sim, run, 0, n), capacities=[100] * len(devices))
    assert res2.totals["hits"] == res.totals["hits"] and res2.totals["reserved"] == res.totals["reserved"]
    for d in range(len(devices)):
        assert res2.hits_numpy(d).tobytes() == res.hits_numpy(d).tobytes()
    # ample slabs, scratch rows one short of the need on every plane: same retry, same tables
    short = [max(r - 1, 1) for r in res.totals["reserved"]]
    res4 = sim.run(0, n, injected=H.injected_for(sim, run, 0, n), capacities=[64 * n] * len(devices),
                   scratch_rows=short)
    for d in range(len(devices)):
        assert res4.hits_numpy(d).tobytes() == res.hits_numpy(d).tobytes()


def _random_setup(rng, n):
    """A telescope drawn at random inside the schema of setup.yml: 2..9 planes, pitches from 3 um to centimetres
    (single-pixel absorbers included), matrices from 1 x 1 to thousands of pixels, offsets of either sign up to the
    beam width, thresholds from 0 to far above any charge, both trigger modes, both materials, tilted / displaced beam,
    energies from 20 MeV to 6 GeV, rates on both branches of the Poisson sampler."""
    from pytestbeam_b200 import config as cfg
    n_planes = int(rng.integers(2, 10))
    z, devices = 0, {}
    for i in range(n_planes):
        kind = rng.integers(0, 4)
        if kind == 0:      # absorber: one big pixel
            col = row = 1
            pc = pr = float(rng.choice([20000.0, 40000.0, 5000.0]))
            thick, mat, thr = float(rng.choice([200.0, 1000.0])), "lead", 1e9
        else:
            col, row = int(rng.integers(1, 3000)), int(rng.integers(1, 3000))
            pc, pr = float(np.round(rng.uniform(3.0, 120.0), 2)), float(np.round(rng.uniform(3.0, 120.0), 2))
            thick, mat = float(rng.choice([25.0, 50.0, 100.0, 300.0])), "silicon"
            thr = float(rng.choice([0.0, 0.0, 1.0, 10.0, 100.0, 1000.0]))
        devices[f"plane{i}"] = {
            "column": col, "row": row, "column_pitch": pc, "row_pitch": pr, "thickness": thick, "z_position": z,
            "delta_x": float(np.round(rng.normal(0, 1500), 1)), "delta_y": float(np.round(rng.normal(0, 1500), 1)),
            "threshold": thr, "trigger": str(rng.choice(["triggered", "untriggered"])), "material": mat}
        z += int(rng.integers(1000, 80000))
    beam = {"energy": float(rng.choice([20.0, 100.0, 1000.0, 6000.0])), "sigma_x": float(rng.uniform(50, 6000)),
            "sigma_y": float(rng.uniform(50, 6000)), "loc_x": float(rng.normal(0, 500)), "loc_y": float(rng.normal(0, 500)),
            "x_disp": float(rng.uniform(0, 2e-3)), "y_disp": float(rng.uniform(0, 2e-3)),
            "x_angle": float(rng.normal(0, 2e-3)), "y_angle": float(rng.normal(0, 2e-3)), "nmb_particles": n,
            "particle_rate": float(rng.choice([5000.0, 2e8, 5e7]))}
    return cfg.flatten({"devices": devices, "beam": beam, "data_output": None, "plotting": False, "saving_tracks": False})


@pytest.mark.parametrize("seed", range(40))
def test_random_geometries_deterministic(torch_cuda, seed):
    """Deterministic-mode parity on telescopes drawn at random (see _random_setup): every event / column / row equal to
    the oracle's, track state within 1e-6 relative, census equal."""
    from oracle import cpu_oracle as co
    from pytestbeam_b200.engine import Simulator
    rng = np.random.default_rng(1000 + seed)
    n = 4000
    beam, devices, materials, names = _random_setup(rng, n)
    rs = np.random.RandomState(77 + seed)
    table = co.landau_table(rs, beam, devices, materials, 1 << 15)[:, :n]   # the pool of small runs can be empty (Q15)
    run = co.run(beam, devices, materials, n, 300 + seed, 400 + seed, table_override=table)
    sim = Simulator(beam, devices, materials)
    caps = [max(2 * len(h), 1000) for h in run.hits]
    res = sim.run(0, n, injected=H.injected_for(sim, run, 0, n), write_tracks=True, keep_charge64=True, capacities=caps)
    assert res.totals["error"] == 0
    _compare_hits(res, run, len(devices), exact_charge_fraction=0.995)
    for i, k in enumerate(TRACK_KEYS):
        want, got = run.tracks[i], res.tracks[k].cpu().numpy()
        ok = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), ok), k                 # NaN below the electron mass stays NaN (Q17)
        assert H.rel_err(got[ok], want[ok]) < 1e-6, k
    assert np.array_equal(res.tracks["time_stamp"].cpu().numpy(), run.tracks[6])
    for k in ("pairs", "inside", "pixels"):
        assert res.totals[k] == run.census[k], k
    # production mode on the same telescope: the tables do not depend on how the particle range is cut
    whole = sim.run(500, n, seed=seed, capacities=caps)
    cuts = [500, 500 + int(rng.integers(1, 200)), 500 + int(rng.integers(200, n - 1)), 500 + n]
    bases, parts = None, []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        r = sim.run(lo, hi - lo, seed=seed, bases_in=bases, capacities=caps)
        bases = r.bases_out
        parts.append([r.hits_numpy(d) for d in range(sim.D)])
    for d in range(sim.D):
        assert np.concatenate([p[d] for p in parts]).tobytes() == whole.hits_numpy(d).tobytes(), (cuts, names[d])


def test_clusters_wider_than_the_profile_pool(torch_cuda):
    """0.25 um pixels: practically every cluster spans more columns + rows than the 128-entry profile pool of a whole
    group holds; 1 um pixels: some do.  Such (group, plane) pairs are served by k_digitise_big (both profiles per
    pixel, no pool).  Deterministic mode must equal the oracle; in production mode the tables must not depend on which
    of the two kernels served a pair (pools of 128 and 96 entries draw the line differently)."""
    from pytestbeam_b200 import config as cfg
    from oracle import cpu_oracle as co
    from pytestbeam_b200.engine import Simulator
    n = 600
    setup = cfg.c1(n)
    setup["beam"].update(sigma_x=1500.0, sigma_y=1500.0)
    setup["devices"]["mimosa26_p2"].update(column=60000, row=40000, column_pitch=0.25, row_pitch=0.25, threshold=0)
    setup["devices"]["mimosa26_p4"].update(column=20000, row=10000, column_pitch=1.0, row_pitch=1.0, threshold=0)
    beam, devices, materials, names = cfg.flatten(setup)
    run = co.run(beam, devices, materials, n, 81, 82)
    assert len(run.hits[1]) > 2000 * n and len(run.hits[3]) > 100 * n     # thousands of hits per particle at 0.25 um
    sim = Simulator(beam, devices, materials)
    caps = [max(2 * len(h), 1000) for h in run.hits]
    res = sim.run(0, n, injected=H.injected_for(sim, run, 0, n), keep_charge64=True, capacities=caps)
    assert res.totals["error"] == 0
    _compare_hits(res, run, len(devices))
    assert res.totals["pixels"] == run.census["pixels"]
    # production mode: 128- and 96-entry pools split the pairs differently between the two kernels
    a = sim.run(1000, n, seed=5, capacities=caps)
    small = Simulator(beam, devices, materials, pool_entries=96)
    b = small.run(1000, n, seed=5, capacities=caps)
    assert a.totals["hits"][1] > 2000 * n
    for d in range(len(devices)):
        assert a.hits_numpy(d).tobytes() == b.hits_numpy(d).tobytes(), names[d]


def test_one_million_electrons_deterministic_in_chunks(torch_cuda):
    """BASELINE config 2's deterministic leg at the size the CPU oracle affords here: 1e6 electrons of the default
    telescope, fed to CUDA in four chunks of recorded variates with chained bases.  Every event / column / row must
    equal the oracle's; mismatching f4 charges are counted, not hidden."""
    n, chunk = 1_000_000, 250_000
    setup, run = H.oracle_run("c1", n, 101, 102)
    sim = _sim(setup)
    D = sim.D
    bases = None
    got = [[] for _ in range(D)]
    for lo in range(0, n, chunk):
        r = sim.run(lo, chunk, injected=H.injected_for(sim, run, lo, chunk), bases_in=bases)
        bases = r.bases_out
        for d in range(D):
            got[d].append(r.hits_numpy(d))
    rows = charge_diff = 0
    for d in range(D):
        g, w = np.concatenate(got[d]), run.hits[d]
        assert len(g) == len(w), d
        for f in ("event_number", "column", "row"):
            assert np.array_equal(g[f], w[f]), (d, f)
        ulp = np.abs(g["charge"].view(np.int32).astype(np.int64) - w["charge"].view(np.int32).astype(np.int64))
        assert ulp.max() <= 1
        charge_diff += int((ulp != 0).sum())
        rows += len(w)
    print(f"\n1e6 electrons: {rows} hit rows, IDs identical, {charge_diff} f4 charges off by one ulp")
    assert rows > 7_000_000 and charge_diff <= rows * 1e-4


def test_ten_million_electrons_deterministic_in_parallel_runs(torch_cuda):
    """The deterministic leg beyond what one sequential oracle pass affords: sixteen complete reference replays of
    625 000 electrons each (own seeds; every one is exactly what the reference computes for those seeds, event 0 and
    all), produced by the C oracle on all host cores at once (the oracle keeps no global state and its ctypes calls
    release the GIL), each fed to CUDA with its recorded variates: 1e7 electrons, 7.4e7 hit rows.  IDs must be
    identical everywhere; f4 charges that differ by one ulp are counted and reported."""
    from concurrent.futures import ThreadPoolExecutor
    import os
    runs, n = int(os.environ.get("PTB_DET_RUNS", 16)), 625_000
    setup = H.setup("c1", n)
    sim = _sim(setup)
    D = sim.D
    rows = charge_diff = 0
    with ThreadPoolExecutor(max_workers=min(runs, os.cpu_count() or 1)) as pool:
        futures = [pool.submit(lambda s=s: co_run(setup, n, 5000 + 2 * s, 5001 + 2 * s)) for s in range(runs)]
        for s, fut in enumerate(futures):
            run = fut.result()
            futures[s] = None
            r = sim.run(0, n, injected=H.injected_for(sim, run, 0, n))
            for d in range(D):
                g, w = r.hits_numpy(d), run.hits[d]
                assert len(g) == len(w), (s, d)
                for f in ("event_number", "column", "row"):
                    assert np.array_equal(g[f], w[f]), (s, d, f)
                ulp = np.abs(g["charge"].view(np.int32).astype(np.int64) - w["charge"].view(np.int32).astype(np.int64))
                assert ulp.max() <= 1
                charge_diff += int((ulp != 0).sum())
                rows += len(w)
            del run, r
    print(f"\n{runs * n} electrons in {runs} reference replays: {rows} hit rows, IDs identical, "
          f"{charge_diff} f4 charges off by one ulp ({charge_diff / rows:.2e})")
    assert rows > 7.3 * runs * n and charge_diff <= rows * 1e-4


def co_run(setup, n, seed_interp, seed_numba):
    from oracle import cpu_oracle as co
    beam, devices, materials, names = setup
    return co.run(beam, devices, materials, n, seed_interp, seed_numba)


@pytest.mark.parametrize("n_planes", [1, 2, 16, 32])
def test_plane_count_extremes(torch_cuda, n_planes):
    """One plane (nothing to propagate to), two, and the maximum the ABI supports (PTB_MAX_PLANES = 32)."""
    from pytestbeam_b200 import config as cfg
    from oracle import cpu_oracle as co
    from pytestbeam_b200.engine import Simulator
    n = 3000
    setup = cfg.default_setup(n)
    base = setup["devices"]["mimosa26_p2"]
    devs = {}
    for i in range(n_planes):
        d = dict(base)
        d["z_position"] = 0 if i == 0 else 10000 * i + 37 * (i % 3)
        d["delta_x"], d["delta_y"] = 50 * (i % 5) - 100, 30 * (i % 4)
        d["trigger"] = "triggered" if i % 3 == 2 else "untriggered"
        if i % 7 == 6:
            d.update(column=400, row=384, column_pitch=50, row_pitch=50, thickness=300, threshold=1000)
        devs[f"plane{i}"] = d
    setup["devices"] = devs
    beam, devices, materials, names = cfg.flatten(setup)
    run = co.run(beam, devices, materials, n, 91, 92)
    sim = Simulator(beam, devices, materials)
    res = sim.run(0, n, injected=H.injected_for(sim, run, 0, n), write_tracks=True, keep_charge64=True)
    assert H.rel_err(res.tracks["x"].cpu().numpy(), run.tracks[3]) < 1e-6
    assert np.array_equal(res.tracks["time_stamp"].cpu().numpy(), run.tracks[6])
    _compare_hits(res, run, n_planes)
    prod = sim.run(0, 20000, seed=3)
    assert prod.totals["error"] == 0 and sum(prod.totals["hits"]) > 0


def test_too_many_planes_is_refused(torch_cuda):
    from pytestbeam_b200 import config as cfg
    from pytestbeam_b200.engine import Simulator
    setup = cfg.default_setup(10)
    base = setup["devices"]["mimosa26_p2"]
    setup["devices"] = {f"p{i}": dict(base, z_position=1000 * i) for i in range(33)}
    beam, devices, materials, names = cfg.flatten(setup)
    with pytest.raises(ValueError):
        Simulator(beam, devices, materials)


def test_tilted_offset_beam_low_energy(torch_cuda):
    """Non-zero beam angles, offsets, unequal widths and dispersions, 40 MeV: exercises every beam scalar, the
    large-angle branch of the flight tangent and big scattering kicks."""
    from pytestbeam_b200 import config as cfg
    from oracle import cpu_oracle as co
    from pytestbeam_b200.engine import Simulator
    n = 20000
    setup = cfg.c4(n)
    setup["beam"].update(energy=40.0, sigma_x=3000, sigma_y=4000, loc_x=500, loc_y=-300, x_disp=2e-4, y_disp=5e-5,
                         x_angle=1e-3, y_angle=-2e-3, particle_rate=3.3e7)
    beam, devices, materials, names = cfg.flatten(setup)
    run = co.run(beam, devices, materials, n, 111, 112)
    sim = Simulator(beam, devices, materials)
    res = sim.run(0, n, injected=H.injected_for(sim, run, 0, n), write_tracks=True, keep_charge64=True)
    for i, k in enumerate(TRACK_KEYS):
        assert H.rel_err(res.tracks[k].cpu().numpy(), run.tracks[i]) < 1e-6, k
    assert np.array_equal(res.tracks["time_stamp"].cpu().numpy(), run.tracks[6])
    assert np.abs(run.tracks[1]).max() > 0.03125        # some angles beyond the polynomial's range
    _compare_hits(res, run, sim.D)
