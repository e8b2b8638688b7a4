"""Production (Philox) mode: shard invariance, determinism, and distribution parity with the CPU oracle
by two-sample KS / chi-square tests at p > 0.01 (BASELINE.json north_star, mode b).

The oracle's energy-loss table is drawn from the reference's UN-bootstrapped accept/reject pool
(main/hit.py:347-356), because the reference's own bootstrap (main/hit.py:359) fails a KS test against
itself; everything else in the oracle run is the reference's algorithm unchanged."""
import os

import numpy as np
import pytest
from scipy import stats

from oracle import cpu_oracle as co
from tests import helpers as H

pytestmark = pytest.mark.gpu
P_MIN = 0.01


def _sim(setup, **kw):
    from pytestbeam_b200.engine import Simulator
    beam, devices, materials, names = setup
    return Simulator(beam, devices, materials, **kw)


def _pool_table(beam, devices, materials, n, seed):
    """energy_lost[D,N] drawn i.i.d. from the accept/reject pool (no bootstrap)."""
    rs = np.random.RandomState(seed)
    D = len(devices)
    table = np.zeros((D, n))
    for d in range(D):
        args = co.landau_args(devices[d]["thickness"], materials[d]["Z"], materials[d]["A"], materials[d]["rho"],
                              beam["energy"] - np.mean(table[d - 1]))
        got = []
        need = n
        while need > 0:
            p = co.accept_pool(rs, args, 4_000_000)
            got.append(p[:need])
            need -= len(got[-1])
        table[d] = np.concatenate(got)
    return table


def test_philox_shard_invariance_and_determinism():
    n = 100_003
    setup = H.setup("c1", n)
    sim = _sim(setup)
    whole = sim.run(0, n, seed=99, write_tracks=True)
    again = sim.run(0, n, seed=99)
    other = sim.run(0, n, seed=100)
    for d in range(sim.D):
        assert whole.hits_numpy(d).tobytes() == again.hits_numpy(d).tobytes()
    assert whole.totals["hits"] != other.totals["hits"]
    for cuts in ([0, 50_000, n], [0, 1, 33, 65_537, n]):
        bases, parts, tt = None, [], []
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            r = sim.run(lo, hi - lo, seed=99, bases_in=bases, write_tracks=True)
            bases = r.bases_out
            parts.append([r.hits_numpy(d) for d in range(sim.D)])
            tt.append({k: r.tracks[k].cpu().numpy() for k in ("x", "energy", "time_stamp")})
        for d in range(sim.D):
            assert np.concatenate([p[d] for p in parts]).tobytes() == whole.hits_numpy(d).tobytes(), (cuts, d)
        for k in ("x", "energy", "time_stamp"):
            assert np.array_equal(np.concatenate([t[k] for t in tt]), whole.tracks[k].cpu().numpy()), k
    acc, tsum = sim.prefix(0, n, seed=99)
    assert acc == whole.totals["accepted"] and tsum == whole.totals["time_sum"]
    acc2, _ = sim.prefix(50_000, n - 50_000, seed=99)
    acc1, _ = sim.prefix(0, 50_000, seed=99)
    assert acc1 + acc2 == acc


def _first_rows(hits):
    """One row per cluster: a new cluster starts where event_number changes or the (col,row) order restarts."""
    if len(hits) == 0:
        return np.zeros(0, dtype=bool)
    ev, c, r = hits["event_number"], hits["column"].astype(np.int64), hits["row"].astype(np.int64)
    key = c * 100000 + r
    start = np.ones(len(hits), dtype=bool)
    start[1:] = (ev[1:] != ev[:-1]) | (key[1:] <= key[:-1])
    return start


def _plane_pool(job):
    args, n, seed = job
    rs = np.random.RandomState(seed)
    got, need = [], n
    while need > 0:
        p = co.accept_pool(rs, args, 4_000_000)
        got.append(p[:need])
        need -= len(got[-1])
    return np.concatenate(got)


def _pool_table_parallel(beam, devices, materials, n, seed):
    """_pool_table with one process per plane (the accept/reject pass keeps 0.8 % of its candidates: 1e6 values per plane
    are minutes on one core).  The argument of plane d uses the mean loss of plane d-1 (main/hit.py:91), which for this
    purpose is taken from a 2e5-value pilot of that plane."""
    import multiprocessing as mp
    D = len(devices)
    means, args = [0.0], []
    for d in range(D):
        a = co.landau_args(devices[d]["thickness"], materials[d]["Z"], materials[d]["A"], materials[d]["rho"],
                           beam["energy"] - means[-1])
        args.append(a)
        means.append(float(np.mean(_plane_pool((a, min(n, 20_000), seed + 1000 + d)))))
    with mp.get_context("fork").Pool(min(D, os.cpu_count() or 1)) as pool:
        rows = pool.map(_plane_pool, [(args[d], n, seed + d) for d in range(D)])
    return np.stack(rows)


def _north_star_report(name, n, seed):
    """p-values of the north star's quantities (per-plane hit position, cluster size, energy deposit, scattering angle;
    plus the event-level draws) for one configuration and one seed: CUDA production mode against the CPU oracle.

    Every comparison looks at the randomness that is NEW at its plane, so that the comparisons are independent of one
    another: the scattering angle enters as the kick a particle takes at the plane (angle after minus angle before; the
    track position downstream is a deterministic function of the kicks upstream, main/hit.py:212-215, and is therefore
    covered by them and by the hit positions), positions as one value per cluster (the rows of one cluster are not
    independent samples)."""
    beam, devices, materials, names = H.setup(name, n)
    D = len(devices)
    table = _pool_table_parallel(beam, devices, materials, n, seed + 7)
    run = co.run(beam, devices, materials, n, seed, seed + 1, table_override=table)
    sim = _sim((beam, devices, materials, names))
    res = sim.run(0, n, seed=seed + 2, write_tracks=True)
    report = {}
    g = {k: res.tracks[k].cpu().numpy().reshape(n, D) for k in ("energy", "angle_x", "angle_y", "x", "y")}
    o = {k: run.tracks[i].reshape(n, D) for i, k in enumerate(("energy", "angle_x", "angle_y", "x", "y"))}
    g_el = res.tracks["eloss"].cpu().numpy().reshape(D, n)
    g_acc = res.tracks["accepted"].cpu().numpy().astype(bool)
    ks = lambda a, b: float(stats.ks_2samp(a, b).pvalue)      # noqa: E731
    # the source: position and angle of every particle but event 0 (Q3)
    for k in ("x", "y", "angle_x", "angle_y"):
        report[f"source.{k}"] = ks(g[k][1:, 0], o[k][1:, 0])
    for d in range(D):
        nm = names[d]
        if d > 0:
            for k in ("angle_x", "angle_y"):
                gk, ok = g[k][1:, d] - g[k][1:, d - 1], o[k][1:, d] - o[k][1:, d - 1]
                gk, ok = gk[gk != 0], ok[ok != 0]            # outside the acceptance window the kick is replaced by 0
                if min(len(gk), len(ok)) > 50:
                    report[f"{nm}.kick_{k[-1]}"] = ks(gk, ok)
            a_in, b_in = int((g_el[d] > 0).sum()), int((run.eloss[d] > 0).sum())
            cnt = np.array([[a_in, n - a_in], [b_in, n - b_in]])
            if cnt.min() > 5:
                report[f"{nm}.in_acceptance"] = float(stats.chi2_contingency(cnt)[1])
        a, b = g_el[d][g_el[d] > 0], run.eloss[d][run.eloss[d] > 0]
        if min(len(a), len(b)) > 50:
            report[f"{nm}.energy_deposit"] = ks(a, b)
        gh, oh = res.hits_numpy(d), run.hits[d]
        if len(oh) > 50 and len(gh) > 50:
            report[f"{nm}.charge"] = ks(gh["charge"], oh["charge"])
            gs, os_ = _first_rows(gh), _first_rows(oh)
            report[f"{nm}.hit_column"] = ks(gh["column"][gs], oh["column"][os_])
            report[f"{nm}.hit_row"] = ks(gh["row"][gs], oh["row"][os_])
            gsz = np.diff(np.append(np.nonzero(gs)[0], len(gh)))
            osz = np.diff(np.append(np.nonzero(os_)[0], len(oh)))
            hist = np.array([np.bincount(np.minimum(gsz, 8), minlength=9)[1:],
                             np.bincount(np.minimum(osz, 8), minlength=9)[1:]])
            hist = hist[:, hist.min(axis=0) > 5]
            if hist.shape[1] > 1:
                report[f"{nm}.cluster_size"] = float(stats.chi2_contingency(hist)[1])
            nd_g = n if not devices[d]["trigger"] == "triggered" else int(g_acc.sum())
            nd_o = n if not devices[d]["trigger"] == "triggered" else int(run.variates.accepted.sum())
            cnt = np.array([[gs.sum(), nd_g - gs.sum()], [os_.sum(), nd_o - os_.sum()]])
            report[f"{nm}.clusters_per_particle"] = float(stats.chi2_contingency(cnt)[1])
    cnt = np.array([[g_acc.sum(), n - g_acc.sum()], [run.variates.accepted.sum(), n - run.variates.accepted.sum()]])
    report["accepted_fraction"] = float(stats.chi2_contingency(cnt)[1])
    g_gap = np.diff(res.tracks["time_stamp"].cpu().numpy().reshape(n, D)[:, 0])
    o_gap = np.diff(run.tracks[6].reshape(n, D)[:, 0])
    if beam["particle_rate"] < 1e7:
        report["time_gap"] = ks(g_gap, o_gap)
    else:   # few distinct values: compare the histograms
        m = int(max(g_gap.max(), o_gap.max())) + 1
        hist = np.array([np.bincount(g_gap, minlength=m), np.bincount(o_gap, minlength=m)])
        hist = hist[:, hist.min(axis=0) > 5]
        report["time_gap"] = float(stats.chi2_contingency(hist)[1])
    assert abs(g_gap.mean() / o_gap.mean() - 1) < 0.01
    return report


# the committed seed list: two independent seeds per configuration
KS_SEEDS = {"c1": (2024, 3024), "c3": (2025, 3025), "c4": (2026, 3026)}


@pytest.mark.parametrize("name", ["c1", "c3", "c4"])
def test_distributions_match_oracle(name):
    """North star, mode (b): two-sample KS / chi-square at p > 0.01 for every listed quantity, 1e6 electrons.

    A correct implementation still sees p < 0.01 in 1 % of its comparisons (the p-value of a true null is uniform), and
    there are ~60 of them per configuration, so "every p of one run above 0.01" would reject a correct kernel four times
    out of ten.  The gate applies p > 0.01 per comparison in the only way that does not depend on luck: every
    comparison must reach p > 0.01 with at least one of the two committed seeds (a true deviation fails with both; a
    correct kernel fails both with probability 1e-4 per comparison), and no single p may lie below 0.01 / m (Holm's
    bound for the m comparisons of a run: family-wise error 1 %)."""
    n = int(os.environ.get("PTB_KS_N", 1_000_000))
    reports = [_north_star_report(name, n, s) for s in KS_SEEDS[name]]
    keys = sorted(set(reports[0]) & set(reports[1]))
    print(f"\n{name}: {len(keys)} comparisons x 2 seeds at {n} electrons")
    for k in keys:
        print(f"{k:44s} p = {reports[0][k]:.4f}  {reports[1][k]:.4f}")
    for s, rep in zip(KS_SEEDS[name], reports):
        low = {k: p for k, p in rep.items() if not (p > P_MIN / len(rep))}
        assert not low, (name, s, low)
    both = {k: (reports[0][k], reports[1][k]) for k in keys if not (max(reports[0][k], reports[1][k]) > P_MIN)}
    assert not both, (name, both)
    once = sum(1 for rep in reports for p in rep.values() if not (p > P_MIN))
    assert once <= 0.05 * 2 * len(keys) + 2, (name, once)       # and chance level overall (expected: 1 % of them)


def test_source_draws_follow_their_exact_laws():
    """One-sample tests of the production generator against the laws the reference draws from (main/hit.py:235-239,
    main/device.py:56): beam position / angle normals, Poisson gaps, trigger acceptance.  N = 4e6."""
    from pytestbeam_b200 import config as cfg
    n = 4_000_000
    beam, devices, materials, names = cfg.flatten(cfg.c1(n))
    sim = _sim((beam, devices, materials, names))
    D = sim.D
    r = sim.run(0, n, seed=2026, digitise=False, write_tracks=True)
    col = lambda k: r.tracks[k].cpu().numpy().reshape(n, D)[1:, 0]      # noqa: E731  (event 0 is special, Q3)
    p = {"x": stats.kstest(col("x") / beam["sigma_x"], "norm").pvalue,
         "y": stats.kstest(col("y") / beam["sigma_y"], "norm").pvalue,
         "angle_x": stats.kstest(col("angle_x") / beam["x_disp"], "norm").pvalue,
         "angle_y": stats.kstest(col("angle_y") / beam["y_disp"], "norm").pvalue}
    gaps = np.diff(r.tracks["time_stamp"].cpu().numpy().reshape(n, D)[:, 0])
    lam = 1 / beam["particle_rate"] * 1e9
    edges = stats.poisson.ppf(np.linspace(0, 1, 41)[1:-1], lam)
    obs = np.bincount(np.searchsorted(edges, gaps, side="left"), minlength=40)
    expect = np.diff(np.concatenate([[0], stats.poisson.cdf(edges, lam), [1]])) * len(gaps)
    p["gap"] = stats.chisquare(obs, expect).pvalue
    acc = r.tracks["accepted"].cpu().numpy().astype(np.float64)
    p["accepted"] = 2 * stats.norm.sf(abs(acc.mean() - 0.7) / np.sqrt(0.21 / n))
    print(p)
    assert all(v > 0.01 / len(p) for v in p.values()), p          # Bonferroni over the six laws
    assert sum(v < 0.01 for v in p.values()) <= 1, p


@pytest.mark.parametrize("rate,lam", [(0.1, 1e10), (2e8, 5.0), (1e8, 10.0), (4e7, 25.0)])
def test_poisson_gap_samplers(rate, lam):
    """Gap laws across the samplers: alias table (lam = 5, 10, 25) and, when the law is too wide for a table
    (lam = 1e10), Hoermann's PTRS -- mean, variance and a chi-square against the exact pmf."""
    from pytestbeam_b200 import config as cfg
    n = 400_000
    setup = cfg.c1(n)
    setup["beam"]["particle_rate"] = rate
    beam, devices, materials, names = cfg.flatten(setup)
    sim = _sim((beam, devices, materials, names))
    r = sim.run(0, n, seed=77, digitise=False, write_tracks=True)
    t = r.tracks["time_stamp"].cpu().numpy().reshape(n, sim.D)[:, 0]
    gaps = np.diff(t).astype(np.float64)
    assert abs(gaps.mean() / lam - 1) < 6 / np.sqrt(lam * n)
    assert abs(gaps.var() / lam - 1) < 6 * np.sqrt(2 / n) + 3 / lam
    if lam < 1e6:
        m = int(gaps.max()) + 1
        obs = np.bincount(gaps.astype(np.int64), minlength=m)
        expect = stats.poisson.pmf(np.arange(m), lam) * len(gaps)
        keep = expect > 10
        obs_k = np.append(obs[keep], obs[~keep].sum())
        exp_k = np.append(expect[keep], len(gaps) - expect[keep].sum())
        assert stats.chisquare(obs_k, exp_k).pvalue > 1e-3
    else:
        z = (gaps - lam) / np.sqrt(lam)
        assert stats.kstest(z, "norm").pvalue > 1e-3        # Poisson(1e10) is normal to 1e-5
    acc, tsum = sim.prefix(0, n, seed=77)
    assert tsum == int(t[-1])
