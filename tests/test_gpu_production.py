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


@pytest.mark.parametrize("name,seed", [("c1", 2024), ("c3", 2025), ("c4", 2026)])
def test_distributions_match_oracle(name, seed):
    n = int(os.environ.get("PTB_KS_N", 100_000))      # the suite runs 1e5; larger values for one-off checks
    beam, devices, materials, names = H.setup(name, n)
    D = len(devices)
    table = _pool_table(beam, devices, materials, n, seed + 7)
    run = co.run(beam, devices, materials, n, seed, seed + 1, table_override=table)
    sim = _sim((beam, devices, materials, names))
    res = sim.run(0, n, seed=seed + 2, write_tracks=True)
    report = []

    def check(label, p):
        report.append((label, float(p)))

    g = {k: res.tracks[k].cpu().numpy().reshape(n, D) for k in ("energy", "angle_x", "angle_y", "x", "y")}
    o = {k: run.tracks[i].reshape(n, D) for i, k in enumerate(("energy", "angle_x", "angle_y", "x", "y"))}
    g_el = res.tracks["eloss"].cpu().numpy().reshape(D, n)
    g_acc = res.tracks["accepted"].cpu().numpy().astype(bool)
    for d in range(D):
        for k in ("x", "y", "angle_x", "angle_y", "energy"):
            check(f"{names[d]}.{k}", stats.ks_2samp(g[k][1:, d], o[k][1:, d]).pvalue)
        a, b = g_el[d][g_el[d] > 0], run.eloss[d][run.eloss[d] > 0]
        check(f"{names[d]}.eloss", stats.ks_2samp(a, b).pvalue)
        # fraction of particles inside the acceptance (eloss zeroed otherwise), planes >= 1
        if d > 0:
            cnt = np.array([[len(a), n - len(a)], [len(b), n - len(b)]])
            if cnt.min() > 5:
                check(f"{names[d]}.in_acceptance", stats.chi2_contingency(cnt)[1])
        gh, oh = res.hits_numpy(d), run.hits[d]
        if len(oh) > 50 and len(gh) > 50:
            check(f"{names[d]}.charge", stats.ks_2samp(gh["charge"], oh["charge"]).pvalue)
            gs, os_ = _first_rows(gh), _first_rows(oh)
            check(f"{names[d]}.cluster_col", stats.ks_2samp(gh["column"][gs], oh["column"][os_]).pvalue)
            check(f"{names[d]}.cluster_row", stats.ks_2samp(gh["row"][gs], oh["row"][os_]).pvalue)
            # cluster-size histogram (discrete): chi-square on counts, sizes >= 6 pooled
            gsz = np.diff(np.append(np.nonzero(gs)[0], len(gh)))
            osz = np.diff(np.append(np.nonzero(os_)[0], len(oh)))
            hist = np.array([np.bincount(np.minimum(gsz, 6), minlength=7)[1:],
                             np.bincount(np.minimum(osz, 6), minlength=7)[1:]])
            hist = hist[:, hist.min(axis=0) > 5]
            if hist.shape[1] > 1:
                check(f"{names[d]}.cluster_size", stats.chi2_contingency(hist)[1])
            # hits per digitised particle
            nd_g = n if not devices[d]["trigger"] == "triggered" else int(g_acc.sum())
            nd_o = n if not devices[d]["trigger"] == "triggered" else int(run.variates.accepted.sum())
            cnt = np.array([[gs.sum(), nd_g - gs.sum()], [os_.sum(), nd_o - os_.sum()]])
            check(f"{names[d]}.clusters_per_particle", stats.chi2_contingency(cnt)[1])
    # trigger mask and time gaps
    cnt = np.array([[g_acc.sum(), n - g_acc.sum()], [run.variates.accepted.sum(), n - run.variates.accepted.sum()]])
    check("accepted_fraction", stats.chi2_contingency(cnt)[1])
    g_t = res.tracks["time_stamp"].cpu().numpy().reshape(n, D)[:, 0]
    o_t = run.tracks[6].reshape(n, D)[:, 0]
    g_gap, o_gap = np.diff(g_t), np.diff(o_t)
    if beam["particle_rate"] < 1e7:
        check("time_gap", stats.ks_2samp(g_gap, o_gap).pvalue)
    else:   # few distinct values: compare the histograms
        m = int(max(g_gap.max(), o_gap.max())) + 1
        hist = np.array([np.bincount(g_gap, minlength=m), np.bincount(o_gap, minlength=m)])
        hist = hist[:, hist.min(axis=0) > 5]
        check("time_gap", stats.chi2_contingency(hist)[1])
    assert abs(g_gap.mean() / o_gap.mean() - 1) < 0.01
    bad = [(k, p) for k, p in report if not (p > P_MIN)]
    print("\n".join(f"{k:40s} p={p:.4f}" for k, p in report))
    # with ~100 independent tests per configuration a couple below 0.01 are expected by chance alone; the gate
    # is therefore the north star's p > 0.01 applied with a Bonferroni-Holm style allowance: no p below
    # 0.01/len(report), and at most 3 % of the tests below 0.01
    assert all(p > P_MIN / len(report) for _, p in report), bad
    assert len(bad) <= max(1, int(0.03 * len(report))), bad


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
