"""Developer probe: one-sample KS of the production generator's source draws against their exact laws (N = 1e7)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy import stats
from pytestbeam_b200 import config as cfg
from pytestbeam_b200.engine import Simulator
n = 10_000_000
beam, devices, materials, names = cfg.flatten(cfg.c1(n))
sim = Simulator(beam, devices, materials)
D = sim.D
for seed in (1, 2):
    r = sim.run(0, n, seed=seed, digitise=False, write_tracks=True)
    x = r.tracks["x"].cpu().numpy().reshape(n, D)[1:, 0]
    y = r.tracks["y"].cpu().numpy().reshape(n, D)[1:, 0]
    ax = r.tracks["angle_x"].cpu().numpy().reshape(n, D)[1:, 0]
    ay = r.tracks["angle_y"].cpu().numpy().reshape(n, D)[1:, 0]
    t = r.tracks["time_stamp"].cpu().numpy().reshape(n, D)[:, 0]
    gaps = np.diff(t)
    acc = r.tracks["accepted"].cpu().numpy()
    print(f"seed {seed}: x p={stats.kstest(x / 5000, 'norm').pvalue:.3f}  y p={stats.kstest(y / 5000, 'norm').pvalue:.3f}  "
          f"ax p={stats.kstest(ax / 1e-4, 'norm').pvalue:.3f}  ay p={stats.kstest(ay / 1e-4, 'norm').pvalue:.3f}  "
          f"gap mean {gaps.mean():.2f} var {gaps.var():.1f} (Poisson 200000)  accepted {acc.mean():.5f}")
    # Poisson: chi-square of the gap histogram against the exact pmf in 40 equiprobable-ish bins
    lam = 200000.0
    edges = stats.poisson.ppf(np.linspace(0, 1, 41)[1:-1], lam)
    obs = np.bincount(np.searchsorted(edges, gaps, side="left"), minlength=40)
    cdf = np.concatenate([[0], stats.poisson.cdf(edges, lam), [1]])
    exp = np.diff(cdf) * len(gaps)
    print("   gap chi2 p =", stats.chisquare(obs, exp).pvalue)
    el = r.tracks["eloss"].cpu().numpy().reshape(D, n)
    print("   eloss mean plane0 %.6f plane6 %.6f" % (el[0].mean(), el[6][el[6] > 0].mean()))
