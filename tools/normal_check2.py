"""Developer probe: p-values of the one-sample KS of x / y source draws over many seeds (N = 1e7 each)."""
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
px, py = [], []
for seed in range(10, 22):
    r = sim.run(0, n, seed=seed, digitise=False, write_tracks=True)
    x = r.tracks["x"].cpu().numpy().reshape(n, D)[1:, 0] / 5000
    y = r.tracks["y"].cpu().numpy().reshape(n, D)[1:, 0] / 5000
    px.append(stats.kstest(x, 'norm').pvalue); py.append(stats.kstest(y, 'norm').pvalue)
    print(seed, "x p=%.3f y p=%.3f  mean %.2e %.2e  var-1 %.2e %.2e" % (px[-1], py[-1], x.mean(), y.mean(), x.var() - 1, y.var() - 1))
print("KS of the p-values against uniform: x", stats.kstest(px, 'uniform').pvalue, "y", stats.kstest(py, 'uniform').pvalue)
