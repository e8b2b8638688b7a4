"""Developer probe: is a low KS p-value of the x draw tied to a seed (structure) or a fluke (other ranges fine)?"""
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
for seed in (10, 1):
    for first in (0, n, 2 * n, 3 * n):
        r = sim.run(first, n, seed=seed, digitise=False, write_tracks=True)
        x = r.tracks["x"].cpu().numpy().reshape(n, D)[1:, 0] / 5000
        y = r.tracks["y"].cpu().numpy().reshape(n, D)[1:, 0] / 5000
        rr = x * x + y * y
        phi = (np.arctan2(y, x) / (2 * np.pi)) % 1.0
        print(seed, first, "x p=%.4f y p=%.4f  r2~Exp(2) p=%.4f  phi~U p=%.4f" % (
            stats.kstest(x, 'norm').pvalue, stats.kstest(y, 'norm').pvalue,
            stats.kstest(rr / 2, 'expon').pvalue, stats.kstest(phi, 'uniform').pvalue))
