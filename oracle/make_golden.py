"""TEST INFRASTRUCTURE ONLY -- generate the golden fixtures under tests/golden/ by executing the
UNMODIFIED reference (``/root/reference``, build container only; it cannot travel to the GPU box).

    python -m oracle.make_golden            # writes tests/golden/*.npz, *.json

Fixtures:
  full_<cfg>_n<N>_s<a>_<b>.npz   every output of hit.tracks + device.calculate_device_hit for a small N
  digest_<cfg>_n<N>_s<a>_<b>.json  sha256 of every output array for N=1e5 (too large to commit verbatim)
  helpers.json                   known-answer values of the reference's helper functions (hex floats)

Seeds: ``a`` seeds the interpreter's numpy RNG, ``b`` Numba's MT19937 (two independent streams,
SURVEY.md H1).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import reference_shim  # noqa: E402
from pytestbeam_b200 import config as cfg  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TRACK_NAMES = ("energy", "anglex", "angley", "x", "y", "z", "time_stamp")

FULL_CASES = [("c1", 1000, 11, 22), ("c3", 1000, 33, 44), ("c4", 1000, 55, 66)]
DIGEST_CASES = [("c1", 100_000, 1, 2), ("c3", 100_000, 3, 4), ("c4", 100_000, 5, 6)]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(name, n, s_interp, s_numba):
    setup = getattr(cfg, name)(n)
    beam, devices, materials, names = cfg.flatten(setup)
    tracks, eloss, hits = reference_shim.run_reference(beam, devices, materials, names, s_interp, s_numba)
    return names, tracks, eloss, hits


def helper_kats():
    hit, device, _, _ = reference_shim.load()
    out = {"highland": [], "cloud_sigma_terms": [], "pixel_index": [], "pixel_centre": [], "gauss": [],
           "moyal_pdf": [], "noise_sigma": []}
    for thick, x0, e in [(50, 93650, 5000.0), (300, 93650, 4999.7), (1000, 5612, 100.0), (50, 93650, 0.6),
                         (100, 93650, 99.2)]:
        out["highland"].append([thick, x0, float(e).hex(), float(hit.highlander_fast_electrons(thick, x0, e)).hex()])
    # calc_cluster_radius draws a normal; pin it through a seeded stream: |sigma * z| with known z
    from numba import njit

    @njit
    def radius_with_seed(seed, energy):
        np.random.seed(seed)
        return device.calc_cluster_radius(energy)

    for e in [0.0, 0.0137, 0.083, 0.5, 0.03286]:
        z = np.random.RandomState(7).standard_normal()
        r = radius_with_seed(7, e)
        out["cloud_sigma_terms"].append([float(e).hex(), float(z).hex(), float(r).hex()])
    for x, delta, pitch, number in [(0.0, 0, 18.4, 1152), (-10.0, 100, 18.4, 1152), (5298.0, 100, 18.4, 1152),
                                    (-10700.0, 0, 18.4, 1152), (-10600.1, 0, 18.4, 576), (9999.99, -2500, 50, 400),
                                    (-30000.0, 0, 40000, 1), (19999.0, 0, 40000, 1), (20001.0, 0, 40000, 1),
                                    (-59999.0, 0, 40000, 1), (1234.5678, 6000, 33.04, 512)]:
        out["pixel_index"].append([float(x).hex(), delta, pitch, number,
                                   int(device._row_col_from_hit(x, delta, pitch, number))])
    for idx, delta, pitch, number in [(1, 0, 18.4, 1152), (577, 100, 18.4, 1152), (1152, -2500, 18.4, 1152),
                                      (0, 0, 50, 400), (-3, 6000, 50, 384), (1, 0, 40000, 1), (300, 0, 33.04, 512)]:
        out["pixel_centre"].append([idx, delta, pitch, number,
                                    float(device._hit_from_row_col(idx, delta, pitch, number)).hex()])
    for x, s in [(0.0, 20.4), (9.2, 3.3), (-18.4, 0.3989422804014327), (55.0, 41.0), (1e-3, 1.0)]:
        out["gauss"].append([float(x).hex(), float(s).hex(), float(device.gauss(x, 0, s)).hex()])
    for delta in [0.0, 0.0315, 0.0752, 0.2, 0.5]:
        for args in [(50 * 10 ** (-4), -1, 14, 24, 5000.0, 2.33, 1), (300 * 10 ** (-4), -1, 14, 24, 4999.8, 2.33, 1),
                     (1000 * 10 ** (-4), -1, 82, 207, 99.9, 11.35, 1)]:
            v = hit.landau_approx(np.array([delta]), *args)[0]
            out["moyal_pdf"].append([float(delta).hex(), [float(a).hex() for a in args], float(v).hex()])

    @njit
    def noise_with_seed(seed, rp, cp, th):
        np.random.seed(seed)
        return device.draw_noise_charge(300, rp, cp, th)

    for rp, cp, th in [(18.4, 18.4, 50), (50, 50, 300), (40000, 40000, 1000), (33.04, 33.04, 100)]:
        z = np.random.RandomState(9).standard_normal()
        out["noise_sigma"].append([rp, cp, th, float(z).hex(), float(noise_with_seed(9, rp, cp, th)).hex()])
    # whole clusters with their f8 charges (the /Hits table only keeps f4): seeded Numba stream,
    # device.py:139-155 call sequence for a single particle
    @njit
    def cluster_with_seed(seed, pc, nc, dx, x, pr, nr, dy, y, thr, th, e):
        np.random.seed(seed)
        rx = device.calc_cluster_radius(e)
        ry = device.calc_cluster_radius(e)
        c, r, q = device.calc_cluster_hits(pc, nc, dx, x, pr, nr, dy, y, rx, ry, thr * 1e-3, th, e)
        return rx, ry, np.array(c), np.array(r), np.array(q)

    out["clusters"] = []
    rs = np.random.RandomState(2024)
    geos = [(18.4, 1152, 100, 18.4, 576, 0, 100, 50), (50, 400, 0, 50, 384, 0, 1000, 300),
            (33.04, 512, -2500, 33.04, 512, 6000, 200, 100), (40000, 1, 0, 40000, 1, 0, 1e9, 1000),
            (18.4, 1152, 0, 18.4, 576, 0, 0, 50)]
    for i in range(60):
        pc, nc, dx, pr, nr, dy, thr, th = geos[i % len(geos)]
        x = float(rs.uniform(-0.55, 0.55) * pc * nc)
        y = float(rs.uniform(-0.55, 0.55) * pr * nr)
        e = float([0.0, 0.0137, 0.0329, 0.083, 0.31][i % 5] * rs.uniform(0.5, 1.5))
        seed = 1000 + i
        rx, ry, c, r, q = cluster_with_seed(seed, pc, nc, dx, x, pr, nr, dy, y, thr, th, e)
        out["clusters"].append({"seed": seed, "geo": [pc, nc, dx, pr, nr, dy, thr, th], "x": x.hex(), "y": y.hex(),
                                "e": e.hex(), "rx": float(rx).hex(), "ry": float(ry).hex(),
                                "col": [int(v) for v in c], "row": [int(v) for v in r],
                                "q": [float(v).hex() for v in q]})
    # threshold scan (main/threshold_scan.py:42-77): its module imports matplotlib / cmcrameri (absent here), so the
    # loop of create_hits is replayed around the reference's own calc_cluster_hits, under a seeded Numba stream
    @njit
    def s_curve(seed, inj_array, numb_inj, pc, pr, thr, th, nc, nr, column, row):
        np.random.seed(seed)
        res = np.zeros(len(inj_array))
        for j in range(len(inj_array)):
            energy = inj_array[j] * 1e-6 * 3.6
            hits = 0
            for i in range(numb_inj):
                c, r, q = device.calc_cluster_hits(pc, nc, column * pc, 0, pr, nr, row * pr, 0, 0, 0, thr * 1e-3, th,
                                                   energy)
                if len(q) > 0 and q[0] > 0:          # `charge > [0]` of threshold_scan.py:70 on the returned list
                    hits += 1
            res[j] = hits / numb_inj
        return res

    out["threshold_scan"] = []
    for pc, pr, thr, th, nc, nr, lo, hi, bins, ninj, seed in [(50, 50, 1000, 300, 400, 384, 950, 1050, 40, 200, 5),
                                                             (18.4, 18.4, 100, 50, 1152, 576, 60, 140, 32, 150, 6)]:
        arr = np.arange(lo, hi, (hi - lo) / bins)
        frac = s_curve(seed, arr, ninj, float(pc), float(pr), float(thr), float(th), nc, nr, 0, 0)
        out["threshold_scan"].append({"geo": [pc, pr, thr, th, nc, nr], "low": lo, "high": hi, "bins": bins,
                                      "inj": ninj, "seed": seed, "frac": [float(v) for v in frac]})
    return out


def reference_eventloop():
    """Compile the reference's own ``_eventloop_fast`` (main/plotting.py:730-764) without importing its module
    (matplotlib / cmcrameri are absent): the function node is cut out of the parsed source and jitted as is."""
    import ast
    from numba import njit
    path = os.path.join(reference_shim.reference_root(), "pytestbeam", "main", "plotting.py")
    tree = ast.parse(open(path).read())
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "_eventloop_fast")
    node.decorator_list = []
    ns = {"np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return njit(nogil=True)(ns["_eventloop_fast"])


def correlation_golden():
    """Correlation histograms of the first and last device (plot_default, main/plotting.py:47-55) for the full-output
    fixtures, computed with the reference's own event loop on the reference's own hit tables."""
    from numba import int64
    from numba.experimental import jitclass
    from oracle import cpu_oracle as co

    @jitclass([("n", int64)])
    class Proxy:
        def __init__(self):
            self.n = 0

        def update(self, k):
            self.n += k

    loop = reference_eventloop()
    for name, n, a, b in FULL_CASES:
        gold = np.load(os.path.join(GOLD, f"full_{name}_n{n}_s{a}_{b}.npz"))
        names = list(gold["names"])
        dx, dy, max_cols, max_rows = co.correlation_prepare(gold["hits_" + names[0]], gold["hits_" + names[-1]])
        xh, yh = np.zeros(max_cols, dtype=np.int32), np.zeros(max_rows, dtype=np.int32)
        xh, yh = loop(dx["event_number"], dx["row"], dx["column"], dy["event_number"], dy["row"], dy["column"], xh, yh,
                      Proxy())
        xi, yi = np.nonzero(xh), np.nonzero(yh)
        path = os.path.join(GOLD, f"corr_{name}_n{n}_s{a}_{b}.npz")
        np.savez_compressed(path, x_shape=np.array(xh.shape), y_shape=np.array(yh.shape),
                            x_idx=np.stack(xi).astype(np.int32), x_val=xh[xi], y_idx=np.stack(yi).astype(np.int32),
                            y_val=yh[yi])
        print("wrote", path, int(xh.sum()), int(yh.sum()))


def main():
    os.makedirs(GOLD, exist_ok=True)
    for name, n, a, b in FULL_CASES:
        names, tracks, eloss, hits = run_case(name, n, a, b)
        blob = {"names": np.array(names), "eloss": eloss}
        for nm, arr in zip(TRACK_NAMES, tracks):
            blob[nm] = np.asarray(arr)
        for dn in names:
            blob["hits_" + dn] = hits[dn]
        path = os.path.join(GOLD, f"full_{name}_n{n}_s{a}_{b}.npz")
        np.savez_compressed(path, **blob)
        print("wrote", path, os.path.getsize(path))
    for name, n, a, b in DIGEST_CASES:
        names, tracks, eloss, hits = run_case(name, n, a, b)
        dig = {"config": name, "n": n, "seed_interp": a, "seed_numba": b, "names": names,
               "eloss": sha(eloss), "tracks": {nm: sha(np.asarray(arr)) for nm, arr in zip(TRACK_NAMES, tracks)},
               "track_dtypes": {nm: str(np.asarray(arr).dtype) for nm, arr in zip(TRACK_NAMES, tracks)},
               "hits": {dn: {"n": int(len(hits[dn])), "sha256": sha(hits[dn]),
                             "sum_charge": float(hits[dn]["charge"].astype(np.float64).sum())} for dn in names}}
        path = os.path.join(GOLD, f"digest_{name}_n{n}_s{a}_{b}.json")
        with open(path, "w") as f:
            json.dump(dig, f, indent=1)
        print("wrote", path)
    with open(os.path.join(GOLD, "helpers.json"), "w") as f:
        json.dump(helper_kats(), f, indent=1)
    print("wrote helpers.json")
    correlation_golden()


if __name__ == "__main__":
    main()
